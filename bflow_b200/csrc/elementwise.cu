// Memory-bound companions of the convolution kernels: layout changes, InstanceNorm, GRU gates,
// correlation pooling, Bezier evaluation and convex upsampling.  Each cites the reference lines it replaces.
#include "common.cuh"

namespace bflow {

// ---------------------------------------------------------------------------------------------
// NCHW (reference layout) <-> NHWC (kernel layout)
// ---------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C_total, int HW,
                                    int c_off, int c_cnt, int dst_ld, float scale, float shift) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
        int c = c0 + cc, p = p0 + threadIdx.x;
        if (c < c_cnt && p < HW) tile[cc][threadIdx.x] = src[((size_t)n * C_total + c_off + c) * HW + p];
    }
    __syncthreads();
    for (int pp = threadIdx.y; pp < 32; pp += blockDim.y) {
        int p = p0 + pp, c = c0 + threadIdx.x;
        if (c < c_cnt && p < HW) dst[((size_t)n * HW + p) * dst_ld + c] = fmaf(tile[threadIdx.x][pp], scale, shift);
    }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW, int src_ld) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int pp = threadIdx.y; pp < 32; pp += blockDim.y) {
        int p = p0 + pp, c = c0 + threadIdx.x;
        if (c < C && p < HW) tile[pp][threadIdx.x] = src[((size_t)n * HW + p) * src_ld + c];
    }
    __syncthreads();
    for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
        int c = c0 + cc, p = p0 + threadIdx.x;
        if (c < C && p < HW) dst[((size_t)n * C + c) * HW + p] = tile[threadIdx.x][cc];
    }
}

// ---------------------------------------------------------------------------------------------
// InstanceNorm2d statistics and application (models/raft_utils/extractor.py:27-31,47-55)
// ---------------------------------------------------------------------------------------------
constexpr int PS_CHUNK = 512;   // pixels per CTA

__global__ void __launch_bounds__(256) plane_sums_kernel(const float* __restrict__ x, int ld, double* __restrict__ sums,
                                                         int HW, int C) {
    __shared__ double red[2][8][32];
    const int n = blockIdx.y;
    const int lane = threadIdx.x, ty = threadIdx.y;
    const int p0 = blockIdx.x * PS_CHUNK;
    const int p1 = min(p0 + PS_CHUNK, HW);
    for (int cb = 0; cb < C; cb += 32) {
        const int c = cb + lane;
        double s = 0.0, ss = 0.0;
        if (c < C) {
            for (int p = p0 + ty; p < p1; p += 8) {
                float v = x[((size_t)n * HW + p) * ld + c];
                s += (double)v;
                ss += (double)v * (double)v;
            }
        }
        red[0][ty][lane] = s;
        red[1][ty][lane] = ss;
        __syncthreads();
        if (ty == 0 && c < C) {
#pragma unroll
            for (int k = 1; k < 8; ++k) { s += red[0][k][lane]; ss += red[1][k][lane]; }
            atomicAdd(&sums[((size_t)n * C + c) * 2 + 0], s);
            atomicAdd(&sums[((size_t)n * C + c) * 2 + 1], ss);
        }
        __syncthreads();
    }
}

constexpr int IN_CHUNK = 256;   // pixels per CTA
constexpr int IN_MAXC = 512;

__global__ void __launch_bounds__(256) instnorm_relu_kernel(const float* __restrict__ a, int lda, const double* __restrict__ sums_a,
                                                            const float* __restrict__ r, int ldr, const double* __restrict__ sums_r,
                                                            float* __restrict__ out, int ldo, int HW, int C, float eps) {
    __shared__ float mu_a[IN_MAXC], rs_a[IN_MAXC], mu_r[IN_MAXC], rs_r[IN_MAXC];
    const int n = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double s = sums_a[((size_t)n * C + c) * 2], ss = sums_a[((size_t)n * C + c) * 2 + 1];
        double m = s / HW;
        double var = ss / HW - m * m;
        if (var < 0.0) var = 0.0;
        mu_a[c] = (float)m;
        rs_a[c] = (float)(1.0 / sqrt(var + (double)eps));
        if (sums_r != nullptr) {
            s = sums_r[((size_t)n * C + c) * 2]; ss = sums_r[((size_t)n * C + c) * 2 + 1];
            m = s / HW;
            var = ss / HW - m * m;
            if (var < 0.0) var = 0.0;
            mu_r[c] = (float)m;
            rs_r[c] = (float)(1.0 / sqrt(var + (double)eps));
        } else {
            mu_r[c] = 0.f;
            rs_r[c] = 1.f;
        }
    }
    __syncthreads();
    const int p0 = blockIdx.x * IN_CHUNK;
    const int np = min(IN_CHUNK, HW - p0);
    const int C4 = C >> 2;
    for (int idx = threadIdx.x; idx < np * C4; idx += blockDim.x) {
        const int pp = idx / C4, c = (idx - pp * C4) * 4;
        const size_t pix = (size_t)n * HW + p0 + pp;
        float4 v = *reinterpret_cast<const float4*>(a + pix * lda + c);
        float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaxf((o[j] - mu_a[c + j]) * rs_a[c + j], 0.f);
        if (r != nullptr) {
            float4 rv = *reinterpret_cast<const float4*>(r + pix * ldr + c);
            float rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j] + (rr[j] - mu_r[c + j]) * rs_r[c + j], 0.f);
        }
        *reinterpret_cast<float4*>(out + pix * ldo + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void __launch_bounds__(256) instnorm_relu16_kernel(const float* __restrict__ a, int lda, const double* __restrict__ sums_a,
                                                              const float* __restrict__ r, int ldr, const double* __restrict__ sums_r,
                                                              const void* __restrict__ r16h, const void* __restrict__ r16l, int ldr16,
                                                              float* __restrict__ out, int ldo, void* __restrict__ o16h, void* __restrict__ o16l, int ldo16,
                                                              int HW, int C, float eps, unsigned long long* tl) {
    __shared__ float mu_a[IN_MAXC], rs_a[IN_MAXC], mu_r[IN_MAXC], rs_r[IN_MAXC];
    const int n = blockIdx.y;
    tl_begin(tl);
    const double inv_hw = 1.0 / (double)HW;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        // fp64 only where cancellation needs it (E[x^2] - E[x]^2): no fp64 divide / sqrt (slow on this part and paid by every CTA)
        double s = sums_a[((size_t)n * C + c) * 2], ss = sums_a[((size_t)n * C + c) * 2 + 1];
        double m = s * inv_hw;
        double var = ss * inv_hw - m * m;
        if (var < 0.0) var = 0.0;
        mu_a[c] = (float)m;
        rs_a[c] = rsqrtf((float)(var + (double)eps));
        mu_r[c] = 0.f;
        rs_r[c] = 1.f;
        if (sums_r != nullptr) {
            s = sums_r[((size_t)n * C + c) * 2]; ss = sums_r[((size_t)n * C + c) * 2 + 1];
            m = s * inv_hw;
            var = ss * inv_hw - m * m;
            if (var < 0.0) var = 0.0;
            mu_r[c] = (float)m;
            rs_r[c] = rsqrtf((float)(var + (double)eps));
        }
    }
    __syncthreads();
    const int p0 = blockIdx.x * IN_CHUNK;
    const int np = min(IN_CHUNK, HW - p0);
    const int C4 = C >> 2;
    for (int idx = threadIdx.x; idx < np * C4; idx += blockDim.x) {
        const int pp = idx / C4, c = (idx - pp * C4) * 4;
        const size_t pix = (size_t)n * HW + p0 + pp;
        const float4 v = *reinterpret_cast<const float4*>(a + pix * lda + c);
        float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaxf((o[j] - mu_a[c + j]) * rs_a[c + j], 0.f);
        if (r != nullptr || r16h != nullptr) {
            const float4 rv = r != nullptr ? *reinterpret_cast<const float4*>(r + pix * ldr + c) : (r16l != nullptr ? load_split4(r16h, r16l, pix * ldr16 + c) : load_hi4(r16h, pix * ldr16 + c));
            const float rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j] + (rr[j] - mu_r[c + j]) * rs_r[c + j], 0.f);
        }
        if (out != nullptr) *reinterpret_cast<float4*>(out + pix * ldo + c) = make_float4(o[0], o[1], o[2], o[3]);
        if (o16h != nullptr) {
            if (o16l != nullptr) store_split4(o16h, o16l, pix * ldo16 + c, o[0], o[1], o[2], o[3]);
            else store_hi4(o16h, pix * ldo16 + c, o[0], o[1], o[2], o[3]);
        }
    }
    tl_end(tl);
}

// Eight consecutive channels of one pixel: InstanceNorm + ReLU (+ skip) and the hi / lo split.  o16l == nullptr: hi plane only.
struct In8 {
    float4 a0, a1;      // raw conv output
    float4 r0, r1;      // fp32 residual
    uint4 rh, rl;       // split residual
};
template <bool HAS_R32, bool HAS_R16>
__device__ __forceinline__ void in8_load(In8& v, const float* __restrict__ a, size_t ia, const float* __restrict__ r, size_t ir,
                                         const void* __restrict__ r16h, const void* __restrict__ r16l, size_t ir16) {
    v.a0 = __ldcs(reinterpret_cast<const float4*>(a + ia));          // streamed once: do not keep the raw tensor in L1 / L2 ahead of others
    v.a1 = __ldcs(reinterpret_cast<const float4*>(a + ia + 4));
    if (HAS_R32) {
        v.r0 = __ldg(reinterpret_cast<const float4*>(r + ir));
        v.r1 = __ldg(reinterpret_cast<const float4*>(r + ir + 4));
    }
    if (HAS_R16) {
        v.rh = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(r16h) + ir16));
        v.rl = r16l != nullptr ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(r16l) + ir16)) : make_uint4(0u, 0u, 0u, 0u);
    }
}
__device__ __forceinline__ float2 h2f(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }

// Fast path of bflow_instnorm_relu16: C % 8 == 0 and every row 16-byte aligned.  A thread owns 8 channels of a pixel (two 16-byte loads,
// one 16-byte store per output plane) and keeps two pixels in flight; a CTA covers IN_CHUNK pixels of one image.
template <bool HAS_R32, bool HAS_R16>
__global__ void __launch_bounds__(256) instnorm_relu16_v8_kernel(const float* __restrict__ a, int lda, const double* __restrict__ sums_a,
                                                                 const float* __restrict__ r, int ldr, const double* __restrict__ sums_r,
                                                                 const void* __restrict__ r16h, const void* __restrict__ r16l, int ldr16,
                                                                 float* __restrict__ out, int ldo, void* __restrict__ o16h, void* __restrict__ o16l, int ldo16,
                                                                 int HW, int C, float eps, unsigned long long* tl) {
    __shared__ float mu_a[IN_MAXC], rs_a[IN_MAXC], mu_r[IN_MAXC], rs_r[IN_MAXC];
    const int n = blockIdx.y;
    tl_begin(tl);
    const double inv_hw = 1.0 / (double)HW;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double s = sums_a[((size_t)n * C + c) * 2], ss = sums_a[((size_t)n * C + c) * 2 + 1];
        double m = s * inv_hw;
        double var = ss * inv_hw - m * m;
        if (var < 0.0) var = 0.0;
        mu_a[c] = (float)m;
        rs_a[c] = rsqrtf((float)(var + (double)eps));
        mu_r[c] = 0.f;
        rs_r[c] = 1.f;
        if (HAS_R32 && sums_r != nullptr) {
            s = sums_r[((size_t)n * C + c) * 2]; ss = sums_r[((size_t)n * C + c) * 2 + 1];
            m = s * inv_hw;
            var = ss * inv_hw - m * m;
            if (var < 0.0) var = 0.0;
            mu_r[c] = (float)m;
            rs_r[c] = rsqrtf((float)(var + (double)eps));
        }
    }
    __syncthreads();
    const int p0 = blockIdx.x * IN_CHUNK;
    const int np = min(IN_CHUNK, HW - p0);
    const int C8 = C >> 3;
    const int total = np * C8;
    auto finish = [&](const In8& v, size_t pix, int c) {
        float o[8] = {v.a0.x, v.a0.y, v.a0.z, v.a0.w, v.a1.x, v.a1.y, v.a1.z, v.a1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaxf((o[j] - mu_a[c + j]) * rs_a[c + j], 0.f);
        if (HAS_R32) {
            const float rr[8] = {v.r0.x, v.r0.y, v.r0.z, v.r0.w, v.r1.x, v.r1.y, v.r1.z, v.r1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j] + (rr[j] - mu_r[c + j]) * rs_r[c + j], 0.f);
        }
        if (HAS_R16) {
            const uint32_t hw_[4] = {v.rh.x, v.rh.y, v.rh.z, v.rh.w}, lw_[4] = {v.rl.x, v.rl.y, v.rl.z, v.rl.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 hf = h2f(hw_[j]), lf = h2f(lw_[j]);
                o[2 * j] = fmaxf(o[2 * j] + (hf.x + lf.x), 0.f);
                o[2 * j + 1] = fmaxf(o[2 * j + 1] + (hf.y + lf.y), 0.f);
            }
        }
        if (out != nullptr) {
            *reinterpret_cast<float4*>(out + pix * ldo + c) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(out + pix * ldo + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
        if (o16h != nullptr) {
            uint4 h4, l4;
            split2(o[0], o[1], h4.x, l4.x);
            split2(o[2], o[3], h4.y, l4.y);
            split2(o[4], o[5], h4.z, l4.z);
            split2(o[6], o[7], h4.w, l4.w);
            *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(o16h) + pix * ldo16 + c) = h4;
            if (o16l != nullptr) *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(o16l) + pix * ldo16 + c) = l4;
        }
    };
    for (int idx = threadIdx.x; idx < total; idx += 2 * 256) {
        const int idx1 = idx + 256;
        const int pp0 = idx / C8, c0 = (idx - pp0 * C8) * 8;
        const int pp1 = idx1 / C8, c1 = (idx1 - pp1 * C8) * 8;
        const size_t pix0 = (size_t)n * HW + p0 + pp0, pix1 = (size_t)n * HW + p0 + pp1;
        In8 v0, v1;
        in8_load<HAS_R32, HAS_R16>(v0, a, pix0 * lda + c0, r, pix0 * ldr + c0, r16h, r16l, pix0 * ldr16 + c0);
        if (idx1 < total) in8_load<HAS_R32, HAS_R16>(v1, a, pix1 * lda + c1, r, pix1 * ldr + c1, r16h, r16l, pix1 * ldr16 + c1);
        finish(v0, pix0, c0);
        if (idx1 < total) finish(v1, pix1, c1);
    }
    tl_end(tl);
}

// ---------------------------------------------------------------------------------------------
// SepConvGRU gate arithmetic (models/raft_spline/update.py:37-40,44-47)
// ---------------------------------------------------------------------------------------------
__global__ void gru_rh_kernel(const float* __restrict__ zr, int ldzr, const float* __restrict__ h, int ldh,
                              float* __restrict__ rh, int ldrh, long long rows, int C) {
    const int C4 = C >> 2;
    const long long total = rows * C4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / C4;
        const int c = (int)(idx - row * C4) * 4;
        float4 rv = *reinterpret_cast<const float4*>(zr + row * ldzr + C + c);
        float4 hv = *reinterpret_cast<const float4*>(h + row * ldh + c);
        *reinterpret_cast<float4*>(rh + row * ldrh + c) = make_float4(rv.x * hv.x, rv.y * hv.y, rv.z * hv.z, rv.w * hv.w);
    }
}

__global__ void gru_update_kernel(const float* __restrict__ zr, int ldzr, const float* __restrict__ q, int ldq,
                                  float* __restrict__ h, int ldh, long long rows, int C) {
    const int C4 = C >> 2;
    const long long total = rows * C4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / C4;
        const int c = (int)(idx - row * C4) * 4;
        float4 z = *reinterpret_cast<const float4*>(zr + row * ldzr + c);
        float4 qv = *reinterpret_cast<const float4*>(q + row * ldq + c);
        float4 hv = *reinterpret_cast<const float4*>(h + row * ldh + c);
        hv.x = (1.f - z.x) * hv.x + z.x * qv.x;
        hv.y = (1.f - z.y) * hv.y + z.y * qv.y;
        hv.z = (1.f - z.z) * hv.z + z.z * qv.z;
        hv.w = (1.f - z.w) * hv.w + z.w * qv.w;
        *reinterpret_cast<float4*>(h + row * ldh + c) = hv;
    }
}

// ---------------------------------------------------------------------------------------------
// avg_pool2d(2, 2) of correlation planes (models/raft_utils/corr.py:119)
// ---------------------------------------------------------------------------------------------
__global__ void corr_pool_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int H, int W) {
    const int Ho = H >> 1, Wo = W >> 1;
    const long long total = planes * Ho * Wo;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int xo = (int)(idx % Wo);
        const long long t = idx / Wo;
        const int yo = (int)(t % Ho);
        const long long p = t / Ho;
        const float* s = in + ((size_t)p * H + 2 * yo) * W + 2 * xo;
        out[idx] = (((s[0] + s[1]) + s[W]) + s[W + 1]) * 0.25f;
    }
}

// the same on 4x4-tiled planes: one thread per output tile row (float4 = 4 pixels of one row); its 8x2 input pixels are the
// same row pair of two horizontally adjacent input tiles
__global__ void corr_pool_tiled_kernel(const float* __restrict__ in, float* __restrict__ out, long long planes, int H, int W) {
    const int Ho = H >> 1, Wo = W >> 1;
    const int twi = (W + 3) >> 2, thi = (H + 3) >> 2, two = (Wo + 3) >> 2, tho = (Ho + 3) >> 2;
    const long long per_plane = (long long)two * tho * 4;
    const long long total = planes * per_plane;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long p = idx / per_plane;
        const int rem = (int)(idx - p * per_plane);
        const int tile = rem >> 2, rr = rem & 3;
        const int tyo = tile / two, txo = tile - tyo * two;
        const int yo = tyo * 4 + rr;
        const float* ip = in + (size_t)p * ((size_t)twi * thi * 16);
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        if (yo < Ho) {
            const int yi = 2 * yo;                       // rows yi, yi+1 share an input tile (yi is even)
            const int tyi = yi >> 2, ri = yi & 3;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int txi = 2 * txo + half;
                if (txi < twi) {
                    const float4 a = *reinterpret_cast<const float4*>(ip + ((size_t)(tyi * twi + txi) << 4) + (ri << 2));
                    const float4 b = *reinterpret_cast<const float4*>(ip + ((size_t)(tyi * twi + txi) << 4) + ((ri + 1) << 2));
                    o[2 * half + 0] = (((a.x + a.y) + b.x) + b.y) * 0.25f;
                    o[2 * half + 1] = (((a.z + a.w) + b.z) + b.w) * 0.25f;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) if (txo * 4 + j >= Wo) o[j] = 0.f;
        }
        *reinterpret_cast<float4*>(out + (size_t)p * ((size_t)two * tho * 16) + ((size_t)tile << 4) + (rr << 2)) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// Bezier evaluation at T timestamps (models/raft_spline/bezier.py:165-186)
// ---------------------------------------------------------------------------------------------
struct BezierCoef { float c[32][BFLOW_MAX_DEGREE]; };

__global__ void bezier_eval_kernel(const float* __restrict__ params, BezierCoef coef, float* __restrict__ flows,
                                   int T, int B, int degree, int HW) {
    const long long total = (long long)B * 2 * HW;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(idx % HW);
        const long long bd = idx / HW;   // b*2 + dim
        float pv[BFLOW_MAX_DEGREE];
        for (int k = 0; k < degree; ++k) pv[k] = params[((size_t)bd * degree + k) * HW + p];
        for (int t = 0; t < T; ++t) {
            float acc = 0.f;
            for (int k = 0; k < degree; ++k) acc = fmaf(coef.c[t][k], pv[k], acc);
            flows[((size_t)t * B * 2 + bd) * HW + p] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Convex 8x upsampling (models/raft_utils/utils.py:33-48)
// CTA = 4 horizontally adjacent low-res pixels x 64 sub-pixels; a warp = one sub-row i of the 4
// pixels, so every store instruction writes 32 consecutive floats of one output row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cvx_upsample_kernel(const float* __restrict__ data, int ldd, int data_nchw,
                                                           const float* __restrict__ mask, int ldm, int mask_nchw,
                                                           float* __restrict__ out, int C, int h, int w) {
    const int n = blockIdx.z, y = blockIdx.y;
    const int j = threadIdx.x & 7, px = (threadIdx.x >> 3) & 3, i = threadIdx.x >> 5;
    const int x = blockIdx.x * 4 + px;
    if (x >= w) return;
    const int sub = i * 8 + j;
    const size_t pix = ((size_t)n * h + y) * w + x;
    float m[9];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        m[k] = mask_nchw ? mask[(((size_t)n * 576 + k * 64 + sub) * h + y) * w + x] : mask[pix * ldm + k * 64 + sub];
        mx = fmaxf(mx, m[k]);
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { m[k] = expf(m[k] - mx); sum += m[k]; }
    const float inv = 1.f / sum;
    const int H8 = 8 * h, W8 = 8 * w;
    for (int c = 0; c < C; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
            float v = 0.f;
            if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
                v = data_nchw ? data[(((size_t)n * C + c) * h + yy) * w + xx]
                              : data[(((size_t)n * h + yy) * w + xx) * ldd + c];
            }
            acc = fmaf(m[k] * inv, 8.f * v, acc);
        }
        out[(((size_t)n * C + c) * H8 + 8 * y + i) * W8 + 8 * x + j] = acc;
    }
}

static inline unsigned grid_1d(long long total, int block) {
    long long g = ceil_div_ll(total, block);
    const long long cap = (long long)num_sms() * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace bflow

using namespace bflow;

extern "C" int bflow_nchw_to_nhwc(const float* src, float* dst, int N, int C_total, int H, int W, int c_off, int c_cnt,
                                  int dst_ld, float scale, float shift, void* stream) {
    BFLOW_REQUIRE(src != nullptr && dst != nullptr, "nchw_to_nhwc: null tensor");
    BFLOW_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && c_cnt > 0 && c_off >= 0 && c_off + c_cnt <= C_total && dst_ld >= c_cnt,
                  "nchw_to_nhwc: bad shape");
    dim3 grid(ceil_div(H * W, 32), ceil_div(c_cnt, 32), N), block(32, 8);
    nchw_to_nhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, dst, C_total, H * W, c_off, c_cnt, dst_ld, scale, shift);
    return check_launch("bflow_nchw_to_nhwc");
}

extern "C" int bflow_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, int src_ld, void* stream) {
    BFLOW_REQUIRE(src != nullptr && dst != nullptr, "nhwc_to_nchw: null tensor");
    BFLOW_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && C > 0 && src_ld >= C, "nhwc_to_nchw: bad shape");
    dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), N), block(32, 8);
    nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, dst, C, H * W, src_ld);
    return check_launch("bflow_nhwc_to_nchw");
}

extern "C" int bflow_plane_sums(const float* x, int ld, double* sums, int N, int HW, int C, void* stream) {
    BFLOW_REQUIRE(x != nullptr && sums != nullptr, "plane_sums: null tensor");
    BFLOW_REQUIRE(N > 0 && N <= 65535 && HW > 0 && C > 0 && ld >= C, "plane_sums: bad shape");
    dim3 grid(ceil_div(HW, PS_CHUNK), N), block(32, 8);
    plane_sums_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, ld, sums, HW, C);
    return check_launch("bflow_plane_sums");
}

extern "C" int bflow_instnorm_relu(const float* a, int lda, const double* sums_a, const float* r, int ldr, const double* sums_r,
                                   float* out, int ldo, int N, int HW, int C, float eps, void* stream) {
    BFLOW_REQUIRE(a != nullptr && sums_a != nullptr && out != nullptr, "instnorm: null tensor");
    BFLOW_REQUIRE(N > 0 && N <= 65535 && HW > 0 && C > 0 && C <= IN_MAXC && C % 4 == 0, "instnorm: bad shape (C%4==0, C<=512)");
    BFLOW_REQUIRE(lda >= C && lda % 4 == 0 && ldo >= C && ldo % 4 == 0 && aligned16(a) && aligned16(out), "instnorm: alignment");
    BFLOW_REQUIRE(r == nullptr || (ldr >= C && ldr % 4 == 0 && aligned16(r)), "instnorm: residual alignment");
    BFLOW_REQUIRE(r != nullptr || sums_r == nullptr, "instnorm: residual sums without residual");
    dim3 grid(ceil_div(HW, IN_CHUNK), N);
    instnorm_relu_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, lda, sums_a, r, ldr, sums_r, out, ldo, HW, C, eps);
    return check_launch("bflow_instnorm_relu");
}

extern "C" int bflow_instnorm_relu16(const float* a, int lda, const double* sums_a, const float* r, int ldr, const double* sums_r,
                                     const void* r16_hi, const void* r16_lo, int ldr16, float* out, int ldo, void* out16_hi, void* out16_lo,
                                     int ldo16, int N, int HW, int C, float eps, void* stream) {
    BFLOW_REQUIRE(a != nullptr && sums_a != nullptr && (out != nullptr || out16_hi != nullptr), "instnorm16: null tensor");
    BFLOW_REQUIRE(N > 0 && N <= 65535 && HW > 0 && C > 0 && C <= IN_MAXC && C % 4 == 0, "instnorm16: bad shape (C%4==0, C<=512)");
    BFLOW_REQUIRE(lda >= C && lda % 4 == 0 && aligned16(a), "instnorm16: alignment");
    BFLOW_REQUIRE(out == nullptr || (ldo >= C && ldo % 4 == 0 && aligned16(out)), "instnorm16: fp32 output alignment");
    BFLOW_REQUIRE(out16_hi == nullptr || (ldo16 >= C && ldo16 % 4 == 0), "instnorm16: split output");
    BFLOW_REQUIRE(r == nullptr || (r16_hi == nullptr && ldr >= C && ldr % 4 == 0 && aligned16(r)), "instnorm16: residual alignment");
    BFLOW_REQUIRE(r16_hi == nullptr || (ldr16 >= C && ldr16 % 4 == 0 && sums_r == nullptr), "instnorm16: split residual");
    BFLOW_REQUIRE(r != nullptr || sums_r == nullptr, "instnorm16: residual sums without fp32 residual");
    dim3 grid(ceil_div(HW, IN_CHUNK), N);
    unsigned long long* tls = timeline_next_slot("instnorm_relu16");
    cudaStream_t st = (cudaStream_t)stream;
    const bool v8 = C % 8 == 0 && lda % 4 == 0 && (out == nullptr || ldo % 4 == 0) &&
                    (out16_hi == nullptr || (ldo16 % 8 == 0 && aligned16(out16_hi) && (out16_lo == nullptr || aligned16(out16_lo)))) &&
                    (r16_hi == nullptr || (ldr16 % 8 == 0 && aligned16(r16_hi) && (r16_lo == nullptr || aligned16(r16_lo))));
    if (v8 && r == nullptr && r16_hi == nullptr)
        instnorm_relu16_v8_kernel<false, false><<<grid, 256, 0, st>>>(a, lda, sums_a, r, ldr, sums_r, r16_hi, r16_lo, ldr16, out, ldo, out16_hi, out16_lo, ldo16, HW, C, eps, tls);
    else if (v8 && r != nullptr)
        instnorm_relu16_v8_kernel<true, false><<<grid, 256, 0, st>>>(a, lda, sums_a, r, ldr, sums_r, r16_hi, r16_lo, ldr16, out, ldo, out16_hi, out16_lo, ldo16, HW, C, eps, tls);
    else if (v8)
        instnorm_relu16_v8_kernel<false, true><<<grid, 256, 0, st>>>(a, lda, sums_a, r, ldr, sums_r, r16_hi, r16_lo, ldr16, out, ldo, out16_hi, out16_lo, ldo16, HW, C, eps, tls);
    else
        instnorm_relu16_kernel<<<grid, 256, 0, st>>>(a, lda, sums_a, r, ldr, sums_r, r16_hi, r16_lo, ldr16, out, ldo, out16_hi, out16_lo, ldo16, HW, C, eps, tls);
    return check_launch("bflow_instnorm_relu16");
}

extern "C" int bflow_gru_rh(const float* zr, int ldzr, const float* h, int ldh, float* rh, int ldrh, long long rows, int C, void* stream) {
    BFLOW_REQUIRE(zr != nullptr && h != nullptr && rh != nullptr, "gru_rh: null tensor");
    BFLOW_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && ldzr >= 2 * C && ldh >= C && ldrh >= C, "gru_rh: bad shape");
    BFLOW_REQUIRE(ldzr % 4 == 0 && ldh % 4 == 0 && ldrh % 4 == 0 && aligned16(zr) && aligned16(h) && aligned16(rh), "gru_rh: alignment");
    gru_rh_kernel<<<grid_1d(rows * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(zr, ldzr, h, ldh, rh, ldrh, rows, C);
    return check_launch("bflow_gru_rh");
}

extern "C" int bflow_gru_update(const float* zr, int ldzr, const float* q, int ldq, float* h, int ldh, long long rows, int C, void* stream) {
    BFLOW_REQUIRE(zr != nullptr && q != nullptr && h != nullptr, "gru_update: null tensor");
    BFLOW_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && ldzr >= C && ldq >= C && ldh >= C, "gru_update: bad shape");
    BFLOW_REQUIRE(ldzr % 4 == 0 && ldq % 4 == 0 && ldh % 4 == 0 && aligned16(zr) && aligned16(q) && aligned16(h), "gru_update: alignment");
    gru_update_kernel<<<grid_1d(rows * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(zr, ldzr, q, ldq, h, ldh, rows, C);
    return check_launch("bflow_gru_update");
}

extern "C" int bflow_corr_pool(const float* in, float* out, long long planes, int H, int W, void* stream) {
    BFLOW_REQUIRE(in != nullptr && out != nullptr, "corr_pool: null tensor");
    BFLOW_REQUIRE(planes > 0 && H >= 2 && W >= 2, "corr_pool: bad shape");
    corr_pool_kernel<<<grid_1d(planes * (H / 2) * (W / 2), 256), 256, 0, (cudaStream_t)stream>>>(in, out, planes, H, W);
    return check_launch("bflow_corr_pool");
}

extern "C" int bflow_corr_pool_tiled(const float* in, float* out, long long planes, int H, int W, void* stream) {
    BFLOW_REQUIRE(in != nullptr && out != nullptr, "corr_pool_tiled: null tensor");
    BFLOW_REQUIRE(planes > 0 && H >= 2 && W >= 2 && aligned16(in) && aligned16(out), "corr_pool_tiled: bad shape");
    const long long total = planes * (((H / 2 + 3) / 4) * ((W / 2 + 3) / 4)) * 4;
    corr_pool_tiled_kernel<<<grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, planes, H, W);
    return check_launch("bflow_corr_pool_tiled");
}

extern "C" int bflow_bezier_eval(const float* params_nchw, const float* coef_host, float* flows, int T, int B, int degree, int H, int W,
                                 void* stream) {
    BFLOW_REQUIRE(params_nchw != nullptr && coef_host != nullptr && flows != nullptr, "bezier_eval: null tensor");
    BFLOW_REQUIRE(T > 0 && T <= 32 && B > 0 && degree >= 1 && degree <= BFLOW_MAX_DEGREE && H > 0 && W > 0, "bezier_eval: bad shape");
    BezierCoef coef;
    for (int t = 0; t < T; ++t)
        for (int k = 0; k < BFLOW_MAX_DEGREE; ++k) coef.c[t][k] = k < degree ? coef_host[t * degree + k] : 0.f;
    bezier_eval_kernel<<<grid_1d((long long)B * 2 * H * W, 256), 256, 0, (cudaStream_t)stream>>>(params_nchw, coef, flows, T, B, degree, H * W);
    return check_launch("bflow_bezier_eval");
}

extern "C" int bflow_cvx_upsample(const float* data, int ldd, int data_nchw, const float* mask, int ldm, int mask_nchw, float* out,
                                  int N, int C, int h, int w, void* stream) {
    BFLOW_REQUIRE(data != nullptr && mask != nullptr && out != nullptr, "cvx_upsample: null tensor");
    BFLOW_REQUIRE(N > 0 && N <= 65535 && h > 0 && h <= 65535 && w > 0 && C > 0, "cvx_upsample: bad shape");
    BFLOW_REQUIRE(data_nchw || ldd >= C, "cvx_upsample: ldd < C");
    BFLOW_REQUIRE(mask_nchw || ldm >= 576, "cvx_upsample: ldm < 576");
    dim3 grid(ceil_div(w, 4), h, N);
    cvx_upsample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(data, ldd, data_nchw, mask, ldm, mask_nchw, out, C, h, w);
    return check_launch("bflow_cvx_upsample");
}
