// On-the-fly correlation lookup (scope row f3): the per-iteration window lookup WITHOUT a materialised correlation volume.
//
// Replaces CorrComputation + CorrData.get_downsampled + CorrBlockParallelMultiTarget.__call__ (models/raft_utils/corr.py:108-125,
// 264-272, 307-350) by computing, for every unit (query pixel, slot), the 10 x 10 footprint of correlation values directly as dot
// products <f1[q], f2_level[p]> / sqrt(D) against an average-pooled TARGET FEATURE pyramid.  avg_pool2d is linear, so pooling the
// target features with the same 2x2 / stride 2 / floor rule gives exactly the pooled correlation planes (corr.py:119); only the
// fp32 summation order differs from the volume path.
//
//   memory      T x (B*Q) x Q fp32 volume (369 MB at config D, 906 MB at M batch 4)  ->  T x B x Q x D pooled features (+33 %)
//   arithmetic  2*100*D flops per unit per iteration on fp32 CUDA cores (20.6 GF per frame at D for 12 iterations, against 47.2 GF
//               once for the all-pairs GEMM) -- but every unit gathers 100 feature rows of D floats (100 KB) through L1 / L2
//
// One warp per unit, lanes split the channels (D = 128 * ND, a float4 per lane per 128 channels).  32 footprint positions are
// accumulated at a time into 32 per-lane partial sums and reduced with ONE butterfly transpose (31 shuffles per 32 positions instead
// of 5 per position); the 10 x 10 patch goes through shared memory and the 81 taps are blended exactly as in corr_lookup.cu.
// A CTA takes 8 horizontally adjacent query pixels of one slot, so that with a coherent flow field their windows overlap in L1.
#include "common.cuh"

namespace bflow {

constexpr int OTF_WARPS = 8;
constexpr int OTF_PITCH = 12;      // floats per footprint row in shared memory (10 used)

template <int ND>
__global__ void __launch_bounds__(OTF_WARPS * 32) corr_lookup_otf_kernel(const bflow_lookup_otf_desc d, const FastDiv div_q, const FastDiv div_w,
                                                                        unsigned long long* tl) {
    __shared__ __align__(16) float patch[OTF_WARPS][10 * OTF_PITCH + 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.y;
    const int Q = d.h * d.w;
    const unsigned bq = blockIdx.x * OTF_WARPS + warp;
    tl_begin(tl);
    if (bq >= (unsigned)(d.B * Q)) {
        tl_end(tl);
        return;
    }
    const unsigned b = fastdiv(bq, div_q), q = bq - b * (unsigned)Q;
    const unsigned qy = fastdiv(q, div_w), qx = q - qy * (unsigned)d.w;
    const int hl = d.hl[slot], wl = d.wl[slot], t = d.target[slot];
    float cx, cy;
    if (d.coords != nullptr) {
        const float* c = d.coords + (((size_t)t * d.B + b) * 2) * Q + q;
        cx = __ldg(c);
        cy = __ldg(c + Q);
    } else {      // coords1 = pixel grid + sum_i coef[t][i] * P_i   (raft.py:180-181, bezier.py:165-186)
        const float* prm = d.params + (size_t)bq * d.params_ld;
        float fxv = 0.f, fyv = 0.f;
        for (int k = 0; k < d.degree; ++k) {
            const float ck = d.coef[t][k];
            fxv = fmaf(ck, prm[k], fxv);
            fyv = fmaf(ck, prm[d.degree + k], fyv);
        }
        cx = (float)qx + fxv;
        cy = (float)qy + fyv;
    }
    cx = fminf(fmaxf(cx * d.inv_scale[slot], -16.f), (float)wl + 16.f);
    cy = fminf(fmaxf(cy * d.inv_scale[slot], -16.f), (float)hl + 16.f);
    const float flx = floorf(cx), fly = floorf(cy);
    const float fx = cx - flx, fy = cy - fly;
    const int x0 = (int)flx - 4, y0 = (int)fly - 4;

    // query feature: lane holds channels [128 j + 4 lane, +4)
    float4 f1[ND];
    bool on[ND];                     // D need not fill the last 128-channel pass (D % 4 == 0)
    const float* f1row = d.f1[slot] + (size_t)bq * d.ld1 + lane * 4;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
        on[j] = 128 * j + 4 * lane < d.D;
        f1[j] = on[j] ? __ldg(reinterpret_cast<const float4*>(f1row + 128 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float* f2b = d.f2[slot] + (size_t)b * hl * wl * d.ld2 + lane * 4;
    float* ps = patch[warp];

    // 100 footprint positions, 32 at a time; position i of a group lands in lane i after the butterfly
#pragma unroll 1
    for (int g0 = 0; g0 < 100; g0 += 32) {
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int pos = g0 + i;
            const int py = pos / 10, px = pos - py * 10;
            const int y = y0 + py, x = x0 + px;
            float s = 0.f;
            if (pos < 100 && (unsigned)y < (unsigned)hl && (unsigned)x < (unsigned)wl) {      // warp-uniform; outside the plane = zero padding
                const float* r = f2b + ((size_t)y * wl + x) * d.ld2;
#pragma unroll
                for (int j = 0; j < ND; ++j) {
                    const float4 v = on[j] ? __ldg(reinterpret_cast<const float4*>(r + 128 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    s = fmaf(f1[j].x, v.x, s);
                    s = fmaf(f1[j].y, v.y, s);
                    s = fmaf(f1[j].z, v.z, s);
                    s = fmaf(f1[j].w, v.w, s);
                }
            }
            acc[i] = s;
        }
        // butterfly transpose-reduce: after the stage with `bit`, a lane keeps the half of the positions whose index has that bit equal to its own
#pragma unroll
        for (int width = 16, bit = 16; width >= 1; width >>= 1, bit >>= 1) {
            const bool upper = (lane & bit) != 0;
#pragma unroll
            for (int i = 0; i < width; ++i) {
                const float keep = upper ? acc[width + i] : acc[i], send = upper ? acc[i] : acc[width + i];
                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
        }
        // lane l now holds position g0 + bitrev-free index: stage order (16, 8, 4, 2, 1) keeps position bits in place, i.e. position g0 + l
        const int pos = g0 + lane;
        if (pos < 100) {
            const int py = pos / 10;
            ps[py * OTF_PITCH + (pos - py * 10)] = acc[0] * d.scale;
        }
    }
    __syncwarp();
    const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int k = lane + 32 * j;
        if (k < 81) {
            const int iy = k / 9, ix = k - iy * 9;
            const float* f = ps + iy * OTF_PITCH + ix;
            const float val = w00 * f[0] + w01 * f[1] + w10 * f[OTF_PITCH] + w11 * f[OTF_PITCH + 1];
            const size_t e = (size_t)bq * (d.out16_hi != nullptr ? d.out16_ld : d.out_ld) + slot * 81 + k;
            if (d.out16_hi != nullptr) {
                if (d.out16_lo != nullptr) store_split1(d.out16_hi, d.out16_lo, e, val);
                else reinterpret_cast<__half*>(d.out16_hi)[e] = __float2half_rn(fminf(fmaxf(val, -65504.f), 65504.f));
            } else {
                d.out[e] = val;
            }
        }
    }
    tl_end(tl);
}

// avg_pool2d(2, stride 2, floor) of NHWC features: (N, H, W, C) -> (N, H/2, W/2, C); one float4 of channels per thread
__global__ void feat_pool_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W, int C4, int ld_in, int ld_out) {
    const int Ho = H >> 1, Wo = W >> 1;
    const long long total = (long long)N * Ho * Wo * C4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C4) * 4;
        long long r = idx / C4;
        const int xo = (int)(r % Wo);
        r /= Wo;
        const int yo = (int)(r % Ho), n = (int)(r / Ho);
        const float* p = in + (((size_t)n * H + 2 * yo) * W + 2 * xo) * ld_in + c;
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + ld_in);
        const float4 e = *reinterpret_cast<const float4*>(p + (size_t)W * ld_in), f = *reinterpret_cast<const float4*>(p + (size_t)W * ld_in + ld_in);
        // the summation order of avg_pool2d: row-major over the 2 x 2 window, then * 0.25
        float4 o;
        o.x = (a.x + b.x + e.x + f.x) * 0.25f;
        o.y = (a.y + b.y + e.y + f.y) * 0.25f;
        o.z = (a.z + b.z + e.z + f.z) * 0.25f;
        o.w = (a.w + b.w + e.w + f.w) * 0.25f;
        *reinterpret_cast<float4*>(out + (((size_t)n * Ho + yo) * Wo + xo) * ld_out + c) = o;
    }
}

}  // namespace bflow

extern "C" int bflow_feat_pool(const float* in, float* out, int N, int H, int W, int C, int ld_in, int ld_out, void* stream) {
    BFLOW_REQUIRE(in != nullptr && out != nullptr, "feat_pool: null tensor");
    BFLOW_REQUIRE(N > 0 && H >= 2 && W >= 2 && C > 0 && C % 4 == 0 && ld_in >= C && ld_out >= C && ld_in % 4 == 0 && ld_out % 4 == 0, "feat_pool: bad shape");
    BFLOW_REQUIRE(bflow::aligned16(in) && bflow::aligned16(out), "feat_pool: alignment");
    const long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
    long long g = (total + 255) / 256;
    const long long cap = (long long)bflow::num_sms() * 16;
    if (g > cap) g = cap;
    bflow::feat_pool_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(in, out, N, H, W, C / 4, ld_in, ld_out);
    return bflow::check_launch("bflow_feat_pool");
}

extern "C" int bflow_corr_lookup_otf(const bflow_lookup_otf_desc* dp, void* stream) {
    BFLOW_CHECK_DESC(dp, bflow_lookup_otf_desc, "lookup_otf");
    const bflow_lookup_otf_desc& d = *dp;
    BFLOW_REQUIRE(d.n_slots > 0 && d.n_slots <= BFLOW_MAX_SLOTS && d.n_targets > 0 && d.n_targets <= BFLOW_MAX_TARGETS, "lookup_otf: bad slot / target count");
    BFLOW_REQUIRE(d.B > 0 && d.h > 0 && d.w > 0 && d.radius == 4, "lookup_otf: bad shape (radius is fixed to 4, raft.py:38-40)");
    BFLOW_REQUIRE(d.D >= 4 && d.D <= 512 && d.D % 4 == 0, "lookup_otf: feature dimension must be a multiple of 4, at most 512");
    BFLOW_REQUIRE(d.ld1 >= d.D && d.ld1 % 4 == 0 && d.ld2 >= d.D && d.ld2 % 4 == 0, "lookup_otf: feature layout");
    BFLOW_REQUIRE(d.out != nullptr || d.out16_hi != nullptr, "lookup_otf: null output");
    BFLOW_REQUIRE(d.out16_hi != nullptr ? d.out16_ld >= d.n_slots * 81 : d.out_ld >= d.n_slots * 81, "lookup_otf: output row stride too small");
    BFLOW_REQUIRE(d.coords != nullptr || (d.params != nullptr && d.degree >= 1 && d.degree <= BFLOW_MAX_DEGREE && d.params_ld >= 2 * d.degree),
                  "lookup_otf: need coords or Bezier params");
    for (int s = 0; s < d.n_slots; ++s) {
        BFLOW_REQUIRE(d.f1[s] != nullptr && bflow::aligned16(d.f1[s]), "lookup_otf: query features");
        BFLOW_REQUIRE(d.f2[s] != nullptr && bflow::aligned16(d.f2[s]) && d.hl[s] > 0 && d.wl[s] > 0, "lookup_otf: bad pyramid level");
        BFLOW_REQUIRE(d.target[s] >= 0 && d.target[s] < d.n_targets, "lookup_otf: bad slot target");
    }
    const long long BQ = (long long)d.B * d.h * d.w;
    BFLOW_REQUIRE(BQ < (1ll << 31), "lookup_otf: too many query pixels");
    dim3 grid((unsigned)bflow::ceil_div_ll(BQ, bflow::OTF_WARPS), (unsigned)d.n_slots);
    const bflow::FastDiv dq = bflow::make_fastdiv((unsigned)(d.h * d.w)), dw = bflow::make_fastdiv((unsigned)d.w);
    unsigned long long* tls = bflow::timeline_next_slot("corr_lookup_otf");
    cudaStream_t st = (cudaStream_t)stream;
    switch ((d.D + 127) / 128) {
        case 1: bflow::corr_lookup_otf_kernel<1><<<grid, bflow::OTF_WARPS * 32, 0, st>>>(d, dq, dw, tls); break;
        case 2: bflow::corr_lookup_otf_kernel<2><<<grid, bflow::OTF_WARPS * 32, 0, st>>>(d, dq, dw, tls); break;
        case 3: bflow::corr_lookup_otf_kernel<3><<<grid, bflow::OTF_WARPS * 32, 0, st>>>(d, dq, dw, tls); break;
        default: bflow::corr_lookup_otf_kernel<4><<<grid, bflow::OTF_WARPS * 32, 0, st>>>(d, dq, dw, tls); break;
    }
    return bflow::check_launch("bflow_corr_lookup_otf");
}
