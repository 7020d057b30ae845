// fp32 implicit-GEMM convolution on NHWC activations (CUDA-core path).
//
// Replaces every nn.Conv2d on the reference path (models/raft_utils/extractor.py:49-53,112,120;
// models/raft_spline/update.py:17-18,36-45,89-96,112-114) and, through bflow_corr_volume, the
// all-pairs matmul of models/raft_utils/corr.py:264-272.
//
//   GEMM view:  M = N*Ho*Wo output pixels,  Ngemm = Cout,  K = KH*KW*Cin  (k = (kh*KW+kw)*Cin + c)
//   CTA tile :  BM x 64 outputs, BK = 16, 256 threads, each thread TM x 4 accumulators (BM = 16*TM)
//   A tile   :  gathered from up to two channel-concatenated NHWC sources, zero filled outside the
//               image (padding) and beyond M/K; stored transposed in shared memory As[k][m]
//   B tile   :  packed weights w[k][ldw] (Cout contiguous) -> Bs[k][n]
//   pipeline :  next tile's global loads are issued into registers before the FMA loop of the
//               current tile (register double buffering), one shared buffer, two barriers per k-block
//
// This is the exact-fp32 path: it is the numerical anchor for the tensor-core kernels and the
// fallback for shapes they do not cover (odd channel counts, tiny Cout).
#include "common.cuh"

namespace bflow {

constexpr int BN = 64;
constexpr int BK = 16;
constexpr int NTHREADS = 256;

template <int TM, bool VEC>
__global__ void __launch_bounds__(NTHREADS) conv_simt_kernel(const bflow_conv_desc d, const int M, const int K) {
    constexpr int BM = 16 * TM;
    constexpr int LDA = BM + 4;
    constexpr int NA_VEC = BM * 4 / NTHREADS;   // float4 loads per thread per k-block (2 or 1)
    constexpr int NA_GEN = BM * BK / NTHREADS;  // scalar loads per thread per k-block (8 or 4)

    __shared__ __align__(16) float As[BK][LDA];
    __shared__ __align__(16) float Bs[BK][BN];
    __shared__ int row_n[BM], row_ih0[BM], row_iw0[BM];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int Cin = d.c0 + d.c1;

    // decompose this CTA's output pixels once
    for (int r = tid; r < BM; r += NTHREADS) {
        int m = m0 + r;
        if (m < M) {
            int ow = m % d.Wo;
            int t = m / d.Wo;
            int oh = t % d.Ho;
            int n = t / d.Ho;
            row_n[r] = n;
            row_ih0[r] = oh * d.stride - d.pad_h;
            row_iw0[r] = ow * d.stride - d.pad_w;
        } else {
            row_n[r] = -1;
            row_ih0[r] = 0;
            row_iw0[r] = 0;
        }
    }
    __syncthreads();

    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int tx = tid % 16;
    const int ty = tid / 16;

    // register staging
    float4 a_vec[NA_VEC];
    float a_gen[NA_GEN];
    float4 b_reg;

    const int b_kk = tid / 16;
    const int b_nq = tid % 16;
    const bool b_col_ok = (n0 + b_nq * 4) < d.ldw;

    auto load_tile = [&](int k0) {
        // ---- B (weights) ----
        {
            int k = k0 + b_kk;
            if (k < K && b_col_ok)
                b_reg = __ldg(reinterpret_cast<const float4*>(d.w + (size_t)k * d.ldw + n0 + b_nq * 4));
            else
                b_reg = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // ---- A (activations) ----
        if (VEC) {
            // the whole 16-wide k-block lies inside one tap and one source
            int tap = k0 / Cin;
            int c = k0 - tap * Cin;
            int kh = tap / d.KW;
            int kw = tap - kh * d.KW;
            const float* src;
            int ld;
            if (c < d.c0) { src = d.x0; ld = d.ld0; } else { src = d.x1; ld = d.ld1; c -= d.c0; }
            const int kq = tid % 4;
#pragma unroll
            for (int j = 0; j < NA_VEC; ++j) {
                int r = tid / 4 + j * (NTHREADS / 4);
                int n = row_n[r];
                int ih = row_ih0[r] + kh;
                int iw = row_iw0[r] + kw;
                if (n >= 0 && ih >= 0 && ih < d.H && iw >= 0 && iw < d.W) {
                    size_t pix = ((size_t)n * d.H + ih) * d.W + iw;
                    a_vec[j] = __ldg(reinterpret_cast<const float4*>(src + pix * ld + c + kq * 4));
                } else {
                    a_vec[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        } else {
            const int kk = tid % 16;
            int k = k0 + kk;
            bool k_ok = k < K;
            int tap = k_ok ? k / Cin : 0;
            int c = k - tap * Cin;
            int kh = tap / d.KW;
            int kw = tap - kh * d.KW;
            const float* src;
            int ld;
            if (c < d.c0) { src = d.x0; ld = d.ld0; } else { src = d.x1; ld = d.ld1; c -= d.c0; }
#pragma unroll
            for (int j = 0; j < NA_GEN; ++j) {
                int r = tid / 16 + j * (NTHREADS / 16);
                int n = row_n[r];
                int ih = row_ih0[r] + kh;
                int iw = row_iw0[r] + kw;
                float v = 0.f;
                if (k_ok && n >= 0 && ih >= 0 && ih < d.H && iw >= 0 && iw < d.W) {
                    size_t pix = ((size_t)n * d.H + ih) * d.W + iw;
                    v = __ldg(src + pix * ld + c);
                }
                a_gen[j] = v;
            }
        }
    };

    auto store_tile = [&]() {
        *reinterpret_cast<float4*>(&Bs[b_kk][b_nq * 4]) = b_reg;
        if (VEC) {
            const int kq = tid % 4;
#pragma unroll
            for (int j = 0; j < NA_VEC; ++j) {
                int r = tid / 4 + j * (NTHREADS / 4);
                As[kq * 4 + 0][r] = a_vec[j].x;
                As[kq * 4 + 1][r] = a_vec[j].y;
                As[kq * 4 + 2][r] = a_vec[j].z;
                As[kq * 4 + 3][r] = a_vec[j].w;
            }
        } else {
            const int kk = tid % 16;
#pragma unroll
            for (int j = 0; j < NA_GEN; ++j) {
                int r = tid / 16 + j * (NTHREADS / 16);
                As[kk][r] = a_gen[j];
            }
        }
    };

    const int nkb = (K + BK - 1) / BK;
    load_tile(0);
    for (int kb = 0; kb < nkb; ++kb) {
        store_tile();
        __syncthreads();
        if (kb + 1 < nkb) load_tile((kb + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                float4 t = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
                a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
            }
            float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
                acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
                acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
            }
        }
        __syncthreads();
    }

    // ---- epilogue ----
    const int nbase = n0 + tx * 4;
    float bias[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bias[j] = (d.bias != nullptr && nbase + j < d.Cout) ? __ldg(d.bias + nbase + j) : 0.f;
    const bool vec_store = (d.y == nullptr || (((d.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.y) & 15) == 0))) && (nbase + 3 < d.Cout) &&
                           (d.res == nullptr || (((d.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.res) & 15) == 0)));
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= M) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = d.scale * (acc[i][j] + bias[j]);
        conv_epilogue4(d, m, nbase, v, vec_store);
    }
}

static int launch_conv(const bflow_conv_desc& d, cudaStream_t stream) {
    const long long Mll = (long long)d.N * d.Ho * d.Wo;
    const int Cin = d.c0 + d.c1;
    const long long Kll = (long long)d.KH * d.KW * Cin;
    const int M = (int)Mll, K = (int)Kll;
    bool vec = (d.c0 % 16 == 0) && (d.c1 % 16 == 0) && (d.ld0 % 4 == 0) && aligned16(d.x0) &&
               (d.c1 == 0 || ((d.ld1 % 4 == 0) && aligned16(d.x1)));
    // small problems: 64-row tiles fill more SMs
    const long long ctas128 = ceil_div_ll(Mll, 128) * ceil_div(d.Cout, BN);
    const bool small = ctas128 < 2 * 148;
    dim3 block(NTHREADS);
    if (small) {
        dim3 grid((unsigned)ceil_div_ll(Mll, 64), (unsigned)ceil_div(d.Cout, BN));
        if (vec) conv_simt_kernel<4, true><<<grid, block, 0, stream>>>(d, M, K);
        else conv_simt_kernel<4, false><<<grid, block, 0, stream>>>(d, M, K);
    } else {
        dim3 grid((unsigned)ceil_div_ll(Mll, 128), (unsigned)ceil_div(d.Cout, BN));
        if (vec) conv_simt_kernel<8, true><<<grid, block, 0, stream>>>(d, M, K);
        else conv_simt_kernel<8, false><<<grid, block, 0, stream>>>(d, M, K);
    }
    return check_launch("bflow_conv2d_nhwc");
}

}  // namespace bflow

extern "C" int bflow_conv2d_nhwc(const bflow_conv_desc* dp, void* stream) {
    BFLOW_REQUIRE(dp != nullptr, "conv: null descriptor");
    const bflow_conv_desc& d = *dp;
    BFLOW_REQUIRE(d.x0 != nullptr && d.w != nullptr, "conv: null tensor");
    BFLOW_REQUIRE(d.c0 > 0 && d.c1 >= 0 && d.ld0 >= d.c0, "conv: bad source 0");
    BFLOW_REQUIRE(d.c1 == 0 || (d.x1 != nullptr && d.ld1 >= d.c1), "conv: bad source 1");
    BFLOW_REQUIRE(d.N > 0 && d.H > 0 && d.W > 0 && d.Cout > 0, "conv: bad shape");
    BFLOW_REQUIRE(d.KH > 0 && d.KW > 0 && d.stride > 0 && d.pad_h >= 0 && d.pad_w >= 0, "conv: bad window");
    BFLOW_REQUIRE(d.Ho == (d.H + 2 * d.pad_h - d.KH) / d.stride + 1, "conv: Ho does not match");
    BFLOW_REQUIRE(d.Wo == (d.W + 2 * d.pad_w - d.KW) / d.stride + 1, "conv: Wo does not match");
    BFLOW_REQUIRE(d.ldw >= d.Cout && d.ldw % 4 == 0 && bflow::aligned16(d.w), "conv: packed weights need ldw%4==0, 16B aligned");
    BFLOW_REQUIRE(d.y == nullptr || d.ldy >= d.Cout, "conv: ldy < Cout");
    BFLOW_REQUIRE(d.res == nullptr || d.ldr >= d.Cout, "conv: ldr < Cout");
    BFLOW_REQUIRE(d.act1 >= 0 && d.act1 <= 3 && d.act2 >= 0 && d.act2 <= 3, "conv: bad activation");
    if (const char* msg = bflow::check_epilogue(d)) { bflow::set_error(msg); return BFLOW_ERR_INVALID; }
    BFLOW_REQUIRE((long long)d.N * d.Ho * d.Wo < (1ll << 31) && (long long)d.KH * d.KW * (d.c0 + d.c1) < (1ll << 31), "conv: too large");
    return bflow::launch_conv(d, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Direct convolution for tiny Cout: one warp per output pixel, lanes split the (tap, channel) axis in float4 steps,
// shuffle reduction at the end.  Bezier head conv2 (256 -> 2*degree, update.py:18) fused with the delta update.
// ---------------------------------------------------------------------------------------------
namespace bflow {
template <int NV>   // NV = ceil(Cout/4) float4 accumulators per lane
__global__ void __launch_bounds__(256) conv_small_n_kernel(const bflow_conv_desc d, const int M) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * 8 + warp;
    if (m >= M) return;
    const int ow = m % d.Wo;
    const int t = m / d.Wo;
    const int oh = t % d.Ho;
    const int n = t / d.Ho;
    const int Cin = d.c0;
    float4 acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kh = 0; kh < d.KH; ++kh) {
        const int ih = oh * d.stride - d.pad_h + kh;
        if (ih < 0 || ih >= d.H) continue;
        for (int kw = 0; kw < d.KW; ++kw) {
            const int iw = ow * d.stride - d.pad_w + kw;
            if (iw < 0 || iw >= d.W) continue;
            const float* xp = d.x0 + (((size_t)n * d.H + ih) * d.W + iw) * d.ld0;
            const float* wp = d.w + (size_t)((kh * d.KW + kw) * Cin) * d.ldw;
            for (int c = lane * 4; c < Cin; c += 128) {
                const float4 xv = __ldg(reinterpret_cast<const float4*>(xp + c));
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        const float4 wv = __ldg(reinterpret_cast<const float4*>(wp + (size_t)(c + e) * d.ldw + 4 * j));
                        acc[j].x = fmaf(xs[e], wv.x, acc[j].x);
                        acc[j].y = fmaf(xs[e], wv.y, acc[j].y);
                        acc[j].z = fmaf(xs[e], wv.z, acc[j].z);
                        acc[j].w = fmaf(xs[e], wv.w, acc[j].w);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc[j].x += __shfl_xor_sync(0xffffffffu, acc[j].x, o);
            acc[j].y += __shfl_xor_sync(0xffffffffu, acc[j].y, o);
            acc[j].z += __shfl_xor_sync(0xffffffffu, acc[j].z, o);
            acc[j].w += __shfl_xor_sync(0xffffffffu, acc[j].w, o);
        }
    }
    if (lane < NV) {
        float4 a = acc[0];
#pragma unroll
        for (int j = 1; j < NV; ++j) if (lane == j) a = acc[j];
        const int nb = lane * 4;
        float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = d.scale * (v[j] + ((d.bias != nullptr && nb + j < d.Cout) ? __ldg(d.bias + nb + j) : 0.f));
        const bool vec = (d.y == nullptr || (((d.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.y) & 15) == 0))) && (nb + 3 < d.Cout) &&
                         (d.res == nullptr || (((d.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.res) & 15) == 0)));
        conv_epilogue4(d, m, nb, v, vec);
    }
}
}  // namespace bflow

extern "C" int bflow_conv2d_small_n(const bflow_conv_desc* dp, void* stream) {
    BFLOW_REQUIRE(dp != nullptr, "conv_small_n: null descriptor");
    const bflow_conv_desc& d = *dp;
    BFLOW_REQUIRE(d.x0 != nullptr && d.w != nullptr, "conv_small_n: null tensor");
    BFLOW_REQUIRE(d.c1 == 0 && d.c0 > 0 && d.c0 % 4 == 0 && d.ld0 % 4 == 0 && bflow::aligned16(d.x0), "conv_small_n: one aligned source, Cin % 4 == 0");
    BFLOW_REQUIRE(d.Cout > 0 && d.Cout <= 32 && d.ldw % 4 == 0 && d.ldw >= d.Cout && bflow::aligned16(d.w), "conv_small_n: Cout <= 32, packed weights");
    BFLOW_REQUIRE(d.epi == BFLOW_EPI_STD, "conv_small_n: standard epilogue only");
    BFLOW_REQUIRE(d.Ho == (d.H + 2 * d.pad_h - d.KH) / d.stride + 1 && d.Wo == (d.W + 2 * d.pad_w - d.KW) / d.stride + 1, "conv_small_n: Ho/Wo mismatch");
    BFLOW_REQUIRE((d.y == nullptr || d.ldy >= d.Cout) && (d.res == nullptr || d.ldr >= d.Cout), "conv_small_n: bad output stride");
    if (const char* msg = bflow::check_epilogue(d)) { bflow::set_error(msg); return BFLOW_ERR_INVALID; }
    const long long Mll = (long long)d.N * d.Ho * d.Wo;
    BFLOW_REQUIRE(Mll > 0 && Mll < (1ll << 31), "conv_small_n: bad shape");
    const int M = (int)Mll;
    const int nv = (d.Cout + 3) / 4;
    dim3 grid((unsigned)bflow::ceil_div(M, 8));
    cudaStream_t st = (cudaStream_t)stream;
    switch (nv) {
        case 1: bflow::conv_small_n_kernel<1><<<grid, 256, 0, st>>>(d, M); break;
        case 2: bflow::conv_small_n_kernel<2><<<grid, 256, 0, st>>>(d, M); break;
        case 3: bflow::conv_small_n_kernel<3><<<grid, 256, 0, st>>>(d, M); break;
        case 4: bflow::conv_small_n_kernel<4><<<grid, 256, 0, st>>>(d, M); break;
        case 5: bflow::conv_small_n_kernel<5><<<grid, 256, 0, st>>>(d, M); break;
        case 6: bflow::conv_small_n_kernel<6><<<grid, 256, 0, st>>>(d, M); break;
        case 7: bflow::conv_small_n_kernel<7><<<grid, 256, 0, st>>>(d, M); break;
        default: bflow::conv_small_n_kernel<8><<<grid, 256, 0, st>>>(d, M); break;
    }
    return bflow::check_launch("bflow_conv2d_small_n");
}

// corr[bq, p] = sum_d f1[bq, d] * f2[b, d, p] / sqrt(D)   (models/raft_utils/corr.py:264-272)
// = a 1x1 "convolution" per sample whose weight matrix is the target feature map in NCHW.
extern "C" int bflow_corr_volume(const float* f1, int ld1, const float* f2_nchw, float* corr, int B, int D, int Q, void* stream) {
    BFLOW_REQUIRE(f1 != nullptr && f2_nchw != nullptr && corr != nullptr, "corr_volume: null tensor");
    BFLOW_REQUIRE(B > 0 && D > 0 && Q > 0 && ld1 >= D, "corr_volume: bad shape");
    BFLOW_REQUIRE(Q % 4 == 0, "corr_volume: h*w must be a multiple of 4");
    for (int b = 0; b < B; ++b) {
        bflow_conv_desc d{};
        d.x0 = f1 + (size_t)b * Q * ld1; d.c0 = D; d.ld0 = ld1;
        d.x1 = nullptr; d.c1 = 0; d.ld1 = 0;
        d.w = f2_nchw + (size_t)b * D * Q; d.ldw = Q;
        d.bias = nullptr; d.res = nullptr; d.ldr = 0;
        d.y = corr + (size_t)b * Q * Q; d.ldy = Q;
        d.N = 1; d.H = 1; d.W = Q; d.Ho = 1; d.Wo = Q; d.Cout = Q;
        d.KH = 1; d.KW = 1; d.stride = 1; d.pad_h = 0; d.pad_w = 0;
        d.act1 = BFLOW_ACT_NONE; d.act2 = BFLOW_ACT_NONE;
        d.epi = BFLOW_EPI_STD;
        d.scale = 1.0f / sqrtf((float)D);
        BFLOW_REQUIRE(bflow::aligned16(d.w), "corr_volume: f2 must be 16B aligned per sample");
        int rc = bflow::launch_conv(d, (cudaStream_t)stream);
        if (rc != BFLOW_OK) return rc;
    }
    return BFLOW_OK;
}
