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
__global__ void __launch_bounds__(NTHREADS) conv_simt_kernel(const bflow_conv_desc d, const int M, const int K, unsigned long long* tl) {
    tl_begin(tl);
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
    tl_end(tl);
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
    unsigned long long* tls = timeline_next_slot("conv_simt");
    if (small) {
        dim3 grid((unsigned)ceil_div_ll(Mll, 64), (unsigned)ceil_div(d.Cout, BN));
        if (vec) conv_simt_kernel<4, true><<<grid, block, 0, stream>>>(d, M, K, tls);
        else conv_simt_kernel<4, false><<<grid, block, 0, stream>>>(d, M, K, tls);
    } else {
        dim3 grid((unsigned)ceil_div_ll(Mll, 128), (unsigned)ceil_div(d.Cout, BN));
        if (vec) conv_simt_kernel<8, true><<<grid, block, 0, stream>>>(d, M, K, tls);
        else conv_simt_kernel<8, false><<<grid, block, 0, stream>>>(d, M, K, tls);
    }
    return check_launch("bflow_conv2d_nhwc");
}

}  // namespace bflow

extern "C" int bflow_conv2d_nhwc(const bflow_conv_desc* dp, void* stream) {
    BFLOW_CHECK_DESC(dp, bflow_conv_desc, "conv");
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
__global__ void __launch_bounds__(256) conv_small_n_kernel(const bflow_conv_desc d, const int M, unsigned long long* tl) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * 8 + warp;
    tl_begin(tl);
    pdl_trigger();
    pdl_wait();
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
    tl_end(tl);
}
// Same operator for the shape the Bezier head really has (3x3, stride 1, pad 1, Cin = 128 or 256, Cout <= 8).  Measured on B200: per-lane
// loads from the 8 warps of a CTA move only ~9 B/clk per SM, and the warp-per-pixel kernel above reads every input row nine times.  Here a CTA
// owns a 5x8 patch of output pixels: the 7x10 halo of input rows (1 KB each) and the whole weight matrix arrive by cp.async.bulk (the copy
// engine is not bound by per-thread load slots), a warp owns 4 consecutive pixels, lane l owns channels l + 32 i -- activations are read with
// conflict-free LDS.32, the weights of a (tap, channel) with one conflict-free LDS.128 shared by the 4 pixels.
// NV = ceil(Cout/4), CB = Cin / 128
constexpr int HD_ROWS = 5;        // tile rows: 60x80 -> 12 x 10 = 120 CTAs, one wave on 148 SMs (4-row tiles: 150 CTAs, two SMs ran two CTAs in turn)
template <int NV, int CB>
__global__ void __launch_bounds__(64 * HD_ROWS) conv_head3x3_kernel(const bflow_conv_desc d, const int M, unsigned long long* tl) {
    extern __shared__ __align__(128) float hs[];          // [9*Cin][NV*4] weights (global order), then the [HD_ROWS + 2][10][Cin] input patch
    __shared__ __align__(8) unsigned long long bar;
    constexpr int Cin = CB * 128;
    constexpr int PR = HD_ROWS + 2, PC = 10;
    float* s_w = hs;
    float* s_x = hs + 9 * Cin * NV * 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (d.Wo + 7) / 8, tiles_y = (d.Ho + HD_ROWS - 1) / HD_ROWS;
    int bid = blockIdx.x;
    const int tx = bid % tiles_x; bid /= tiles_x;
    const int ty = bid % tiles_y;
    const int n = bid / tiles_y;
    const int oy0 = ty * HD_ROWS, ox0 = tx * 8;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    tl_begin(tl);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // which patch pixels exist (the others are the zero padding): pixel i = py * PC + px
    if (threadIdx.x < PR * PC) {
        const int py = threadIdx.x / PC, px = threadIdx.x - py * PC;
        const int iy = oy0 + py - 1, ix = ox0 + px - 1;
        const bool ok = iy >= 0 && iy < d.H && ix >= 0 && ix < d.W;
        float* dst = s_x + threadIdx.x * Cin;
        if (ok) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(dst)),
                         "l"(d.x0 + (((size_t)n * d.H + iy) * d.W + ix) * d.ld0), "r"(Cin * 4), "r"(bar_a)
                         : "memory");
        } else {
            for (int c = 0; c < Cin; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (threadIdx.x == 64) {
        // expected bytes: the weights + every in-image patch pixel
        const int y_lo = max(oy0 - 1, 0), y_hi = min(oy0 + PR - 2, d.H - 1), x_lo = max(ox0 - 1, 0), x_hi = min(ox0 + PC - 2, d.W - 1);
        const uint32_t npix = (uint32_t)(max(y_hi - y_lo + 1, 0) * max(x_hi - x_lo + 1, 0));
        const uint32_t wbytes = 9u * Cin * NV * 16u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(wbytes + npix * Cin * 4u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(s_w)),
                     "l"(d.w), "r"(wbytes), "r"(bar_a)
                     : "memory");
    }
    __syncthreads();                                        // zero-filled padding pixels are visible
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(bar_a)
                : "memory");
        }
    }
    // warp -> 4 consecutive pixels of one tile row
    const int prow = warp >> 1, pcol0 = (warp & 1) * 4;
    float4 acc[4][NV];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NV; ++j) acc[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap - ky * 3;
        const float* xrow = s_x + ((prow + ky) * PC + pcol0 + kx) * Cin + lane;
        const float4* wrow = reinterpret_cast<const float4*>(s_w) + (size_t)(tap * Cin + lane) * NV;
#pragma unroll
        for (int i = 0; i < Cin / 32; ++i) {
            float xv[4];
#pragma unroll
            for (int px = 0; px < 4; ++px) xv[px] = xrow[px * Cin + i * 32];
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const float4 wv = wrow[(size_t)i * 32 * NV + j];
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                    acc[px][j].x = fmaf(xv[px], wv.x, acc[px][j].x);
                    acc[px][j].y = fmaf(xv[px], wv.y, acc[px][j].y);
                    acc[px][j].z = fmaf(xv[px], wv.z, acc[px][j].z);
                    acc[px][j].w = fmaf(xv[px], wv.w, acc[px][j].w);
                }
            }
        }
    }
#pragma unroll
    for (int px = 0; px < 4; ++px) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc[px][j].x += __shfl_xor_sync(0xffffffffu, acc[px][j].x, o);
                acc[px][j].y += __shfl_xor_sync(0xffffffffu, acc[px][j].y, o);
                acc[px][j].z += __shfl_xor_sync(0xffffffffu, acc[px][j].z, o);
                acc[px][j].w += __shfl_xor_sync(0xffffffffu, acc[px][j].w, o);
            }
        }
    }
    // lane (px * NV + j) finishes output group j of pixel px
    if (lane < 4 * NV) {
        const int px = lane / NV, j = lane - px * NV;
        float4 a = acc[0][0];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp)
#pragma unroll
            for (int jj = 0; jj < NV; ++jj)
                if (pp == px && jj == j) a = acc[pp][jj];
        const int oy = oy0 + prow, ox = ox0 + pcol0 + px;
        if (oy < d.Ho && ox < d.Wo) {
            const int m = (n * d.Ho + oy) * d.Wo + ox;
            const int nb = j * 4;
            float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = d.scale * (v[e] + ((d.bias != nullptr && nb + e < d.Cout) ? __ldg(d.bias + nb + e) : 0.f));
            const bool vec = (d.y == nullptr || (((d.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.y) & 15) == 0))) && (nb + 3 < d.Cout) &&
                             (d.res == nullptr || (((d.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.res) & 15) == 0)));
            conv_epilogue4(d, m, nb, v, vec);
        }
    }
    (void)M;
    tl_end(tl);
}

template <int NV, int CB>
static cudaError_t launch_head3x3(const bflow_conv_desc& d, int M, cudaStream_t st, unsigned long long* tls) {
    const size_t smem = (size_t)9 * CB * 128 * NV * 16 + (size_t)(HD_ROWS + 2) * 10 * CB * 128 * 4;
    static PerDeviceFlag configured;
    if (!configured.get()) {
        cudaError_t e = cudaFuncSetAttribute(conv_head3x3_kernel<NV, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured.set();
    }
    const int g = d.N * ((d.Ho + HD_ROWS - 1) / HD_ROWS) * ((d.Wo + 7) / 8);
    conv_head3x3_kernel<NV, CB><<<(unsigned)g, 64 * HD_ROWS, smem, st>>>(d, M, tls);
    return cudaGetLastError();
}
}  // namespace bflow

extern "C" int bflow_conv2d_small_n(const bflow_conv_desc* dp, void* stream) {
    BFLOW_CHECK_DESC(dp, bflow_conv_desc, "conv_small_n");
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
    cudaError_t le = cudaSuccess;
    unsigned long long* tls = bflow::timeline_next_slot("conv_small_n");
    if (d.KH == 3 && d.KW == 3 && d.stride == 1 && d.pad_h == 1 && d.pad_w == 1 && nv <= 2 && (d.c0 == 128 || d.c0 == 256) && d.ldw == nv * 4 &&
        d.Ho == d.H && d.Wo == d.W) {
        if (nv == 1) le = d.c0 == 128 ? bflow::launch_head3x3<1, 1>(d, M, st, tls) : bflow::launch_head3x3<1, 2>(d, M, st, tls);
        else le = d.c0 == 128 ? bflow::launch_head3x3<2, 1>(d, M, st, tls) : bflow::launch_head3x3<2, 2>(d, M, st, tls);
        if (le != cudaSuccess) {
            bflow::set_error(cudaGetErrorString(le));
            return BFLOW_ERR_CUDA;
        }
        return bflow::check_launch("bflow_conv2d_small_n(head3x3)");
    }
    switch (nv) {
        case 1: le = bflow::launch_pdl(bflow::conv_small_n_kernel<1>, grid, dim3(256), 0, st, d, M, tls); break;
        case 2: le = bflow::launch_pdl(bflow::conv_small_n_kernel<2>, grid, dim3(256), 0, st, d, M, tls); break;
        case 3: le = bflow::launch_pdl(bflow::conv_small_n_kernel<3>, grid, dim3(256), 0, st, d, M, tls); break;
        case 4: le = bflow::launch_pdl(bflow::conv_small_n_kernel<4>, grid, dim3(256), 0, st, d, M, tls); break;
        case 5: le = bflow::launch_pdl(bflow::conv_small_n_kernel<5>, grid, dim3(256), 0, st, d, M, tls); break;
        case 6: le = bflow::launch_pdl(bflow::conv_small_n_kernel<6>, grid, dim3(256), 0, st, d, M, tls); break;
        case 7: le = bflow::launch_pdl(bflow::conv_small_n_kernel<7>, grid, dim3(256), 0, st, d, M, tls); break;
        default: le = bflow::launch_pdl(bflow::conv_small_n_kernel<8>, grid, dim3(256), 0, st, d, M, tls); break;
    }
    if (le != cudaSuccess) {
        bflow::set_error(cudaGetErrorString(le));
        return BFLOW_ERR_CUDA;
    }
    return bflow::check_launch("bflow_conv2d_small_n");
}

// ---------------------------------------------------------------------------------------------
// 7x7 stride-1 convolution of a THIN input (Cin = 4, 8, ... 32) to 128 channels: convf1 of the motion encoder (update.py:91,
// Bezier parameters 2*degree -> 128).  K = 49*Cin is too small and too ragged for the tensor-core tiles, and the generic CUDA-core
// kernel gathers it element by element.  Here a CTA owns a 5x8 patch of output pixels (60x80 -> 120 CTAs: one wave); per chunk of 4 input channels the 49*4*128
// weights (100 KB) sit in shared memory next to the (5+6)x(8+6) input patch; a thread owns 4 consecutive pixels x 4 output channels and,
// per (channel, filter row), loads the 10 inputs its pixels need once and reuses them across the 7 filter columns: 112 FMAs per 17
// shared-memory loads.
// ---------------------------------------------------------------------------------------------
namespace bflow {
constexpr int TH_ROWS = 5, TH_COLS = 8, TH_CO = 128, TH_K = 7, TH_PR = TH_ROWS + TH_K - 1, TH_PC = TH_COLS + TH_K - 1;
__global__ void __launch_bounds__(64 * TH_ROWS) conv_thin7_kernel(const bflow_conv_desc d, unsigned long long* tl) {
    extern __shared__ __align__(16) float th_smem[];
    float* s_w = th_smem;                                    // [49][4][128]
    float* s_x = th_smem + TH_K * TH_K * 4 * TH_CO;          // [4][TH_PR][TH_PC]
    const int tid = threadIdx.x;
    const int cg = tid & 31, r = tid >> 5;
    const int prow = r >> 1, px0 = (r & 1) * 4;
    const int tiles_x = (d.Wo + TH_COLS - 1) / TH_COLS, tiles_y = (d.Ho + TH_ROWS - 1) / TH_ROWS;
    int bid = blockIdx.x;
    const int tx = bid % tiles_x; bid /= tiles_x;
    const int ty = bid % tiles_y;
    const int n = bid / tiles_y;
    const int oy0 = ty * TH_ROWS, ox0 = tx * TH_COLS;
    tl_begin(tl);
    __shared__ __align__(8) unsigned long long wbar_storage;
    const uint32_t wbar = (uint32_t)__cvta_generic_to_shared(&wbar_storage);
    uint32_t wphase = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(wbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int c0 = 0; c0 < d.c0; c0 += 4) {
        __syncthreads();
        // weights of this channel chunk: rows (tap*Cin + c0 + ch) of the packed [K][ldw] matrix -- per tap 4 consecutive rows (2 KB), one
        // cp.async.bulk each (per-lane copies of the 100 KB cost ~5 us of LSU time per CTA)
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic reads of s_w (previous chunk) are ordered before the refill
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wbar), "r"(TH_K * TH_K * 4 * TH_CO * 4) : "memory");
        }
        if (tid < TH_K * TH_K) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_w + tid * 4 * TH_CO);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(d.w + (size_t)(tid * d.c0 + c0) * d.ldw), "r"(4 * TH_CO * 4), "r"(wbar)
                         : "memory");
        }
        for (int i = tid; i < TH_PR * TH_PC; i += 64 * TH_ROWS) {
            const int yy = i / TH_PC, xx = i - yy * TH_PC;
            const int iy = oy0 + yy - 3, ix = ox0 + xx - 3;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (iy >= 0 && iy < d.H && ix >= 0 && ix < d.W) v = *reinterpret_cast<const float4*>(d.x0 + (((size_t)n * d.H + iy) * d.W + ix) * d.ld0 + c0);
            s_x[0 * TH_PR * TH_PC + i] = v.x;
            s_x[1 * TH_PR * TH_PC + i] = v.y;
            s_x[2 * TH_PR * TH_PC + i] = v.z;
            s_x[3 * TH_PR * TH_PC + i] = v.w;
        }
        {
            uint32_t ok = 0;
            while (!ok) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok)
                    : "r"(wbar), "r"(wphase)
                    : "memory");
            }
            wphase ^= 1u;
        }
        __syncthreads();
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
#pragma unroll 1
            for (int ky = 0; ky < TH_K; ++ky) {
                float xr[10];
                const float* xp = s_x + (ch * TH_PR + prow + ky) * TH_PC + px0;
#pragma unroll
                for (int i = 0; i < 10; ++i) xr[i] = xp[i];
#pragma unroll
                for (int kx = 0; kx < TH_K; ++kx) {
                    const float4 w4 = *reinterpret_cast<const float4*>(s_w + (((ky * TH_K + kx) * 4 + ch) * TH_CO) + cg * 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[i][0] = fmaf(xr[i + kx], w4.x, acc[i][0]);
                        acc[i][1] = fmaf(xr[i + kx], w4.y, acc[i][1]);
                        acc[i][2] = fmaf(xr[i + kx], w4.z, acc[i][2]);
                        acc[i][3] = fmaf(xr[i + kx], w4.w, acc[i][3]);
                    }
                }
            }
        }
    }
    const int oy = oy0 + prow;
    const float4 b4 = d.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(d.bias + cg * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const bool vec = (d.y == nullptr || (((d.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.y) & 15) == 0))) &&
                     (d.res == nullptr || (((d.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.res) & 15) == 0)));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ox = ox0 + px0 + i;
        if (oy < d.Ho && ox < d.Wo) {
            float v[4] = {d.scale * (acc[i][0] + b4.x), d.scale * (acc[i][1] + b4.y), d.scale * (acc[i][2] + b4.z), d.scale * (acc[i][3] + b4.w)};
            conv_epilogue4(d, (n * d.Ho + oy) * d.Wo + ox, cg * 4, v, vec);
        }
    }
    tl_end(tl);
}
}  // namespace bflow

extern "C" int bflow_conv2d_thin7(const bflow_conv_desc* dp, void* stream) {
    BFLOW_CHECK_DESC(dp, bflow_conv_desc, "conv_thin7");
    const bflow_conv_desc& d = *dp;
    BFLOW_REQUIRE(d.x0 != nullptr && d.w != nullptr, "conv_thin7: null tensor");
    BFLOW_REQUIRE(d.c1 == 0 && d.c0 > 0 && d.c0 % 4 == 0 && d.ld0 % 4 == 0 && bflow::aligned16(d.x0), "conv_thin7: one aligned source, Cin % 4 == 0");
    BFLOW_REQUIRE(d.KH == 7 && d.KW == 7 && d.stride == 1 && d.pad_h == 3 && d.pad_w == 3 && d.Ho == d.H && d.Wo == d.W, "conv_thin7: 7x7, stride 1, pad 3");
    BFLOW_REQUIRE(d.Cout == bflow::TH_CO && d.ldw == bflow::TH_CO && bflow::aligned16(d.w), "conv_thin7: Cout == 128, packed weights with ldw == 128");
    BFLOW_REQUIRE(d.bias == nullptr || bflow::aligned16(d.bias), "conv_thin7: bias alignment");
    BFLOW_REQUIRE(d.epi == BFLOW_EPI_STD, "conv_thin7: standard epilogue only");
    BFLOW_REQUIRE((d.y == nullptr || d.ldy >= d.Cout) && (d.res == nullptr || d.ldr >= d.Cout), "conv_thin7: bad output stride");
    if (const char* msg = bflow::check_epilogue(d)) { bflow::set_error(msg); return BFLOW_ERR_INVALID; }
    const long long tiles = (long long)d.N * ((d.Ho + bflow::TH_ROWS - 1) / bflow::TH_ROWS) * ((d.Wo + bflow::TH_COLS - 1) / bflow::TH_COLS);
    BFLOW_REQUIRE(tiles > 0 && tiles < (1ll << 31), "conv_thin7: bad shape");
    const size_t smem = (size_t)(bflow::TH_K * bflow::TH_K * 4 * bflow::TH_CO + 4 * bflow::TH_PR * bflow::TH_PC) * sizeof(float);
    static bflow::PerDeviceFlag configured;
    if (!configured.get()) {
        cudaError_t e = cudaFuncSetAttribute(bflow::conv_thin7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            bflow::set_error(cudaGetErrorString(e));
            return BFLOW_ERR_CUDA;
        }
        configured.set();
    }
    bflow::conv_thin7_kernel<<<(unsigned)tiles, 64 * bflow::TH_ROWS, smem, (cudaStream_t)stream>>>(d, bflow::timeline_next_slot("conv_thin7"));
    return bflow::check_launch("bflow_conv2d_thin7");
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
