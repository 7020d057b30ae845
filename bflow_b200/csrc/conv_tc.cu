// Tensor-core implicit-GEMM convolution for sm_100a: tcgen05.mma with TMEM accumulators, split-bf16 operands.
//
// Same contract as bflow_conv2d_nhwc (conv_simt.cu):  out = act2(res + act1(scale * (conv(x, w) + bias)))
// on NHWC fp32 activations, up to two channel-concatenated sources.  Replaces the nn.Conv2d calls of
// models/raft_utils/extractor.py:49-53,112,120 and models/raft_spline/update.py:17-18,36-45,89-96,112-114.
//
// Precision: the parity bar is 1e-3 px against an fp32 reference and single-pass bf16/TF32 operands miss it
// (SURVEY.md §7.1), so every operand is split x = hi + lo (two bf16) and each product is issued as
// hi*hi + hi*lo + lo*hi into one fp32 TMEM accumulator (~16 mantissa bits per operand).
//
// One CTA = one 128 x BN output tile (BN in {64,128,256} TMEM columns), 5 warps:
//   warps 0-3  A producers, then epilogue.  Thread t owns 16-byte swizzle chunk (t % 8) of rows t/8 + 16 i:
//              it gathers 8 fp32 channels of the input pixel that row/tap maps to (zero outside the image),
//              splits them into bf16 hi/lo and stores them straight into the SWIZZLE_128B K-major layout the
//              UMMA descriptor expects; fence.proxy.async, then mbarrier arrive on full[stage].
//              Thread 0 also streams the stage's weight tile (host-packed image of the swizzled smem
//              tile, hi then lo) with one cp.async.bulk (UBLKCP) that completes on the same barrier.
//   warp 4     allocates TMEM; one elected lane issues 4 k-steps x 3 tcgen05.mma per stage and
//              tcgen05.commit's the stage's empty barrier (and the accumulator barrier after the last).
//   epilogue   tcgen05.ld 32 lanes x 16 columns -> bias / scale / activation / residual -> float4 stores.
//
// K is the flattened (tap, channel) axis k = (kh*KW + kw)*Cin + c in blocks of 64; a block may straddle taps
// (each 8-channel chunk resolves its own tap), so Cin only has to be a multiple of 8.
#include "common.cuh"
#include <cuda_bf16.h>

namespace bflow {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                 // bf16 elements per k-block = one 128-byte swizzle row
constexpr int TC_THREADS = 160;           // 4 producer/epilogue warps + 1 MMA warp
constexpr int TC_A_BYTES = TC_BM * 128;   // one bf16 A tile (hi or lo)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU.  On timeout the error word is set and the wait returns.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it)
        if (mbar_try_wait(bar, parity)) return;
    if (err != nullptr) atomicExch(err, 1);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);       // start address
    d |= (uint64_t)0 << 16;                        // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// split two floats into packed bf16 hi and lo words (element 0 in the low half)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    float ra = a - __low2float(h), rb = b - __high2float(h);
    __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
}

struct TcSmemTail {
    uint64_t full[4];
    uint64_t empty[4];
    uint64_t accum;
    uint32_t tmem_base;
    int row_n[TC_BM];
    int row_ih0[TC_BM];
    int row_iw0[TC_BM];
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const bflow_conv_desc d, const uint8_t* __restrict__ wtc, const int M, const int K, const int nkb, int* err) {
    constexpr int B_BYTES = BN * 128;                          // one bf16 B tile (hi or lo)
    constexpr int STAGE_BYTES = 2 * TC_A_BYTES + 2 * B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    TcSmemTail* tail = reinterpret_cast<TcSmemTail*>(smem + STAGES * STAGE_BYTES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int m0 = blockIdx.x * TC_BM;
    const int n_tile = blockIdx.y;
    const int n0 = n_tile * BN;
    const int Cin = d.c0 + d.c1;

    if (tid < TC_BM) {
        int m = m0 + tid;
        if (m < M) {
            int ow = m % d.Wo;
            int t = m / d.Wo;
            int oh = t % d.Ho;
            tail->row_n[tid] = t / d.Ho;
            tail->row_ih0[tid] = oh * d.stride - d.pad_h;
            tail->row_iw0[tid] = ow * d.stride - d.pad_w;
        } else {
            tail->row_n[tid] = -1;
            tail->row_ih0[tid] = 0;
            tail->row_iw0[tid] = 0;
        }
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&tail->full[s]), TC_BM);     // 128 producer arrivals (+ weight bytes via complete_tx)
            mbar_init(smem_u32(&tail->empty[s]), 1);        // one tcgen05.commit
        }
        mbar_init(smem_u32(&tail->accum), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)), "r"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;

    if (warp < 4) {
        // ------------------------------------------------ A producers ------------------------------------------------
        const int chunk = tid & 7;                 // 16-byte chunk (8 bf16 = 8 source floats) within the 128-byte row
        const int rbase = tid >> 3;                // rows rbase + 16 i
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
            mbar_wait(smem_u32(&tail->empty[s]), ph ^ 1u, err);
            uint8_t* stage = smem + s * STAGE_BYTES;
            if (tid == 0) {
                const uint32_t bar = smem_u32(&tail->full[s]);
                mbar_expect_tx(bar, 2 * B_BYTES);
                bulk_g2s(smem_u32(stage + 2 * TC_A_BYTES), wtc + ((size_t)n_tile * nkb + kb) * (2 * B_BYTES), 2 * B_BYTES, bar);
            }
            // which tap / channel does this thread's chunk of the k-block fall on?
            const int kflat = kb * TC_BK + chunk * 8;
            const bool k_ok = kflat < K;
            const int tap = k_ok ? kflat / Cin : 0;
            int c = kflat - tap * Cin;
            const int kh = tap / d.KW;
            const int kw = tap - kh * d.KW;
            const float* src;
            int ld;
            if (c < d.c0) { src = d.x0; ld = d.ld0; } else { src = d.x1; ld = d.ld1; c -= d.c0; }
            float4 va[8], vb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rbase + 16 * i;
                const int n = tail->row_n[r];
                const int ih = tail->row_ih0[r] + kh;
                const int iw = tail->row_iw0[r] + kw;
                if (k_ok && n >= 0 && ih >= 0 && ih < d.H && iw >= 0 && iw < d.W) {
                    const float4* p = reinterpret_cast<const float4*>(src + (((size_t)n * d.H + ih) * d.W + iw) * ld + c);
                    va[i] = __ldg(p);
                    vb[i] = __ldg(p + 1);
                } else {
                    va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    vb[i] = va[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rbase + 16 * i;
                uint4 hi, lo;
                split2(va[i].x, va[i].y, hi.x, lo.x);
                split2(va[i].z, va[i].w, hi.y, lo.y);
                split2(vb[i].x, vb[i].y, hi.z, lo.z);
                split2(vb[i].z, vb[i].w, hi.w, lo.w);
                const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
                *reinterpret_cast<uint4*>(stage + off) = hi;
                *reinterpret_cast<uint4*>(stage + TC_A_BYTES + off) = lo;
            }
            fence_proxy_async();                    // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(smem_u32(&tail->full[s]));
        }
        // ------------------------------------------------ epilogue ------------------------------------------------
        mbar_wait(smem_u32(&tail->accum), 0, err);
        tc_fence_after();
        const int row = tid;                        // TMEM lane == tile row; warp w may only touch lanes 32w..32w+31
        const int m = m0 + row;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const bool vec_ok = ((d.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.y) & 15) == 0) &&
                            (d.res == nullptr || (((d.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.res) & 15) == 0)));
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            if (n0 + c0 >= d.Cout) break;           // warp-uniform
            float v[16];
            tmem_ld16(taddr + (uint32_t)c0, v);
            if (m < M) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = n0 + c0 + j;
                    const float b = (d.bias != nullptr && n < d.Cout) ? __ldg(d.bias + n) : 0.f;
                    v[j] = apply_act(d.scale * (v[j] + b), d.act1);
                }
                float* yrow = d.y + (size_t)m * d.ldy + n0 + c0;
                const float* rrow = d.res != nullptr ? d.res + (size_t)m * d.ldr + n0 + c0 : nullptr;
                if (vec_ok && n0 + c0 + 15 < d.Cout) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        if (rrow != nullptr) {
                            float4 r4 = *reinterpret_cast<const float4*>(rrow + j);
                            o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                        }
                        o.x = apply_act(o.x, d.act2); o.y = apply_act(o.y, d.act2);
                        o.z = apply_act(o.z, d.act2); o.w = apply_act(o.w, d.act2);
                        *reinterpret_cast<float4*>(yrow + j) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (n0 + c0 + j < d.Cout) {
                            float o = v[j];
                            if (rrow != nullptr) o += rrow[j];
                            yrow[j] = apply_act(o, d.act2);
                        }
                    }
                }
            }
        }
        tc_fence_before();
    } else {
        // ------------------------------------------------ MMA issuer ------------------------------------------------
        // instruction descriptor: D fp32, A/B bf16, both K-major, N = BN, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        const bool leader = (tid & 31) == 0;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
            mbar_wait(smem_u32(&tail->full[s]), ph, err);
            tc_fence_after();
            if (leader) {
                const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
                const uint32_t a_lo = a_hi + TC_A_BYTES;
                const uint32_t b_hi = a_hi + 2 * TC_A_BYTES;
                const uint32_t b_lo = b_hi + B_BYTES;
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) {
                    const uint32_t ko = (uint32_t)k * 32u;      // 16 bf16 = 32 bytes along K inside the swizzle atom
                    const uint64_t dah = umma_desc(a_hi + ko), dal = umma_desc(a_lo + ko);
                    const uint64_t dbh = umma_desc(b_hi + ko), dbl = umma_desc(b_lo + ko);
                    umma_bf16(tmem_base, dah, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_bf16(tmem_base, dah, dbl, idesc, 1u);
                    umma_bf16(tmem_base, dal, dbh, idesc, 1u);
                }
                umma_commit(smem_u32(&tail->empty[s]));         // frees the stage when these MMAs have read it
                if (kb == nkb - 1) umma_commit(smem_u32(&tail->accum));
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
    }
}

template <int BN, int STAGES>
static int launch_tc(const bflow_conv_desc& d, const void* wtc, int M, int K, int nkb, int* err, cudaStream_t stream) {
    constexpr int smem = STAGES * (2 * TC_A_BYTES + 2 * BN * 128) + (int)sizeof(TcSmemTail) + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error(cudaGetErrorString(e));
            return BFLOW_ERR_CUDA;
        }
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(M, TC_BM), (unsigned)ceil_div(d.Cout, BN));
    conv_tc_kernel<BN, STAGES><<<grid, TC_THREADS, smem, stream>>>(d, reinterpret_cast<const uint8_t*>(wtc), M, K, nkb, err);
    return check_launch("bflow_conv2d_nhwc_tc");
}

}  // namespace bflow

extern "C" int bflow_conv2d_tc_supported(const bflow_conv_desc* dp) {
    if (dp == nullptr) return 0;
    const bflow_conv_desc& d = *dp;
    if (d.c0 <= 0 || d.c0 % 8 != 0 || d.c1 % 8 != 0) return 0;
    if (d.ld0 % 4 != 0 || !bflow::aligned16(d.x0)) return 0;
    if (d.c1 > 0 && (d.ld1 % 4 != 0 || !bflow::aligned16(d.x1))) return 0;
    return 1;
}

// w_tc: host-packed weight image [ceil(Cout/bn)][nkb][hi|lo][bn rows][64 bf16, 16-byte chunks XOR-swizzled by row%8]
// (bflow_b200/ops.py: pack_conv_weight_tc).  err: optional device int set to 1 if a pipeline wait timed out.
extern "C" int bflow_conv2d_nhwc_tc(const bflow_conv_desc* dp, const void* w_tc, int bn, int* err, void* stream) {
    BFLOW_REQUIRE(dp != nullptr && w_tc != nullptr, "conv_tc: null argument");
    const bflow_conv_desc& d = *dp;
    BFLOW_REQUIRE(d.x0 != nullptr && d.y != nullptr, "conv_tc: null tensor");
    BFLOW_REQUIRE(bflow_conv2d_tc_supported(dp) == 1, "conv_tc: needs channels % 8 == 0 and 16-byte aligned rows");
    BFLOW_REQUIRE(d.c1 == 0 || d.x1 != nullptr, "conv_tc: bad source 1");
    BFLOW_REQUIRE(d.N > 0 && d.H > 0 && d.W > 0 && d.Cout > 0, "conv_tc: bad shape");
    BFLOW_REQUIRE(d.Ho == (d.H + 2 * d.pad_h - d.KH) / d.stride + 1 && d.Wo == (d.W + 2 * d.pad_w - d.KW) / d.stride + 1, "conv_tc: Ho/Wo mismatch");
    BFLOW_REQUIRE(d.ldy >= d.Cout && (d.res == nullptr || d.ldr >= d.Cout), "conv_tc: bad output stride");
    BFLOW_REQUIRE((reinterpret_cast<uintptr_t>(w_tc) & 15) == 0, "conv_tc: packed weights must be 16-byte aligned");
    const long long Mll = (long long)d.N * d.Ho * d.Wo;
    const long long Kll = (long long)d.KH * d.KW * (d.c0 + d.c1);
    BFLOW_REQUIRE(Mll < (1ll << 31) && Kll < (1ll << 31), "conv_tc: too large");
    const int M = (int)Mll, K = (int)Kll, nkb = (K + bflow::TC_BK - 1) / bflow::TC_BK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (bn) {
        case 64: return bflow::launch_tc<64, 4>(d, w_tc, M, K, nkb, err, st);
        case 128: return bflow::launch_tc<128, 3>(d, w_tc, M, K, nkb, err, st);
        case 256: return bflow::launch_tc<256, 2>(d, w_tc, M, K, nkb, err, st);
        default: bflow::set_error("conv_tc: bn must be 64, 128 or 256"); return BFLOW_ERR_INVALID;
    }
}
