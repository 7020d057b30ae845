// Shared device helpers of the tcgen05 / TMA kernels (conv_tc3.cu, conv_stem7.cu): mbarrier, bulk-copy, UMMA descriptor, TMEM access.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace bflow {

constexpr int T3_BM = 128;
constexpr int T3_A_BYTES = T3_BM * 128;   // one fp16 A tile (hi or lo): 128 rows x 64 channels
constexpr int T3_EPI_WARPS = 8;            // two per TMEM lane quadrant, each takes half of the tile's columns
constexpr int T3_THREADS = 64 + 32 * T3_EPI_WARPS;

__device__ __forceinline__ uint32_t t3_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t3_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void t3_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void t3_mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool t3_mbar_try_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ bool t3_mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// development: per-role clock64 stamps of CTA 0 (bflow_tc3_trace); the pointer travels as a kernel parameter
#define T3_CTA(i) do { if (p.cta != nullptr) p.cta[blockIdx.x * 16 + (i)] = global_ns(); } while (0)
#define T3_TRACE(slot, idx) do { if (p.trace != nullptr && blockIdx.x == 0 && (idx) < 256) p.trace[(slot) * 256 + (idx)] = clock64(); } while (0)
// bounded: a protocol bug cannot hang the GPU.  After 2^24 failed polls (>= 0.3 s; no legitimate wait is longer than a few hundred
// microseconds) the error word is set and the role carries on, so the kernel still terminates; the result of that forward is
// garbage and the HOST makes it loud: the engine copies the word back with every forward's outputs and raises (Engine._poll_err,
// BezierCurves of a pipelined call, _Plan.check).  A __trap() here was measured: even out of line it costs 1.2 % of the whole
// forward (4.18 -> 4.23 ms, code layout of the ~20 inlined waits), which buys nothing over the host-side check.
__device__ __forceinline__ void t3_mbar_wait(uint32_t bar, uint32_t parity, int* err) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it)
        if (t3_mbar_try_wait(bar, parity)) return;
    if (err != nullptr) atomicExch(err, 1);
}
// one lane of a CONVERGED warp (elect.sync): the branch it guards is the idiom under which nvcc keeps warp-uniform operands in uniform
// registers and issues UTCHMMA / UTCBAR directly.  Under `if (lane == 0)` it wraps every tcgen05.mma in an ELECT + 5 x R2UR + loop
// "waterfall" -- ~70 cycles per instruction on the one thread that feeds the tensor pipe (measured: 576 of 1094 cycles per k-block).
__device__ __forceinline__ bool t3_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void t3_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void t3_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void t3_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
// im2col tile load: {c, w, h, n} = first channel and the input coordinates of the first output pixel's tap (0,0);
// {off_w, off_h} = filter tap
__device__ __forceinline__ void t3_tma_im2col(uint32_t dst, const CUtensorMap* map, int c, int w, int h, int n, uint16_t off_w, uint16_t off_h,
                                              uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}
__device__ __forceinline__ uint64_t t3_umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;     // 8-row groups are 1024 bytes apart
    d |= (uint64_t)1 << 46;               // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;               // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void t3_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void t3_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void t3_tmem_ld16(uint32_t taddr, float* v) {
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

__device__ __forceinline__ void t3_tmem_ld16_nowait(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive accumulator columns of this thread's TMEM lane (= tile row) <- registers
__device__ __forceinline__ void t3_tmem_st16(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
        "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])),
        "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
        "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}

__host__ __device__ constexpr int t3_area_bytes(int bn, int stages) {
    return bn <= 128 ? 224 * 1024 : stages * (2 * T3_A_BYTES + 2 * bn * 128);
}
__device__ __forceinline__ void t3_bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void sl_tma_tile(uint32_t dst, const CUtensorMap* map, int c, int x, int y, int n, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(x), "r"(y), "r"(n)
                 : "memory");
}

__device__ __forceinline__ void t3_tma_store2d(const CUtensorMap* map, uint32_t src, int c, int r) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c),
                 "r"(r)
                 : "memory");
}

__device__ __forceinline__ void t3_tma_load2d(uint32_t dst, const CUtensorMap* map, int c, int r, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(r)
                 : "memory");
}

}  // namespace bflow
