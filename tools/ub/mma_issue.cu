// Micro-benchmark: what one tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) costs on sm_100a as a function of N and of where A comes from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue mma_issue.cu && ./mma_issue
// One CTA per SM (all 148 busy, as in the convolution kernels); warp 0 issues R back-to-back MMAs under elect.sync (operands in uniform
// registers, the idiom of csrc/conv_tc3.cu), commits to an mbarrier and waits; cycles per instruction = (clock64 after - before) / R.
//   SS   A and B from SWIZZLE_128B shared-memory tiles (what every kernel of this repository does)
//   TS   A from tensor memory (8 columns per k-step), B from shared memory
// `acc2` alternates between two accumulators (no dependency between consecutive instructions).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc) : "memory");
}

constexpr int R = 2048;

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// FILL: warp 2 keeps `FILL` 16 KB bulk copies (global -> a spare shared-memory ring) in flight for as long as the MMAs run: the TMA fills of a real
// pipeline, on the same shared-memory port as the operand reads
// SYNC (bit mask), after every group of 8 instructions = one k-block of the convolution kernels: 1 = tcgen05.commit to a second mbarrier,
// 2 = mbarrier.try_wait on an already completed barrier (the stage-full wait), 4 = tcgen05.fence::after_thread_sync, 8 = leave and re-enter the
// elect.sync region (__syncwarp + elect per group, as the real loops do)
template <bool TS, bool ACC2, int M, int FILL = 0, int SYNC = 0>
__global__ void __launch_bounds__(128, 1) k(long long* out, int N, const uint8_t* gsrc = nullptr) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint64_t fbar[8];
    __shared__ uint32_t slot;
    __shared__ volatile int done_flag;
    __shared__ long long fill_count;
    const uint32_t base = (s32(raw) + 1023u) & ~1023u;
    const uint32_t sa = base, sb = base + 16384;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - s32(raw)))[i] = 0x3c003c00u;      // 1.0h
    if (threadIdx.x == 0) {
        done_flag = 0;
        fill_count = 0;
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&fbar[i])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t da = umma_desc(sa), db = umma_desc(sb);
        const uint32_t a_tmem = tmem + 480u;                   // 32 columns behind the accumulators (garbage values: timing only)
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2; ++rep) {                    // rep 0 warms up
            __syncwarp();
            t0 = clock64();
            if (SYNC & 8) {
#pragma unroll 1
                for (int i = 0; i < R; i += 8) {
                    if (SYNC & 2) while (!try_wait(s32(&fbar[7]), 1u)) {}             // a fresh barrier: the "previous phase" parity completes at once
                    if (SYNC & 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint64_t ko = (uint64_t)((j & 3) * 2);
                            if (TS) mma_ts(tmem, a_tmem + (uint32_t)((j & 3) * 8), db + ko, idesc);
                            else mma_ss(tmem, da + ko, db + ko, idesc);
                        }
                        if (SYNC & 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&fbar[6])) : "memory");
                    }
                    __syncwarp();
                }
                if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
            } else
            if (elect_one()) {
#pragma unroll 1
                for (int i = 0; i < R; i += 8) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint64_t ko = (uint64_t)((j & 3) * 2);
                        const uint32_t d = tmem + ((ACC2 && (j & 1)) ? 256u : 0u);
                        if (TS) mma_ts(d, a_tmem + (uint32_t)((j & 3) * 8), db + ko, idesc);
                        else mma_ss(d, da + ko, db + ko, idesc);
                    }
                    if (SYNC & 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&fbar[6])) : "memory");
                    if (SYNC & 2) while (!try_wait(s32(&fbar[7]), 1u)) {}
                    if (SYNC & 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
            }
            __syncwarp();
            while (!try_wait(s32(&bar), (uint32_t)rep & 1u)) {}
            t1 = clock64();
        }
        if (threadIdx.x == 0) done_flag = 1;
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    } else if (FILL > 0 && warp == 2 && (threadIdx.x & 31) == 0) {
        const uint32_t ring = base + 16384 + 32768;              // FILL x 16 KB behind the operand tiles
        const uint8_t* src = gsrc + (size_t)blockIdx.x * (1u << 20);
        uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long n = 0;
        for (int i = 0; i < FILL; ++i) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&fbar[i])), "r"(16384) : "memory");
            bulk_g2s(ring + (uint32_t)i * 16384u, src + ((n * 16384) & ((1u << 20) - 1)), 16384, s32(&fbar[i]));
            ++n;
        }
        int i = 0;
        while (!done_flag) {
            while (!try_wait(s32(&fbar[i]), ph[i])) {}
            ph[i] ^= 1u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&fbar[i])), "r"(16384) : "memory");
            bulk_g2s(ring + (uint32_t)i * 16384u, src + ((n * 16384) & ((1u << 20) - 1)), 16384, s32(&fbar[i]));
            ++n;
            i = (i + 1 == FILL) ? 0 : i + 1;
        }
        for (int j = 0; j < FILL; ++j) while (!try_wait(s32(&fbar[j]), ph[j])) {}      // drain before the CTA exits
        if (blockIdx.x == 0) out[1] = n * 16384;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static uint8_t* g_src = nullptr;
static double g_fill_bytes_per_cycle = 0.0;
template <bool TS, bool ACC2, int M, int FILL = 0, int SYNC = 0>
static double run(int N, int grid, long long* dout) {
    const int smem = 16384 + 32768 + 1024 + FILL * 16384;
    cudaFuncSetAttribute(k<TS, ACC2, M, FILL, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<TS, ACC2, M, FILL, SYNC><<<grid, 128, smem>>>(dout, N, g_src);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("error: %s\n", cudaGetErrorString(e));
        return -1.0;
    }
    long long c[2] = {0, 0};
    cudaMemcpy(c, dout, sizeof(c), cudaMemcpyDeviceToHost);
    g_fill_bytes_per_cycle = c[0] > 0 ? (double)c[1] / (2.0 * (double)c[0]) : 0.0;      // the copier ran through both repetitions
    return (double)c[0] / R;
}

int main() {
    long long* dout;
    cudaMalloc(&dout, 64);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("tcgen05.mma kind::f16 cta_group::1, K = 16: SM cycles per instruction (%d back-to-back, %d CTAs = one per SM)\n", R, sms);
    printf("%5s %5s | %9s %9s | %9s %9s | %10s\n", "M", "N", "SS", "SS acc2", "TS", "TS acc2", "N/2 (math)");
    const int Ns[] = {16, 32, 64, 96, 128, 192, 256};
    for (int N : Ns) {
        const double a = run<false, false, 128>(N, sms, dout), b = run<false, true, 128>(N, sms, dout);
        const double c = run<true, false, 128>(N, sms, dout), d = run<true, true, 128>(N, sms, dout);
        printf("%5d %5d | %9.1f %9.1f | %9.1f %9.1f | %10.1f\n", 128, N, a, b, c, d, N / 2.0);
    }
    for (int N : {64, 128, 256}) {
        const double a = run<false, false, 64>(N, sms, dout), c = run<true, false, 64>(N, sms, dout);
        printf("%5d %5d | %9.1f %9s | %9.1f %9s | %10.1f\n", 64, N, a, "-", c, "-", N / 4.0);
    }
    cudaMalloc(&g_src, (size_t)sms << 20);
    cudaMemset(g_src, 0, (size_t)sms << 20);
    printf("with TMA-style fills running beside the MMAs (bulk copies of 16 KB global -> shared memory, k in flight, L2-resident source):\n");
    printf("%5s %5s | %9s %12s | %9s %12s | %9s %12s\n", "M", "N", "SS k=2", "fill B/clk", "SS k=6", "fill B/clk", "TS k=6", "fill B/clk");
    for (int N : {64, 128, 256}) {
        const double a = run<false, false, 128, 2>(N, sms, dout);
        const double fa = g_fill_bytes_per_cycle;
        const double b = run<false, false, 128, 6>(N, sms, dout);
        const double fb = g_fill_bytes_per_cycle;
        const double c = run<true, false, 128, 6>(N, sms, dout);
        const double fc = g_fill_bytes_per_cycle;
        printf("%5d %5d | %9.1f %12.1f | %9.1f %12.1f | %9.1f %12.1f\n", 128, N, a, fa, b, fb, c, fc);
    }
    printf("pipeline bookkeeping after every 8 instructions (SS, M 128): cycles per instruction\n");
    printf("%5s | %8s %8s %8s %8s %10s %12s %14s\n", "N", "none", "commit", "wait", "fence", "c+w+f", "elect/group", "elect+c+w+f");
    for (int N : {64, 128, 256}) {
        printf("%5d | %8.1f %8.1f %8.1f %8.1f %10.1f %12.1f %14.1f\n", N, run<false, false, 128, 0, 0>(N, sms, dout), run<false, false, 128, 0, 1>(N, sms, dout),
               run<false, false, 128, 0, 2>(N, sms, dout), run<false, false, 128, 0, 4>(N, sms, dout), run<false, false, 128, 0, 7>(N, sms, dout),
               run<false, false, 128, 0, 8>(N, sms, dout), run<false, false, 128, 0, 15>(N, sms, dout));
    }
    const double one = run<false, false, 128>(128, 1, dout);
    printf("one CTA alone, M 128 N 128 SS: %.1f\n", one);
    return 0;
}
