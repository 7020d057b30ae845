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

template <bool TS, bool ACC2, int M>
__global__ void __launch_bounds__(128, 1) k(long long* out, int N) {
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (s32(raw) + 1023u) & ~1023u;
    const uint32_t sa = base, sb = base + 16384;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - s32(raw)))[i] = 0x3c003c00u;      // 1.0h
    if (threadIdx.x == 0) {
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
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
            }
            __syncwarp();
            while (!try_wait(s32(&bar), (uint32_t)rep & 1u)) {}
            t1 = clock64();
        }
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <bool TS, bool ACC2, int M>
static double run(int N, int grid, long long* dout) {
    const int smem = 16384 + 32768 + 1024;
    cudaFuncSetAttribute(k<TS, ACC2, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<TS, ACC2, M><<<grid, 128, smem>>>(dout, N);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("error: %s\n", cudaGetErrorString(e));
        return -1.0;
    }
    long long c = 0;
    cudaMemcpy(&c, dout, sizeof(c), cudaMemcpyDeviceToHost);
    return (double)c / R;
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
    const double one = run<false, false, 128>(128, 1, dout);
    printf("one CTA alone, M 128 N 128 SS: %.1f\n", one);
    return 0;
}
