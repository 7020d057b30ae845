// Micro-benchmark: latency of mbarrier try_wait / arrive and of tcgen05.commit -> mbarrier on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_latency mbar_latency.cu && ./mbar_latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__global__ void k(long long* out) {
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(s32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // (a) try_wait on a fresh barrier with parity 1 ("previous phase"), 16 times
        long long t0 = clock64();
        int ok = 0;
        for (int i = 0; i < 16; ++i) ok += try_wait(s32(&bar[0]), 1);
        long long t1 = clock64();
        out[0] = (t1 - t0) / 16; out[1] = ok;
        // (b) arrive (count 1 -> completes phase 0), then try_wait parity 0
        t0 = clock64();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar[1])) : "memory");
        t1 = clock64();
        int spins = 0;
        while (!try_wait(s32(&bar[1]), 0)) ++spins;
        long long t2 = clock64();
        out[2] = t1 - t0; out[3] = t2 - t1; out[4] = spins;
        // (c) tcgen05.commit with nothing pending -> wait
        t0 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar[2])) : "memory");
        t1 = clock64();
        spins = 0;
        while (!try_wait(s32(&bar[2]), 0)) ++spins;
        t2 = clock64();
        out[5] = t1 - t0; out[6] = t2 - t1; out[7] = spins;
        // (d) try_wait on an already completed phase, 16 times
        t0 = clock64();
        ok = 0;
        for (int i = 0; i < 16; ++i) ok += try_wait(s32(&bar[2]), 0);
        t1 = clock64();
        out[8] = (t1 - t0) / 16; out[9] = ok;
    }
    __syncthreads();
    if (warp == 1) { uint32_t a = slot; asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(a) : "memory"); }
    (void)lane;
}
int main() {
    long long* d; cudaMalloc(&d, 16 * 8); cudaMemset(d, 0, 128);
    k<<<1, 64>>>(d); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
    printf("try_wait fresh parity1: %lld cyc each (ok=%lld/16)\n", h[0], h[1]);
    printf("arrive: %lld cyc; then wait: %lld cyc (%lld spins)\n", h[2], h[3], h[4]);
    printf("tcgen05.commit issue: %lld cyc; commit->barrier visible: %lld cyc (%lld spins)\n", h[5], h[6], h[7]);
    printf("try_wait completed phase: %lld cyc each (ok=%lld/16)\n", h[8], h[9]);
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
