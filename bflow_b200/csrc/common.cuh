// Shared helpers for the bflow_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/bflow_b200.h"

namespace bflow {

void set_error(const char* msg);
int check_launch(const char* what);

#define BFLOW_REQUIRE(cond, msg)                       \
    do {                                               \
        if (!(cond)) {                                 \
            ::bflow::set_error(msg " [" #cond "]");    \
            return BFLOW_ERR_INVALID;                  \
        }                                              \
    } while (0)

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case BFLOW_ACT_RELU: return fmaxf(v, 0.f);
        case BFLOW_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case BFLOW_ACT_TANH: return tanhf(v);
        default: return v;
    }
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

__host__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace bflow
