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

// Epilogue shared by the CUDA-core and tensor-core convolutions for 4 consecutive output channels n..n+3 of row m.
// v[] = scale * (acc + bias).  `vec` = all four channels valid and every pointer involved is 16-byte aligned.
__device__ __forceinline__ void conv_epilogue4(const bflow_conv_desc& d, int m, int n, float* v, bool vec) {
    if (d.epi == BFLOW_EPI_STD) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], d.act1);
        if (vec) {
            if (d.res != nullptr) {
                const float4 r = *reinterpret_cast<const float4*>(d.res + (size_t)m * d.ldr + n);
                v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], d.act2);
            *reinterpret_cast<float4*>(d.y + (size_t)m * d.ldy + n) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (n + j < d.Cout) {
                    float o = v[j];
                    if (d.res != nullptr) o += d.res[(size_t)m * d.ldr + n + j];
                    d.y[(size_t)m * d.ldy + n + j] = apply_act(o, d.act2);
                }
            }
        }
    } else if (d.epi == BFLOW_EPI_GRU_ZR) {
        // host guarantees Cout % 8 == 0 and aligned rows, so a group of 4 never straddles the z | r boundary
        const int C = d.Cout >> 1;
        if (n + 3 >= d.Cout) return;
        float4 g = make_float4(v[0], v[1], v[2], v[3]);
        if (d.res != nullptr) {
            const float4 r = *reinterpret_cast<const float4*>(d.res + (size_t)m * d.ldr + n);
            g.x += r.x; g.y += r.y; g.z += r.z; g.w += r.w;
        }
        g.x = apply_act(g.x, BFLOW_ACT_SIGMOID); g.y = apply_act(g.y, BFLOW_ACT_SIGMOID);
        g.z = apply_act(g.z, BFLOW_ACT_SIGMOID); g.w = apply_act(g.w, BFLOW_ACT_SIGMOID);
        *reinterpret_cast<float4*>(d.y + (size_t)m * d.ldy + n) = g;
        if (n >= C) {
            const float4 hv = *reinterpret_cast<const float4*>(d.aux0 + (size_t)m * d.ld_aux0 + (n - C));
            *reinterpret_cast<float4*>(d.aux1 + (size_t)m * d.ld_aux1 + (n - C)) = make_float4(g.x * hv.x, g.y * hv.y, g.z * hv.z, g.w * hv.w);
        }
    } else {   // BFLOW_EPI_GRU_Q
        if (n + 3 >= d.Cout) return;
        float4 q = make_float4(v[0], v[1], v[2], v[3]);
        if (d.res != nullptr) {
            const float4 r = *reinterpret_cast<const float4*>(d.res + (size_t)m * d.ldr + n);
            q.x += r.x; q.y += r.y; q.z += r.z; q.w += r.w;
        }
        q.x = tanhf(q.x); q.y = tanhf(q.y); q.z = tanhf(q.z); q.w = tanhf(q.w);
        const float4 z = *reinterpret_cast<const float4*>(d.aux0 + (size_t)m * d.ld_aux0 + n);
        float4 hv = *reinterpret_cast<const float4*>(d.y + (size_t)m * d.ldy + n);
        hv.x = (1.f - z.x) * hv.x + z.x * q.x;
        hv.y = (1.f - z.y) * hv.y + z.y * q.y;
        hv.z = (1.f - z.z) * hv.z + z.z * q.z;
        hv.w = (1.f - z.w) * hv.w + z.w * q.w;
        *reinterpret_cast<float4*>(d.y + (size_t)m * d.ldy + n) = hv;
    }
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

__host__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// host-side contract of the fused GRU epilogues; returns nullptr when fine
__host__ inline const char* check_epilogue(const bflow_conv_desc& d) {
    if (d.epi == BFLOW_EPI_STD) return nullptr;
    if (d.epi != BFLOW_EPI_GRU_ZR && d.epi != BFLOW_EPI_GRU_Q) return "conv: unknown epilogue mode";
    if (d.Cout % 8 != 0 || d.ldy % 4 != 0 || !aligned16(d.y)) return "conv: GRU epilogue needs Cout % 8 == 0 and aligned output";
    if (d.res != nullptr && (d.ldr % 4 != 0 || !aligned16(d.res))) return "conv: GRU epilogue needs aligned res";
    if (d.aux0 == nullptr || d.ld_aux0 % 4 != 0 || !aligned16(d.aux0)) return "conv: GRU epilogue needs aligned aux0";
    if (d.epi == BFLOW_EPI_GRU_ZR && (d.aux1 == nullptr || d.ld_aux1 % 4 != 0 || !aligned16(d.aux1))) return "conv: GRU_ZR needs aligned aux1";
    return nullptr;
}

}  // namespace bflow
