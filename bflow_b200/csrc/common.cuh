// Shared helpers for the bflow_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
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
// a descriptor built against another version of the header is refused before any field behind struct_size is read
#define BFLOW_CHECK_DESC(dp, type, what)                                                                                          \
    do {                                                                                                                           \
        BFLOW_REQUIRE((dp) != nullptr, what ": null descriptor");                                                                  \
        BFLOW_REQUIRE((dp)->struct_size == (int)sizeof(type), what ": descriptor size mismatch (caller built against another bflow_b200.h; set struct_size = sizeof(" #type "))"); \
    } while (0)

// Per-device host state.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are per DEVICE, the process may drive several:
// every cache on the launch path is keyed by the current device ordinal.
int current_device();      // runtime.cu
int num_sms();             // SMs of the current device (148 on B200)
struct PerDeviceFlag {
    bool done[64] = {};
    bool get() const { const int d = current_device(); return d >= 0 && d < 64 && done[d]; }
    void set() { const int d = current_device(); if (d >= 0 && d < 64) done[d] = true; }
};

// Gate activations on the SFU: ex2.approx + a fast divide (absolute error ~2e-7, the size of fp32 rounding of the gate itself).  The libm
// versions cost 30-50 dependent instructions per element, and the 8 epilogue warps of a tensor-core CTA are latency-bound on exactly that chain
// (measured: 1.6-2.8 us per 1024 four-channel groups).
__device__ __forceinline__ float fast_sigmoid(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }
__device__ __forceinline__ float fast_tanh(float v) {
    // 1 - 2 / (e^{2v} + 1): -> -1 for e^{2v} = 0, -> 1 for e^{2v} = inf (2 / inf = 0)
    const float e = __expf(2.f * fminf(v, 40.f));
    return 1.f - __fdividef(2.f, e + 1.f);
}
__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == BFLOW_ACT_RELU) return fmaxf(v, 0.f);
    if (act == BFLOW_ACT_NONE) return v;
    return act == BFLOW_ACT_SIGMOID ? fast_sigmoid(v) : fast_tanh(v);
}

// split two floats into packed fp16 "hi" and "lo" words (element 0 in the low half): x = hi + lo with hi = fp16(x) and
// lo = fp16(x - hi): 11 + 11 mantissa bits.  Conversions saturate (no inf): |x| up to 1.3e5 is represented, beyond that it
// clamps; residuals below 6e-5 are fp16-subnormal with 6e-8 absolute spacing.  (tcgen05 kind::f16 rejects mixed bf16 x fp16
// operands with an illegal-instruction fault, so a bf16 hi / fp16 lo split is not available.)
__device__ __forceinline__ uint32_t pack_f16x2_sat(float e0, float e1) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
    return r;
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_f16x2_sat(a, b);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    lo = pack_f16x2_sat(a - hf.x, b - hf.y);
}
// 4 floats -> 8 bytes of hi and 8 bytes of lo
__device__ __forceinline__ void store_split4(void* hi_base, void* lo_base, size_t elem, float a, float b, float c, float d) {
    uint2 h, l;
    split2(a, b, h.x, l.x);
    split2(c, d, h.y, l.y);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(hi_base) + elem) = h;
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(lo_base) + elem) = l;
}
__device__ __forceinline__ void store_split1(void* hi_base, void* lo_base, size_t elem, float a) {
    const __half h = __float2half_rn(fminf(fmaxf(a, -65504.f), 65504.f));
    reinterpret_cast<__half*>(hi_base)[elem] = h;
    reinterpret_cast<__half*>(lo_base)[elem] = __float2half_rn(a - __half2float(h));
}
__device__ __forceinline__ float4 load_split4(const void* hi_base, const void* lo_base, size_t elem) {
    const uint2 h = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(hi_base) + elem);
    const uint2 l = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(lo_base) + elem);
    const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
    const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
    return make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
}
// hi plane alone (BFLOW_PREC_F16 tensors)
__device__ __forceinline__ void store_hi4(void* hi_base, size_t elem, float a, float b, float c, float d) {
    uint2 h;
    h.x = pack_f16x2_sat(a, b);
    h.y = pack_f16x2_sat(c, d);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(hi_base) + elem) = h;
}
__device__ __forceinline__ float4 load_hi4(const void* hi_base, size_t elem) {
    const uint2 h = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(hi_base) + elem);
    const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
    return make_float4(h0.x, h0.y, h1.x, h1.y);
}
__device__ __forceinline__ float load_split1(const void* hi_base, const void* lo_base, size_t elem) {
    return __half2float(reinterpret_cast<const __half*>(hi_base)[elem]) + __half2float(reinterpret_cast<const __half*>(lo_base)[elem]);
}

// split residual of a convolution: hi + lo, or the hi plane alone when the producer ran in BFLOW_PREC_F16 (its lo plane is never written)
__device__ __forceinline__ float4 load_res16_4(const bflow_conv_desc& d, size_t elem) {
    return d.precision == BFLOW_PREC_F16 ? load_hi4(d.res16_hi, elem) : load_split4(d.res16_hi, d.res16_lo, elem);
}
__device__ __forceinline__ float load_res16_1(const bflow_conv_desc& d, size_t elem) {
    return d.precision == BFLOW_PREC_F16 ? __half2float(reinterpret_cast<const __half*>(d.res16_hi)[elem]) : load_split1(d.res16_hi, d.res16_lo, elem);
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
            } else if (d.res16_hi != nullptr) {
                const float4 r = load_res16_4(d, (size_t)m * d.ldr16 + n);
                v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], d.act2);
            if (d.y != nullptr) *reinterpret_cast<float4*>(d.y + (size_t)m * d.ldy + n) = make_float4(v[0], v[1], v[2], v[3]);
            if (d.y16_hi != nullptr) store_split4(d.y16_hi, d.y16_lo, (size_t)m * d.ldy16 + n, v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (n + j < d.Cout) {
                    float o = v[j];
                    if (d.res != nullptr) o += d.res[(size_t)m * d.ldr + n + j];
                    else if (d.res16_hi != nullptr) o += load_res16_1(d, (size_t)m * d.ldr16 + n + j);
                    o = apply_act(o, d.act2);
                    if (d.y != nullptr) d.y[(size_t)m * d.ldy + n + j] = o;
                    if (d.y16_hi != nullptr) store_split1(d.y16_hi, d.y16_lo, (size_t)m * d.ldy16 + n + j, o);
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
            if (d.aux1 != nullptr)
                *reinterpret_cast<float4*>(d.aux1 + (size_t)m * d.ld_aux1 + (n - C)) = make_float4(g.x * hv.x, g.y * hv.y, g.z * hv.z, g.w * hv.w);
            if (d.aux1_16_hi != nullptr)
                store_split4(d.aux1_16_hi, d.aux1_16_lo, (size_t)m * d.ld_aux1_16 + (n - C), g.x * hv.x, g.y * hv.y, g.z * hv.z, g.w * hv.w);
        }
    } else {   // BFLOW_EPI_GRU_Q
        if (n + 3 >= d.Cout) return;
        float4 q = make_float4(v[0], v[1], v[2], v[3]);
        if (d.res != nullptr) {
            const float4 r = *reinterpret_cast<const float4*>(d.res + (size_t)m * d.ldr + n);
            q.x += r.x; q.y += r.y; q.z += r.z; q.w += r.w;
        }
        q.x = fast_tanh(q.x); q.y = fast_tanh(q.y); q.z = fast_tanh(q.z); q.w = fast_tanh(q.w);
        const float4 z = *reinterpret_cast<const float4*>(d.aux0 + (size_t)m * d.ld_aux0 + n);
        float4 hv = *reinterpret_cast<const float4*>(d.y + (size_t)m * d.ldy + n);
        hv.x = (1.f - z.x) * hv.x + z.x * q.x;
        hv.y = (1.f - z.y) * hv.y + z.y * q.y;
        hv.z = (1.f - z.z) * hv.z + z.z * q.z;
        hv.w = (1.f - z.w) * hv.w + z.w * q.w;
        *reinterpret_cast<float4*>(d.y + (size_t)m * d.ldy + n) = hv;
        if (d.y16_hi != nullptr) store_split4(d.y16_hi, d.y16_lo, (size_t)m * d.ldy16 + n, hv.x, hv.y, hv.z, hv.w);
    }
}

// The same epilogue in two steps, for callers that batch several groups per thread: every global LOAD a group needs is issued by
// conv_epilogue4_prefetch (so the loads of a whole batch are in flight together), conv_epilogue4_finish does the arithmetic and the
// stores.  `vec` as above; ragged / unaligned groups fall back to conv_epilogue4 inside finish.
struct EpiPre {
    float4 r, a, y;
};
__device__ __forceinline__ void conv_epilogue4_prefetch(const bflow_conv_desc& d, int m, int n, bool vec, EpiPre& e) {
    e.r = make_float4(0.f, 0.f, 0.f, 0.f);
    e.a = e.r;
    e.y = e.r;
    if (!vec) return;
    if (d.epi == BFLOW_EPI_STD) {
        if (d.res != nullptr) e.r = *reinterpret_cast<const float4*>(d.res + (size_t)m * d.ldr + n);
        else if (d.res16_hi != nullptr) e.r = load_res16_4(d, (size_t)m * d.ldr16 + n);
    } else if (d.epi == BFLOW_EPI_GRU_ZR) {
        if (d.res != nullptr) e.r = *reinterpret_cast<const float4*>(d.res + (size_t)m * d.ldr + n);
        const int C = d.Cout >> 1;
        if (n >= C) e.a = *reinterpret_cast<const float4*>(d.aux0 + (size_t)m * d.ld_aux0 + (n - C));
    } else {
        if (d.res != nullptr) e.r = *reinterpret_cast<const float4*>(d.res + (size_t)m * d.ldr + n);
        e.a = *reinterpret_cast<const float4*>(d.aux0 + (size_t)m * d.ld_aux0 + n);
        e.y = *reinterpret_cast<const float4*>(d.y + (size_t)m * d.ldy + n);
    }
}
__device__ __forceinline__ void conv_epilogue4_finish(const bflow_conv_desc& d, int m, int n, float* v, bool vec, const EpiPre& e) {
    if (!vec) {
        conv_epilogue4(d, m, n, v, false);
        return;
    }
    if (d.epi == BFLOW_EPI_STD) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], d.act1);
        v[0] += e.r.x; v[1] += e.r.y; v[2] += e.r.z; v[3] += e.r.w;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], d.act2);
        if (d.y != nullptr) *reinterpret_cast<float4*>(d.y + (size_t)m * d.ldy + n) = make_float4(v[0], v[1], v[2], v[3]);
        if (d.y16_hi != nullptr) store_split4(d.y16_hi, d.y16_lo, (size_t)m * d.ldy16 + n, v[0], v[1], v[2], v[3]);
    } else if (d.epi == BFLOW_EPI_GRU_ZR) {
        const int C = d.Cout >> 1;
        float4 g = make_float4(v[0] + e.r.x, v[1] + e.r.y, v[2] + e.r.z, v[3] + e.r.w);
        g.x = apply_act(g.x, BFLOW_ACT_SIGMOID); g.y = apply_act(g.y, BFLOW_ACT_SIGMOID);
        g.z = apply_act(g.z, BFLOW_ACT_SIGMOID); g.w = apply_act(g.w, BFLOW_ACT_SIGMOID);
        *reinterpret_cast<float4*>(d.y + (size_t)m * d.ldy + n) = g;
        if (n >= C) {
            if (d.aux1 != nullptr)
                *reinterpret_cast<float4*>(d.aux1 + (size_t)m * d.ld_aux1 + (n - C)) = make_float4(g.x * e.a.x, g.y * e.a.y, g.z * e.a.z, g.w * e.a.w);
            if (d.aux1_16_hi != nullptr)
                store_split4(d.aux1_16_hi, d.aux1_16_lo, (size_t)m * d.ld_aux1_16 + (n - C), g.x * e.a.x, g.y * e.a.y, g.z * e.a.z, g.w * e.a.w);
        }
    } else {
        float4 q = make_float4(fast_tanh(v[0] + e.r.x), fast_tanh(v[1] + e.r.y), fast_tanh(v[2] + e.r.z), fast_tanh(v[3] + e.r.w));
        float4 hv = e.y;
        hv.x = (1.f - e.a.x) * hv.x + e.a.x * q.x;
        hv.y = (1.f - e.a.y) * hv.y + e.a.y * q.y;
        hv.z = (1.f - e.a.z) * hv.z + e.a.z * q.z;
        hv.w = (1.f - e.a.w) * hv.w + e.a.w * q.w;
        *reinterpret_cast<float4*>(d.y + (size_t)m * d.ldy + n) = hv;
        if (d.y16_hi != nullptr) store_split4(d.y16_hi, d.y16_lo, (size_t)m * d.ldy16 + n, hv.x, hv.y, hv.z, hv.w);
    }
}

// Development timeline (bflow_timeline): every instrumented launch owns a slot {first CTA start, last CTA end} in globaltimer
// nanoseconds, baked into the launch at record time -- so a CUDA-graph replay leaves the true in-graph schedule behind.
unsigned long long* timeline_next_slot(const char* name);      // runtime.cu; nullptr when the timeline is off
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tl_begin(unsigned long long* slot) {
    if (slot != nullptr && threadIdx.x == 0) atomicMin(slot, global_ns());
}
__device__ __forceinline__ void tl_end(unsigned long long* slot) {
    if (slot != nullptr && (threadIdx.x & 31) == 0) atomicMax(slot + 1, global_ns());
}

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may begin while its predecessor on the stream is
// still running; it must call pdl_wait() before its first access to memory the predecessor touches.  pdl_trigger() lets the
// successor's CTAs be scheduled (their prologue -- barrier init, TMEM allocation, descriptor fetch -- then overlaps our tail).
// Both are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();      // runtime.cu: BFLOW_PDL != 0 (default on)
// CTA budget of a persistent launch: one per SM, or fewer when the descriptor says so (bflow_conv_desc::max_ctas)
inline int grid_cap(const bflow_conv_desc& d) { const int s = num_sms(); return (d.max_ctas > 0 && d.max_ctas < s) ? d.max_ctas : s; }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool on, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (on && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    return launch_pdl_if(true, kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

// Division of a 32-bit unsigned by a launch-time constant without the ~150-cycle IDIV sequence (Granlund-Montgomery): the host
// precomputes (mul, shift), the device spends a multiply-high, a subtract and two shifts.  Exact for every 32-bit dividend.
struct FastDiv {
    unsigned mul, shift, d;
};
__host__ inline FastDiv make_fastdiv(unsigned d) {
    FastDiv f;
    f.d = d < 1u ? 1u : d;
    unsigned l = 0;
    while ((1ull << l) < f.d) ++l;                       // ceil(log2 d)
    f.shift = l;
    f.mul = (unsigned)(((1ull << 32) * ((1ull << l) - f.d)) / f.d + 1ull);
    return f;
}
__device__ __forceinline__ unsigned fastdiv(unsigned n, const FastDiv& f) {
    const unsigned t = __umulhi(n, f.mul);
    return f.shift == 0 ? n : (t + ((n - t) >> 1)) >> (f.shift - 1);
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

__host__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// host-side contract of the fused GRU epilogues; returns nullptr when fine
__host__ inline const char* check_epilogue(const bflow_conv_desc& d) {
    if (d.y == nullptr && d.y16_hi == nullptr) return "conv: no output (y and y16 both null)";
    if (d.y16_hi != nullptr && (d.y16_lo == nullptr || d.ldy16 < d.Cout || d.ldy16 % 4 != 0 ||
                                (reinterpret_cast<uintptr_t>(d.y16_hi) & 7) != 0 || (reinterpret_cast<uintptr_t>(d.y16_lo) & 7) != 0))
        return "conv: split output needs both planes, ldy16 >= Cout, ldy16 % 4 == 0, 8-byte alignment";
    if (d.res16_hi != nullptr && (d.res != nullptr || d.res16_lo == nullptr || d.ldr16 < d.Cout || d.ldr16 % 4 != 0)) return "conv: bad split residual";
    if (d.epi == BFLOW_EPI_STD) return nullptr;
    if (d.epi != BFLOW_EPI_GRU_ZR && d.epi != BFLOW_EPI_GRU_Q) return "conv: unknown epilogue mode";
    if (d.Cout % 8 != 0 || d.ldy % 4 != 0 || !aligned16(d.y)) return "conv: GRU epilogue needs Cout % 8 == 0 and aligned output";
    if (d.res != nullptr && (d.ldr % 4 != 0 || !aligned16(d.res))) return "conv: GRU epilogue needs aligned res";
    if (d.aux0 == nullptr || d.ld_aux0 % 4 != 0 || !aligned16(d.aux0)) return "conv: GRU epilogue needs aligned aux0";
    if (d.epi == BFLOW_EPI_GRU_ZR && d.aux1 == nullptr && d.aux1_16_hi == nullptr) return "conv: GRU_ZR needs aux1 (fp32 or split)";
    if (d.epi == BFLOW_EPI_GRU_ZR && d.aux1 != nullptr && (d.ld_aux1 % 4 != 0 || !aligned16(d.aux1))) return "conv: GRU_ZR needs aligned aux1";
    if (d.aux1_16_hi != nullptr && (d.aux1_16_lo == nullptr || d.ld_aux1_16 % 4 != 0)) return "conv: bad split aux1";
    return nullptr;
}

}  // namespace bflow
