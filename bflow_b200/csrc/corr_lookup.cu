// Correlation-pyramid window lookup — the HBM-roofline kernel of the path.
//
// Replaces CorrBlockParallelMultiTarget.__call__ (models/raft_utils/corr.py:307-350) and
// bilinear_sampler / F.grid_sample (models/raft_utils/utils.py:5-21): ~20 ATen launches per pyramid
// level per iteration in the reference, ONE launch here for all level-slots.
//
// Unit of work = (query pixel, slot).  All 81 taps of a unit share one fractional offset, so the unit
// reads a 10x10 footprint (<= 400 B, 10 row segments of 40 B) from the query's private plane and writes
// 81 values: 724 algorithmic bytes per unit.
//
// Mapping: a CTA owns 32 consecutive queries of one slot; each of its 8 warps handles 4 queries.
//   phase 1  every lane issues its 4x4 footprint loads back to back (16 independent LDGs in flight per
//            lane; rows outside the plane are predicated off = zero padding per corner)
//   phase 2  footprints are exchanged through shared memory; lane k computes taps k, k+32, k+64
//   phase 3  NHWC: the warp writes its query's 81 contiguous floats (coalesced 324 B)
//            NCHW: taps are transposed through shared memory so that each channel row receives 32
//            consecutive pixels (128 B segments)
#include "common.cuh"

namespace bflow {

constexpr int LK_QPB = 32;        // queries per CTA
constexpr int LK_WARPS = 8;
constexpr int LK_QPW = LK_QPB / LK_WARPS;   // 4 queries per warp
constexpr int LK_FP = 104;        // footprint stride (100 used)

__global__ void __launch_bounds__(LK_WARPS * 32) corr_lookup_kernel(const bflow_lookup_desc d) {
    __shared__ float fp[LK_QPB][LK_FP];
    __shared__ float tile[81][LK_QPB + 1];
    __shared__ float s_fx[LK_QPB], s_fy[LK_QPB];

    const int slot = blockIdx.y;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int Q = d.h * d.w;
    const long long BQ = (long long)d.B * Q;
    const long long bq0 = (long long)blockIdx.x * LK_QPB;

    const float* __restrict__ vol = d.vol[slot];
    const int hl = d.hl[slot], wl = d.wl[slot];
    const int t = d.target[slot];
    const float inv_scale = d.inv_scale[slot];
    const int plane = hl * wl;

    float v[LK_QPW][4];
#pragma unroll
    for (int i = 0; i < LK_QPW; ++i) {
        const int ql = warp * LK_QPW + i;
        const long long bq = bq0 + ql;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][j] = 0.f;
        if (bq >= BQ) continue;
        const int b = (int)(bq / Q);
        const int q = (int)(bq - (long long)b * Q);
        float cx, cy;
        if (d.coords != nullptr) {
            const float* c = d.coords + (((size_t)t * d.B + b) * 2) * Q + q;
            cx = __ldg(c);
            cy = __ldg(c + Q);
        } else {
            // coords1 = pixel grid + sum_i coef[t][i] * P_i   (raft.py:180-181, bezier.py:165-186)
            const float* p = d.params + (size_t)bq * d.params_ld;
            float fxv = 0.f, fyv = 0.f;
            for (int k = 0; k < d.degree; ++k) {
                float ck = d.coef[t][k];
                fxv = fmaf(ck, __ldg(p + k), fxv);
                fyv = fmaf(ck, __ldg(p + d.degree + k), fyv);
            }
            cx = (float)(q % d.w) + fxv;
            cy = (float)(q / d.w) + fyv;
        }
        cx *= inv_scale;
        cy *= inv_scale;
        // anything farther than the window from the plane samples zeros; clamp to keep the int cast defined
        cx = fminf(fmaxf(cx, -16.f), (float)wl + 16.f);
        cy = fminf(fmaxf(cy, -16.f), (float)hl + 16.f);
        const float flx = floorf(cx), fly = floorf(cy);
        const int x0 = (int)flx - d.radius;
        const int y0 = (int)fly - d.radius;
        if (lane == 0) {
            s_fx[ql] = cx - flx;
            s_fy[ql] = cy - fly;
        }
        const float* pl = vol + (size_t)bq * plane;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = lane + 32 * j;
            if (e < 100) {
                const int r = e / 10, c = e - r * 10;
                const int yy = y0 + r, xx = x0 + c;
                if (yy >= 0 && yy < hl && xx >= 0 && xx < wl) v[i][j] = __ldg(pl + yy * wl + xx);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < LK_QPW; ++i) {
        const int ql = warp * LK_QPW + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = lane + 32 * j;
            if (e < 100) fp[ql][e] = v[i][j];
        }
    }
    __syncwarp();

    const int n = 2 * d.radius + 1;      // 9
    const int ntap = n * n;              // 81
#pragma unroll
    for (int i = 0; i < LK_QPW; ++i) {
        const int ql = warp * LK_QPW + i;
        const long long bq = bq0 + ql;
        if (bq >= BQ) continue;
        const float fx = s_fx[ql], fy = s_fy[ql];
        const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int k = lane + 32 * j;
            if (k < ntap) {
                const int iy = k / n, ix = k - iy * n;
                const float* f = &fp[ql][iy * 10 + ix];
                float o = w00 * f[0] + w01 * f[1] + w10 * f[10] + w11 * f[11];
                if (d.out_nhwc)
                    d.out[(size_t)bq * d.out_ld + slot * ntap + k] = o;
                else
                    tile[k][ql] = o;
            }
        }
    }
    if (!d.out_nhwc) {
        __syncthreads();
        const int S = d.n_slots;
        for (int idx = threadIdx.x; idx < ntap * LK_QPB; idx += LK_WARPS * 32) {
            const int k = idx / LK_QPB, ql = idx - k * LK_QPB;
            const long long bq = bq0 + ql;
            if (bq < BQ) {
                const int b = (int)(bq / Q);
                const int q = (int)(bq - (long long)b * Q);
                d.out[((size_t)b * S * ntap + (size_t)slot * ntap + k) * Q + q] = tile[k][ql];
            }
        }
    }
}

}  // namespace bflow

extern "C" int bflow_corr_lookup(const bflow_lookup_desc* dp, void* stream) {
    BFLOW_REQUIRE(dp != nullptr, "lookup: null descriptor");
    const bflow_lookup_desc& d = *dp;
    BFLOW_REQUIRE(d.n_slots > 0 && d.n_slots <= BFLOW_MAX_SLOTS, "lookup: bad slot count");
    BFLOW_REQUIRE(d.n_targets > 0 && d.n_targets <= BFLOW_MAX_TARGETS, "lookup: bad target count");
    BFLOW_REQUIRE(d.B > 0 && d.h > 0 && d.w > 0, "lookup: bad shape");
    BFLOW_REQUIRE(d.radius == 4, "lookup: radius is fixed to 4 (raft.py:38-40, corr.py:279)");
    BFLOW_REQUIRE(d.out != nullptr, "lookup: null output");
    BFLOW_REQUIRE(d.coords != nullptr || (d.params != nullptr && d.degree >= 1 && d.degree <= BFLOW_MAX_DEGREE &&
                                          d.params_ld >= 2 * d.degree),
                  "lookup: need coords or Bezier params");
    BFLOW_REQUIRE(!d.out_nhwc || d.out_ld >= d.n_slots * 81, "lookup: out_ld too small");
    for (int s = 0; s < d.n_slots; ++s) {
        BFLOW_REQUIRE(d.vol[s] != nullptr && d.hl[s] > 0 && d.wl[s] > 0, "lookup: bad pyramid level");
        BFLOW_REQUIRE(d.target[s] >= 0 && d.target[s] < d.n_targets, "lookup: bad slot target");
    }
    const long long BQ = (long long)d.B * d.h * d.w;
    dim3 grid((unsigned)bflow::ceil_div_ll(BQ, bflow::LK_QPB), (unsigned)d.n_slots);
    bflow::corr_lookup_kernel<<<grid, bflow::LK_WARPS * 32, 0, (cudaStream_t)stream>>>(d);
    return bflow::check_launch("bflow_corr_lookup");
}
