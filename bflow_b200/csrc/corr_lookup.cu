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
#include <stdlib.h>
#include "common.cuh"

namespace bflow {

// 4x4-pixel-tiled plane layout: tile (y>>2, x>>2) row-major over ceil(w/4) tiles, 16 floats per tile
__host__ __device__ __forceinline__ int tiled_index(int y, int x, int w) {
    return ((((y >> 2) * ((w + 3) >> 2)) + (x >> 2)) << 4) + ((y & 3) << 2) + (x & 3);
}
__host__ __device__ __forceinline__ int tiled_plane_size(int h, int w) { return ((h + 3) >> 2) * ((w + 3) >> 2) * 16; }

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
    const int plane = d.tiled ? tiled_plane_size(hl, wl) : hl * wl;

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
                if (yy >= 0 && yy < hl && xx >= 0 && xx < wl) v[i][j] = __ldg(pl + (d.tiled ? tiled_index(yy, xx, wl) : yy * wl + xx));
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

// -------------------------------------------------------------------------------------------------------------------
// v2: granule-tiled volume.  Every plane is stored as 4x4-pixel tiles (64 B = one DRAM access granule, one float4 per tile
// row, zero padded to multiples of 4), so a 10x10 window touches a 3x3 .. 4x4 block of tiles (10.6 on average, ~680 B)
// instead of ten 40-byte row segments that each cost one or two 64-byte granules (~1100 B measured with ncu).
// One warp per unit: 64 lane-slots = 16 tile rows x 4 tile columns, two LDG.128 per lane, the 16x16 patch goes through
// shared memory, then lane k computes taps k, k+32, k+64 with tap offsets precomputed once per warp.
// -------------------------------------------------------------------------------------------------------------------
constexpr int LK2_WARPS = 8;
constexpr int LK2_PITCH = 20;      // floats per patch row in shared memory (16 used; 20 spreads the tap reads over banks)

// Host-precomputed divisors (query pixel index -> sample, row, column; unit index -> pixel, slot): the integer divisions were a
// quarter of the 346 instructions per unit that made this kernel issue-bound (ncu, round 1: 80 % issue slots busy at 38 % DRAM).
struct LookupDivs {
    FastDiv S, Q, w;
};

// one unit = (query pixel bq, slot): fetch the 3x3..4x4 tiles under the 10x10 window, blend the 81 taps, store them
__device__ __forceinline__ void lookup_unit(const bflow_lookup_desc& d, const unsigned bq, const unsigned b, const unsigned q, const float gx,
                                            const float gy, const int slot, float* ps, const int (&soff)[3], const int lane, const int Q) {
    const int tcA = lane & 3, rgA = lane >> 2;            // tile column, patch row 0..7   (second float4: patch row + 8)
    const int trA = rgA >> 2, rrA = rgA & 3;              // tile row 0..1 (+2 for the second float4), row inside the tile
    const int hl = d.hl[slot], wl = d.wl[slot];
    const int t = d.target[slot];
    float cx, cy;
    if (d.coords != nullptr) {
        const float* c = d.coords + (((size_t)t * d.B + b) * 2) * Q + q;
        cx = __ldg(c);
        cy = __ldg(c + Q);
    } else {
        // coords1 = pixel grid + sum_i coef[t][i] * P_i   (raft.py:180-181, bezier.py:165-186)
        const float* prm = d.params + (size_t)bq * d.params_ld;
        float fxv, fyv;
        if (d.degree == 2 && (d.params_ld & 3) == 0) {    // [P1x P2x P1y P2y]: one 16-byte load (host contract: 16-byte aligned rows)
            const float4 p4 = *reinterpret_cast<const float4*>(prm);
            const float c0 = d.coef[t][0], c1 = d.coef[t][1];
            fxv = fmaf(c1, p4.y, c0 * p4.x);
            fyv = fmaf(c1, p4.w, c0 * p4.z);
        } else {
            fxv = 0.f;
            fyv = 0.f;
            for (int k = 0; k < d.degree; ++k) {
                const float ck = d.coef[t][k];
                fxv = fmaf(ck, prm[k], fxv);
                fyv = fmaf(ck, prm[d.degree + k], fyv);
            }
        }
        cx = gx + fxv;
        cy = gy + fyv;
    }
    const float inv_scale = d.inv_scale[slot];
    // anything farther than the window from the plane samples zeros; clamp to keep the int cast defined
    cx = fminf(fmaxf(cx * inv_scale, -16.f), (float)wl + 16.f);
    cy = fminf(fmaxf(cy * inv_scale, -16.f), (float)hl + 16.f);
    const float flx = floorf(cx), fly = floorf(cy);
    const float fx = cx - flx, fy = cy - fly;
    const int x0 = (int)flx - 4, y0 = (int)fly - 4;
    const int tx0 = x0 >> 2, ty0 = y0 >> 2, ox = x0 & 3, oy = y0 & 3;
    const int ntx = ((ox + 9) >> 2) + 1, nty = ((oy + 9) >> 2) + 1;   // 3 or 4 tiles per axis
    const int tw = (wl + 3) >> 2, th = (hl + 3) >> 2;
    const float* pl = d.vol[slot] + (size_t)bq * (size_t)(tw * th * 16);

    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
    {
        const int txx = tx0 + tcA;
        const bool colok = tcA < ntx && (unsigned)txx < (unsigned)tw;
        const int tya = ty0 + trA, tyb = tya + 2;
        const float* pa = pl + ((tya * tw + txx) << 4) + (rrA << 2);
        if (colok && (unsigned)tya < (unsigned)th) va = __ldg(reinterpret_cast<const float4*>(pa));
        if (colok && (trA + 2) < nty && (unsigned)tyb < (unsigned)th) vb = __ldg(reinterpret_cast<const float4*>(pa + (tw << 5)));
    }
    __syncwarp();                                     // previous unit's tap reads are done
    *reinterpret_cast<float4*>(ps + rgA * LK2_PITCH + tcA * 4) = va;
    *reinterpret_cast<float4*>(ps + (rgA + 8) * LK2_PITCH + tcA * 4) = vb;
    __syncwarp();
    const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
    const float* f0 = ps + oy * LK2_PITCH + ox;
    float val[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float* f = f0 + soff[j];                // j == 2: lanes >= 17 read inside the patch (soff clamped) and do not store
        val[j] = w00 * f[0] + w01 * f[1] + w10 * f[LK2_PITCH] + w11 * f[LK2_PITCH + 1];
    }
    if (d.out16_hi != nullptr) {
        const size_t e = (size_t)bq * d.out16_ld + slot * 81 + lane;
        __half* ph = reinterpret_cast<__half*>(d.out16_hi) + e;
        if (d.out16_lo != nullptr) {
            __half* plo = reinterpret_cast<__half*>(d.out16_lo) + e;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if (j < 2 || lane < 17) {
                    const __half h = __float2half_rn(fminf(fmaxf(val[j], -65504.f), 65504.f));
                    ph[32 * j] = h;
                    plo[32 * j] = __float2half_rn(val[j] - __half2float(h));
                }
            }
        } else {                                      // BFLOW_PREC_F16 consumers: hi plane only
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (j < 2 || lane < 17) ph[32 * j] = __float2half_rn(fminf(fmaxf(val[j], -65504.f), 65504.f));
        }
    } else {
        float* po = d.out + (size_t)bq * d.out_ld + slot * 81 + lane;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (j < 2 || lane < 17) po[32 * j] = val[j];
    }
}

// FLAT = false: a warp walks query pixels (stride = warps in the grid) and, per pixel, the slots of its blockIdx.y group (no division
//               in the inner loop) -- the bandwidth regime (many pixels).
// FLAT = true:  units (pixel, slot) are dealt round-robin to the warps of a one-wave grid -- the batch-1 regime, where the launch is a few
//               latency chains long and an uneven tail (a partial second wave of CTAs) would double it.
template <bool FLAT>
__global__ void __launch_bounds__(LK2_WARPS * 32) corr_lookup_tiled_kernel(const bflow_lookup_desc d, const LookupDivs dv, const long long n_units, const int slots_per_group,
                                                                           unsigned long long* tl) {
    __shared__ __align__(16) float patch[LK2_WARPS][16 * LK2_PITCH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = d.n_slots;
    const int Q = d.h * d.w;
    float* ps = patch[warp];
    tl_begin(tl);
    pdl_trigger();
    pdl_wait();
    int soff[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int k = min(lane + 32 * j, 80);
        const int iy = k / 9, ix = k - iy * 9;
        soff[j] = iy * LK2_PITCH + ix;
    }
    if (FLAT) {
        const unsigned nu = (unsigned)n_units, nw = gridDim.x * LK2_WARPS;
#pragma unroll 1
        for (unsigned u = blockIdx.x * LK2_WARPS + warp; u < nu; u += nw) {
            const unsigned bq = fastdiv(u, dv.S);
            const int slot = (int)(u - bq * (unsigned)S);
            const unsigned b = fastdiv(bq, dv.Q);
            const unsigned q = bq - b * (unsigned)Q;
            const unsigned qy = fastdiv(q, dv.w);
            lookup_unit(d, bq, b, q, (float)(q - qy * (unsigned)d.w), (float)qy, slot, ps, soff, lane, Q);
        }
    } else {
        const unsigned n_bq = (unsigned)(n_units / S);
        for (unsigned bq = blockIdx.x * LK2_WARPS + warp; bq < n_bq; bq += gridDim.x * LK2_WARPS) {
            const unsigned b = fastdiv(bq, dv.Q);
            const unsigned q = bq - b * (unsigned)Q;
            const unsigned qy = fastdiv(q, dv.w);
            const float gx = (float)(q - qy * (unsigned)d.w), gy = (float)qy;
            const int s_beg = (int)blockIdx.y * slots_per_group, s_end = min(S, s_beg + slots_per_group);
            // (a two-deep software pipeline over the slots was measured 10 % slower: 60 registers cost more occupancy than the
            //  extra loads in flight gain)
#pragma unroll 1
            for (int slot = s_beg; slot < s_end; ++slot) lookup_unit(d, bq, b, q, gx, gy, slot, ps, soff, lane, Q);
        }
    }
    tl_end(tl);
}

}  // namespace bflow

extern "C" int bflow_corr_lookup(const bflow_lookup_desc* dp, void* stream) {
    BFLOW_CHECK_DESC(dp, bflow_lookup_desc, "lookup");
    const bflow_lookup_desc& d = *dp;
    BFLOW_REQUIRE(d.n_slots > 0 && d.n_slots <= BFLOW_MAX_SLOTS, "lookup: bad slot count");
    BFLOW_REQUIRE(d.n_targets > 0 && d.n_targets <= BFLOW_MAX_TARGETS, "lookup: bad target count");
    BFLOW_REQUIRE(d.B > 0 && d.h > 0 && d.w > 0, "lookup: bad shape");
    BFLOW_REQUIRE(d.radius == 4, "lookup: radius is fixed to 4 (raft.py:38-40, corr.py:279)");
    BFLOW_REQUIRE(d.out != nullptr || d.out16_hi != nullptr, "lookup: null output");
    BFLOW_REQUIRE(d.out16_hi == nullptr || (d.tiled && d.out_nhwc && d.out16_ld >= d.n_slots * 81),
                  "lookup: split-fp16 output needs the tiled NHWC path");
    BFLOW_REQUIRE(d.coords != nullptr || (d.params != nullptr && d.degree >= 1 && d.degree <= BFLOW_MAX_DEGREE &&
                                          d.params_ld >= 2 * d.degree),
                  "lookup: need coords or Bezier params");
    BFLOW_REQUIRE(!d.out_nhwc || d.out16_hi != nullptr || d.out_ld >= d.n_slots * 81, "lookup: out_ld too small");
    for (int s = 0; s < d.n_slots; ++s) {
        BFLOW_REQUIRE(d.vol[s] != nullptr && d.hl[s] > 0 && d.wl[s] > 0, "lookup: bad pyramid level");
        BFLOW_REQUIRE(d.target[s] >= 0 && d.target[s] < d.n_targets, "lookup: bad slot target");
    }
    const long long BQ = (long long)d.B * d.h * d.w;
    if (d.tiled && d.out_nhwc) {
        BFLOW_REQUIRE(BQ * d.n_slots < (1ll << 31), "lookup: too many units");
        const long long n_units = BQ * d.n_slots;
        static int mode = -1, occ = 0;
        if (mode < 0) {
            const char* e = getenv("BFLOW_LK_FLAT");
            mode = (e != nullptr && e[0] == '0') ? 0 : 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bflow::corr_lookup_tiled_kernel<true>, bflow::LK2_WARPS * 32, 0);
            if (occ <= 0) occ = 6;
        }
        const int sms = bflow::num_sms();
        cudaError_t le;
        static const bool lk_pdl = [] { const char* e = getenv("BFLOW_LOOKUP_PDL"); return !(e != nullptr && e[0] == '0'); }();
        unsigned long long* tls = bflow::timeline_next_slot("corr_lookup");
        bflow::LookupDivs dv;
        dv.S = bflow::make_fastdiv((unsigned)d.n_slots);
        dv.Q = bflow::make_fastdiv((unsigned)(d.h * d.w));
        dv.w = bflow::make_fastdiv((unsigned)d.w);
        BFLOW_REQUIRE(d.params == nullptr || d.degree != 2 || (d.params_ld & 3) != 0 || (reinterpret_cast<uintptr_t>(d.params) & 15) == 0,
                      "lookup: degree-2 Bezier parameters with a row stride that is a multiple of 4 floats must be 16-byte aligned");
        const long long resident_warps = (long long)sms * occ * bflow::LK2_WARPS;
        if (mode == 1 && n_units <= 8 * resident_warps) {
            // one wave of CTAs, units dealt round-robin (at most 8 units per warp; beyond that the per-pixel walk below is as balanced)
            long long g = bflow::ceil_div_ll(n_units, bflow::LK2_WARPS);
            if (g > (long long)sms * occ) g = (long long)sms * occ;
            le = bflow::launch_pdl_if(lk_pdl, bflow::corr_lookup_tiled_kernel<true>, dim3((unsigned)g), dim3(bflow::LK2_WARPS * 32), 0, (cudaStream_t)stream, d, dv, n_units, 0, tls);
        } else {
            long long g = bflow::ceil_div_ll(BQ, bflow::LK2_WARPS);
            const long long cap = (long long)sms * 8 * 8;
            if (g > cap) g = cap;
            // few query pixels: split the slots over blockIdx.y so that every SM still holds a full set of warps
            long long groups = bflow::ceil_div_ll((long long)sms * 48, BQ);
            if (groups < 1) groups = 1;
            if (groups > d.n_slots) groups = d.n_slots;
            const int spg = (int)bflow::ceil_div_ll(d.n_slots, groups);
            dim3 grid2((unsigned)g, (unsigned)bflow::ceil_div(d.n_slots, spg));
            le = bflow::launch_pdl_if(lk_pdl, bflow::corr_lookup_tiled_kernel<false>, grid2, dim3(bflow::LK2_WARPS * 32), 0, (cudaStream_t)stream, d, dv, n_units, spg, tls);
        }
        if (le != cudaSuccess) {
            bflow::set_error(cudaGetErrorString(le));
            return BFLOW_ERR_CUDA;
        }
        return bflow::check_launch("bflow_corr_lookup(tiled)");
    }
    dim3 grid((unsigned)bflow::ceil_div_ll(BQ, bflow::LK_QPB), (unsigned)d.n_slots);
    bflow::corr_lookup_kernel<<<grid, bflow::LK_WARPS * 32, 0, (cudaStream_t)stream>>>(d);
    return bflow::check_launch("bflow_corr_lookup");
}
