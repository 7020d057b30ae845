// Fused 7x7 / stride 2 / pad 3 stem (extractor.py:112) for thin inputs (K = 49 * Cin <= 256: event windows of 5 bins, RGB images).
//
// The patch matrix never leaves the SM.  Per 8 x 16 tile of output pixels:
//   builder warps (8)   fetch the (2*8+5) x (2*16+5+3) x Cin fp32 input footprint with 16-byte cp.async (out-of-image = zero fill; a 3-D
//                       fp32 TMA box faulted on B200) into one of TWO footprint buffers, and expand it k-block by k-block into the
//                       SWIZZLE_128B A tile (4 k-blocks, hi | lo).  Cin is a template parameter, so every one of the 256 footprint offsets
//                       of a row is an immediate of its LDS: no offset table, no per-element branch (the first version of this kernel spent
//                       96 of its 194 us in this expansion);
//   warp 1              issues the MMAs of k-block kb as soon as that block is built (per-k-block full / empty barriers: the expansion
//                       of tile i+1 overwrites block kb while the MMAs of tile i still read blocks kb+1..3) against the resident weights;
//   epilogue warps (8)  drain the other TMEM accumulator; an 8 x 8 in-warp transpose turns TMEM's row-per-thread layout into whole 128-byte
//                       rows per store instruction (row-per-thread stores touched 32 lines per instruction: 33 us of 109), and leaves the
//                       InstanceNorm column sums two shuffle steps away; bias / ReLU, fp32 and split-fp16 stores.
// Before (one set of 8 worker warps doing expansion and epilogue in turn, one A buffer handed over whole): 9.5 us per tile, 194 us per
// 5 x 480 x 640 window batch; measured split of that: expansion 96 us, epilogue 59 us, MMA + hand-overs 44 us, all serial.
// Now 87 us (MMAs alone: 36 us; what remains is shared-memory traffic -- 128 KB of A written and 224 KB of operands read per tile --
// shared by the three roles).
#include <cuda.h>
#include <stdlib.h>
#include "tc3_common.cuh"

namespace bflow {

constexpr int ST_BN = 64;
constexpr int ST_PW = 40, ST_PH = 21;               // footprint: x in [32 tx - 4, 32 tx + 36) (10 aligned float4), y in [16 ty - 3, 16 ty + 18)
constexpr int ST_MAXC = 5;
constexpr int ST_XSHIFT = 1;                        // the footprint starts one pixel left of the first tap so that its rows are 16-byte aligned
constexpr int ST_NKB = 4;                           // K padded to 256
constexpr int ST_A_BYTES = ST_NKB * 2 * T3_A_BYTES; // 131072
constexpr int ST_B_BYTES = ST_NKB * 2 * ST_BN * 128;// 65536
constexpr int ST_PATCH_BYTES = ST_MAXC * ST_PH * ST_PW * 4;   // 16800
constexpr int ST_SMALL_BYTES = 1024;                // barriers, bias, statistics
constexpr int ST_THREADS = 64 + 256 + 256;          // warp 0: weights, warp 1: MMA, warps 2-9: builders, warps 10-17: epilogue
constexpr int ST_SMEM = ST_A_BYTES + ST_B_BYTES + 2 * ST_PATCH_BYTES + ST_SMALL_BYTES + 1008;
static_assert(ST_SMEM <= 232448, "stem: shared memory budget");

struct StemParams {
    int tiles_x, tiles_y, n_tiles, c_total;
    int n_win, ns, c_off[8];      // image n = window * ns + sample: channels [c_off[window], +cin) of input sample `sample`
    int f16;
    float acc_scale, in_scale, in_shift;
    unsigned long long* tl;
};

// footprint offset (relative to the pixel's tap (0, 0) of channel 0) of K index k = (c*7 + kh)*7 + kw (the natural OIHW order of the
// weights: consecutive k walk along a footprint row, so two neighbouring taps are one 8-byte shared-memory load)
__host__ __device__ constexpr int st_koff(int k) { return ((k / 49) * ST_PH + (k % 49) / 7) * ST_PW + k % 7 + ST_XSHIFT; }

// four 16-byte chunks (j = 4 Q .. 4 Q + 3) of k-block KB of one A row: 32 footprint values, split, swizzled stores.  px (the pixel's
// tap (0, 0) of channel 0) is 8-byte aligned and st_koff is even for odd kw: taps (1,2), (3,4), (5,6) of a row are loaded as pairs.
template <int CIN, bool AFFINE, int KB, int Q>
__device__ __forceinline__ void st_build(const float* __restrict__ px, uint8_t* a_row, const int row7, const bool f16, const unsigned rowmask,
                                         const unsigned colmask, const float in_scale, const float in_shift) {
    constexpr int K = 49 * CIN, k0 = KB * 64 + Q * 32;
    float xv[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        const int k = k0 + e, kw = k % 7;
        if (k >= K) {
            xv[e] = 0.f;
        } else if (kw >= 2 && (kw & 1) == 0 && e >= 1) {
            // second half of the pair loaded at e - 1
        } else if ((kw & 1) && e + 1 < 32 && k + 1 < K) {
            const float2 t = *reinterpret_cast<const float2*>(px + st_koff(k));
            xv[e] = t.x;
            xv[e + 1] = t.y;
        } else {
            xv[e] = px[st_koff(k)];
        }
    }
    if (AFFINE) {
        // padding must stay zero AFTER the affine input map (raft.py:134 normalises before the conv pads)
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int k = k0 + e, kh = (k % 49) / 7, kw = k % 7;
            if (k < K) xv[e] = (((rowmask >> kh) & (colmask >> kw)) & 1u) ? fmaf(xv[e], in_scale, in_shift) : 0.f;
        }
    }
#pragma unroll
    for (int cj = 0; cj < 4; ++cj) {
        const int j = Q * 4 + cj;
        const float* x8 = xv + cj * 8;
        uint4 h4, l4;
        split2(x8[0], x8[1], h4.x, l4.x);
        split2(x8[2], x8[3], h4.y, l4.y);
        split2(x8[4], x8[5], h4.z, l4.z);
        split2(x8[6], x8[7], h4.w, l4.w);
        uint8_t* dst = a_row + KB * (2 * T3_A_BYTES) + ((j ^ row7) << 4);
        *reinterpret_cast<uint4*>(dst) = h4;
        if (!f16) *reinterpret_cast<uint4*>(dst + T3_A_BYTES) = l4;
    }
}

template <int CIN, bool AFFINE>
__global__ void __launch_bounds__(ST_THREADS, 1)
conv_stem7_kernel(const bflow_conv_desc d, const uint8_t* __restrict__ wtc, const StemParams p, int* err) {
    constexpr int ACC_COLS = 2 * ST_BN;
    constexpr int TMEM_COLS = 2 * ACC_COLS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (t3_smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base;                       // [kb][hi | lo][128][128 B]
    const uint32_t b_base = a_base + ST_A_BYTES;             // resident weights
    const uint32_t patch_u32 = b_base + ST_B_BYTES;          // 2 x [cin][21][40] fp32
    const uint32_t bars = patch_u32 + 2 * ST_PATCH_BYTES;
    auto afull_bar = [&](int kb) { return bars + 8u * (uint32_t)kb; };
    auto aempty_bar = [&](int kb) { return bars + 32u + 8u * (uint32_t)kb; };
    const uint32_t wbar = bars + 64u;
    auto tfull_bar = [&](uint32_t a) { return bars + 72u + 8u * a; };
    auto tempty_bar = [&](uint32_t a) { return bars + 88u + 8u * a; };
    const uint32_t tmem_slot = bars + 104u;
    uint8_t* gen = smem_raw + (smem_base - t3_smem_u32(smem_raw));
    const float* patch = reinterpret_cast<const float*>(gen + ST_A_BYTES + ST_B_BYTES);
    float* s_bias = reinterpret_cast<float*>(gen + ST_A_BYTES + ST_B_BYTES + 2 * ST_PATCH_BYTES + 128);     // [64]
    float* s_stat = s_bias + ST_BN;                                                                          // [2][64]
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;      // warp-uniform role index (see conv_tc3_kernel)
    tl_begin(p.tl);
    if (tid == 0) {
        for (int kb = 0; kb < ST_NKB; ++kb) {
            t3_mbar_init(afull_bar(kb), 256);      // every builder thread
            t3_mbar_init(aempty_bar(kb), 1);       // one tcgen05.commit
        }
        t3_mbar_init(wbar, 1);
        for (uint32_t a = 0; a < 2; ++a) {
            t3_mbar_init(tfull_bar(a), 1);         // one tcgen05.commit
            t3_mbar_init(tempty_bar(a), 8);        // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    t3_fence_before();
    __syncthreads();
    t3_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        if (lane == 0) {
            t3_mbar_arrive_expect_tx(wbar, ST_B_BYTES);
            for (int kb = 0; kb < ST_NKB; ++kb) t3_bulk_g2s(b_base + (uint32_t)kb * (2 * ST_BN * 128), wtc + (size_t)kb * (2 * ST_BN * 128), 2 * ST_BN * 128, wbar);
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer ------------------------------------------------
        const uint32_t idesc = (1u << 4) | ((uint32_t)(ST_BN >> 3) << 17) | ((uint32_t)(T3_BM >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * ST_BN) >> 3) << 17) | ((uint32_t)(T3_BM >> 4) << 24);
        t3_mbar_wait(wbar, 0u, err);
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++lt) {
            const uint32_t acc = lt & 1u, aph = (lt >> 1) & 1u;
            t3_mbar_wait(tempty_bar(acc), aph ^ 1u, err);          // the epilogue has drained this accumulator
            const uint32_t tacc = tmem_base + acc * ACC_COLS;
#pragma unroll
            for (int kb = 0; kb < ST_NKB; ++kb) {
                t3_mbar_wait(afull_bar(kb), lt & 1u, err);
                t3_fence_after();
                if (t3_elect_one()) {
                    const uint32_t a_hi = a_base + (uint32_t)kb * (2 * T3_A_BYTES), a_lo = a_hi + T3_A_BYTES;
                    const uint32_t b = b_base + (uint32_t)kb * (2 * ST_BN * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t ko = (uint32_t)k * 32u;
                        const uint64_t dbh = t3_umma_desc(b + ko);
                        if (p.f16) {
                            t3_umma(tacc, t3_umma_desc(a_hi + ko), dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        } else {
                            t3_umma(tacc, t3_umma_desc(a_hi + ko), dbh, idesc2, (kb > 0 || k > 0) ? 1u : 0u);     // hi*hi | hi*lo
                            t3_umma(tacc, t3_umma_desc(a_lo + ko), dbh, idesc, 1u);                               // lo*hi
                        }
                    }
                    t3_commit(aempty_bar(kb));
                    if (kb == ST_NKB - 1) t3_commit(tfull_bar(acc));
                }
                __syncwarp();
            }
        }
        t3_fence_before();
    } else if (warp < 10) {
        // ------------------------------------------------ builders ------------------------------------------------
        const int btid = tid - 64;
        // footprint loader: 16-byte cp.async with zero fill outside the image (W % 4 == 0, so a float4 is entirely in or out)
        auto load_patch = [&](int tile, uint32_t buf) {
            const int n = tile / tiles_per_img, r = tile - n * tiles_per_img;
            const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
            const int gx0 = tx * 32 - 4, gy0 = ty * 16 - 3;
            const int win = n / p.ns, nl = n - win * p.ns;
            const float* src0 = d.x0 + ((size_t)nl * p.c_total + p.c_off[win]) * d.H * d.W;
            const uint32_t dst0 = patch_u32 + buf * ST_PATCH_BYTES;
            constexpr int total = CIN * ST_PH * (ST_PW / 4);
            for (int i = btid; i < total; i += 256) {
                const int x4 = i % (ST_PW / 4);
                const int t = i / (ST_PW / 4);
                const int yy = t % ST_PH, c = t / ST_PH;
                const int gx = gx0 + x4 * 4, gy = gy0 + yy;
                const bool ok = gx >= 0 && gx < d.W && gy >= 0 && gy < d.H;
                const float* src = src0 + ((size_t)c * d.H + (ok ? gy : 0)) * d.W + (ok ? gx : 0);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (uint32_t)(i * 16)), "l"(src), "r"(ok ? 16 : 0) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // thread -> tile row (btid & 127: pixel (row >> 4, row & 15)) and one half (Q) of the eight 16-byte chunks of every k-block
        const int brow = btid & 127, q = btid >> 7;
        const int bpy = brow >> 4, bpx = brow & 15;
        const int pbase = (2 * bpy) * ST_PW + 2 * bpx;                       // footprint offset of tap (0, 0), channel 0
        uint8_t* a_row = gen + brow * 128;
        const int row7 = brow & 7;
        const bool f16 = p.f16 != 0;
        uint32_t lt = 0;
        if ((int)blockIdx.x < p.n_tiles) load_patch(blockIdx.x, 0u);
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++lt) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");        // this tile's footprint has landed; every builder is done with the other buffer
            if (tile + (int)gridDim.x < p.n_tiles) load_patch(tile + gridDim.x, (lt + 1u) & 1u);
            const float* px = patch + (lt & 1u) * (ST_PATCH_BYTES / 4) + pbase;
            unsigned rowmask = 0x7fu, colmask = 0x7fu;
            if (AFFINE) {
                const int n = tile / tiles_per_img, r = tile - n * tiles_per_img;
                const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
                const int iy0 = 2 * (ty * 8 + bpy) - 3, ix0 = 2 * (tx * 16 + bpx) - 3;
                rowmask = colmask = 0u;
#pragma unroll
                for (int t = 0; t < 7; ++t) {
                    rowmask |= (iy0 + t >= 0 && iy0 + t < d.H) ? (1u << t) : 0u;
                    colmask |= (ix0 + t >= 0 && ix0 + t < d.W) ? (1u << t) : 0u;
                }
            }
            const uint32_t eph = (lt & 1u) ^ 1u;          // the MMAs of the previous tile have finished reading the k-block
#define ST_BUILD_KB(KB)                                                                                                       \
            t3_mbar_wait(aempty_bar(KB), eph, err);                                                                           \
            if (q == 0) st_build<CIN, AFFINE, KB, 0>(px, a_row, row7, f16, rowmask, colmask, p.in_scale, p.in_shift);         \
            else st_build<CIN, AFFINE, KB, 1>(px, a_row, row7, f16, rowmask, colmask, p.in_scale, p.in_shift);                \
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      /* generic-proxy writes of A -> visible to the tensor core */ \
            t3_mbar_arrive(afull_bar(KB));
            ST_BUILD_KB(0)
            ST_BUILD_KB(1)
            ST_BUILD_KB(2)
            ST_BUILD_KB(3)
#undef ST_BUILD_KB
        }
    } else {
        // ------------------------------------------------ epilogue ------------------------------------------------
        const int quad = warp & 3, chalf = (warp - 10) >> 2, etid = tid - 320;
        for (int j = etid; j < ST_BN; j += 256) {
            s_bias[j] = d.bias != nullptr ? __ldg(d.bias + j) : 0.f;
            s_stat[j] = 0.f;
            s_stat[ST_BN + j] = 0.f;
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const float lo1 = d.act1 == BFLOW_ACT_RELU ? 0.f : -INFINITY;
        const float post = d.scale;
        int cur_img = -1;
        auto flush_stats = [&](int img) {
            asm volatile("bar.sync 2, 256;" ::: "memory");
            for (int j = etid; j < ST_BN; j += 256) {
                if (img >= 0) {
                    atomicAdd(d.stats + ((size_t)img * d.Cout + j) * 2, (double)s_stat[j]);
                    atomicAdd(d.stats + ((size_t)img * d.Cout + j) * 2 + 1, (double)s_stat[ST_BN + j]);
                }
                s_stat[j] = 0.f;
                s_stat[ST_BN + j] = 0.f;
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
        };
        const int nb0 = chalf * 32;
        const int b4 = (lane & 7) * 4;                 // after the in-warp transpose a lane owns columns nb0 + b4 .. + 3
        const float bias4[4] = {s_bias[nb0 + b4], s_bias[nb0 + b4 + 1], s_bias[nb0 + b4 + 2], s_bias[nb0 + b4 + 3]};
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++lt) {
            // TMEM: thread = pixel row of the tile (quad*32 + lane), half of the 64 channels
            const int n = tile / tiles_per_img, r = tile - n * tiles_per_img;
            const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
            if (d.stats != nullptr && n != cur_img) {
                flush_stats(cur_img);
                cur_img = n;
            }
            const uint32_t acc = lt & 1u;
            t3_mbar_wait(tfull_bar(acc), (lt >> 1) & 1u, err);
            t3_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * ACC_COLS + (uint32_t)nb0;
            float v[32];
            {
                float u[32];
                t3_tmem_ld16_nowait(taddr, v);
                t3_tmem_ld16_nowait(taddr + 16u, v + 16);
                if (!p.f16) {
                    t3_tmem_ld16_nowait(taddr + (uint32_t)ST_BN, u);
                    t3_tmem_ld16_nowait(taddr + (uint32_t)ST_BN + 16u, u + 16);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) u[c] = 0.f;
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                t3_fence_before();
                __syncwarp();
                if (lane == 0) t3_mbar_arrive(tempty_bar(acc));        // the MMAs of tile lt + 2 may overwrite this accumulator
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] += u[c];
            }
            // 8 x 8 transpose inside every group of 8 lanes (3 butterfly steps over the 16-byte chunks): afterwards lane (g, b) holds columns
            // nb0 + 4b .. 4b + 3 of the tile rows quad*32 + 8g + c, c = 0..7, so that one store instruction writes whole 128-byte rows
            // (row-per-thread stores touched 32 lines per instruction; measured 33 of the kernel's 109 us)
#pragma unroll
            for (int bit = 4; bit >= 1; bit >>= 1) {
                const bool upper = (lane & bit) != 0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c & bit) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float send = upper ? v[4 * c + j] : v[4 * (c | bit) + j];
                        const float recv = __shfl_xor_sync(0xffffffffu, send, bit);
                        if (upper) v[4 * c + j] = recv;
                        else v[4 * (c | bit) + j] = recv;
                    }
                }
            }
            const int g = lane >> 3;
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) v[4 * c + j] = post * fmaf(v[4 * c + j], p.acc_scale, bias4[j]);
            const int yy = ty * 8 + quad * 2 + (g >> 1), x0 = tx * 16 + (g & 1) * 8;
            const size_t m0 = ((size_t)n * d.Ho + yy) * d.Wo + x0;
            const int nvalid = yy < d.Ho ? min(8, d.Wo - x0) : 0;             // rows c < nvalid are inside the image
            if (d.stats != nullptr) {
                float s[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c < nvalid) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            s[j] += v[4 * c + j];
                            sq[j] = fmaf(v[4 * c + j], v[4 * c + j], sq[j]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 8);
                    sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], 8);
                    s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16);
                    sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], 16);
                }
                if (g == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        atomicAdd(s_stat + nb0 + b4 + j, s[j]);
                        atomicAdd(s_stat + ST_BN + nb0 + b4 + j, sq[j]);
                    }
                }
            }
            {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c < nvalid) {
                        const float o0 = fmaxf(v[4 * c], lo1), o1 = fmaxf(v[4 * c + 1], lo1), o2 = fmaxf(v[4 * c + 2], lo1), o3 = fmaxf(v[4 * c + 3], lo1);
                        if (d.y != nullptr) *reinterpret_cast<float4*>(d.y + (m0 + c) * d.ldy + nb0 + b4) = make_float4(o0, o1, o2, o3);
                        if (d.y16_hi != nullptr) {
                            uint2 h2, l2;
                            split2(o0, o1, h2.x, l2.x);
                            split2(o2, o3, h2.y, l2.y);
                            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(d.y16_hi) + (m0 + c) * d.ldy16 + nb0 + b4) = h2;
                            if (!p.f16) *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(d.y16_lo) + (m0 + c) * d.ldy16 + nb0 + b4) = l2;
                        }
                    }
                }
            }
        }
        if (d.stats != nullptr) flush_stats(cur_img);
    }
    __syncthreads();
    tl_end(p.tl);
    if (warp == 1) {
        t3_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

template <int CIN, bool AFFINE>
static cudaError_t stem7_launch(int grid, cudaStream_t stream, const bflow_conv_desc& d, const uint8_t* wtc, const StemParams& p, int* err) {
    static PerDeviceFlag configured;
    if (!configured.get()) {
        cudaError_t e = cudaFuncSetAttribute(conv_stem7_kernel<CIN, AFFINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM);
        if (e != cudaSuccess) return e;
        configured.set();
    }
    conv_stem7_kernel<CIN, AFFINE><<<grid, ST_THREADS, ST_SMEM, stream>>>(d, wtc, p, err);
    return cudaSuccess;
}

}  // namespace bflow

// Fused stem: 7x7 / stride 2 / pad 3, channels [c_off, c_off + cin) of an fp32 NCHW input (49 * cin <= 256) -> 64 channels, input mapped
// x -> in_scale * x + in_shift first (raft.py:134).  d carries the output side (N, H, W, Ho, Wo, Cout = 64, bias, act1, y / y16, stats);
// x0: the fp32 NCHW input; w_tc: tc3 weight image (bn 64) of the [64][256] matrix with K = (c*7+kh)*7+kw (OIHW order).
extern "C" int bflow_conv2d_stem7(const bflow_conv_desc* dp, const void* w_tc, int c_total, const int* c_offs, int n_windows, float in_scale,
                                  float in_shift, float acc_scale, int* err, void* stream) {
    BFLOW_CHECK_DESC(dp, bflow_conv_desc, "conv_stem7");
    BFLOW_REQUIRE(w_tc != nullptr, "conv_stem7: null argument");
    BFLOW_REQUIRE(dp->x0 != nullptr && bflow::aligned16(dp->x0) && dp->W % 4 == 0, "conv_stem7: x0 = 16-byte aligned fp32 NCHW input, W % 4 == 0");
    const bflow_conv_desc& d = *dp;
    BFLOW_REQUIRE(d.c0 > 0 && d.c0 <= bflow::ST_MAXC && 49 * d.c0 <= 256 && d.c1 == 0 && d.Cout == 64, "conv_stem7: cin <= 5, Cout == 64");
    BFLOW_REQUIRE(d.KH == 7 && d.KW == 7 && d.stride == 2 && d.pad_h == 3 && d.pad_w == 3, "conv_stem7: 7x7 / 2 / pad 3 only");
    BFLOW_REQUIRE(d.N > 0 && d.H > 0 && d.W > 0 && d.Ho == (d.H - 1) / 2 + 1 && d.Wo == (d.W - 1) / 2 + 1, "conv_stem7: Ho/Wo mismatch");
    BFLOW_REQUIRE(c_offs != nullptr && n_windows >= 1 && n_windows <= 8 && d.N % n_windows == 0, "conv_stem7: 1..8 channel windows, N = windows * samples");
    for (int i = 0; i < n_windows; ++i) BFLOW_REQUIRE(c_offs[i] >= 0 && c_offs[i] + d.c0 <= c_total, "conv_stem7: channel window");
    BFLOW_REQUIRE(d.epi == BFLOW_EPI_STD && d.res == nullptr && d.res16_hi == nullptr && d.act1 <= BFLOW_ACT_RELU && d.act2 == BFLOW_ACT_NONE, "conv_stem7: plain epilogue (none / relu)");
    BFLOW_REQUIRE((d.y == nullptr || (d.ldy >= 64 && d.ldy % 4 == 0 && bflow::aligned16(d.y))), "conv_stem7: fp32 output alignment");
    BFLOW_REQUIRE(d.y16_hi == nullptr || (d.y16_lo != nullptr && d.ldy16 % 8 == 0 && bflow::aligned16(d.y16_hi) && bflow::aligned16(d.y16_lo)), "conv_stem7: split output alignment");
    BFLOW_REQUIRE(d.stats == nullptr || d.act1 == BFLOW_ACT_NONE, "conv_stem7: fused statistics need the plain epilogue");
    if (const char* msg = bflow::check_epilogue(d)) { bflow::set_error(msg); return BFLOW_ERR_INVALID; }
    bflow::StemParams p;
    p.tiles_x = (d.Wo + 15) / 16;
    p.tiles_y = (d.Ho + 7) / 8;
    const long long nt = (long long)d.N * p.tiles_x * p.tiles_y;
    BFLOW_REQUIRE(nt < (1ll << 31), "conv_stem7: too large");
    p.n_tiles = (int)nt;
    p.c_total = c_total;
    p.n_win = n_windows;
    p.ns = d.N / n_windows;
    for (int i = 0; i < 8; ++i) p.c_off[i] = i < n_windows ? c_offs[i] : 0;
    p.acc_scale = acc_scale;
    p.in_scale = in_scale;
    p.in_shift = in_shift;
    BFLOW_REQUIRE(d.precision == BFLOW_PREC_SPLIT3 || d.precision == BFLOW_PREC_F16, "conv_stem7: unknown precision");
    p.f16 = d.precision == BFLOW_PREC_F16 ? 1 : 0;
    p.tl = bflow::timeline_next_slot("stem7");
    const int grid = p.n_tiles < bflow::grid_cap(d) ? p.n_tiles : bflow::grid_cap(d);
    const bool affine = in_scale != 1.f || in_shift != 0.f;
    const uint8_t* w8 = reinterpret_cast<const uint8_t*>(w_tc);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t le = cudaSuccess;
#define BFLOW_STEM_CASE(C)                                                                     \
    case C:                                                                                    \
        le = affine ? bflow::stem7_launch<C, true>(grid, st, d, w8, p, err) : bflow::stem7_launch<C, false>(grid, st, d, w8, p, err); \
        break;
    switch (d.c0) {
        BFLOW_STEM_CASE(1)
        BFLOW_STEM_CASE(2)
        BFLOW_STEM_CASE(3)
        BFLOW_STEM_CASE(4)
        default:
        BFLOW_STEM_CASE(5)
    }
#undef BFLOW_STEM_CASE
    if (le != cudaSuccess) {
        bflow::set_error(cudaGetErrorString(le));
        return BFLOW_ERR_CUDA;
    }
    return bflow::check_launch("bflow_conv2d_stem7");
}
