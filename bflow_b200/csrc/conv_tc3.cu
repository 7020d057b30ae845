// TMA-fed, persistent tensor-core convolution for sm_100a (bflow_conv2d_nhwc_tc3).
//
// Same operator as conv_tc.cu / conv_simt.cu, but the A operand never touches registers: activations are stored as
// split-fp16 planes (x = hi + lo) by whoever produced them, and the TMA unit gathers the implicit-GEMM tile
// [128 output pixels] x [64 channels of one filter tap] straight into the SWIZZLE_128B shared-memory layout with
// cp.async.bulk.tensor.4d...im2col (zero fill for the padding halo, for channels beyond Cin and for rows beyond M).
//
//   grid      one CTA per SM, looping over (m_tile, n_tile) output tiles (n fastest, so co-running CTAs share A in L2)
//   warp 0    TMA producer: per k-block one arrive.expect_tx + two im2col loads (hi, lo) + one bulk copy of the weight tile
//   warp 1    MMA issuer: 4 k-steps x 3 tcgen05.mma (hi*hi, hi*lo, lo*hi) per k-block into TMEM accumulator `acc`,
//             tcgen05.commit -> empty[stage]; after the last k-block tcgen05.commit -> tmem_full[acc]
//   warps 2-5 epilogue: tcgen05.ld of accumulator `acc` (two accumulators: the epilogue of tile i overlaps the MMAs of
//             tile i+1), bias / scale / activation / residual / GRU gates, fp32 and/or split-fp16 stores
//
// K order: k-block = (tap, 64-channel block); the two concatenated sources are two tensor-map pairs.
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include <type_traits>
#include "tc3_common.cuh"

namespace bflow {


struct T3Params {
    int M, n_mtiles, n_ntiles, ntaps, ncb0, ncb1, nkb;
    float acc_scale;
    long long* trace;
    unsigned long long* tl;
    unsigned long long* cta;   // development (bflow_tc3_cta_trace): [gridDim.x][16] globaltimer stamps of ONE chosen launch
    // slab mode (stride-1 3x3 / 1x5 / 5x1): an output tile is an 8 x 16 pixel patch and one halo slab per slow filter index serves all taps
    // along the other axis (see conv_slab64_kernel).  1: slab rows = (y, x) with x fastest (taps along y), 2: rows = (x, y) (taps along x).
    int slab, tiles_x, tiles_y, n_slabs, tps;
    int f16;      // BFLOW_PREC_F16: one MMA per k-step on the hi planes (lo planes / the lo half of the weight tiles are not loaded)
    FastDiv fd_ntiles, fd_wo, fd_ho;      // tile -> (m_tile, n_tile), first output pixel -> (n, oh, ow) on the producer's critical path
    int staged;   // every CTA owns exactly one tile: the epilogue goes through shared memory (coalesced, batched global accesses)
    int dbg;      // development switch (bflow_tc3_debug): 1 = no TMA loads (the producer only arrives: MMA + epilogue path alone)
    // 64-channel block index (within a tap) whose channels 32..63 lie beyond the source's channel count (c % 64 in 1..32), or -1: the TMA
    // unit zero-fills them and the weight image holds zeros there, so the MMA warp issues only the first two of the block's four k-steps
    // (the main loops are bound by the number of tcgen05.mma instructions: 96-channel layers otherwise spend a quarter of them on zeros)
    int half0, half1;
    // rows of one half (hi or lo) of a weight tile = N of the MMAs = TMEM column of the accumulator's second half.  Normally BN; a launch with a
    // single column tile and Cout < BN (the 96-channel encoder layers under BN = 128) packs and multiplies only bn_eff = Cout rows: N 192 + 96
    // instead of 256 + 128 per k-step, and a quarter less weight traffic
    int bn_eff;
};

template <int BN, int STAGES, bool F16>
__global__ void __launch_bounds__(T3_THREADS, 1)
conv_tc3_kernel(const __grid_constant__ CUtensorMap map0h, const __grid_constant__ CUtensorMap map0l, const __grid_constant__ CUtensorMap map1h,
                const __grid_constant__ CUtensorMap map1l, const __grid_constant__ CUtensorMap omap_hi, const __grid_constant__ CUtensorMap omap_lo,
                const __grid_constant__ CUtensorMap omap32, const __grid_constant__ CUtensorMap rmap, const __grid_constant__ CUtensorMap amap,
                const bflow_conv_desc d, const uint8_t* __restrict__ wtc, const T3Params p, int* err) {
    constexpr int B_BYTES = BN * 128;
    constexpr int STAGE_BYTES = 2 * T3_A_BYTES + 2 * B_BYTES;
    constexpr uint32_t TX_BYTES = 2 * T3_A_BYTES + 2 * B_BYTES;
    // BN <= 128: the weight tile's hi and lo halves are contiguous in shared memory, so A_hi x [B_hi; B_lo] is ONE tcgen05.mma with
    // N = 2 BN (columns [0,BN) = hi*hi, [BN,2BN) = hi*lo) and A_lo x B_hi a second one into [0,BN): 2 instructions per k-step
    // instead of 3 (the single issuing thread, not the tensor pipe, is the limiter for small N).  The epilogue adds the two halves.
    constexpr bool STACK = BN <= 128;
    constexpr int ACC_COLS = STACK ? 2 * BN : BN;
    // two accumulators; 512 columns in every instantiation: single-tile GRU launches (staged 6) park their per-element operands behind
    // accumulator 0 (BN = 64: 128 + 3 x 64 columns)
    constexpr int TMEM_COLS = 512;
    // data area: the pipeline stages; for BN <= 128 it is stretched to 224 KB so that the bulk-copy epilogue of a single-tile CTA (below) can
    // park the fp32 tile and up to three operand tiles in it once the pipeline has drained
    constexpr int AREA_BYTES = t3_area_bytes(BN, STAGES);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (t3_smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + AREA_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bars + 32u + 8u * (uint32_t)s; };
    auto tfull_bar = [&](int a) { return bars + 64u + 8u * (uint32_t)a; };
    auto tempty_bar = [&](int a) { return bars + 80u + 8u * (uint32_t)a; };
    const uint32_t tmem_slot = bars + 96u;
    const uint32_t ebar = bars + 104u;          // bulk-copy epilogue: operand tiles have landed
    constexpr int SL_NSA = 3, SL_NSB = BN <= 64 ? 4 : 3;                 // slab mode: halo-slab ring and weight-tile ring
    constexpr int SL_SLAB_MAX = 2 * 8 * 20 * 128;                        // hi + lo planes of an 8 x (16 + 4) slab
    const uint32_t sl_bars = bars + 128u + 3u * BN * 4u;                 // [afull 3][aempty 3]
    auto afull_bar = [&](int s_) { return sl_bars + 8u * (uint32_t)s_; };
    auto aempty_bar = [&](int s_) { return sl_bars + 24u + 8u * (uint32_t)s_; };
    const uint32_t sl_bring = smem_base + SL_NSA * SL_SLAB_MAX;          // weight-tile ring behind the slab ring

    const int tid = threadIdx.x;
    // warp index through a shuffle: nvcc then treats it (and every branch on it) as warp-uniform, which is what lets the role loops below keep
    // their operands in uniform registers
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const int n_tiles = p.n_mtiles * p.n_ntiles;

    if (tid == 0) T3_TRACE(5, 255);              // CTA start
    if (tid == 0) T3_CTA(0);
    tl_begin(p.tl);
    // Prologue.  The producer warp only needs the mbarriers: it initialises them, signals named barrier 2 WITHOUT waiting and goes
    // straight to its first TMA issue, while warp 1 allocates TMEM (~0.3 us) and everybody else waits on barrier 2 for both.  Measured
    // before: first TMA issue 1.25 us after the CTA start (0.5 us of common prologue + 0.7 us of integer divisions in the producer).
    uint32_t tmem_base = 0;
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map0h)) : "memory");
            if (!F16) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map0l)) : "memory");
            if (p.ncb1 > 0) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map1h)) : "memory");
                if (!F16) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map1l)) : "memory");
            }
            for (int s = 0; s < 4; ++s) {
                t3_mbar_init(full_bar(s), 1);      // the producer's arrive.expect_tx (+ TMA bytes)
                t3_mbar_init(empty_bar(s), 1);     // one tcgen05.commit
            }
            for (int s = 0; s < SL_NSA; ++s) {
                t3_mbar_init(afull_bar(s), 1);
                t3_mbar_init(aempty_bar(s), 1);
            }
            t3_mbar_init(ebar, 1);
            for (int a = 0; a < 2; ++a) {
                t3_mbar_init(tfull_bar(a), 1);     // one tcgen05.commit
                t3_mbar_init(tempty_bar(a), T3_EPI_WARPS);    // one arrive per epilogue warp
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __threadfence_block();
        __syncwarp();
        asm volatile("bar.arrive 2, %0;" ::"n"(T3_THREADS) : "memory");
    } else {
        if (warp == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        if (tid == 64 && p.staged >= 4) {        // tensor-map epilogues: their descriptors (a cold descriptor costs ~1 us at first use)
            if (d.y16_hi != nullptr || d.aux1_16_hi != nullptr) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&omap_hi)) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&omap_lo)) : "memory");
            }
            if (d.y != nullptr) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&omap32)) : "memory");
            if (p.staged == 5) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&rmap)) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&amap)) : "memory");
            }
        }
        t3_fence_before();
        asm volatile("bar.sync 2, %0;" ::"n"(T3_THREADS) : "memory");
        t3_fence_after();
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    }
    // PDL: everything above touched only shared memory / TMEM; let the next kernel's CTAs set themselves up on idle SMs, then
    // wait for the previous kernel's results before the first global access
    pdl_trigger();
    pdl_wait();
    if (tid == 0) T3_TRACE(3, 255);              // prologue done
    if (tid == 0) T3_CTA(1);

    if (warp == 0) {
        // ------------------------------------------------ TMA producer ------------------------------------------------
        if (lane == 0 && p.slab) {
            // slab mode: per 64-channel block and slow filter index one halo slab (hi, lo), then the weight tiles of its taps
            uint32_t ia = 0, ib = 0;
            const int tiles_per_img = p.tiles_x * p.tiles_y;
            const uint32_t slab_plane = (uint32_t)(8 * (16 + p.tps - 1) * 128);
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m_tile = tile / p.n_ntiles, n_tile = tile - m_tile * p.n_ntiles;
                const int n = m_tile / tiles_per_img, r = m_tile - n * tiles_per_img;
                const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
                const int x0 = tx * (p.slab == 1 ? 8 : 16) - d.pad_w, y0 = ty * (p.slab == 1 ? 16 : 8) - d.pad_h;
                const uint8_t* wt = wtc + (size_t)n_tile * p.nkb * (2 * B_BYTES);
                const int ncb = p.ncb0 + p.ncb1;
                for (int cb = 0; cb < ncb; ++cb) {
                    const bool src0 = cb < p.ncb0;
                    const int cc = (src0 ? cb : cb - p.ncb0) * 64;
                    for (int sl = 0; sl < p.n_slabs; ++sl, ++ia) {
                        const int sa = (int)(ia % SL_NSA);
                        t3_mbar_wait(aempty_bar(sa), ((ia / SL_NSA) & 1u) ^ 1u, err);
                        const uint32_t dst = smem_base + (uint32_t)sa * SL_SLAB_MAX;
                        t3_mbar_arrive_expect_tx(afull_bar(sa), 2u * slab_plane);
                        // slab == 1: map dims {C, W, H, N}, slabs indexed by kw; slab == 2: map dims {C, H, W, N}, slabs indexed by kh
                        const int c1_ = p.slab == 1 ? x0 + sl : y0 + sl, c2_ = p.slab == 1 ? y0 : x0;
                        sl_tma_tile(dst, src0 ? &map0h : &map1h, cc, c1_, c2_, n, afull_bar(sa));
                        sl_tma_tile(dst + slab_plane, src0 ? &map0l : &map1l, cc, c1_, c2_, n, afull_bar(sa));
                        for (int t = 0; t < p.tps; ++t, ++ib) {
                            const int sb = (int)(ib % SL_NSB);
                            t3_mbar_wait(empty_bar(sb), ((ib / SL_NSB) & 1u) ^ 1u, err);
                            const int tap = p.slab == 1 ? t * d.KW + sl : sl * d.KW + t;
                            t3_mbar_arrive_expect_tx(full_bar(sb), 2 * B_BYTES);
                            t3_bulk_g2s(sl_bring + (uint32_t)sb * (2 * B_BYTES), wt + (size_t)(tap * ncb + cb) * (2 * B_BYTES), 2 * B_BYTES, full_bar(sb));
                        }
                    }
                }
            }
        } else if (!p.slab) {
            // converged warp: every lane waits on the stage barrier, one elected lane issues (operands stay in uniform registers)
            uint32_t it = 0, ps_ = 0, pph = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m_tile = (int)fastdiv((unsigned)tile, p.fd_ntiles), n_tile = tile - m_tile * p.n_ntiles;
                const int m0 = m_tile * T3_BM;
                const int t = (int)fastdiv((unsigned)m0, p.fd_wo);
                const int ow = m0 - t * d.Wo;
                const int n = (int)fastdiv((unsigned)t, p.fd_ho);
                const int oh = t - n * d.Ho;
                const int cw = ow * d.stride - d.pad_w, ch = oh * d.stride - d.pad_h;
                const uint32_t bb = (uint32_t)p.bn_eff * 128u;              // bytes of one half of a weight tile
                const uint8_t* wt = wtc + (size_t)n_tile * p.nkb * (2 * bb);
                int kb = 0, kh = 0, kw = 0;
                for (int tap = 0; tap < p.ntaps; ++tap, ++kw) {
                    if (kw == d.KW) {
                        kw = 0;
                        ++kh;
                    }
                    for (int cb = 0; cb < p.ncb0 + p.ncb1; ++cb, ++kb, ++it) {
                        const int s = (int)ps_;
                        t3_mbar_wait(empty_bar(s), pph ^ 1u, err);
                        if (++ps_ == (uint32_t)STAGES) {
                            ps_ = 0;
                            pph ^= 1u;
                        }
                        if (it == 0 && lane == 0) T3_CTA(2);
                        const uint32_t stage = smem_base + (uint32_t)s * STAGE_BYTES;
                        const uint32_t bar = full_bar(s);
                        const bool src0 = cb < p.ncb0;
                        const int cc = (src0 ? cb : cb - p.ncb0) * 64;
                        if (t3_elect_one()) {
                            if (p.dbg & 1) {                 // development: no loads (MMA + epilogue path alone)
                                t3_mbar_arrive(bar);
                            } else {
                                if (F16) {               // hi plane and the hi half of the weight tile only
                                    t3_mbar_arrive_expect_tx(bar, T3_A_BYTES + bb);
                                    t3_tma_im2col(stage, src0 ? &map0h : &map1h, cc, cw, ch, n, (uint16_t)kw, (uint16_t)kh, bar);
                                    t3_bulk_g2s(stage + 2 * T3_A_BYTES, wt + (size_t)kb * (2 * bb), bb, bar);
                                } else {
                                    t3_mbar_arrive_expect_tx(bar, 2 * T3_A_BYTES + 2 * bb);
                                    t3_tma_im2col(stage, src0 ? &map0h : &map1h, cc, cw, ch, n, (uint16_t)kw, (uint16_t)kh, bar);
                                    t3_tma_im2col(stage + T3_A_BYTES, src0 ? &map0l : &map1l, cc, cw, ch, n, (uint16_t)kw, (uint16_t)kh, bar);
                                    t3_bulk_g2s(stage + 2 * T3_A_BYTES, wt + (size_t)kb * (2 * bb), 2 * bb, bar);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer ------------------------------------------------
        // fp16 x fp16 -> fp32, K-major, M 128, N = BN (idesc) or 2 BN (idesc2)
        const uint32_t idesc = (1u << 4) | ((uint32_t)(p.bn_eff >> 3) << 17) | ((uint32_t)(T3_BM >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * p.bn_eff) >> 3) << 17) | ((uint32_t)(T3_BM >> 4) << 24);
        uint32_t lt = 0, ia = 0, ib = 0, ms = 0, mph = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            const uint32_t acc = lt & 1u, aph = (lt >> 1) & 1u;
            t3_mbar_wait(tempty_bar(acc), aph ^ 1u, err);          // epilogue has drained this accumulator
            t3_fence_after();
            const uint32_t tacc = tmem_base + acc * ACC_COLS;
            if (p.slab) {
                if (STACK) {
                    const int ncb = p.ncb0 + p.ncb1;
                    const uint32_t slab_plane = (uint32_t)(8 * (16 + p.tps - 1) * 128);
                    uint32_t first = 1u;
                    for (int cb = 0; cb < ncb; ++cb) {
                        for (int sl = 0; sl < p.n_slabs; ++sl, ++ia) {
                            const int sa = (int)(ia % SL_NSA);
                            t3_mbar_wait(afull_bar(sa), (ia / SL_NSA) & 1u, err);
                            const uint32_t a_hi0 = smem_base + (uint32_t)sa * SL_SLAB_MAX, a_lo0 = a_hi0 + slab_plane;
                            for (int t = 0; t < p.tps; ++t, ++ib) {
                                const int sb = (int)(ib % SL_NSB);
                                t3_mbar_wait(full_bar(sb), (ib / SL_NSB) & 1u, err);
                                t3_fence_after();
                                if (t3_elect_one()) {
                                    const uint32_t b_hi = sl_bring + (uint32_t)sb * (2 * B_BYTES);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const uint32_t ko = (uint32_t)k * 32u + (uint32_t)t * 1024u;      // tap t = the slab advanced by 8 rows
                                        const uint64_t dbh = t3_umma_desc(b_hi + (uint32_t)k * 32u);
                                        t3_umma(tacc, t3_umma_desc(a_hi0 + ko), dbh, idesc2, first ? 0u : 1u);
                                        t3_umma(tacc, t3_umma_desc(a_lo0 + ko), dbh, idesc, 1u);
                                        first = 0u;
                                    }
                                    t3_commit(empty_bar(sb));
                                    if (t == p.tps - 1) t3_commit(aempty_bar(sa));
                                    if (t == p.tps - 1 && sl == p.n_slabs - 1 && cb == ncb - 1) {
                                        t3_commit(tfull_bar(acc));
                                        T3_CTA(4);
                                    }
                                }
                                __syncwarp();
                            }
                        }
                    }
                }
                continue;
            }
            // The issuing thread is the bottleneck of this kernel's main loop, so the loop runs warp-converged (every lane waits on the barrier,
            // stage index and phase are carried, descriptors are base + constant) and only the MMA / commit instructions sit under elect.sync.
            // The first k-block is peeled off: it alone starts the accumulator (accumulate = 0 on its first MMA), so the steady-state loop
            // issues eight unconditional MMAs per k-block.  Measured: ONE more uniform predicate per MMA pair in this loop cost 10 % of
            // the update block -- the issue sequence of the single MMA thread is on the critical path of these short-N kernels.
            auto kblock = [&](auto first, auto half) {
                constexpr bool FIRST = decltype(first)::value;
                constexpr int NK = decltype(half)::value ? 2 : 4;
                t3_mbar_wait(full_bar(ms), mph, err);
                t3_fence_after();
                if (FIRST && lt == 0 && lane == 0) T3_CTA(3);
                const uint32_t a_hi = smem_base + ms * (uint32_t)STAGE_BYTES;
                const uint64_t dah0 = t3_umma_desc(a_hi);
                const uint64_t dal0 = dah0 + (uint64_t)(T3_A_BYTES >> 4);
                const uint64_t dbh0 = dah0 + (uint64_t)((2 * T3_A_BYTES) >> 4);
                const uint64_t dbl0 = dbh0 + (uint64_t)(B_BYTES >> 4);
                const uint32_t ebar_s = empty_bar(ms);
                if (t3_elect_one()) {
#pragma unroll
                    for (int k = 0; k < NK; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);             // 32 bytes along K, in 16-byte descriptor units
                        const uint32_t acc0 = (FIRST && k == 0) ? 0u : 1u;
                        if (F16) {
                            t3_umma(tacc, dah0 + ko, dbh0 + ko, idesc, acc0);       // hi*hi alone
                        } else if (STACK) {
                            t3_umma(tacc, dah0 + ko, dbh0 + ko, idesc2, acc0);      // hi*hi | hi*lo
                            t3_umma(tacc, dal0 + ko, dbh0 + ko, idesc, 1u);         // lo*hi
                        } else {
                            t3_umma(tacc, dah0 + ko, dbh0 + ko, idesc, acc0);
                            t3_umma(tacc, dah0 + ko, dbl0 + ko, idesc, 1u);
                            t3_umma(tacc, dal0 + ko, dbh0 + ko, idesc, 1u);
                        }
                    }
                    t3_commit(ebar_s);
                }
                __syncwarp();
                if (++ms == (uint32_t)STAGES) {
                    ms = 0;
                    mph ^= 1u;
                }
            };
            if (p.half0 == 0) kblock(std::true_type{}, std::true_type{});
            else kblock(std::true_type{}, std::false_type{});
            const int ncb_ = p.ncb0 + p.ncb1;
            int cbi = ncb_ > 1 ? 1 : 0;
#pragma unroll 1
            for (int kb = 1; kb < p.nkb; ++kb) {
                if (cbi == p.half0 || cbi == p.half1) kblock(std::false_type{}, std::true_type{});
                else kblock(std::false_type{}, std::false_type{});
                if (++cbi == ncb_) cbi = 0;
            }
            if (t3_elect_one()) {
                t3_commit(tfull_bar(acc));      // arrives when every MMA issued above has completed
                T3_CTA(4);
            }
            __syncwarp();
        }
        t3_fence_before();
    } else {
        // ------------------------------------------------ epilogue ------------------------------------------------
        // 8 warps: quadrant (warp & 3) of the TMEM lanes = 32 tile rows, column half ((warp - 2) >> 2).  The tile's bias slice is
        // staged in shared memory once per n_tile; all tcgen05.ld of a thread's columns are issued before the single wait.
        constexpr int HALF = BN / 2;                  // columns per thread
        constexpr bool two_halves = STACK && !F16;    // the accumulator is [hi*hi + lo*hi | hi*lo]: the epilogue adds the halves
        constexpr bool wlo = !F16;                    // BFLOW_PREC_F16: nobody reads the lo planes, the hot store paths skip them
        const int quad = warp & 3;
        const int chalf = (warp - 2) >> 2;
        const int etid = tid - 64;                   // 0 .. 255
        const int hoff = p.bn_eff;                   // TMEM column of the accumulator's second half (hi*lo)
        float* s_bias = reinterpret_cast<float*>(smem_raw + (bars - t3_smem_u32(smem_raw)) + 128);   // BN floats behind the barriers
        const float lo1 = d.act1 == BFLOW_ACT_RELU ? 0.f : -INFINITY, lo2 = d.act2 == BFLOW_ACT_RELU ? 0.f : -INFINITY;
        const bool slow1 = d.act1 >= BFLOW_ACT_SIGMOID, slow2 = d.act2 >= BFLOW_ACT_SIGMOID;
        const bool aligned = (d.y == nullptr || (((d.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.y) & 15) == 0))) &&
                             (d.res == nullptr || (((d.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(d.res) & 15) == 0)));
        const float post = d.scale;
        const bool wide16 = d.y16_hi != nullptr && (d.ldy16 & 7) == 0 && ((reinterpret_cast<uintptr_t>(d.y16_hi) | reinterpret_cast<uintptr_t>(d.y16_lo)) & 15) == 0;
        if (p.staged == 6) {
            // GRU gate epilogues of a single-tile CTA (update.py:36-45), everything per-element fetched BEFORE the accumulator is ready.
            // Measured (tools/timeline.py --cta-iter): the z|r and candidate launches spent 7-13 us after their last MMA, longer than
            // their main loops, waiting for the hoisted term, h and z to arrive and walking staging tiles.  The 8 epilogue warps are idle
            // while the pipeline fills and the MMAs run, and a single-tile CTA uses only one of the two TMEM accumulators.  So, right
            // after the prologue, each thread loads its row's slices of the hoisted term `res`, of h (r gate) or of z and h (candidate)
            // and parks them with tcgen05.st in the TMEM columns BEHIND accumulator 0 (thread = TMEM lane = tile row, the layout
            // tcgen05.ld hands back): no registers held across the main loop, no shared memory (the pipeline owns it), and the slow
            // row-wise loads (~3.6 us per 64 KB through the LSU) hide under the MMAs.  (Writing `res` into the accumulator itself and letting
            // the MMAs add to it was measured first: it delays the first MMA by 1-3 us.)  After the last MMA only TMEM -> registers ->
            // gate arithmetic -> SWIZZLE_128B boxes in the idle pipeline stages -> one thread issues the tensor-map stores remains.
            // r itself is never stored (only r * h is consumed, update.py:39).
            const int tile = blockIdx.x;
            const int m_tile = tile / p.n_ntiles, n_tile = tile - m_tile * p.n_ntiles;
            const int n0 = n_tile * BN, mt0 = m_tile * T3_BM;
            const int row = quad * 32 + lane;
            const int colbase = chalf * HALF;
            const int mm = mt0 + row;
            const bool row_ok = mm < p.M;
            const bool is_zr = d.epi == BFLOW_EPI_GRU_ZR;
            const int Cg = d.Cout >> 1;
            const bool r_tile = is_zr && n0 >= Cg;
            const int acol0 = is_zr ? n0 - Cg : n0;              // column of the aux0 / fp16 tile
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)colbase;
            const uint32_t t_res = taddr + (uint32_t)ACC_COLS, t_a = t_res + (uint32_t)BN, t_y = t_a + (uint32_t)BN;      // parked operands
            const bool hasA = r_tile || !is_zr, hasY = !is_zr;
            // Parking one operand: the warp owns rows [quad*32, +32) x columns [colbase, +HALF) of the tile.  Loading thread = row would
            // touch 32 different lines per instruction (measured: the LSU then moves ~9 B/clk and the TMA loads of the main loop
            // queue behind it), so the warp loads PK_ROWS whole row slices per pass with lanes along the columns (a few full lines per
            // instruction), bounces them through a private slab of the 32 KB of shared memory behind the pipeline stages, and every
            // lane picks its own row up from there.
            constexpr int PK_ROWS = 512 / HALF;               // rows per pass: 8 (BN 128), 16 (BN 64)
            constexpr int PK_PITCH = HALF + 4;                // floats; +4 keeps the row-wise 16-byte reads conflict-free
            constexpr int PK_C4 = HALF / 4;                   // float4 per row slice
            float* pk = reinterpret_cast<float*>(smem_raw + (smem_base - t3_smem_u32(smem_raw)) + STAGES * STAGE_BYTES) + (warp - 2) * (PK_ROWS * PK_PITCH);
            static_assert(STAGES * STAGE_BYTES + T3_EPI_WARPS * PK_ROWS * PK_PITCH * 4 <= AREA_BYTES || BN > 128, "operand bounce buffer does not fit behind the stages");
            auto park = [&](const float* gtile, int ldg, uint32_t tdst) {      // gtile: (row mt0, column of this warp's slice)
                float v[HALF];
#pragma unroll
                for (int pass = 0; pass < 32 / PK_ROWS; ++pass) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int idx = i * 32 + lane;
                        const int rr = idx / PK_C4, c4 = idx - rr * PK_C4;
                        const int gr = quad * 32 + pass * PK_ROWS + rr;
                        const float4 x4 = (mt0 + gr < p.M) ? *reinterpret_cast<const float4*>(gtile + (size_t)gr * ldg + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                        *reinterpret_cast<float4*>(pk + rr * PK_PITCH + c4 * 4) = x4;
                    }
                    __syncwarp();
                    if (lane / PK_ROWS == pass) {
                        const float* prow = pk + (lane - pass * PK_ROWS) * PK_PITCH;
#pragma unroll
                        for (int c = 0; c < HALF; c += 4) {
                            const float4 x4 = *reinterpret_cast<const float4*>(prow + c);
                            v[c] = x4.x; v[c + 1] = x4.y; v[c + 2] = x4.z; v[c + 3] = x4.w;
                        }
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int c = 0; c < HALF; c += 16) t3_tmem_st16(tdst + (uint32_t)c, v + c);
            };
            if (BN <= 128) {
                park(d.res + (size_t)mt0 * d.ldr + n0 + colbase, d.ldr, t_res);
                if (hasA) park(d.aux0 + (size_t)mt0 * d.ld_aux0 + acol0 + colbase, d.ld_aux0, t_a);
                if (hasY) park(d.y + (size_t)mt0 * d.ldy + n0 + colbase, d.ldy, t_y);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            uint8_t* sb = smem_raw + (smem_base - t3_smem_u32(smem_raw));
            constexpr int NB16 = (BN + 63) / 64, NB32 = (BN + 31) / 32;
            uint8_t* s_32 = sb;                                  // fp32 boxes (z, or the new h)
            uint8_t* s_hi = sb + NB32 * 16384;
            uint8_t* s_lo = s_hi + NB16 * 16384;
            const bool out32 = !r_tile;
            const bool out16 = is_zr ? r_tile : d.y16_hi != nullptr;
            t3_mbar_wait(tfull_bar(0), 0u, err);
            t3_fence_after();
            if (warp == 2 && lane == 0) T3_CTA(5);
#pragma unroll 1
            for (int c0 = 0; c0 < HALF; c0 += 16) {
                float v[16], rs[16], av[16], yv[16];
                t3_tmem_ld16_nowait(taddr + (uint32_t)c0, v);
                t3_tmem_ld16_nowait(t_res + (uint32_t)c0, rs);
                if (hasA) t3_tmem_ld16_nowait(t_a + (uint32_t)c0, av);
                if (hasY) t3_tmem_ld16_nowait(t_y + (uint32_t)c0, yv);
                if (two_halves) {
                    float u[16];
                    t3_tmem_ld16_nowait(taddr + (uint32_t)(hoff + c0), u);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < 16; ++c) v[c] += u[c];
                } else {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                }
                float h[16];
                if (is_zr) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        v[c] = fast_sigmoid(fmaf(v[c], p.acc_scale, rs[c]));
                        h[c] = hasA ? v[c] * av[c] : 0.f;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        v[c] = (1.f - av[c]) * yv[c] + av[c] * fast_tanh(fmaf(v[c], p.acc_scale, rs[c]));
                        h[c] = v[c];
                    }
                }
                const int col = colbase + c0;
                if (out32) {
                    uint8_t* b3 = s_32 + (col >> 5) * 16384 + row * 128;
#pragma unroll
                    for (int c = 0; c < 16; c += 4) {
                        const int chunk = (((col & 31) + c) >> 2) ^ (row & 7);
                        *reinterpret_cast<float4*>(b3 + (chunk << 4)) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                    }
                }
                if (out16) {
                    uint8_t* bh = s_hi + (col >> 6) * 16384 + row * 128;
                    uint8_t* bl = s_lo + (col >> 6) * 16384 + row * 128;
#pragma unroll
                    for (int c = 0; c < 16; c += 8) {
                        uint4 h4, l4;
                        split2(h[c], h[c + 1], h4.x, l4.x);
                        split2(h[c + 2], h[c + 3], h4.y, l4.y);
                        split2(h[c + 4], h[c + 5], h4.z, l4.z);
                        split2(h[c + 6], h[c + 7], h4.w, l4.w);
                        const int chunk = (((col & 63) + c) >> 3) ^ (row & 7);
                        *reinterpret_cast<uint4*>(bh + (chunk << 4)) = h4;
                        if (wlo) *reinterpret_cast<uint4*>(bl + (chunk << 4)) = l4;
                    }
                }
            }
            if (warp == 2 && lane == 0) T3_CTA(8);
            t3_fence_before();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (warp == 2 && t3_elect_one()) {
                if (out32) {
#pragma unroll
                    for (int b = 0; b < NB32; ++b) t3_tma_store2d(&omap32, t3_smem_u32(s_32 + b * 16384), n0 + b * 32, mt0);
                }
                if (out16) {
#pragma unroll
                    for (int b = 0; b < NB16; ++b) {
                        t3_tma_store2d(&omap_hi, t3_smem_u32(s_hi + b * 16384), acol0 + b * 64, mt0);
                        if (wlo) t3_tma_store2d(&omap_lo, t3_smem_u32(s_lo + b * 16384), acol0 + b * 64, mt0);
                    }
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the boxes have left shared memory; the writes complete with the grid
            }
            __syncwarp();
            if (warp == 2 && lane == 0) T3_CTA(6);
        } else {
        uint32_t lt = 0;
        int bias_tile = -1;
        // fused InstanceNorm statistics: per-CTA partial (sum, sum of squares) of the current image in shared memory, flushed with
        // one double atomicAdd pair per column when the CTA moves on to another image (contention on the global table stays low)
        float* s_stat = s_bias + BN;                 // [2][BN]
        int cur_img = -1;
        const int shw = d.stats_hw > 0 ? d.stats_hw : d.Ho * d.Wo;
        auto flush_stats = [&](int img, int n0_) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int j = etid; j < BN; j += 256) {       // every thread reads and clears only its own columns
                if (img >= 0 && n0_ + j < d.Cout) {
                    atomicAdd(d.stats + ((size_t)img * d.Cout + n0_ + j) * 2, (double)s_stat[j]);
                    atomicAdd(d.stats + ((size_t)img * d.Cout + n0_ + j) * 2 + 1, (double)s_stat[BN + j]);
                }
                s_stat[j] = 0.f;
                s_stat[BN + j] = 0.f;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        };
        int stat_n0 = 0;
        if (d.stats != nullptr) flush_stats(-1, 0);
        // column sums of 16 consecutive channels over the warp's 32 tile rows (all of one image): butterfly transpose-reduce, then the lanes with
        // the low bit clear own one column each and add (sum, sum of squares) to the CTA's shared accumulators
        auto stat16 = [&](const float* x, int col0) {
            float sv[16], sq[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                sv[j] = x[j];
                sq[j] = x[j] * x[j];
            }
#pragma unroll
            for (int width = 8, bit = 16; width >= 1; width >>= 1, bit >>= 1) {
                const bool upper = (lane & bit) != 0;
#pragma unroll
                for (int i = 0; i < width; ++i) {
                    const float keep_s = upper ? sv[width + i] : sv[i], send_s = upper ? sv[i] : sv[width + i];
                    const float keep_q = upper ? sq[width + i] : sq[i], send_q = upper ? sq[i] : sq[width + i];
                    sv[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, bit);
                    sq[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, bit);
                }
            }
            const float ts = sv[0] + __shfl_xor_sync(0xffffffffu, sv[0], 1);
            const float tq = sq[0] + __shfl_xor_sync(0xffffffffu, sq[0], 1);
            if ((lane & 1) == 0) {
                const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                atomicAdd(s_stat + col0 + col, ts);
                atomicAdd(s_stat + BN + col0 + col, tq);
            }
        };
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            const int m_tile = tile / p.n_ntiles, n_tile = tile - m_tile * p.n_ntiles;
            // tile row -> output pixel.  im2col mode: 128 consecutive pixels; slab mode: an 8 x 16 patch of image t_n at (t_y0, t_x0)
            int t_n = 0, t_y0 = 0, t_x0 = 0;
            if (p.slab) {
                const int tpi = p.tiles_x * p.tiles_y;
                t_n = m_tile / tpi;
                const int r_ = m_tile - t_n * tpi;
                const int ty_ = r_ / p.tiles_x;
                t_y0 = ty_ * (p.slab == 1 ? 16 : 8);
                t_x0 = (r_ - ty_ * p.tiles_x) * (p.slab == 1 ? 8 : 16);
            }
            auto row_m = [&](int row, bool& ok) -> int {
                if (!p.slab) {
                    const int mm_ = m_tile * T3_BM + row;
                    ok = mm_ < p.M;
                    return mm_;
                }
                const int yy = p.slab == 1 ? (row >> 3) : (row & 7), xx = p.slab == 1 ? (row & 7) : (row >> 3);
                const int y = t_y0 + yy, x = t_x0 + xx;
                ok = y < d.Ho && x < d.Wo;
                return (t_n * d.Ho + y) * d.Wo + x;
            };
            bool m_ok;
            const int m = row_m(quad * 32 + lane, m_ok);
            const int n0 = n_tile * BN;
            if (n_tile != bias_tile) {               // uniform over the epilogue warps
                asm volatile("bar.sync 1, 256;" ::: "memory");      // everyone is done with the previous slice
                for (int j = etid; j < BN; j += 256) s_bias[j] = (d.bias != nullptr && n0 + j < d.Cout) ? __ldg(d.bias + n0 + j) : 0.f;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                bias_tile = n_tile;
            }
            if (d.stats != nullptr) {
                const int tile_img = p.slab ? t_n : (m_tile * T3_BM) / shw;
                if (tile_img != cur_img || n0 != stat_n0) {
                    flush_stats(cur_img, stat_n0);
                    cur_img = tile_img;
                    stat_n0 = n0;
                }
            }
            const uint32_t acc = lt & 1u, aph = (lt >> 1) & 1u;
            t3_mbar_wait(tfull_bar(acc), aph, err);
            t3_fence_after();
            if (warp == 2 && lane == 0) T3_TRACE(3, lt);
            if (warp == 2 && lane == 0) T3_CTA(5);
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * ACC_COLS + (uint32_t)(chalf * HALF);
            const int nb0 = n0 + chalf * HALF;
            if (p.staged == 5) {
                // GRU gate epilogues of a single-tile CTA on tensor maps (update.py:36-45).  One elected thread fetches the tile of every
                // per-element operand (hoisted term `res`; h for the r gate / z for the candidate; the state h itself for the candidate) as
                // SWIZZLE_128B boxes of 32 floats into the idle pipeline stages while the accumulator is being unloaded; each thread then
                // finishes ITS row out of shared memory (chunk j of row r sits at j ^ (r & 7): conflict-free for the 16-byte accesses of a
                // quarter-warp), writes the fp32 result in place of an operand and the split-fp16 planes next to it, and one thread stores
                // the boxes.  No per-lane global access, no staging pass.  (Per-row bulk copies: 7.4 us per 128x128 tile; batched loads: 10.8.)
                uint8_t* sb = smem_raw + (smem_base - t3_smem_u32(smem_raw));
                constexpr int NB16 = (BN + 63) / 64, NB32 = (BN + 31) / 32;
                const int Cg = d.Cout >> 1;
                const bool is_zr = d.epi == BFLOW_EPI_GRU_ZR;
                const bool r_tile = is_zr && n0 >= Cg;
                const bool hasR = d.res != nullptr, hasA = r_tile || !is_zr, hasY = !is_zr;
                const bool out16 = is_zr ? r_tile : d.y16_hi != nullptr;
                uint8_t* s_R = sb;                                   // res, then (GRU_ZR) the fp32 gates in place
                uint8_t* s_A = sb + NB32 * 16384;                    // h (r gate) / z (candidate)
                uint8_t* s_Y = sb + 2 * NB32 * 16384;                // candidate: h in, h out
                uint8_t* s_hi = sb + (is_zr ? 2 : 3) * NB32 * 16384;      // the z|r launch has no Y region
                uint8_t* s_lo = s_hi + NB16 * 16384;
                uint8_t* s_O = is_zr ? s_R : s_Y;
                const int mt0 = m_tile * T3_BM;
                const int acol0 = is_zr ? n0 - Cg : n0;              // column of the aux0 / fp16 tile
                if (warp == 2 && t3_elect_one()) {
                    const uint32_t nbox = (uint32_t)NB32 * ((hasR ? 1u : 0u) + (hasA ? 1u : 0u) + (hasY ? 1u : 0u));
                    t3_mbar_arrive_expect_tx(ebar, nbox * 16384u);
#pragma unroll
                    for (int b = 0; b < NB32; ++b) {
                        if (hasR) t3_tma_load2d(t3_smem_u32(s_R + b * 16384), &rmap, n0 + b * 32, mt0, ebar);
                        if (hasA) t3_tma_load2d(t3_smem_u32(s_A + b * 16384), &amap, acol0 + b * 32, mt0, ebar);
                        if (hasY) t3_tma_load2d(t3_smem_u32(s_Y + b * 16384), &omap32, n0 + b * 32, mt0, ebar);
                    }
                }
                __syncwarp();
                const int row = quad * 32 + lane;
                const int colbase = chalf * HALF;
                bool waited = false;
#pragma unroll 1
                for (int c0 = 0; c0 < HALF; c0 += 32) {
                    float v[32];
                    t3_tmem_ld16_nowait(taddr + (uint32_t)c0, v);
                    t3_tmem_ld16_nowait(taddr + (uint32_t)(c0 + 16), v + 16);
                    if (two_halves) {
                        float u[32];
                        t3_tmem_ld16_nowait(taddr + (uint32_t)(hoff + c0), u);
                        t3_tmem_ld16_nowait(taddr + (uint32_t)(hoff + c0 + 16), u + 16);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] += u[c];
                    } else {
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    }
                    if (!waited) {
                        t3_mbar_wait(ebar, 0u, err);
                        waited = true;
                    }
                    const int col = colbase + c0;
                    const int boff = (col >> 5) * 16384 + row * 128;          // this thread's 128-byte row inside the fp32 box of these 32 columns
                    float h[32];
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const int ch = ((c >> 2) ^ (row & 7)) << 4;
                        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + colbase + c0 + c);
                        float o[4] = {post * fmaf(v[c], p.acc_scale, b4.x), post * fmaf(v[c + 1], p.acc_scale, b4.y), post * fmaf(v[c + 2], p.acc_scale, b4.z),
                                      post * fmaf(v[c + 3], p.acc_scale, b4.w)};
                        if (hasR) {
                            const float4 r4 = *reinterpret_cast<const float4*>(s_R + boff + ch);
                            o[0] += r4.x; o[1] += r4.y; o[2] += r4.z; o[3] += r4.w;
                        }
                        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (hasA) a4 = *reinterpret_cast<const float4*>(s_A + boff + ch);
                        if (is_zr) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) o[j] = fast_sigmoid(o[j]);
                            h[c] = o[0] * a4.x; h[c + 1] = o[1] * a4.y; h[c + 2] = o[2] * a4.z; h[c + 3] = o[3] * a4.w;
                        } else {
                            const float4 y4 = *reinterpret_cast<const float4*>(s_Y + boff + ch);
                            o[0] = (1.f - a4.x) * y4.x + a4.x * fast_tanh(o[0]);
                            o[1] = (1.f - a4.y) * y4.y + a4.y * fast_tanh(o[1]);
                            o[2] = (1.f - a4.z) * y4.z + a4.z * fast_tanh(o[2]);
                            o[3] = (1.f - a4.w) * y4.w + a4.w * fast_tanh(o[3]);
                            h[c] = o[0]; h[c + 1] = o[1]; h[c + 2] = o[2]; h[c + 3] = o[3];
                        }
                        *reinterpret_cast<float4*>(s_O + boff + ch) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                    if (out16) {
                        uint8_t* bh = s_hi + (col >> 6) * 16384 + row * 128;
                        uint8_t* bl = s_lo + (col >> 6) * 16384 + row * 128;
#pragma unroll
                        for (int c = 0; c < 32; c += 8) {
                            uint4 h4, l4;
                            split2(h[c], h[c + 1], h4.x, l4.x);
                            split2(h[c + 2], h[c + 3], h4.y, l4.y);
                            split2(h[c + 4], h[c + 5], h4.z, l4.z);
                            split2(h[c + 6], h[c + 7], h4.w, l4.w);
                            const int chunk = (((col & 63) + c) >> 3) ^ (row & 7);
                            *reinterpret_cast<uint4*>(bh + (chunk << 4)) = h4;
                            *reinterpret_cast<uint4*>(bl + (chunk << 4)) = l4;
                        }
                    }
                }
                if (warp == 2 && lane == 0) T3_CTA(8);
                t3_fence_before();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (warp == 2 && t3_elect_one()) {
#pragma unroll
                    for (int b = 0; b < NB32; ++b) t3_tma_store2d(&omap32, t3_smem_u32(s_O + b * 16384), n0 + b * 32, mt0);
                    if (out16) {
#pragma unroll
                        for (int b = 0; b < NB16; ++b) {
                            t3_tma_store2d(&omap_hi, t3_smem_u32(s_hi + b * 16384), acol0 + b * 64, mt0);
                            t3_tma_store2d(&omap_lo, t3_smem_u32(s_lo + b * 16384), acol0 + b * 64, mt0);
                        }
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                }
                __syncwarp();
                if (lane == 0) t3_mbar_arrive(tempty_bar(acc));
            } else if (p.staged == 4) {
                // Single-tile CTA, plain epilogue (bias, scale, none / relu), outputs through 2-D tensor-map stores: each thread converts its row
                // straight from the accumulator into SWIZZLE_128B boxes in the idle pipeline stages ([128 rows][128 bytes] per 64 halves / 32
                // floats of width; chunk j of row r sits at j ^ (r & 7), so the 16-byte stores of a quarter-warp hit 8 different bank groups), and
                // ONE thread issues a cp.async.bulk.tensor per box.  Rows beyond M and channels beyond Cout are clipped by the tensor map.
                // (The per-row bulk-copy variant below spent 2.3 us issuing 256 row copies and 1.7 us in a separate conversion pass.)
                uint8_t* sb = smem_raw + (smem_base - t3_smem_u32(smem_raw));
                constexpr int NB16 = (BN + 63) / 64, NB32 = (BN + 31) / 32;          // boxes per plane
                uint8_t* s_hi = sb;
                uint8_t* s_lo = sb + NB16 * 16384;
                uint8_t* s_32 = sb + 2 * NB16 * 16384;
                const bool o16 = d.y16_hi != nullptr, o32 = d.y != nullptr;
                const int row = quad * 32 + lane;
                const int colbase = chalf * HALF;
#pragma unroll 1
                for (int c0 = 0; c0 < HALF; c0 += 32) {
                    float v[32];
                    t3_tmem_ld16_nowait(taddr + (uint32_t)c0, v);
                    t3_tmem_ld16_nowait(taddr + (uint32_t)(c0 + 16), v + 16);
                    if (two_halves) {
                        float u[32];
                        t3_tmem_ld16_nowait(taddr + (uint32_t)(hoff + c0), u);
                        t3_tmem_ld16_nowait(taddr + (uint32_t)(hoff + c0 + 16), u + 16);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] += u[c];
                    } else {
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    }
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + colbase + c0 + c);
                        v[c] = fmaxf(fmaxf(post * fmaf(v[c], p.acc_scale, b4.x), lo1), lo2);
                        v[c + 1] = fmaxf(fmaxf(post * fmaf(v[c + 1], p.acc_scale, b4.y), lo1), lo2);
                        v[c + 2] = fmaxf(fmaxf(post * fmaf(v[c + 2], p.acc_scale, b4.z), lo1), lo2);
                        v[c + 3] = fmaxf(fmaxf(post * fmaf(v[c + 3], p.acc_scale, b4.w), lo1), lo2);
                    }
                    const int col = colbase + c0;                  // first of these 32 columns inside the tile
                    if (o16) {
                        uint8_t* bh = s_hi + (col >> 6) * 16384 + row * 128;
                        uint8_t* bl = s_lo + (col >> 6) * 16384 + row * 128;
#pragma unroll
                        for (int c = 0; c < 32; c += 8) {
                            uint4 h4, l4;
                            split2(v[c], v[c + 1], h4.x, l4.x);
                            split2(v[c + 2], v[c + 3], h4.y, l4.y);
                            split2(v[c + 4], v[c + 5], h4.z, l4.z);
                            split2(v[c + 6], v[c + 7], h4.w, l4.w);
                            const int chunk = (((col & 63) + c) >> 3) ^ (row & 7);
                            *reinterpret_cast<uint4*>(bh + (chunk << 4)) = h4;
                            if (wlo) *reinterpret_cast<uint4*>(bl + (chunk << 4)) = l4;
                        }
                    }
                    if (o32) {
                        uint8_t* b3 = s_32 + (col >> 5) * 16384 + row * 128;
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            const int chunk = (c >> 2) ^ (row & 7);
                            *reinterpret_cast<float4*>(b3 + (chunk << 4)) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                        }
                    }
                }
                if (warp == 2 && lane == 0) T3_CTA(8);
                t3_fence_before();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (warp == 2 && t3_elect_one()) {
                    const int mt0 = m_tile * T3_BM;
                    if (o16) {
#pragma unroll
                        for (int b = 0; b < NB16; ++b) {
                            if (n0 + b * 64 < d.Cout) {
                                t3_tma_store2d(&omap_hi, t3_smem_u32(s_hi + b * 16384), n0 + b * 64, mt0);
                                if (wlo) t3_tma_store2d(&omap_lo, t3_smem_u32(s_lo + b * 16384), n0 + b * 64, mt0);
                            }
                        }
                    }
                    if (o32) {
#pragma unroll
                        for (int b = 0; b < NB32; ++b)
                            if (n0 + b * 32 < d.Cout) t3_tma_store2d(&omap32, t3_smem_u32(s_32 + b * 16384), n0 + b * 32, mt0);
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                }
                // Cout % 8 == 4: the fp16 maps cover Cout - 4 channels (a box clipped inside a 16-byte chunk would write the whole chunk); the last
                // four channels of each row go out as one 8-byte store per plane
                const int c8 = d.Cout & ~7;
                if (o16 && c8 < d.Cout && c8 >= n0 && c8 < n0 + BN && etid < T3_BM) {
                    const int mm = m_tile * T3_BM + etid;
                    if (mm < p.M) {
                        const int lc = c8 - n0;
                        const int off = (lc >> 6) * 16384 + etid * 128 + ((((lc & 63) >> 3) ^ (etid & 7)) << 4);
                        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(d.y16_hi) + (size_t)mm * d.ldy16 + c8) = *reinterpret_cast<const uint2*>(s_hi + off);
                        if (wlo) *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(d.y16_lo) + (size_t)mm * d.ldy16 + c8) = *reinterpret_cast<const uint2*>(s_lo + off);
                    }
                }
                __syncwarp();
                if (lane == 0) t3_mbar_arrive(tempty_bar(acc));
            } else if (p.staged == 3) {
                // Multi-tile CTAs with a plain fp32 store epilogue (the all-pairs correlation GEMM, corr.py:264-272: 1444 tiles of 64 KB): the
                // direct path below stores 16 bytes per row per instruction and was measured LSU-bound (~7 us per tile against a 2 us main
                // loop).  Here the pipeline runs with one stage less and the freed shared memory is a [128][BN + 4] fp32 staging tile: the
                // epilogue warps park the tile there and one cp.async.bulk per row streams it out while the next tile's MMAs run.
                constexpr int PITCH = BN + 4;
                float* S = reinterpret_cast<float*>(smem_raw + (smem_base - t3_smem_u32(smem_raw)) + STAGES * STAGE_BYTES);
                const int mt0 = m_tile * T3_BM;
                const int ncols = min(BN, d.Cout - n0);
                if (etid < T3_BM) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // previous tile's rows have left the staging tile
                asm volatile("bar.sync 1, 256;" ::: "memory");
                {
                    float* trow = S + (quad * 32 + lane) * PITCH + chalf * HALF;
#pragma unroll 1
                    for (int c0 = 0; c0 < HALF; c0 += 32) {
                        float v[32];
                        t3_tmem_ld16_nowait(taddr + (uint32_t)c0, v);
                        t3_tmem_ld16_nowait(taddr + (uint32_t)(c0 + 16), v + 16);
                        if (two_halves) {
                            float u[32];
                            t3_tmem_ld16_nowait(taddr + (uint32_t)(hoff + c0), u);
                            t3_tmem_ld16_nowait(taddr + (uint32_t)(hoff + c0 + 16), u + 16);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int c = 0; c < 32; ++c) v[c] += u[c];
                        } else {
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        }
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + chalf * HALF + c0 + c);
                            v[c] = fmaxf(post * fmaf(v[c], p.acc_scale, b4.x), lo1);
                            v[c + 1] = fmaxf(post * fmaf(v[c + 1], p.acc_scale, b4.y), lo1);
                            v[c + 2] = fmaxf(post * fmaf(v[c + 2], p.acc_scale, b4.z), lo1);
                            v[c + 3] = fmaxf(post * fmaf(v[c + 3], p.acc_scale, b4.w), lo1);
                            *reinterpret_cast<float4*>(trow + c0 + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                        }
                        if (d.stats != nullptr) {          // host contract: no activation with statistics, every tile inside one image
                            bool rok_;
                            (void)row_m(quad * 32 + lane, rok_);
                            if (!rok_) {
#pragma unroll
                                for (int c = 0; c < 32; ++c) v[c] = 0.f;
                            }
                            if (nb0 + c0 + 15 < d.Cout) stat16(v, chalf * HALF + c0);
                            if (nb0 + c0 + 31 < d.Cout) stat16(v + 16, chalf * HALF + c0 + 16);
                        }
                    }
                }
                t3_fence_before();
                __syncwarp();
                if (lane == 0) t3_mbar_arrive(tempty_bar(acc));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                bool row_ok = false;
                const int row_mm = etid < T3_BM ? row_m(etid, row_ok) : 0;
                if (etid < T3_BM && row_ok) {
                    t3_bulk_s2g(d.y + (size_t)row_mm * d.ldy + n0, t3_smem_u32(S + etid * PITCH), (uint32_t)ncols * 4u);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else if (p.staged == 2) {
                // Bulk-copy epilogue of a single-tile CTA.  Measured (tools/timeline.py --cta-label): ordinary loads / stores issued by the 8
                // epilogue warps move ~9 bytes per clock per SM -- a 128x128 GRU epilogue took 13 us, longer than its main loop.  The
                // pipeline stages are idle once tmem_full has fired, so: (1) one thread per tile row fetches the row of every per-element
                // operand (hoisted GRU term, gate, state) with cp.async.bulk into the stage area; (2) meanwhile the accumulator goes
                // TMEM -> registers -> T[128][BN+4]; (3) 256 threads walk the tile row-major out of shared memory and leave the results
                // there (fp32 in place of an operand, split-fp16 planes in place of T); (4) one cp.async.bulk per row and output plane
                // writes them back.  Global traffic never touches the LSU.
                constexpr int PITCH = BN + 4;
                constexpr int C4 = BN / 4;
                constexpr int EPI_BATCH = 4;
                constexpr int RPB = 256 * EPI_BATCH / C4;                   // tile rows per batch: 32 (BN 128), 64 (BN 64)
                uint8_t* sb = smem_raw + (smem_base - t3_smem_u32(smem_raw));
                float* T = reinterpret_cast<float*>(sb);
                float* Rg = reinterpret_cast<float*>(sb + T3_BM * PITCH * 4);      // res, then fp32 output (unless GRU_Q)
                float* Ag = Rg + T3_BM * BN;                                        // aux0: h (GRU_ZR, r tile) / z (GRU_Q)
                float* Yg = Ag + T3_BM * BN;                                        // GRU_Q: h in, h out
                const int mt0 = m_tile * T3_BM;
                const int Cg = d.Cout >> 1;
                const bool is_zr = d.epi == BFLOW_EPI_GRU_ZR, is_q = d.epi == BFLOW_EPI_GRU_Q;
                const bool r_tile = is_zr && n0 >= Cg;
                const bool hasR = d.res != nullptr, hasA = r_tile || is_q, hasY = is_q;
                const bool out32 = d.y != nullptr;
                const bool out16 = is_zr ? r_tile : d.y16_hi != nullptr;
                float* Og = is_q ? Yg : Rg;
                const int ncols = min(BN, d.Cout - n0);
                int vrows;                                  // number of tile rows that are output pixels
                if (!p.slab) vrows = min(T3_BM, p.M - mt0);
                else vrows = min(p.slab == 1 ? 16 : 8, d.Ho - t_y0) * min(p.slab == 1 ? 8 : 16, d.Wo - t_x0);
                bool row_ok = false;
                const int row_mm = etid < T3_BM ? row_m(etid, row_ok) : 0;
                if (etid == 0) {
                    const uint32_t per_row = (uint32_t)ncols * 4u * ((hasR ? 1u : 0u) + (hasA ? 1u : 0u) + (hasY ? 1u : 0u));
                    t3_mbar_arrive_expect_tx(ebar, per_row * (uint32_t)vrows);
                }
                if (etid < T3_BM && row_ok) {
                    const size_t mm = (size_t)row_mm;
                    const uint32_t nb = (uint32_t)ncols * 4u;
                    if (hasR) t3_bulk_g2s(t3_smem_u32(Rg + etid * BN), d.res + mm * d.ldr + n0, nb, ebar);
                    if (hasA) t3_bulk_g2s(t3_smem_u32(Ag + etid * BN), d.aux0 + mm * d.ld_aux0 + (is_zr ? n0 - Cg : n0), nb, ebar);
                    if (hasY) t3_bulk_g2s(t3_smem_u32(Yg + etid * BN), d.y + mm * d.ldy + n0, nb, ebar);
                }
                {
                    float* trow = T + (quad * 32 + lane) * PITCH + chalf * HALF;
#pragma unroll 1
                    for (int c0 = 0; c0 < HALF; c0 += 32) {         // 32 columns at a time keeps the register count down
                        float v[32];
                        t3_tmem_ld16_nowait(taddr + (uint32_t)c0, v);
                        t3_tmem_ld16_nowait(taddr + (uint32_t)(c0 + 16), v + 16);
                        if (two_halves) {
                            float u[32];
                            t3_tmem_ld16_nowait(taddr + (uint32_t)(hoff + c0), u);
                            t3_tmem_ld16_nowait(taddr + (uint32_t)(hoff + c0 + 16), u + 16);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int c = 0; c < 32; ++c) v[c] += u[c];
                        } else {
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        }
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + chalf * HALF + c0 + c);
                            *reinterpret_cast<float4*>(trow + c0 + c) = make_float4(post * fmaf(v[c], p.acc_scale, b4.x), post * fmaf(v[c + 1], p.acc_scale, b4.y),
                                                                                    post * fmaf(v[c + 2], p.acc_scale, b4.z), post * fmaf(v[c + 3], p.acc_scale, b4.w));
                        }
                    }
                }
                if (warp == 2 && lane == 0) T3_CTA(8);
                t3_fence_before();
                t3_mbar_wait(ebar, 0u, err);                    // operand rows have landed (async proxy writes, made visible by the barrier)
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (warp == 2 && lane == 0) T3_CTA(9);
#pragma unroll 1
                for (int rb = 0; rb < T3_BM; rb += RPB) {
                    float4 val[EPI_BATCH], r4[EPI_BATCH], a4[EPI_BATCH], y4[EPI_BATCH];
#pragma unroll
                    for (int u = 0; u < EPI_BATCH; ++u) {
                        const int idx = etid + u * 256;
                        const int row = rb + idx / C4, c = (idx % C4) * 4;
                        val[u] = *reinterpret_cast<const float4*>(T + row * PITCH + c);
                        r4[u] = hasR ? *reinterpret_cast<const float4*>(Rg + row * BN + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                        a4[u] = hasA ? *reinterpret_cast<const float4*>(Ag + row * BN + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                        y4[u] = hasY ? *reinterpret_cast<const float4*>(Yg + row * BN + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");       // every T row of this batch has been read: its bytes now hold the fp16 planes
                    uint8_t* o16h = reinterpret_cast<uint8_t*>(T + rb * PITCH);
                    uint8_t* o16l = o16h + RPB * BN * 2;
#pragma unroll
                    for (int u = 0; u < EPI_BATCH; ++u) {
                        const int idx = etid + u * 256;
                        const int lrow = idx / C4, c = (idx % C4) * 4;
                        const int row = rb + lrow;
                        float o[4] = {val[u].x + r4[u].x, val[u].y + r4[u].y, val[u].z + r4[u].z, val[u].w + r4[u].w};
                        float h[4];
                        if (is_zr) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) o[j] = apply_act(o[j], BFLOW_ACT_SIGMOID);
                            h[0] = o[0] * a4[u].x; h[1] = o[1] * a4[u].y; h[2] = o[2] * a4[u].z; h[3] = o[3] * a4[u].w;
                        } else if (is_q) {
                            const float zz[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w}, hh[4] = {y4[u].x, y4[u].y, y4[u].z, y4[u].w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                o[j] = (1.f - zz[j]) * hh[j] + zz[j] * fast_tanh(o[j]);
                                h[j] = o[j];
                            }
                        } else {
                            // standard epilogue without residual (host contract): act2 follows act1 directly
                            if (slow1 || slow2) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) o[j] = apply_act(apply_act(o[j], d.act1), d.act2);      // r4 is zero here
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j) o[j] = fmaxf(fmaxf(o[j], lo1), lo2);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) h[j] = o[j];
                        }
                        if (out32) *reinterpret_cast<float4*>(Og + row * BN + c) = make_float4(o[0], o[1], o[2], o[3]);
                        if (out16) {
                            uint2 hv, lv;
                            split2(h[0], h[1], hv.x, lv.x);
                            split2(h[2], h[3], hv.y, lv.y);
                            *reinterpret_cast<uint2*>(o16h + (lrow * BN + c) * 2) = hv;
                            *reinterpret_cast<uint2*>(o16l + (lrow * BN + c) * 2) = lv;
                        }
                    }
                    if (warp == 2 && lane == 0 && rb / RPB < 5) T3_CTA(10 + rb / RPB);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the bulk-copy engine
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (etid < T3_BM && row_ok) {
                    const size_t mm = (size_t)row_mm;
                    if (out32) t3_bulk_s2g(d.y + mm * d.ldy + n0, t3_smem_u32(Og + etid * BN), (uint32_t)ncols * 4u);
                    if (out16) {
                        const int b = etid / RPB, rr = etid - b * RPB;
                        const uint32_t sh = t3_smem_u32(T + b * RPB * PITCH) + (uint32_t)(rr * BN * 2);
                        __half* gh = reinterpret_cast<__half*>(is_zr ? d.aux1_16_hi : d.y16_hi);
                        __half* gl = reinterpret_cast<__half*>(is_zr ? d.aux1_16_lo : d.y16_lo);
                        const size_t off = mm * (size_t)(is_zr ? d.ld_aux1_16 : d.ldy16) + (is_zr ? n0 - Cg : n0);
                        const int nc8 = ncols & ~7;             // bulk pieces are 16-byte multiples; a ragged tail (Cout % 8 = 4) goes out as plain stores
                        if (nc8 > 0) {
                            t3_bulk_s2g(gh + off, sh, (uint32_t)nc8 * 2u);
                            t3_bulk_s2g(gl + off, sh + (uint32_t)(RPB * BN * 2), (uint32_t)nc8 * 2u);
                        }
                        if (nc8 < ncols) {
                            const uint8_t* th = reinterpret_cast<const uint8_t*>(T + b * RPB * PITCH) + rr * BN * 2;
                            *reinterpret_cast<uint2*>(gh + off + nc8) = *reinterpret_cast<const uint2*>(th + nc8 * 2);
                            *reinterpret_cast<uint2*>(gl + off + nc8) = *reinterpret_cast<const uint2*>(th + RPB * BN * 2 + nc8 * 2);
                        }
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                }
                if (lane == 0) t3_mbar_arrive(tempty_bar(acc));
            } else if (p.staged) {
                // Single-tile CTA: the pipeline stages are idle once tmem_full has fired (every TMA load was consumed, every MMA retired), so the
                // fp32 tile is parked there, [128][BN + 4].  Then the 256 epilogue threads walk it row-major, 4 channels per thread: a warp
                // touches 2-4 whole rows per instruction (full sectors) instead of 32 rows x 16 bytes, and the residual / gate operands of
                // EPI_BATCH groups are loaded together before the first store (one L2 round trip per batch, not per group).
                constexpr int PITCH = BN + 4;
                constexpr int C4 = BN / 4;
                constexpr int EPI_BATCH = 4;
                float* T = reinterpret_cast<float*>(smem_raw + (smem_base - t3_smem_u32(smem_raw)));
                {
                    float* trow = T + (quad * 32 + lane) * PITCH + chalf * HALF;
#pragma unroll 1
                    for (int c0 = 0; c0 < HALF; c0 += 32) {         // 32 columns at a time keeps the register count down
                        float v[32];
                        t3_tmem_ld16_nowait(taddr + (uint32_t)c0, v);
                        t3_tmem_ld16_nowait(taddr + (uint32_t)(c0 + 16), v + 16);
                        if (two_halves) {
                            float u[32];
                            t3_tmem_ld16_nowait(taddr + (uint32_t)(hoff + c0), u);
                            t3_tmem_ld16_nowait(taddr + (uint32_t)(hoff + c0 + 16), u + 16);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int c = 0; c < 32; ++c) v[c] += u[c];
                        } else {
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        }
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + chalf * HALF + c0 + c);
                            *reinterpret_cast<float4*>(trow + c0 + c) = make_float4(post * fmaf(v[c], p.acc_scale, b4.x), post * fmaf(v[c + 1], p.acc_scale, b4.y),
                                                                                    post * fmaf(v[c + 2], p.acc_scale, b4.z), post * fmaf(v[c + 3], p.acc_scale, b4.w));
                        }
                    }
                }
                if (warp == 2 && lane == 0) T3_CTA(8);
                t3_fence_before();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (warp == 2 && lane == 0) T3_CTA(9);
                const int mt0 = m_tile * T3_BM;
#pragma unroll 1
                for (int base = etid; base < T3_BM * C4; base += 256 * EPI_BATCH) {
                    EpiPre pre[EPI_BATCH];
                    float4 val[EPI_BATCH];
                    bool ok[EPI_BATCH], vec[EPI_BATCH];
                    int mrow[EPI_BATCH];
#pragma unroll
                    for (int u = 0; u < EPI_BATCH; ++u) {
                        const int idx = base + u * 256;
                        const int row = idx / C4, c = (idx - row * C4) * 4;
                        bool rok = false;
                        const int mm = idx < T3_BM * C4 ? row_m(row, rok) : 0, nn = n0 + c;
                        mrow[u] = mm;
                        ok[u] = idx < T3_BM * C4 && rok && nn < d.Cout;
                        vec[u] = aligned && nn + 3 < d.Cout;
                        if (ok[u]) {
                            val[u] = *reinterpret_cast<const float4*>(T + row * PITCH + c);
                            conv_epilogue4_prefetch(d, mm, nn, vec[u], pre[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < EPI_BATCH; ++u) {
                        const int idx = base + u * 256;
                        const int row = idx / C4, c = (idx - row * C4) * 4;
                        if (ok[u]) {
                            float t4[4] = {val[u].x, val[u].y, val[u].z, val[u].w};
                            conv_epilogue4_finish(d, mrow[u], n0 + c, t4, vec[u], pre[u]);
                        }
                    }
                    if (warp == 2 && lane == 0 && base < 256 * EPI_BATCH * 5) T3_CTA(10 + base / (256 * EPI_BATCH));
                }
                if (lane == 0) t3_mbar_arrive(tempty_bar(acc));
            } else if (nb0 < d.Cout) {                      // warp-uniform
                float v[HALF];
#pragma unroll
                for (int c = 0; c < HALF; c += 16) t3_tmem_ld16_nowait(taddr + (uint32_t)c, v + c);
                if (two_halves) {
                    float u[HALF];
#pragma unroll
                    for (int c = 0; c < HALF; c += 16) t3_tmem_ld16_nowait(taddr + (uint32_t)(hoff + c), u + c);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int c = 0; c < HALF; ++c) v[c] += u[c];
                } else {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                }
                if (warp == 2 && lane == 0) T3_CTA(8);
                // the accumulator is in registers: hand the TMEM buffer back before the (slow) global stores
                t3_fence_before();
                __syncwarp();
                if (lane == 0) t3_mbar_arrive(tempty_bar(acc));
                if (warp == 2 && lane == 0) T3_TRACE(4, lt);
                if (d.stats != nullptr) {
                    // butterfly transpose-reduce over the warp's 32 rows, 16 columns at a time; lanes with the low bit clear then own one
                    // column's (sum, sum of squares) and add it to the CTA's shared accumulators
                    const int m_first = p.slab ? t_n * shw : m_tile * T3_BM + quad * 32;      // slab mode: any pixel of image t_n
                    const int m_last = min(m_first + 31, p.M - 1);
                    const bool uniform = p.slab ? true : (m_first < p.M && (m_first / shw) == (m_last / shw));       // warp-uniform
#pragma unroll
                    for (int c = 0; c < HALF; c += 16) {
                        const int nb = nb0 + c;
                        if (nb + 15 >= d.Cout) break;
                        float sv[16], sq[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float b = s_bias[chalf * HALF + c + j];
                            const float x = m_ok ? post * fmaf(v[c + j], p.acc_scale, b) : 0.f;
                            sv[j] = x;
                            sq[j] = x * x;
                        }
                        if (uniform) {
#pragma unroll
                            for (int width = 8, bit = 16; width >= 1; width >>= 1, bit >>= 1) {
                                const bool upper = (lane & bit) != 0;
#pragma unroll
                                for (int i = 0; i < width; ++i) {
                                    const float keep_s = upper ? sv[width + i] : sv[i], send_s = upper ? sv[i] : sv[width + i];
                                    const float keep_q = upper ? sq[width + i] : sq[i], send_q = upper ? sq[i] : sq[width + i];
                                    sv[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, bit);
                                    sq[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, bit);
                                }
                            }
                            const float ts = sv[0] + __shfl_xor_sync(0xffffffffu, sv[0], 1);
                            const float tq = sq[0] + __shfl_xor_sync(0xffffffffu, sq[0], 1);
                            if ((lane & 1) == 0) {
                                const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                                if (m_first / shw == cur_img) {
                                    atomicAdd(s_stat + chalf * HALF + c + col, ts);
                                    atomicAdd(s_stat + BN + chalf * HALF + c + col, tq);
                                } else {             // a tile that straddles two images: the minority warps go straight to the table
                                    double* st = d.stats + ((size_t)(m_first / shw) * d.Cout + nb + col) * 2;
                                    atomicAdd(st, (double)ts);
                                    atomicAdd(st + 1, (double)tq);
                                }
                            }
                        } else if (m_ok) {
                            double* st = d.stats + ((size_t)(m / shw) * d.Cout + nb) * 2;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                atomicAdd(st + 2 * j, (double)sv[j]);
                                atomicAdd(st + 2 * j + 1, (double)sq[j]);
                            }
                        }
                    }
                }
                if (m_ok) {
#pragma unroll
                    for (int c = 0; c < HALF; c += 16) {
                        const int nb = nb0 + c;
                        if (nb >= d.Cout) break;
                        float* w = v + c;
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + chalf * HALF + c + j);
                            w[j] = post * fmaf(w[j], p.acc_scale, b4.x); w[j + 1] = post * fmaf(w[j + 1], p.acc_scale, b4.y);
                            w[j + 2] = post * fmaf(w[j + 2], p.acc_scale, b4.z); w[j + 3] = post * fmaf(w[j + 3], p.acc_scale, b4.w);
                        }
                        if (d.epi == BFLOW_EPI_STD && aligned && nb + 15 < d.Cout) {
                            if (slow1) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) w[j] = apply_act(w[j], d.act1);
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j) w[j] = fmaxf(w[j], lo1);
                            }
                            if (d.res != nullptr) {
                                const float* rrow = d.res + (size_t)m * d.ldr + nb;
#pragma unroll
                                for (int j = 0; j < 16; j += 4) {
                                    const float4 r4 = *reinterpret_cast<const float4*>(rrow + j);
                                    w[j] += r4.x; w[j + 1] += r4.y; w[j + 2] += r4.z; w[j + 3] += r4.w;
                                }
                            } else if (d.res16_hi != nullptr) {
#pragma unroll
                                for (int j = 0; j < 16; j += 4) {
                                    const float4 r4 = load_res16_4(d, (size_t)m * d.ldr16 + nb + j);
                                    w[j] += r4.x; w[j + 1] += r4.y; w[j + 2] += r4.z; w[j + 3] += r4.w;
                                }
                            }
                            if (slow2) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) w[j] = apply_act(w[j], d.act2);
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j) w[j] = fmaxf(w[j], lo2);
                            }
                            if (d.y != nullptr) {
                                float* yrow = d.y + (size_t)m * d.ldy + nb;
#pragma unroll
                                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(yrow + j) = make_float4(w[j], w[j + 1], w[j + 2], w[j + 3]);
                            }
                            if (d.y16_hi != nullptr) {
                                if (wide16) {            // 16-byte stores: half as many store requests per row
#pragma unroll
                                    for (int j = 0; j < 16; j += 8) {
                                        uint4 h4, l4;
                                        split2(w[j], w[j + 1], h4.x, l4.x);
                                        split2(w[j + 2], w[j + 3], h4.y, l4.y);
                                        split2(w[j + 4], w[j + 5], h4.z, l4.z);
                                        split2(w[j + 6], w[j + 7], h4.w, l4.w);
                                        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(d.y16_hi) + (size_t)m * d.ldy16 + nb + j) = h4;
                                        if (wlo) *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(d.y16_lo) + (size_t)m * d.ldy16 + nb + j) = l4;
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 16; j += 4) store_split4(d.y16_hi, d.y16_lo, (size_t)m * d.ldy16 + nb + j, w[j], w[j + 1], w[j + 2], w[j + 3]);
                                }
                            }
                        } else {
                            // GRU gate epilogues and ragged / unaligned tails
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                float t4[4] = {w[j], w[j + 1], w[j + 2], w[j + 3]};
                                conv_epilogue4(d, m, nb + j, t4, aligned && (nb + j + 3 < d.Cout));
                            }
                        }
                    }
                }
            } else {
                t3_fence_before();
                __syncwarp();
                if (lane == 0) t3_mbar_arrive(tempty_bar(acc));
            }
            if (warp == 2 && lane == 0) T3_TRACE(5, lt);
            if (warp == 2 && lane == 0) T3_CTA(6);
        }
        if (d.stats != nullptr) flush_stats(cur_img, stat_n0);
        if (p.staged == 3) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
    __syncthreads();
    if (tid == 0) T3_TRACE(4, 255);              // CTA end
    if (tid == 0) T3_CTA(7);
    tl_end(p.tl);
    if (warp == 1) {
        t3_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Slab variant for the layer that dominates the encoders: 3x3 / stride 1 / pad 1, 64 -> 64 channels (extractor.py layer1, 59 % of the
// encoder flops at 240x320).  Measured (tools/timeline.py --cta-label): the im2col kernel above moves 48 KB per k-block through L2 -> SM
// for 384 tensor-pipe cycles of work -- 125 B/clk per SM against the ~42 B/clk every SM gets when all 148 pull -- so it runs at 30 % of
// the tensor pipe.  Two changes cut the traffic 4x:
//   * the whole weight image (9 taps x [hi|lo] x 64 x 64 fp16 = 144 KB) is loaded ONCE per CTA and stays in shared memory;
//   * an output tile is an 8 (x) by 16 (y) patch of pixels, and for each filter column kw ONE halo slab of 8 x 18 pixels x 64 channels
//     (TMA tiled mode, out-of-image rows / columns zero-filled = the convolution's padding) serves the three filter rows: row r of the
//     slab is pixel (r / 8, r % 8), so filter row kh is the same slab advanced by 8 rows = 1024 bytes, which keeps the UMMA descriptor
//     on a swizzle-atom boundary.  3 slabs x 36 KB per tile instead of 9 x 32 KB of im2col rows plus 9 x 16 KB of weights.
// Roles as above: warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (thread = pixel row, half of the 64 channels), two TMEM accumulators.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SL_BN = 64;
constexpr int SL_B_BYTES = 9 * 2 * SL_BN * 128;          // resident weights: 147456
constexpr int SL_PLANE = 8 * 18 * 128;                   // one fp16 slab plane: 18432
constexpr int SL_STAGE = SL_PLANE;                       // one plane (hi or lo) of one slab
constexpr int SL_STAGES = 4;

struct SlabParams {
    int tiles_x, tiles_y, n_tiles, f16;
    float acc_scale;
    unsigned long long* tl;
};

__global__ void __launch_bounds__(T3_THREADS, 1)
conv_slab64_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const bflow_conv_desc d,
                   const uint8_t* __restrict__ wtc, const SlabParams p, int* err) {
    constexpr int ACC_COLS = 2 * SL_BN;            // [hi*hi + lo*hi | hi*lo]
    constexpr int TMEM_COLS = 2 * ACC_COLS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (t3_smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bres = smem_base;
    const uint32_t stage0 = smem_base + SL_B_BYTES;
    const uint32_t bars = stage0 + SL_STAGES * SL_STAGE;
    auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bars + 32u + 8u * (uint32_t)s; };
    auto tfull_bar = [&](int a) { return bars + 64u + 8u * (uint32_t)a; };
    auto tempty_bar = [&](int a) { return bars + 80u + 8u * (uint32_t)a; };
    const uint32_t tmem_slot = bars + 96u;
    const uint32_t wbar = bars + 104u;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;      // warp-uniform role index (see conv_tc3_kernel)
    tl_begin(p.tl);
    if (tid == 0) {
        for (int s = 0; s < SL_STAGES; ++s) {
            t3_mbar_init(full_bar(s), 1);
            t3_mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            t3_mbar_init(tfull_bar(a), 1);
            t3_mbar_init(tempty_bar(a), T3_EPI_WARPS);
        }
        t3_mbar_init(wbar, 1);
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
        if (t3_elect_one()) {
            t3_mbar_arrive_expect_tx(wbar, SL_B_BYTES);
            for (int t = 0; t < 9; ++t) t3_bulk_g2s(bres + (uint32_t)t * (2 * SL_BN * 128), wtc + (size_t)t * (2 * SL_BN * 128), 2 * SL_BN * 128, wbar);
        }
        __syncwarp();
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int n = tile / tiles_per_img, r = tile - n * tiles_per_img;
            const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
            const int x0 = tx * 8, y0 = ty * 16;
            // one stage = ONE plane of a slab (18 KB).  Measured with the epilogue, the MMAs and the loads switched off in turn (128 SMs, 23.4 tiles per
            // CTA): loads alone 40 us (63 GB/s per SM), epilogue alone 52 us, MMAs alone 84 us (~100 cycles per tcgen05.mma whatever its N: the kernel
            // is bound by MMA issue), everything 94 us.  Four small stages instead of two 36 KB ones keep three loads in flight: 99 -> 94 us.
            const int npl = p.f16 ? 1 : 2;
            for (int kw = 0; kw < 3; ++kw) {
                for (int pl = 0; pl < npl; ++pl, ++it) {
                    const int s = (int)(it % SL_STAGES);
                    const uint32_t ph = (it / SL_STAGES) & 1u;
                    t3_mbar_wait(empty_bar(s), ph ^ 1u, err);
                    const uint32_t st = stage0 + (uint32_t)s * SL_STAGE;
                    if (t3_elect_one()) {
                        t3_mbar_arrive_expect_tx(full_bar(s), SL_PLANE);
                        sl_tma_tile(st, pl == 0 ? &map_hi : &map_lo, 0, x0 + kw - 1, y0 - 1, n, full_bar(s));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(SL_BN >> 3) << 17) | ((uint32_t)(T3_BM >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * SL_BN) >> 3) << 17) | ((uint32_t)(T3_BM >> 4) << 24);
        t3_mbar_wait(wbar, 0u, err);
        uint32_t it = 0, lt = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++lt) {
            const uint32_t acc = lt & 1u, aph = (lt >> 1) & 1u;
            t3_mbar_wait(tempty_bar(acc), aph ^ 1u, err);
            t3_fence_after();
            const uint32_t tacc = tmem_base + acc * ACC_COLS;
            const int npl = p.f16 ? 1 : 2;
            for (int kw = 0; kw < 3; ++kw) {
                for (int pl = 0; pl < npl; ++pl, ++it) {
                    const int s = (int)(it % SL_STAGES);
                    const uint32_t ph = (it / SL_STAGES) & 1u;
                    t3_mbar_wait(full_bar(s), ph, err);
                    t3_fence_after();
                    if (t3_elect_one()) {
                        const uint32_t a = stage0 + (uint32_t)s * SL_STAGE;
                        // hi plane: A_hi x [B_hi | B_lo] (N = 128, or N = 64 in f16 mode); lo plane: A_lo x B_hi (N = 64)
                        const uint32_t id = (pl == 0 && !p.f16) ? idesc2 : idesc;
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            const uint32_t b = bres + (uint32_t)(kh * 3 + kw) * (2 * SL_BN * 128);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t ko = (uint32_t)k * 32u;
                                t3_umma(tacc, t3_umma_desc(a + (uint32_t)kh * 1024u + ko), t3_umma_desc(b + ko), id, (kw > 0 || pl > 0 || kh > 0 || k > 0) ? 1u : 0u);
                            }
                        }
                        t3_commit(empty_bar(s));
                        if (kw == 2 && pl == npl - 1) t3_commit(tfull_bar(acc));
                    }
                    __syncwarp();
                }
            }
        }
        t3_fence_before();
    } else {
        const int quad = warp & 3, chalf = (warp - 2) >> 2, etid = tid - 64;
        float* s_bias = reinterpret_cast<float*>(smem_raw + (bars - t3_smem_u32(smem_raw)) + 128);     // [64]
        float* s_stat = s_bias + SL_BN;                                                                    // [2][64]
        for (int j = etid; j < SL_BN; j += 256) {
            s_bias[j] = d.bias != nullptr ? __ldg(d.bias + j) : 0.f;
            s_stat[j] = 0.f;
            s_stat[SL_BN + j] = 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float lo1 = d.act1 == BFLOW_ACT_RELU ? 0.f : -INFINITY, lo2 = d.act2 == BFLOW_ACT_RELU ? 0.f : -INFINITY;
        const float post = d.scale;
        int cur_img = -1;
        auto flush_stats = [&](int img) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int j = etid; j < SL_BN; j += 256) {
                if (img >= 0) {
                    atomicAdd(d.stats + ((size_t)img * d.Cout + j) * 2, (double)s_stat[j]);
                    atomicAdd(d.stats + ((size_t)img * d.Cout + j) * 2 + 1, (double)s_stat[SL_BN + j]);
                }
                s_stat[j] = 0.f;
                s_stat[SL_BN + j] = 0.f;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        };
        const int nb0 = chalf * 32;
        const int b4 = (lane & 7) * 4;                 // after the in-warp transpose a lane owns channels nb0 + b4 .. + 3
        const float bias4[4] = {s_bias[nb0 + b4], s_bias[nb0 + b4 + 1], s_bias[nb0 + b4 + 2], s_bias[nb0 + b4 + 3]};
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++lt) {
            const int n = tile / tiles_per_img, r = tile - n * tiles_per_img;
            const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
            if (d.stats != nullptr && n != cur_img) {
                flush_stats(cur_img);
                cur_img = n;
            }
            // after the in-warp transpose below, lane (g, b) owns channels nb0 + 4b .. + 3 of the 8 pixels of image row ty*16 + quad*4 + g
            const int g = lane >> 3;
            const int y = ty * 16 + quad * 4 + g;
            const bool valid = y < d.H;
            const size_t m0 = ((size_t)n * d.H + y) * d.W + tx * 8;
            // the residual rows are requested BEFORE the wait for the accumulator: their latency (several microseconds when the other
            // stream's InstanceNorm pass saturates HBM) then hides behind this tile's MMAs instead of stretching the epilogue
            // (raw 8-byte loads, converted after the wait: written as load_res16_4 calls the compiler kept one branch per row and the 16
            // loads went out one after the other -- 4.3 us per tile, twice the tile's MMA time)
            uint2 rh[8], rl[8];
            {
                const bool has_res = valid && d.res16_hi != nullptr, has_lo = has_res && d.precision != BFLOW_PREC_F16;
                const __half* rhp = reinterpret_cast<const __half*>(d.res16_hi) + m0 * d.ldr16 + nb0 + b4;
                const __half* rlp = reinterpret_cast<const __half*>(d.res16_lo) + m0 * d.ldr16 + nb0 + b4;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    rh[c] = has_res ? *reinterpret_cast<const uint2*>(rhp + (size_t)c * d.ldr16) : make_uint2(0u, 0u);
                    rl[c] = has_lo ? *reinterpret_cast<const uint2*>(rlp + (size_t)c * d.ldr16) : make_uint2(0u, 0u);
                }
            }
            const uint32_t acc = lt & 1u, aph = (lt >> 1) & 1u;
            t3_mbar_wait(tfull_bar(acc), aph, err);
            t3_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * ACC_COLS + (uint32_t)nb0;
            float v[32];
            {
                float u[32];
                t3_tmem_ld16_nowait(taddr, v);
                t3_tmem_ld16_nowait(taddr + 16u, v + 16);
                if (!p.f16) {
                    t3_tmem_ld16_nowait(taddr + (uint32_t)SL_BN, u);
                    t3_tmem_ld16_nowait(taddr + (uint32_t)SL_BN + 16u, u + 16);
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) u[c] = 0.f;
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                t3_fence_before();
                __syncwarp();
                if (lane == 0) t3_mbar_arrive(tempty_bar(acc));
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] += u[c];
            }
            // TMEM hands a thread one tile row (= pixel) and 32 consecutive channels.  8 x 8 transpose inside every group of 8 lanes (3
            // butterfly steps over the 16-byte chunks): afterwards lane (g, b) holds channels nb0 + 4b .. 4b + 3 of the 8 pixels of image row
            // ty*16 + quad*4 + g, so one store instruction writes whole 128-byte rows instead of touching 32 lines, and the column sums of
            // the InstanceNorm statistics need 2 shuffle steps instead of 5 (measured on the stem kernel: stores 33 -> 11 us of 109)
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
#pragma unroll
            for (int c = 0; c < 8; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) v[4 * c + j] = post * fmaf(v[4 * c + j], p.acc_scale, bias4[j]);
            if (d.stats != nullptr) {
                float s[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
                if (valid) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            s[j] += v[4 * c + j];
                            sq[j] = fmaf(v[4 * c + j], v[4 * c + j], sq[j]);
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
                        atomicAdd(s_stat + SL_BN + nb0 + b4 + j, sq[j]);
                    }
                }
            }
            if (valid) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = fmaxf(v[4 * c + j], lo1);
                    {
                        const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&rh[c].x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&rh[c].y));
                        const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&rl[c].x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&rl[c].y));
                        o[0] += h0.x + l0.x; o[1] += h0.y + l0.y; o[2] += h1.x + l1.x; o[3] += h1.y + l1.y;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], lo2);
                    if (d.y != nullptr) *reinterpret_cast<float4*>(d.y + (m0 + c) * d.ldy + nb0 + b4) = make_float4(o[0], o[1], o[2], o[3]);
                    if (d.y16_hi != nullptr) {
                        uint2 h2, l2;
                        split2(o[0], o[1], h2.x, l2.x);
                        split2(o[2], o[3], h2.y, l2.y);
                        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(d.y16_hi) + (m0 + c) * d.ldy16 + nb0 + b4) = h2;
                        if (!p.f16) *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(d.y16_lo) + (m0 + c) * d.ldy16 + nb0 + b4) = l2;
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

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeIm2colFn get_encode_im2col() {
    static EncodeIm2colFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeIm2colFn>(p);
    }
    return fn;
}

static int g_tc3_debug = 0;
static long long* g_tc3_trace = nullptr;
static unsigned long long* g_tc3_cta = nullptr;
static int g_tc3_cta_nth = -1, g_tc3_cta_span = 1, g_tc3_cta_count = 0;

template <int BN, int STAGES, bool F16>
static int launch_tc3(const CUtensorMap* maps, const bflow_conv_desc& d, const void* wtc, const T3Params& p, int* err, cudaStream_t stream) {      // maps: 4 input + 3 output
    constexpr int smem = t3_area_bytes(BN, STAGES) + 128 + 3 * BN * 4 + 64 + 1024;
    static PerDeviceFlag configured;
    if (!configured.get()) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc3_kernel<BN, STAGES, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error(cudaGetErrorString(e));
            return BFLOW_ERR_CUDA;
        }
        configured.set();
    }
    const int n_tiles = p.n_mtiles * p.n_ntiles;
    const int grid = n_tiles < grid_cap(d) ? n_tiles : grid_cap(d);
    // a CTA budget means another chain owns the other SMs: no early launch then (an early-launched CTA holds a whole SM while it waits)
    cudaError_t le = launch_pdl_if(d.max_ctas <= 0, conv_tc3_kernel<BN, STAGES, F16>, dim3(grid), dim3(T3_THREADS), smem, stream, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5],
                                maps[6], maps[7], maps[8], d, reinterpret_cast<const uint8_t*>(wtc), p, err);
    if (le != cudaSuccess) {
        set_error(cudaGetErrorString(le));
        return BFLOW_ERR_CUDA;
    }
    return check_launch("bflow_conv2d_nhwc_tc3");
}

__global__ void split_f16_kernel(const float* __restrict__ src, int ld, __half* __restrict__ hi, __half* __restrict__ lo, int ld16, long long rows, int C4) {
    const long long total = rows * C4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / C4;
        const int c = (int)(idx - r * C4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + r * ld + c);
        store_split4(hi, lo, (size_t)r * ld16 + c, v.x, v.y, v.z, v.w);
    }
}

// CTA = 32 consecutive patch rows.  Phase 1 gathers column k for the 32 rows with one warp-wide load (consecutive output pixels read
// a stride-`stride` run of the same input row: two 128-byte lines) into shared memory; phase 2 turns every 8 consecutive k of a
// row into one 16-byte hi and one 16-byte lo store (a full 2*ld16-byte row per 32 threads: coalesced).
constexpr int I2C_ROWS = 32;
__global__ void __launch_bounds__(256) im2col_split16_kernel(const float* __restrict__ src, int C_total, int c_off, int cin, int H, int W, int Ho, int Wo,
                                                             int KH, int KW, int stride, int pad_h, int pad_w, float scale, float shift,
                                                             __half* __restrict__ hi, __half* __restrict__ lo, int ld16, long long rows, unsigned long long* tl) {
    extern __shared__ float patch[];
    tl_begin(tl);              // [I2C_ROWS][ld16 + 1] floats, then the tap table: int off[ld16], int khkw[ld16]
    const int pitch = ld16 + 1;
    const int K = KH * KW * cin;
    int* t_off = reinterpret_cast<int*>(patch + I2C_ROWS * pitch);
    int* t_khkw = t_off + ld16;
    for (int k = threadIdx.x; k < ld16; k += blockDim.x) {
        int off = 0, code = -1;
        if (k < K) {
            const int tap = k / cin, c = k - tap * cin;
            const int kh = tap / KW, kw = tap - kh * KW;
            off = c * (H * W) + kh * W + kw;
            code = (kh << 8) | kw;
        }
        t_off[k] = off;
        t_khkw[k] = code;
    }
    __syncthreads();
    const size_t HW = (size_t)H * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row0 = (long long)blockIdx.x * I2C_ROWS;
    const long long row = row0 + lane;
    const bool row_ok = row < rows;
    int ih0 = 0, iw0 = 0;
    const float* base = src;
    if (row_ok) {
        const int ow = (int)(row % Wo);
        const long long t = row / Wo;
        const int oh = (int)(t % Ho);
        const int n = (int)(t / Ho);
        ih0 = oh * stride - pad_h;
        iw0 = ow * stride - pad_w;
        base = src + ((size_t)n * C_total + c_off) * HW;
    }
    const float* p00 = base + (long long)ih0 * W + iw0;      // tap (0,0), channel 0 (may point outside: only dereferenced when in bounds)
    for (int k = warp; k < ld16; k += 8) {        // warp-uniform k
        float x = 0.f;
        const int code = t_khkw[k];
        if (code >= 0 && row_ok) {
            const int ih = ih0 + (code >> 8), iw = iw0 + (code & 255);
            if (ih >= 0 && ih < H && iw >= 0 && iw < W) x = fmaf(__ldg(p00 + t_off[k]), scale, shift);
        }
        patch[lane * pitch + k] = x;
    }
    __syncthreads();
    const int chunks = ld16 >> 3;
    for (int idx = threadIdx.x; idx < I2C_ROWS * chunks; idx += 256) {
        const int r = idx / chunks, ck = idx - r * chunks;
        if (row0 + r >= rows) continue;
        const float* p = patch + r * pitch + ck * 8;
        uint4 h4, l4;
        split2(p[0], p[1], h4.x, l4.x);
        split2(p[2], p[3], h4.y, l4.y);
        split2(p[4], p[5], h4.z, l4.z);
        split2(p[6], p[7], h4.w, l4.w);
        *reinterpret_cast<uint4*>(hi + (size_t)(row0 + r) * ld16 + ck * 8) = h4;
        *reinterpret_cast<uint4*>(lo + (size_t)(row0 + r) * ld16 + ck * 8) = l4;
    }
    tl_end(tl);
}

}  // namespace bflow

extern "C" int bflow_im2col_split16(const float* src, int C_total, int c_off, int cin, int N, int H, int W, int KH, int KW, int stride, int pad_h,
                                    int pad_w, float scale, float shift, void* out_hi, void* out_lo, int ld16, void* stream) {
    BFLOW_REQUIRE(src != nullptr && out_hi != nullptr && out_lo != nullptr, "im2col_split16: null tensor");
    BFLOW_REQUIRE(N > 0 && H > 0 && W > 0 && cin > 0 && c_off >= 0 && c_off + cin <= C_total, "im2col_split16: bad source window");
    BFLOW_REQUIRE(KH > 0 && KW > 0 && stride > 0 && pad_h >= 0 && pad_w >= 0 && ld16 % 8 == 0 && ld16 >= KH * KW * cin, "im2col_split16: bad window / ld16");
    BFLOW_REQUIRE((reinterpret_cast<uintptr_t>(out_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_lo) & 15) == 0, "im2col_split16: alignment");
    const int Ho = (H + 2 * pad_h - KH) / stride + 1, Wo = (W + 2 * pad_w - KW) / stride + 1;
    const long long rows = (long long)N * Ho * Wo;
    BFLOW_REQUIRE(ld16 <= 1024, "im2col_split16: ld16 > 1024");
    const long long g = (rows + bflow::I2C_ROWS - 1) / bflow::I2C_ROWS;
    BFLOW_REQUIRE((long long)cin * H * W < (1ll << 31), "im2col_split16: source too large");
    const size_t smem = (size_t)bflow::I2C_ROWS * (ld16 + 1) * sizeof(float) + (size_t)ld16 * 2 * sizeof(int);
    static bflow::PerDeviceFlag configured;
    if (!configured.get()) {
        cudaFuncSetAttribute(bflow::im2col_split16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1025 * 4 + 1024 * 8);
        configured.set();
    }
    bflow::im2col_split16_kernel<<<(unsigned)g, 256, smem, (cudaStream_t)stream>>>(src, C_total, c_off, cin, H, W, Ho, Wo, KH, KW, stride, pad_h, pad_w, scale, shift,
                                                                                reinterpret_cast<__half*>(out_hi), reinterpret_cast<__half*>(out_lo), ld16, rows,
                                                                                bflow::timeline_next_slot("im2col_split16"));
    return bflow::check_launch("bflow_im2col_split16");
}

extern "C" void bflow_tc3_trace(long long* device_buf_6x256) { bflow::g_tc3_trace = device_buf_6x256; }

// development: the nth tc3 launch from now on writes per-CTA stamps {start, prologue done, first TMA issued, first stage full, last commit,
// accumulator ready, epilogue done, end} (globaltimer ns) into buf[grid][8]
extern "C" void bflow_tc3_cta_trace(void* buf, int nth) {
    bflow::g_tc3_cta = reinterpret_cast<unsigned long long*>(buf);
    bflow::g_tc3_cta_nth = nth;
    bflow::g_tc3_cta_span = 1;
    bflow::g_tc3_cta_count = 0;
}
// the same for `count` consecutive tc3 launches starting at the nth: buf[count][148][16]
extern "C" void bflow_tc3_cta_trace_range(void* buf, int nth, int count) {
    bflow::g_tc3_cta = reinterpret_cast<unsigned long long*>(buf);
    bflow::g_tc3_cta_nth = nth;
    bflow::g_tc3_cta_span = count > 0 ? count : 1;
    bflow::g_tc3_cta_count = 0;
}

extern "C" void bflow_tc3_debug(int flags) { bflow::g_tc3_debug = flags; }

extern "C" int bflow_split_f16(const float* src, int ld, void* hi, void* lo, int ld16, long long rows, int C, void* stream) {
    BFLOW_REQUIRE(src != nullptr && hi != nullptr && lo != nullptr, "split_f16: null tensor");
    BFLOW_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && ld >= C && ld % 4 == 0 && ld16 >= C && ld16 % 4 == 0, "split_f16: bad shape");
    BFLOW_REQUIRE(bflow::aligned16(src) && (reinterpret_cast<uintptr_t>(hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(lo) & 7) == 0, "split_f16: alignment");
    const long long total = rows * (C / 4);
    long long g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    bflow::split_f16_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(src, ld, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), ld16, rows, C / 4);
    return bflow::check_launch("bflow_split_f16");
}

extern "C" int bflow_tma_im2col_map(void* map_out, const void* base, int N, int H, int W, int C, int ld_halves, int KH, int KW, int stride, int pad_h,
                                    int pad_w) {
    BFLOW_REQUIRE(map_out != nullptr && base != nullptr, "tma_im2col_map: null argument");
    BFLOW_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && ld_halves >= C && ld_halves % 8 == 0, "tma_im2col_map: bad shape (ld must be a multiple of 8 halves)");
    BFLOW_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tma_im2col_map: base must be 16-byte aligned");
    BFLOW_REQUIRE(KH > 0 && KW > 0 && stride >= 1 && stride <= 8 && pad_h >= 0 && pad_w >= 0, "tma_im2col_map: bad window");
    bflow::EncodeIm2colFn enc = bflow::get_encode_im2col();
    BFLOW_REQUIRE(enc != nullptr, "tma_im2col_map: cuTensorMapEncodeIm2col not available from the driver");
    alignas(64) CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)ld_halves * 2, (cuuint64_t)W * ld_halves * 2, (cuuint64_t)H * W * ld_halves * 2};
    const int lower[2] = {-pad_w, -pad_h};                                    // {W, H}
    const int upper[2] = {pad_w - (KW - 1), pad_h - (KH - 1)};
    const cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, lower, upper, 64, bflow::T3_BM, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        bflow::set_error("tma_im2col_map: cuTensorMapEncodeIm2col failed");
        return BFLOW_ERR_CUDA;
    }
    // driver workaround carried by CUTLASS (copy_traits_sm90_im2col.hpp): im2col descriptors of tensors smaller than 128 KiB
    int drv = 0;
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && (unsigned long long)N * H * W * ld_halves * 2ull < 131072ull) reinterpret_cast<uint64_t*>(&tm)[1] &= ~(1ull << 21);
    memcpy(map_out, &tm, sizeof(tm));
    return BFLOW_OK;
}

static int tc3_entry(const bflow_conv_desc* dp, const void* maps, const void* omaps, const void* w_tc, int bn, float acc_scale, int slab, int* err,
                     void* stream) {
    BFLOW_CHECK_DESC(dp, bflow_conv_desc, "conv_tc3");
    BFLOW_REQUIRE(maps != nullptr && w_tc != nullptr, "conv_tc3: null argument");
    const bflow_conv_desc& d = *dp;
    // narrow weight tile: bn = 80 / 96 / 112 covering all of Cout -> the 128-column instantiation with bn rows per weight half
    int bn_eff = bn;
    if (bn != 64 && bn != 128 && bn != 256) {
        BFLOW_REQUIRE(bn > 64 && bn < 128 && bn % 16 == 0 && !slab && d.Cout <= bn, "conv_tc3: bn must be 64, 128, 256, or 80 / 96 / 112 with Cout <= bn (im2col mode)");
        bn = 128;
    }
    if (slab) {
        BFLOW_REQUIRE(slab == 1 || slab == 2, "conv_tc3s: orientation must be 1 (taps along y) or 2 (taps along x)");
        BFLOW_REQUIRE(d.stride == 1 && d.pad_h == d.KH / 2 && d.pad_w == d.KW / 2 && (d.KH & 1) && (d.KW & 1), "conv_tc3s: stride 1, odd window, 'same' padding");
        BFLOW_REQUIRE((slab == 1 ? d.KH : d.KW) <= 5 && bn <= 128, "conv_tc3s: at most 5 taps along the slab axis, bn <= 128");
    }
    // channel counts are free (the TMA unit zero-fills channels beyond C); a second source must start on a 64-channel block
    BFLOW_REQUIRE(d.c0 > 0 && d.c1 >= 0 && (d.c1 == 0 || d.c0 % 64 == 0), "conv_tc3: channel counts");
    BFLOW_REQUIRE(d.N > 0 && d.H > 0 && d.W > 0 && d.Cout > 0 && d.KH > 0 && d.KW > 0 && d.stride > 0, "conv_tc3: bad shape");
    BFLOW_REQUIRE(d.Ho == (d.H + 2 * d.pad_h - d.KH) / d.stride + 1 && d.Wo == (d.W + 2 * d.pad_w - d.KW) / d.stride + 1, "conv_tc3: Ho/Wo mismatch");
    BFLOW_REQUIRE((d.y == nullptr || d.ldy >= d.Cout) && (d.res == nullptr || d.ldr >= d.Cout), "conv_tc3: bad output stride");
    BFLOW_REQUIRE((reinterpret_cast<uintptr_t>(w_tc) & 15) == 0 && (reinterpret_cast<uintptr_t>(maps) & 7) == 0, "conv_tc3: alignment");
    if (const char* msg = bflow::check_epilogue(d)) { bflow::set_error(msg); return BFLOW_ERR_INVALID; }
    BFLOW_REQUIRE(d.stats == nullptr || (d.epi == BFLOW_EPI_STD && d.Cout % 16 == 0 && d.res == nullptr && d.res16_hi == nullptr &&
                                         d.act1 == BFLOW_ACT_NONE && d.act2 == BFLOW_ACT_NONE && d.stats_hw >= 0),
                  "conv_tc3: fused statistics need the plain epilogue and Cout % 16 == 0");
    const long long Mll = (long long)d.N * d.Ho * d.Wo;
    BFLOW_REQUIRE(Mll < (1ll << 31), "conv_tc3: too large");
    bflow::T3Params p;
    p.M = (int)Mll;
    p.n_mtiles = (p.M + bflow::T3_BM - 1) / bflow::T3_BM;
    p.slab = slab;
    p.tiles_x = p.tiles_y = p.n_slabs = p.tps = 0;
    if (slab) {
        p.tiles_x = (d.Wo + (slab == 1 ? 8 : 16) - 1) / (slab == 1 ? 8 : 16);
        p.tiles_y = (d.Ho + (slab == 1 ? 16 : 8) - 1) / (slab == 1 ? 16 : 8);
        p.n_slabs = slab == 1 ? d.KW : d.KH;
        p.tps = slab == 1 ? d.KH : d.KW;
        const long long nm = (long long)d.N * p.tiles_x * p.tiles_y;
        BFLOW_REQUIRE(nm < (1ll << 31), "conv_tc3s: too large");
        p.n_mtiles = (int)nm;
    }
    p.n_ntiles = (d.Cout + bn - 1) / bn;
    p.fd_ntiles = bflow::make_fastdiv((unsigned)p.n_ntiles);
    p.fd_wo = bflow::make_fastdiv((unsigned)d.Wo);
    p.fd_ho = bflow::make_fastdiv((unsigned)d.Ho);
    p.ntaps = d.KH * d.KW;
    p.ncb0 = (d.c0 + 63) / 64;
    p.ncb1 = (d.c1 + 63) / 64;
    p.nkb = p.ntaps * (p.ncb0 + p.ncb1);
    p.bn_eff = bn_eff;
    p.half0 = (d.c0 % 64 >= 1 && d.c0 % 64 <= 32) ? p.ncb0 - 1 : -1;
    p.half1 = (d.c1 % 64 >= 1 && d.c1 % 64 <= 32) ? p.ncb0 + p.ncb1 - 1 : -1;
    p.acc_scale = acc_scale;
    BFLOW_REQUIRE(d.precision == BFLOW_PREC_SPLIT3 || d.precision == BFLOW_PREC_F16, "conv_tc3: unknown precision");
    BFLOW_REQUIRE(d.precision == BFLOW_PREC_SPLIT3 || !slab, "conv_tc3s: the halo-slab mode has no single-MMA form");
    p.f16 = d.precision == BFLOW_PREC_F16 ? 1 : 0;
    p.dbg = bflow::g_tc3_debug;
    {
        const int sms = bflow::grid_cap(d);
        static int staged_on = -1;
        if (staged_on < 0) {
            const char* e = getenv("BFLOW_TC3_STAGED");
            staged_on = e == nullptr ? 1 : (e[0] == '0' ? 0 : (e[0] == '2' ? 2 : 1));
        }
        // measured on B200 (tools/timeline.py): the shared-memory pass pays for itself when the epilogue LOADS something per element (GRU
        // gates, residuals) or the tile is ragged; a plain store-only epilogue is ~2 us faster straight from registers
        const bool wants = d.epi != BFLOW_EPI_STD || d.res != nullptr || d.res16_hi != nullptr || (d.Cout % 16) != 0 || staged_on == 2;
        const bool single = d.stats == nullptr && p.n_mtiles * p.n_ntiles <= sms;
        p.staged = (staged_on && wants && single) ? 1 : 0;
        // bulk-copy epilogue: every row piece must be a 16-byte multiple at a 16-byte aligned address
        static int bulk_on = -1;
        if (bulk_on < 0) {
            const char* e = getenv("BFLOW_TC3_BULK");
            bulk_on = (e != nullptr && e[0] == '0') ? 0 : 1;
        }
        auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
        bool bulk = bulk_on && single && bn <= 128 && d.Cout % 4 == 0 && d.res16_hi == nullptr;
        bulk = bulk && (d.y == nullptr || (a16(d.y) && d.ldy % 4 == 0)) && (d.res == nullptr || (a16(d.res) && d.ldr % 4 == 0));
        if (d.epi == BFLOW_EPI_STD) {
            bulk = bulk && d.res == nullptr && (d.y16_hi == nullptr || (a16(d.y16_hi) && a16(d.y16_lo) && d.ldy16 % 8 == 0));
        } else if (d.epi == BFLOW_EPI_GRU_ZR) {
            bulk = bulk && d.y != nullptr && a16(d.aux0) && d.ld_aux0 % 4 == 0 && d.aux1 == nullptr && d.aux1_16_hi != nullptr && a16(d.aux1_16_hi) &&
                   a16(d.aux1_16_lo) && d.ld_aux1_16 % 8 == 0 && (d.Cout / 2) % bn == 0;
        } else {
            static int bulk_q = -1;
            if (bulk_q < 0) {
                const char* e = getenv("BFLOW_TC3_BULK_Q");
                bulk_q = (e != nullptr && e[0] == '1') ? 1 : 0;      // measured: three operand rows in + three out per tile row cost more in bulk-copy issue than the batched loads (22.8 vs 17.1 us)
            }
            bulk = bulk && bulk_q && d.y != nullptr && a16(d.aux0) && d.ld_aux0 % 4 == 0 && (d.y16_hi == nullptr || (a16(d.y16_hi) && a16(d.y16_lo) && d.ldy16 % 8 == 0));
        }
        {   // the fp32 tile plus the operand / output tiles must fit the data area
            const int regions = d.epi == BFLOW_EPI_GRU_Q ? 3 : (d.epi == BFLOW_EPI_GRU_ZR ? 2 : 1);
            bulk = bulk && bflow::T3_BM * (bn + 4) * 4 + regions * bflow::T3_BM * bn * 4 <= bflow::t3_area_bytes(bn, 0);
        }
        if (bulk) p.staged = 2;
        // multi-tile bulk-store epilogue: plain fp32 output only, bn = 128 (runs the <128, 2> instantiation: the third stage's memory is the staging tile)
        static int mstore_on = -1;
        if (mstore_on < 0) {
            const char* e = getenv("BFLOW_TC3_MSTORE");
            mstore_on = (e != nullptr && e[0] == '0') ? 0 : 1;
        }
        const long long shw_ = d.stats_hw > 0 ? d.stats_hw : (long long)d.Ho * d.Wo;
        // with fused statistics: a tile must not straddle two images, and only short main loops (<= 12 k-blocks) are epilogue-bound enough to
        // pay for the lost pipeline stage (measured: 64->96 3x3/2 80 -> 61 us, 1x1/2 59 -> 38 us, but 96->96 3x3 85 -> 90 us)
        const bool stats_ok = d.stats == nullptr || (shw_ % bflow::T3_BM == 0 && d.Cout % 16 == 0 && p.nkb <= 12);
        if (mstore_on && !slab && !single && bn == 128 && d.epi == BFLOW_EPI_STD && stats_ok && d.res == nullptr && d.res16_hi == nullptr &&
            d.y16_hi == nullptr && d.y != nullptr && a16(d.y) && d.ldy % 4 == 0 && d.Cout % 4 == 0 && d.act1 <= BFLOW_ACT_RELU && d.act2 == BFLOW_ACT_NONE)
            p.staged = 3;
    }
    p.trace = bflow::g_tc3_trace;
    p.cta = nullptr;
    if (bflow::g_tc3_cta != nullptr) {
        const int idx = bflow::g_tc3_cta_count++ - bflow::g_tc3_cta_nth;
        if (idx >= 0 && idx < bflow::g_tc3_cta_span) p.cta = bflow::g_tc3_cta + (size_t)idx * 148 * 16;
    }
    p.tl = bflow::timeline_next_slot(bn == 64 ? "tc3_64" : bn == 128 ? "tc3_128" : "tc3_256");
    alignas(64) CUtensorMap tm[9];
    memset(tm, 0, sizeof(tm));
    memcpy(tm, maps, 4 * sizeof(CUtensorMap));
    if (omaps != nullptr) {
        memcpy(tm + 4, omaps, 5 * sizeof(CUtensorMap));
        // tensor-map store epilogue: single-tile CTAs, plain epilogue (none / relu), no residual / statistics
        static int ostore_on = -1;
        if (ostore_on < 0) {
            const char* e = getenv("BFLOW_TC3_OSTORE");
            ostore_on = (e != nullptr && e[0] == '0') ? 0 : 1;
        }
        const bool single1 = d.stats == nullptr && p.n_mtiles * p.n_ntiles <= bflow::grid_cap(d);
        // Cout % 8: measured on B200, a tensor-map store whose box is clipped inside a 16-byte chunk still writes the whole chunk (it
        // overwrote the 4 Bezier-parameter channels behind the motion encoder's 124) -- so the caller's fp16 maps must cover only Cout & ~7 channels;
        // the kernel writes a 4-channel tail itself
        if (ostore_on && single1 && !slab && bn <= 128 && d.Cout % 4 == 0 && d.epi == BFLOW_EPI_STD && d.res == nullptr && d.res16_hi == nullptr && d.act1 <= BFLOW_ACT_RELU &&
            d.act2 <= BFLOW_ACT_RELU)
            p.staged = 4;
        // GRU gate epilogues on tensor maps: omaps = {fp16 hi, fp16 lo (aux1_16 for the z|r launch, y16 for the candidate), y fp32, res, aux0};
        // every tile must be full width (Cout a multiple of bn, the z|r boundary on a tile boundary)
        static int gru_tma = -1;
        if (gru_tma < 0) {
            const char* e = getenv("BFLOW_TC3_GRU_TMA");
            gru_tma = (e != nullptr && e[0] == '1') ? 1 : 0;      // opt-in: measured equal to the per-row bulk / batched-load epilogues -- the 8-12 operand
                                                                    // boxes arrive as 128-byte rows (~4 cycles each through the TMA unit), 6 us before the first use
        }
        // staged 6: GRU gate epilogues with the hoisted term pre-loaded into the accumulator and h / z prefetched into registers (default)
        static int gru6 = -1;
        if (gru6 < 0) {
            const char* e = getenv("BFLOW_TC3_GRU6");
            gru6 = (e != nullptr && e[0] == '0') ? 0 : 1;
        }
        auto a16_ = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
        if (gru6 && ostore_on && single1 && !slab && bn <= 128 && d.epi != BFLOW_EPI_STD && d.Cout % bn == 0 && d.bias == nullptr && d.scale == 1.f &&
            d.res != nullptr && a16_(d.res) && d.ldr % 4 == 0 && d.y != nullptr && a16_(d.y) && d.ldy % 4 == 0 && a16_(d.aux0) && d.ld_aux0 % 4 == 0 &&
            (d.epi == BFLOW_EPI_GRU_Q ? (bn == 64 && d.y16_hi != nullptr) : ((d.Cout / 2) % bn == 0 && d.aux1 == nullptr && d.aux1_16_hi != nullptr)))
            p.staged = 6;
        const int regions = d.epi == BFLOW_EPI_GRU_Q ? 3 : 2;
        if (p.staged != 6 && gru_tma && ostore_on && single1 && !slab && bn <= 128 && d.epi != BFLOW_EPI_STD && d.Cout % bn == 0 && d.y != nullptr && d.res != nullptr &&
            (d.epi == BFLOW_EPI_GRU_Q ? d.y16_hi != nullptr : ((d.Cout / 2) % bn == 0 && d.aux1 == nullptr && d.aux1_16_hi != nullptr)) &&
            (regions * ((bn + 31) / 32) + 2 * ((bn + 63) / 64)) * 16384 <= bflow::t3_area_bytes(bn, 0))
            p.staged = 5;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (p.f16) {
        switch (bn) {
            case 64: return bflow::launch_tc3<64, 4, true>(tm, d, w_tc, p, err, st);
            case 128: return p.staged == 3 ? bflow::launch_tc3<128, 2, true>(tm, d, w_tc, p, err, st) : bflow::launch_tc3<128, 3, true>(tm, d, w_tc, p, err, st);
            case 256: return bflow::launch_tc3<256, 2, true>(tm, d, w_tc, p, err, st);
            default: break;
        }
    } else {
        switch (bn) {
            case 64: return bflow::launch_tc3<64, 4, false>(tm, d, w_tc, p, err, st);
            case 128: return p.staged == 3 ? bflow::launch_tc3<128, 2, false>(tm, d, w_tc, p, err, st) : bflow::launch_tc3<128, 3, false>(tm, d, w_tc, p, err, st);
            case 256: return bflow::launch_tc3<256, 2, false>(tm, d, w_tc, p, err, st);
            default: break;
        }
    }
    bflow::set_error("conv_tc3: bn must be 64, 128 or 256");
    return BFLOW_ERR_INVALID;
}

extern "C" int bflow_conv2d_nhwc_tc3(const bflow_conv_desc* d, const void* maps, const void* w_tc, int bn, float acc_scale, int* err, void* stream) {
    return tc3_entry(d, maps, nullptr, w_tc, bn, acc_scale, 0, err, stream);
}

// The same with tensor maps for the OUTPUTS: omaps = three 128-byte maps {y16 hi plane, y16 lo plane, y fp32} from bflow_tma_out_map (zeroed
// entries for outputs the descriptor does not have).  Single-tile launches with a plain epilogue then store through cp.async.bulk.tensor.
extern "C" int bflow_conv2d_nhwc_tc3o(const bflow_conv_desc* d, const void* maps, const void* omaps /* 5 x 128 bytes */, const void* w_tc, int bn, float acc_scale, int* err,
                                      void* stream) {
    return tc3_entry(d, maps, omaps, w_tc, bn, acc_scale, 0, err, stream);
}

// 2-D tensor map over an output tensor [rows][cols] (row stride ld elements; elem_bytes 2 = one split-fp16 plane, 4 = fp32), box = 128 rows x
// 128 bytes, SWIZZLE_128B: what the store epilogue of bflow_conv2d_nhwc_tc3o writes.
extern "C" int bflow_tma_out_map(void* map_out, const void* base, long long rows, int cols, int ld_elems, int elem_bytes) {
    BFLOW_REQUIRE(map_out != nullptr && base != nullptr, "tma_out_map: null argument");
    BFLOW_REQUIRE(rows > 0 && cols > 0 && ld_elems >= cols && (elem_bytes == 2 || elem_bytes == 4), "tma_out_map: bad shape");
    BFLOW_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ((long long)ld_elems * elem_bytes) % 16 == 0, "tma_out_map: 16-byte aligned base and row stride");
    static bflow::EncodeTiledFn enc = nullptr;
    if (enc == nullptr) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<bflow::EncodeTiledFn>(fp);
    }
    BFLOW_REQUIRE(enc != nullptr, "tma_out_map: cuTensorMapEncodeTiled not available from the driver");
    alignas(64) CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld_elems * elem_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), 128};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        bflow::set_error("tma_out_map: cuTensorMapEncodeTiled failed");
        return BFLOW_ERR_CUDA;
    }
    memcpy(map_out, &tm, sizeof(tm));
    return BFLOW_OK;
}

// Slab mode of the same kernel for stride-1 3x3 / 5x1 (orientation 1) and 1x5 / 3x3 (orientation 2) convolutions: `maps` are TILED tensor
// maps from bflow_tma_tile_map (box 8 x (16 + taps - 1), transposed for orientation 2) instead of im2col maps; everything else as above.
extern "C" int bflow_conv2d_nhwc_tc3s(const bflow_conv_desc* d, const void* maps, const void* w_tc, int bn, float acc_scale, int orientation, int* err,
                                      void* stream) {
    return tc3_entry(d, maps, nullptr, w_tc, bn, acc_scale, orientation, err, stream);
}


// Tiled (not im2col) tensor map over a split-fp16 NHWC plane: dims {C, W, H, N}, box {64 channels, box_w, box_h, 1}, SWIZZLE_128B,
// out-of-bounds elements read as zero.  Used by bflow_conv2d_slab64 with box 8 x 18.
extern "C" int bflow_tma_tile_map(void* map_out, const void* base, int N, int H, int W, int C, int ld_halves, int box_w, int box_h, int transposed) {
    BFLOW_REQUIRE(map_out != nullptr && base != nullptr, "tma_tile_map: null argument");
    BFLOW_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && ld_halves >= C && ld_halves % 8 == 0, "tma_tile_map: bad shape (ld a multiple of 8 halves)");
    BFLOW_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && box_w > 0 && box_w <= 256 && box_h > 0 && box_h <= 256, "tma_tile_map: alignment / box");
    static bflow::EncodeTiledFn enc = nullptr;
    if (enc == nullptr) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<bflow::EncodeTiledFn>(fp);
    }
    BFLOW_REQUIRE(enc != nullptr, "tma_tile_map: cuTensorMapEncodeTiled not available from the driver");
    alignas(64) CUtensorMap tm;
    // transposed: dimension 1 is y and dimension 2 is x, so the box is stored y-fastest (slab rows = (x, y)); box_w / box_h keep their meaning
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld_halves * 2, (cuuint64_t)W * ld_halves * 2, (cuuint64_t)H * W * ld_halves * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    if (transposed) {
        dims[1] = (cuuint64_t)H; dims[2] = (cuuint64_t)W;
        strides[0] = (cuuint64_t)W * ld_halves * 2; strides[1] = (cuuint64_t)ld_halves * 2;
        box[1] = (cuuint32_t)box_h; box[2] = (cuuint32_t)box_w;
    }
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        bflow::set_error("tma_tile_map: cuTensorMapEncodeTiled failed");
        return BFLOW_ERR_CUDA;
    }
    memcpy(map_out, &tm, sizeof(tm));
    return BFLOW_OK;
}

// 3x3 / stride 1 / pad 1 convolution, exactly 64 -> 64 channels, split-fp16 input (ResidualBlock convs of layer1, extractor.py:49-53).
// maps: two 128-byte tensor maps {hi, lo} from bflow_tma_tile_map(..., box_w 8, box_h 18).  w_tc: the bflow_conv2d_nhwc_tc3 weight image for
// bn = 64 (tap-major k-blocks).  Epilogue: bias, act1, split residual, act2 (none / relu), fp32 and / or split output, fused InstanceNorm sums.
extern "C" int bflow_conv2d_slab64(const bflow_conv_desc* dp, const void* maps, const void* w_tc, float acc_scale, int* err, void* stream) {
    BFLOW_CHECK_DESC(dp, bflow_conv_desc, "conv_slab64");
    BFLOW_REQUIRE(maps != nullptr && w_tc != nullptr, "conv_slab64: null argument");
    const bflow_conv_desc& d = *dp;
    BFLOW_REQUIRE(d.c0 == 64 && d.c1 == 0 && d.Cout == 64 && d.KH == 3 && d.KW == 3 && d.stride == 1 && d.pad_h == 1 && d.pad_w == 1, "conv_slab64: 3x3/1 64->64 only");
    BFLOW_REQUIRE(d.N > 0 && d.H > 0 && d.W > 0 && d.W % 8 == 0 && d.Ho == d.H && d.Wo == d.W, "conv_slab64: W must be a multiple of 8");
    BFLOW_REQUIRE(d.epi == BFLOW_EPI_STD && d.res == nullptr && d.act1 <= BFLOW_ACT_RELU && d.act2 <= BFLOW_ACT_RELU, "conv_slab64: standard epilogue, none / relu, split residual only");
    BFLOW_REQUIRE((d.y == nullptr || (d.ldy >= 64 && d.ldy % 4 == 0 && bflow::aligned16(d.y))), "conv_slab64: fp32 output alignment");
    BFLOW_REQUIRE(d.y16_hi == nullptr || (d.y16_lo != nullptr && d.ldy16 % 8 == 0 && bflow::aligned16(d.y16_hi) && bflow::aligned16(d.y16_lo)), "conv_slab64: split output alignment");
    BFLOW_REQUIRE(d.stats == nullptr || (d.act1 == BFLOW_ACT_NONE && d.act2 == BFLOW_ACT_NONE && d.res16_hi == nullptr), "conv_slab64: fused statistics need the plain epilogue");
    if (const char* msg = bflow::check_epilogue(d)) { bflow::set_error(msg); return BFLOW_ERR_INVALID; }
    BFLOW_REQUIRE((reinterpret_cast<uintptr_t>(w_tc) & 15) == 0, "conv_slab64: weight alignment");
    bflow::SlabParams p;
    p.tiles_x = d.W / 8;
    p.tiles_y = (d.H + 15) / 16;
    const long long nt = (long long)d.N * p.tiles_x * p.tiles_y;
    BFLOW_REQUIRE(nt < (1ll << 31), "conv_slab64: too large");
    p.n_tiles = (int)nt;
    p.acc_scale = acc_scale;
    BFLOW_REQUIRE(d.precision == BFLOW_PREC_SPLIT3 || d.precision == BFLOW_PREC_F16, "conv_slab64: unknown precision");
    p.f16 = d.precision == BFLOW_PREC_F16 ? 1 : 0;
    p.tl = bflow::timeline_next_slot("slab64");
    constexpr int smem = bflow::SL_B_BYTES + bflow::SL_STAGES * bflow::SL_STAGE + 128 + 3 * 64 * 4 + 1024;
    static bflow::PerDeviceFlag configured;
    if (!configured.get()) {
        cudaError_t e = cudaFuncSetAttribute(bflow::conv_slab64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            bflow::set_error(cudaGetErrorString(e));
            return BFLOW_ERR_CUDA;
        }
        configured.set();
    }
    alignas(64) CUtensorMap tm[2];
    memcpy(tm, maps, sizeof(tm));
    const int grid = p.n_tiles < bflow::grid_cap(d) ? p.n_tiles : bflow::grid_cap(d);
    bflow::conv_slab64_kernel<<<grid, bflow::T3_THREADS, smem, (cudaStream_t)stream>>>(tm[0], tm[1], d, reinterpret_cast<const uint8_t*>(w_tc), p, err);
    return bflow::check_launch("bflow_conv2d_slab64");
}
