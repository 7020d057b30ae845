/* bflow_b200 — C ABI of the B200-native (sm_100a) RAFT-spline inference hot path.
 *
 * The reference (uzh-rpg/bflow) is 100 % Python and has NO native interface: every entry point
 * below replaces a group of PyTorch/ATen calls at the cited reference location
 * (paths relative to the reference root).  A maintainer binds them with ctypes;
 * bflow_b200/ops.py is that binding.
 *
 * Conventions
 *   - all pointers are DEVICE pointers to fp32 unless stated otherwise; inputs are borrowed and never
 *     written; nothing is allocated or freed inside the library;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); no entry point
 *     synchronises the host;
 *   - every function returns 0 on success, BFLOW_ERR_INVALID for a contract violation detected on the
 *     host (the reference's convention is `assert`), BFLOW_ERR_CUDA when the launch failed
 *     (bflow_last_error() then holds cudaGetErrorString);
 *   - "NHWC" activations are addressed as  base[pixel * ld + channel]  with pixel = (n*H + y)*W + x,
 *     so a tensor may be a channel slice of a wider buffer (ld >= channels).  "NCHW" is the
 *     reference's layout.
 */
#ifndef BFLOW_B200_H
#define BFLOW_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define BFLOW_OK 0
#define BFLOW_ERR_INVALID 1
#define BFLOW_ERR_CUDA 2

#define BFLOW_ACT_NONE 0
#define BFLOW_ACT_RELU 1
#define BFLOW_ACT_SIGMOID 2
#define BFLOW_ACT_TANH 3

/* epilogue modes of bflow_conv_desc.epi
 *   STD     out = act2(res + act1(scale*(conv + bias)))
 *   GRU_ZR  Cout = 2C, columns [z | r]:  g = sigmoid(conv + bias + res);  y = g;  for the r half also
 *           aux1[m, n-C] = g * aux0[m, n-C]            (aux0 = h, aux1 = r*h : update.py:37-39)
 *   GRU_Q   q = tanh(conv + bias + res);  z = aux0[m, n];  y[m, n] = (1-z)*y[m, n] + z*q   (y = h in place, update.py:39-40) */
#define BFLOW_EPI_STD 0
#define BFLOW_EPI_GRU_ZR 1
#define BFLOW_EPI_GRU_Q 2

#define BFLOW_MAX_SLOTS 16
#define BFLOW_MAX_TARGETS 8
#define BFLOW_MAX_DEGREE 16

/* ABI version 2: both descriptors start with `struct_size` (= sizeof of the struct the CALLER was compiled against; the library
 * refuses a descriptor of another size with BFLOW_ERR_INVALID instead of reading past it), bflow_conv_desc carries `precision`,
 * bflow_sizeof_*() / bflow_source_hash() exist.  Bumped whenever a descriptor layout or an entry-point signature changes. */
#define BFLOW_ABI_VERSION 2
int bflow_abi_version(void);
const char* bflow_last_error(void);
/* compute capability the library was built for (100 for sm_100a) */
int bflow_built_for_sm(void);
/* sizeof(bflow_conv_desc) / sizeof(bflow_lookup_desc) as the LIBRARY was compiled: a binding checks its own struct against these */
int bflow_sizeof_conv_desc(void);
int bflow_sizeof_lookup_desc(void);
/* hex sha256 over the library sources (every .cu and .cuh under csrc, then include/bflow_b200.h; sorted) baked in at build time:
 * bflow_b200/build.py recomputes it from the working tree, so a stale binary is detected by content, not by mtime */
const char* bflow_source_hash(void);

/* arithmetic of the tensor-core convolutions (bflow_conv_desc.precision) */
#define BFLOW_PREC_SPLIT3 0   /* x = hi + lo (two fp16 planes), hi*hi + hi*lo + lo*hi, fp32 accumulate: fp32-equivalent (default) */
#define BFLOW_PREC_F16 1      /* single fp16 MMA on the hi planes only (lo planes neither read nor written), fp32 accumulate */

/* ---------------------------------------------------------------------------------------------
 * Layout plumbing.  Replaces torch slicing/cat of the voxel grid windows (models/raft_spline/raft.py:88-99),
 * the image normalisation 2*(x/255)-1 (raft.py:134) and the NCHW views the reference API exposes.
 * dst[pix*dst_ld + c] = src[n, c_off + c, y, x] * scale + shift     for c in [0, c_cnt)
 * ------------------------------------------------------------------------------------------- */
int bflow_nchw_to_nhwc(const float* src, float* dst, int N, int C_total, int H, int W,
                       int c_off, int c_cnt, int dst_ld, float scale, float shift, void* stream);
/* dst[n, c, y, x] = src[pix*src_ld + c]   for c in [0, C) */
int bflow_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, int src_ld, void* stream);

/* ---------------------------------------------------------------------------------------------
 * 2-D convolution as implicit GEMM on NHWC activations.  Replaces every nn.Conv2d call on the path:
 * models/raft_utils/extractor.py:49-53,112,120 and models/raft_spline/update.py:17-18,36-45,89-96,112-114.
 *   out = act2( res + act1( scale * (conv(x, w) + bias) ) )
 * The input is the channel concatenation of up to two sources (torch.cat at update.py:35,38,42,45,94,120).
 * Weights are pre-packed K-major: w[((kh*KW + kw)*Cin + c) * ldw + o], zero padded to ldw (ldw % 4 == 0).
 * ------------------------------------------------------------------------------------------- */
typedef struct bflow_conv_desc {
    int struct_size;                       /* sizeof(bflow_conv_desc) of the caller */
    int precision;                         /* BFLOW_PREC_*: tensor-core entry points only */
    const float* x0; int c0; int ld0;
    const float* x1; int c1; int ld1;      /* x1 == NULL / c1 == 0: single source */
    const float* w;  int ldw;
    const float* bias;                     /* [Cout] or NULL */
    const float* res; int ldr;             /* optional residual, NHWC with pixel stride ldr */
    float* y; int ldy;
    int N, H, W, Ho, Wo, Cout;
    int KH, KW, stride, pad_h, pad_w;
    int act1, act2;
    float scale;
    /* fused SepConvGRU epilogues (update.py:36-40); epi = BFLOW_EPI_STD ignores aux0/aux1 */
    int epi;
    const float* aux0; int ld_aux0;
    float* aux1; int ld_aux1;
    /* split-fp16 twins (x = hi + lo, two fp16 planes with pixel stride in halves): what the TMA-fed tensor-core kernel
     * (bflow_conv2d_nhwc_tc3) reads.  Every kernel that produces a convolution input can emit them next to, or instead of
     * (y == NULL), the fp32 tensor; res16 lets a residual be read from a split tensor; aux1_16 is the split form of aux1. */
    void* y16_hi; void* y16_lo; int ldy16;
    const void* res16_hi; const void* res16_lo; int ldr16;
    void* aux1_16_hi; void* aux1_16_lo; int ld_aux1_16;
    /* InstanceNorm statistics fused into the epilogue (bflow_conv2d_nhwc_tc3 only, standard epilogue, Cout % 16 == 0):
     * stats[(m / stats_hw) * Cout + n] += (v, v*v) of every value v written to y — the (sum, sum of squares) table that
     * bflow_instnorm_relu(16) consumes, so no separate bflow_plane_sums pass.  stats_hw = rows per image (0: Ho*Wo). */
    double* stats; int stats_hw;
    /* persistent tensor-core kernels (tc3 family, slab64, stem7): at most this many CTAs (0 = one per SM).  Lets two independent chains of
     * launches on two streams own disjoint sets of SMs instead of queueing behind each other's one-CTA-per-SM grids. */
    int max_ctas;
} bflow_conv_desc;
int bflow_conv2d_nhwc(const bflow_conv_desc* d, void* stream);

/* Tensor-core forms of the same operator (tcgen05.mma, TMEM accumulators, split 16-bit operands: every fp32
 * operand x = hi + lo with hi = fp16(x), lo = fp16(x - hi), saturating at |x| = 1.3e5; products hi*hi + hi*lo + lo*hi
 * accumulated in fp32).
 * `d->w/ldw` are ignored; `w_tc` is the packed image of W / acc_scale
 *   [ceil(Cout/bn)][ceil(K/64)][hi | lo (fp16)][bn rows][64 elements]
 * whose 16-byte chunks are XOR-swizzled by (row % 8), i.e. byte for byte the SWIZZLE_128B shared-memory tile
 * (bflow_b200/ops.py pack_conv_weight_tc); acc_scale (a power of two) is multiplied back onto the accumulator.
 * bn in {64,128,256}; for a launch with Cout <= bn < 128 also bn in {80,96,112}: the 128-column kernel with bn weight rows per half (w_tc packed
 * for that bn), i.e. MMAs of N = 2*bn and bn instead of 256 and 128 -- the 96-channel encoder layers.  `err`: optional device int, set to 1 if an in-kernel pipeline wait timed out (never expected; the
 * waits are bounded so that a bug cannot hang the GPU).  The output of such a launch is invalid: the caller must read the word
 * back (bflow_b200/engine.py does so with every forward's results and raises).
 * precision = BFLOW_PREC_F16 runs ONE tcgen05.mma per k-step on the hi planes / the hi half of the weight image. */
/* TMA-fed, persistent tensor-core convolution.  Activations are read as split-fp16 planes (hi, lo) through
 * im2col tensor maps (cp.async.bulk.tensor.4d...im2col: the TMA unit does the implicit-GEMM gather and the zero padding),
 * weights as the image above with K ordered (tap, 64-channel block) — pack_conv_weight_tc(block_per_tap=True).
 * One CTA per SM loops over 128 x bn tiles; TMEM holds two accumulators so the epilogue of a tile overlaps the MMAs of the
 * next.  Warp roles: TMA producer / MMA issuer / 4 epilogue warps.  d->x0/x1 are ignored (c0/c1 and the geometry are used).
 * maps: host array of four 128-byte tensor maps {source0 hi, source0 lo, source1 hi, source1 lo} from bflow_tma_im2col_map
 * (source1 entries unused when c1 == 0).  Channel counts are free (the TMA unit zero-fills beyond C); needs c1 == 0 or c0 % 64 == 0,
 * channel offsets that are multiples of 8 and row strides that are multiples of 8 halves. */
int bflow_tma_im2col_map(void* map_out_128B, const void* base_fp16, int N, int H, int W, int C, int ld_halves,
                         int KH, int KW, int stride, int pad_h, int pad_w);
int bflow_conv2d_nhwc_tc3(const bflow_conv_desc* d, const void* maps, const void* w_tc, int bn, float acc_scale, int* err, void* stream);
/* Slab variant for 3x3 / stride 1 / pad 1, exactly 64 -> 64 channels, W % 8 == 0 (ResidualBlock convs of layer1, extractor.py:49-53):
 * weights resident in shared memory, one 8 x 18 halo slab per filter column serves the three filter rows (4x less L2 -> SM traffic than
 * the im2col kernel).  maps: {hi, lo} tensor maps from bflow_tma_tile_map(..., box_w 8, box_h 18); w_tc: the tc3 weight image for bn = 64.
 * Standard epilogue only (none / relu, split residual, fp32 and / or split output, fused InstanceNorm sums). */
int bflow_tma_tile_map(void* map_out_128B, const void* base_fp16, int N, int H, int W, int C, int ld_halves, int box_w, int box_h, int transposed);
/* bflow_conv2d_nhwc_tc3 in slab mode (stride-1 3x3, 5x1: orientation 1; 1x5, 3x3: orientation 2): output tiles are 8 x 16 pixel patches and one halo
 * slab per slow filter index serves all taps along the other axis; `maps` = tiled maps (bflow_tma_tile_map, box 8 x (16 + taps - 1) pixels,
 * transposed = orientation 2) in the same {source0 hi, source0 lo, source1 hi, source1 lo} order.  Same weights, epilogues and outputs. */
/* bflow_conv2d_nhwc_tc3 with tensor maps for the OUTPUTS: omaps = three 128-byte maps {y16 hi plane, y16 lo plane, y fp32} from
 * bflow_tma_out_map (zero bytes for outputs the descriptor does not have).  Single-tile launches with a plain epilogue (none / relu, no
 * residual, no statistics) then leave through cp.async.bulk.tensor stores of SWIZZLE_128B boxes; every other case behaves like _tc3. */
int bflow_tma_out_map(void* map_out_128B, const void* base, long long rows, int cols, int ld_elems, int elem_bytes);
int bflow_conv2d_nhwc_tc3o(const bflow_conv_desc* d, const void* maps, const void* omaps, const void* w_tc, int bn, float acc_scale, int* err, void* stream);
int bflow_conv2d_nhwc_tc3s(const bflow_conv_desc* d, const void* maps, const void* w_tc, int bn, float acc_scale, int orientation, int* err, void* stream);
int bflow_conv2d_slab64(const bflow_conv_desc* d, const void* maps, const void* w_tc, float acc_scale, int* err, void* stream);
/* Fused encoder stem (extractor.py:112): 7x7 / stride 2 / pad 3 over n_windows (<= 8) channel windows [c_offs[i], c_offs[i] + cin) of an fp32
 * NCHW input of N / n_windows samples (cin <= 5, W % 4 == 0; output image i * samples + s = window i of sample s: the torch.cat of
 * extractor.py:106-110),
 * input first mapped x -> in_scale * x + in_shift (raft.py:134), 64 output channels.  The patch matrix is built in shared memory from the
 * input footprint of each 8 x 16 output tile and multiplied on tcgen05 against the resident weights (w_tc: tc3 image, bn 64, of the
 * [64][256] matrix, K = (kh*7+kw)*cin + c).  d: x0 = the NCHW input, c0 = cin, N/H/W/Ho/Wo, Cout = 64, bias, act1, y / y16, stats; plain
 * epilogue (none / relu, fp32 and / or split output, fused InstanceNorm sums). */
int bflow_conv2d_stem7(const bflow_conv_desc* d, const void* w_tc, int c_total, const int* c_offs, int n_windows, float in_scale, float in_shift,
                       float acc_scale, int* err, void* stream);
/* im2col of a channel window of an NCHW fp32 tensor straight into split-fp16 rows (the 7x7 stride-2 encoder stems on few input
 * channels, extractor.py:112: K = KH*KW*cin is too thin per tap for 64-channel TMA boxes, so the patch matrix is materialised
 * once and the stem becomes a 1x1 tensor-core GEMM).  out[row, (kh*KW+kw)*cin + c] = scale*src[n, c_off+c, oh*s-ph+kh, ow*s-pw+kw] + shift
 * (zero outside the image), row = (n*Ho + oh)*Wo + ow; columns K..ld16-1 are written as zeros.  ld16 % 8 == 0. */
int bflow_im2col_split16(const float* src_nchw, int C_total, int c_off, int cin, int N, int H, int W, int KH, int KW, int stride,
                         int pad_h, int pad_w, float scale, float shift, void* out_hi, void* out_lo, int ld16, void* stream);
/* fp32 NHWC rows -> split-fp16 planes (x = hi + lo): operand staging for tensors produced outside this library */
int bflow_split_f16(const float* src, int ld, void* hi, void* lo, int ld16, long long rows, int C, void* stream);

/* Packs NHWC fp32 rows (rows x K at stride ld) into the tensor-core B-operand image above (rows play the role of
 * output channels): used for the correlation volume, whose "weights" are the target feature map.  dst must be
 * zero-initialised once (rows beyond `rows` in the last tile stay zero).  Row r of the source lands at image row
 * perm(r): identity when tile_w == 0; otherwise the source rows are pixels (y, x) of a tile_h x tile_w plane and
 * land in 4x4-pixel-tiled order (the layout bflow_corr_lookup reads with tiled = 1). */
int bflow_pack_b_tc(const float* src, int ld, void* dst, int rows, int K, int bn, int plane_h, int plane_w, void* stream);
/* Direct convolution for tiny Cout (<= 32): one warp per output pixel, K split over lanes, shuffle reduction.
 * Same descriptor and packed weights as bflow_conv2d_nhwc (Bezier head conv2: 256 -> 2*degree, update.py:18). */
int bflow_conv2d_small_n(const bflow_conv_desc* d, void* stream);
/* 7x7 / stride 1 / pad 3 convolution of a thin input (Cin % 4 == 0) to exactly 128 channels on CUDA cores with the weights in
 * shared memory: convf1 of the motion encoder, Bezier parameters -> 128 (update.py:91).  Same descriptor and packed fp32
 * weights as bflow_conv2d_nhwc; standard epilogue only. */
int bflow_conv2d_thin7(const bflow_conv_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * InstanceNorm2d (biased variance, eps, no affine; extractor.py:27-31) in two passes:
 * per-(n,c) sums, then  out = relu( (a-mu_a)*rstd_a )                      if r == NULL
 *                       out = relu( relu((a-mu_a)*rstd_a) + R )            otherwise,
 *                       R = r (identity skip) or (r-mu_r)*rstd_r (norm3 of the 1x1 downsample;
 *                       extractor.py:43-44,47-55).
 * sums: double[N][C][2] = (sum x, sum x^2), must be zeroed by the caller (bflow_zero).
 * ------------------------------------------------------------------------------------------- */
int bflow_plane_sums(const float* x, int ld, double* sums, int N, int HW, int C, void* stream);
int bflow_instnorm_relu(const float* a, int lda, const double* sums_a,
                        const float* r, int ldr, const double* sums_r,
                        float* out, int ldo, int N, int HW, int C, float eps, void* stream);
/* split-fp16 form: the residual may come from a split tensor (r16_hi/lo, identity skip of an encoder block whose input is
 * stored split; r16_lo == NULL: the hi plane alone) and the result may be written as fp32 (out, may be NULL) and/or split
 * planes (out16_hi/lo, may be NULL; out16_lo == NULL with out16_hi set: hi plane only, BFLOW_PREC_F16 consumers). */
int bflow_instnorm_relu16(const float* a, int lda, const double* sums_a,
                          const float* r, int ldr, const double* sums_r,
                          const void* r16_hi, const void* r16_lo, int ldr16,
                          float* out, int ldo, void* out16_hi, void* out16_lo, int ldo16,
                          int N, int HW, int C, float eps, void* stream);
int bflow_zero(void* ptr, unsigned long long bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Correlation volume (models/raft_utils/corr.py:264-272):
 *   corr[bq, p] = sum_d f1[bq, d] * f2[b, d, p] / sqrt(D)       for one target
 * f1: NHWC rows (B*Q, D) with pixel stride ld1;  f2: NCHW (B, D, Q);  corr: (B*Q, Q) row-major, i.e. the
 * reference's (B*h*w, 1, h, w) plane stack of that target.
 * ------------------------------------------------------------------------------------------- */
int bflow_corr_volume(const float* f1, int ld1, const float* f2_nchw, float* corr,
                      int B, int D, int Q, void* stream);
/* avg_pool2d(2, stride 2) with floor on a stack of planes (corr.py:119): (P,H,W) -> (P,H/2,W/2) */
int bflow_corr_pool(const float* in, float* out, long long planes, int H, int W, void* stream);
/* the same pooling on 4x4-tiled planes (in: ceil4(H) x ceil4(W) tiled, out: ceil4(H/2) x ceil4(W/2) tiled, pad = 0) */
int bflow_corr_pool_tiled(const float* in, float* out, long long planes, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pyramid lookup (corr.py:307-350 + models/raft_utils/utils.py:5-21).  One unit = (query pixel, slot):
 * a (2r+1)^2 window sampled bilinearly (pixel units, align_corners=True, per-corner zero padding) around
 * coords/2^level in the query's private plane.  Output channel = slot*81 + iy*9 + ix (dy = iy-4 major).
 * Centre coordinates come either from `coords` (T,B,2,h,w — the reference's argument) or, when
 * coords == NULL, are formed in-kernel as pixel grid + Bezier flow (raft.py:180-181, bezier.py:165-186)
 * from `params` (NHWC rows (B*Q) x 2*degree at pixel stride params_ld; channel = dim*degree + i-1).
 * ------------------------------------------------------------------------------------------- */
typedef struct bflow_lookup_desc {
    int struct_size;                       /* sizeof(bflow_lookup_desc) of the caller */
    int n_slots, n_targets, B, h, w, radius;
    const float* vol[BFLOW_MAX_SLOTS];     /* (B*Q, hl, wl) planes of this slot's (level, target) */
    int hl[BFLOW_MAX_SLOTS], wl[BFLOW_MAX_SLOTS];
    int target[BFLOW_MAX_SLOTS];
    float inv_scale[BFLOW_MAX_SLOTS];      /* 1 / 2^level */
    const float* coords;                   /* (T,B,2,h,w) or NULL */
    const float* params; int params_ld; int degree;
    float coef[BFLOW_MAX_TARGETS][BFLOW_MAX_DEGREE];   /* Bernstein weights per target timestamp */
    float* out;
    int out_nhwc;                          /* 0: (B, S*81, h, w) like the reference; 1: rows (B*Q) x out_ld */
    int out_ld;
    void* out16_hi; void* out16_lo;        /* when non-NULL (NHWC only): write split-fp16 planes (x = hi + lo) instead of `out`, */
    int out16_ld;                          /* row stride in halves — the form the TMA-fed convolution reads; out16_lo == NULL: hi plane
                                              only (BFLOW_PREC_F16 consumers) */
    int tiled;                             /* 0: planes row-major (hl x wl) like the reference; 1: planes stored as 4x4-pixel
                                              tiles (64-byte DRAM granules), ceil(hl/4) x ceil(wl/4) tiles of 16 floats, zero padded */
} bflow_lookup_desc;
int bflow_corr_lookup(const bflow_lookup_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * On-the-fly correlation lookup (scope row f3): the same output as bflow_corr_lookup WITHOUT a materialised volume.  For every unit
 * (query pixel, slot) the 10 x 10 footprint of correlation values is computed directly as <f1[q], f2_level[p]> * scale against an
 * average-pooled TARGET FEATURE pyramid (bflow_feat_pool: avg_pool2d(2, 2, floor) of NHWC features; pooling is linear, so this equals
 * the pooled correlation planes of corr.py:119) and blended exactly like bflow_corr_lookup.  Replaces corr.py:108-125,264-272,307-350.
 * f1[s]: query features of slot s, NHWC rows (B*Q) x D at stride ld1; f2[s]: the target features of slot s's (target, level), NHWC (B, hl, wl, D)
 * rows at stride ld2; D % 4 == 0, D <= 512; scale = 1/sqrt(D) (corr.py:267).  Output NHWC rows: fp32 `out` (stride out_ld) or, when
 * out16_hi != NULL, split-fp16 planes (out16_lo == NULL: hi plane only).
 * ------------------------------------------------------------------------------------------- */
typedef struct bflow_lookup_otf_desc {
    int struct_size;                       /* sizeof(bflow_lookup_otf_desc) of the caller */
    int n_slots, n_targets, B, h, w, radius, D;
    const float* f1[BFLOW_MAX_SLOTS]; int ld1;   /* per slot: event targets and the image target have different query feature maps */
    const float* f2[BFLOW_MAX_SLOTS]; int ld2;
    int hl[BFLOW_MAX_SLOTS], wl[BFLOW_MAX_SLOTS];
    int target[BFLOW_MAX_SLOTS];
    float inv_scale[BFLOW_MAX_SLOTS];      /* 1 / 2^level */
    float scale;
    const float* coords;                   /* (T,B,2,h,w) or NULL */
    const float* params; int params_ld; int degree;
    float coef[BFLOW_MAX_TARGETS][BFLOW_MAX_DEGREE];
    float* out; int out_ld;
    void* out16_hi; void* out16_lo; int out16_ld;
} bflow_lookup_otf_desc;
int bflow_sizeof_lookup_otf_desc(void);
int bflow_corr_lookup_otf(const bflow_lookup_otf_desc* d, void* stream);
/* avg_pool2d(2, stride 2, floor) of NHWC fp32 features: (N, H, W, C) rows at ld_in -> (N, H/2, W/2, C) rows at ld_out; C % 4 == 0 */
int bflow_feat_pool(const float* in, float* out, int N, int H, int W, int C, int ld_in, int ld_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SepConvGRU gate arithmetic (update.py:37-40,44-47) on NHWC rows:
 *   bflow_gru_rh:      rh = r * h            (r = zr[:, C:2C])
 *   bflow_gru_update:  h  = (1-z)*h + z*q    (z = zr[:, 0:C]), in place
 * ------------------------------------------------------------------------------------------- */
int bflow_gru_rh(const float* zr, int ldzr, const float* h, int ldh, float* rh, int ldrh,
                 long long rows, int C, void* stream);
int bflow_gru_update(const float* zr, int ldzr, const float* q, int ldq, float* h, int ldh,
                     long long rows, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Bezier evaluation (bezier.py:165-186): flows[t,b,d,y,x] = sum_i coef[t][i] * P_i,  params NCHW
 * (B, 2*degree, H, W), coef host array [T][degree] (fp32), T <= 32.
 * ------------------------------------------------------------------------------------------- */
int bflow_bezier_eval(const float* params_nchw, const float* coef_host, float* flows,
                      int T, int B, int degree, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Convex 8x upsampling (utils.py:33-48 via bezier.py:81-84).  data: NHWC rows (N*h*w) x C at pixel stride
 * ldd (or NCHW when data_nchw != 0); mask: NHWC rows x 576 at stride ldm (or NCHW (N,576,h,w) when
 * mask_nchw != 0), channel = k*64 + i*8 + j.  out: NCHW (N, C, 8h, 8w).
 * ------------------------------------------------------------------------------------------- */
int bflow_cvx_upsample(const float* data, int ldd, int data_nchw, const float* mask, int ldm, int mask_nchw,
                       float* out, int N, int C, int h, int w, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Scope rows (f1)/(f2): the steps on either side of forward().
 * bflow_voxelize: VoxelGrid.convert (data/utils/representations.py:64-111).  x, y: int64 pixel coordinates (xy_is_float = 0) or
 * fp32 sub-pixel coordinates (1); pol: uint8/bool 0|1; time: int64; out: (channels, H, W) fp32, ACCUMULATED into (zero it first).
 * Integer coordinates outside [0, W) x [0, H) make the reference's put_ raise; here such events are dropped and counted in
 * *oob_count (device int, may be NULL) so that the binding can raise the same IndexError without corrupting memory.
 * bflow_voxel_norm: norm_voxel_grid (representations.py:9-18), in place; stats3: 3 doubles of scratch.
 * bflow_epe_masked: epe_masked (utils/metrics.py:196-213) as (sum, count): sum_count[0] += sum of sqrt(sum_c (src-tgt)^2) over the
 * valid pixels, sum_count[1] += their number (valid: uint8/bool (N, HW) or NULL); src/tgt NCHW (N, C, HW).
 * ------------------------------------------------------------------------------------------- */
int bflow_voxelize(const void* x, const void* y, int xy_is_float, const unsigned char* pol, const long long* time, long long n_events,
                   long long t0_center, long long t1_center, int channels, int H, int W, float* out, int* oob_count, void* stream);
int bflow_voxel_norm(float* voxel, long long numel, double* stats3, void* stream);
int bflow_epe_masked(const float* src, const float* tgt, const unsigned char* valid, int N, int C, long long HW, double* sum_count, void* stream);
/* Row (f2), the remaining flow metrics in ONE pass over NCHW (N, C, HW) tensors — n_pixel_error_masked (utils/metrics.py:161-193),
 * epe_masked (:196-213), ae_masked (:259-296), and, through src_scale, the linear-assumption baseline (:298-305: prediction at
 * time t = t * final flow).  With s = src_scale * src, e = |s - tgt|_2 and the sums running over the valid pixels:
 *   out[0] += sum e            out[1] += number of valid pixels
 *   out[2] += sum acos(clamp((<s, tgt> + 1) / (sqrt(|s|^2 + 1) * sqrt(|tgt|^2 + 1)), -1, 1))        (radians)
 *   out[3 + k] += #{ e > thresholds[k]  and  e / max(|tgt|_2, 1e-6) >= 0.05 },  k < n_thresholds <= 4   (host array)
 * out: 8 doubles, zeroed by the caller.  valid: uint8/bool (N, HW) or NULL. */
int bflow_flow_metrics(const float* src, const float* tgt, const unsigned char* valid, int N, int C, long long HW, float src_scale,
                       const float* thresholds_host, int n_thresholds, double* out8, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-forward native entry: RAFTSpline.forward(voxel_grid, images, iters, test_mode=True) (models/raft_spline/raft.py:101-200) for ONE
 * (batch, height, width, iterations), replayed from a plan file without Python or torch.  `python -m bflow_b200.export` records the launch
 * list the Python engine replays (the same C entry points as above, in order, with their descriptors and tensor maps), packs the weights and
 * writes everything to the plan file; every device allocation of the exporter lives in one arena mapped at a FIXED virtual address with
 * the CUDA virtual-memory API, so the loader maps fresh memory at the same address and no pointer needs relocating.
 *   bflow_forward_load     maps the arena, restores the weights, replays the launch list once (warm-up) and captures it in a CUDA graph
 *                          (two-stream fork / join branches included).  Fails with BFLOW_ERR_CUDA if the address range is not free.
 *   bflow_forward_info     info16 = {B, voxel channels, H, W, h, w, 2*degree, iters, use_events, use_images, precision, correlation, 0, 0, 0, ABI}
 *   bflow_forward_run      copies the inputs in (host or device pointers; NULL flow_init = zeros; images NULL for an events-only plan),
 *                          launches the graph and copies low (B, 2*deg, H/8, W/8) and up (B, 2*deg, H, W) out (NULL = skip), all on
 *                          `stream`, asynchronously: synchronise the stream before reading host results.
 * The arena functions are the exporter's side (a torch.cuda.memory.CUDAPluggableAllocator binds bflow_arena_alloc / bflow_arena_free).
 * ------------------------------------------------------------------------------------------- */
typedef struct bflow_forward bflow_forward;
int bflow_forward_load(const char* plan_path, bflow_forward** out);
int bflow_forward_info(const bflow_forward* f, int* info16);
int bflow_forward_run(bflow_forward* f, const float* voxel, const float* image0, const float* image1, const float* flow_init,
                      float* low_out, float* up_out, void* stream);
void bflow_forward_destroy(bflow_forward* f);
int bflow_arena_open(unsigned long long base_address, unsigned long long reserve_bytes);
void* bflow_arena_alloc(long size, int device, void* stream);
void bflow_arena_free(void* ptr, long size, int device, void* stream);
unsigned long long bflow_arena_used(void);
int bflow_arena_read(unsigned long long offset, void* dst_host, unsigned long long bytes);
int bflow_arena_close(void);

#ifdef __cplusplus
}
#endif
#endif /* BFLOW_B200_H */
