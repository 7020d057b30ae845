"""Launch recording for the TMA-fed tensor-core path (the default on sm_100a).

Every tensor that feeds a convolution is kept as split-fp16 planes (x = hi + lo, ``_S16``): whoever produces it (conv
epilogues, InstanceNorm, the correlation lookup) writes the two planes directly, and ``bflow_conv2d_nhwc_tc3`` reads them
through im2col tensor maps.  fp32 copies exist only where something other than a convolution consumes the tensor: raw
pre-norm conv outputs, the GRU state ``h``, the Bezier parameters, z|r gates, the upsampling mask and the correlation
volume.  The structure of the forward pass is the same as in ``engine._Plan._record`` (models/raft_spline/raft.py:101-200).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACT, EPI, ConvDesc, check
from .ops import make_lookup_desc, make_lookup_otf_desc, tiled_plane_size, dev_zeros


class _S16:
    """Split-fp16 activation: two fp16 planes (hi, lo) of ``rows x ld`` halves.  ``base``: device pointer of a caller-owned
    region of ``rows*ld*4`` bytes, or None to allocate (zero-filled)."""

    def __init__(self, rows: int, ld: int, dev, base: Optional[int] = None, lo_base: Optional[int] = None):
        self.rows, self.ld = rows, ld
        if base is None:
            self.t = dev_zeros(2, rows, ld, device=dev, dtype=torch.float16)
            base = self.t.data_ptr()
        self.base = base
        self.lo_base = base + 2 * rows * ld if lo_base is None else lo_base

    def hi(self, c_off: int = 0) -> int:
        return self.base + 2 * c_off

    def lo(self, c_off: int = 0) -> int:
        return self.lo_base + 2 * c_off

    def rows_from(self, r0: int, nrows: int) -> '_S16':
        """View of rows [r0, r0 + nrows) (same planes)."""
        return _S16(nrows, self.ld, None, base=self.base + 2 * r0 * self.ld, lo_base=self.lo_base + 2 * r0 * self.ld)


def choose_bn(cout: int, n_mtiles: int) -> int:
    """Tile width: wide tiles keep the single MMA-issuing thread off the critical path; with few row tiles (the update
    block at batch 1: 38) narrower tiles put more CTAs to work."""
    if cout <= 64:
        return 64
    if n_mtiles >= 148:
        # one column tile that is narrower than 128: pack and multiply only `cout` weight rows (bflow_conv2d_nhwc_tc3: bn = 80 / 96 / 112
        # runs the 128-column kernel with N = 2*bn + bn per k-step -- the 96-channel encoder layers: 192 + 96 instead of 256 + 128)
        if 64 < cout < 128 and cout % 16 == 0:
            return cout
        return 128
    return 128 if (cout >= 256 and cout % 128 == 0) else 64


class S16Recorder:
    """Mixin of engine._Plan."""

    # ---- helpers -------------------------------------------------------------------------------------------------------
    def _maps(self, srcs: Sequence[Tuple[_S16, int, int]], N, H, W, wt) -> C.Array:
        buf = (C.c_uint8 * 512)()
        ph, pw = wt.pad
        for i, (s, c_off, cc) in enumerate(srcs):
            for j, base in enumerate((s.hi(c_off), s.lo(c_off))):
                check(self.eng.lib.bflow_tma_im2col_map(C.addressof(buf) + 128 * (2 * i + j), base, N, H, W, cc, s.ld, wt.kh, wt.kw, wt.stride, ph, pw),
                      'tma_im2col_map')
        self.keep.append(buf)
        return buf

    def _conv3(self, wt, srcs: Sequence[Tuple[_S16, int, int]], N, H, W, y=None, ldy=0, y16: Optional[Tuple[_S16, int]] = None,
               act1='none', act2='none', res=None, ldr=0, res16: Optional[Tuple[_S16, int]] = None, scale=1.0, bias=True,
               epi='std', aux0=None, ld_aux0=0, aux1_16: Optional[Tuple[_S16, int]] = None, stats=None, stats_hw=0):
        """Tensor-core convolution on split-fp16 sources [(tensor, channel offset, channels), ...] (channel concatenation)."""
        lib = self.eng.lib
        c0 = srcs[0][2]
        c1 = srcs[1][2] if len(srcs) > 1 else 0
        assert c0 + c1 == wt.cin, (c0, c1, wt.cin)
        d = ConvDesc()
        d.precision = self.eng.prec
        d.max_ctas = getattr(self, '_max_ctas', 0)
        d.x0, d.c0, d.ld0 = None, c0, srcs[0][0].ld
        d.x1, d.c1, d.ld1 = None, c1, (srcs[1][0].ld if c1 else 0)
        d.w, d.ldw, d.bias = None, 0, (wt.b.data_ptr() if bias else None)
        d.res, d.ldr = res, ldr
        d.y, d.ldy = y, ldy
        ph, pw = wt.pad
        Ho, Wo = (H + 2 * ph - wt.kh) // wt.stride + 1, (W + 2 * pw - wt.kw) // wt.stride + 1
        d.N, d.H, d.W, d.Ho, d.Wo, d.Cout = N, H, W, Ho, Wo, wt.cout
        d.KH, d.KW, d.stride, d.pad_h, d.pad_w = wt.kh, wt.kw, wt.stride, ph, pw
        d.act1, d.act2, d.scale = ACT[act1], ACT[act2], scale
        d.epi, d.aux0, d.ld_aux0, d.aux1, d.ld_aux1 = EPI[epi], aux0, ld_aux0, None, 0
        if y16 is not None:
            d.y16_hi, d.y16_lo, d.ldy16 = y16[0].hi(y16[1]), y16[0].lo(y16[1]), y16[0].ld
        if res16 is not None:
            d.res16_hi, d.res16_lo, d.ldr16 = res16[0].hi(res16[1]), res16[0].lo(res16[1]), res16[0].ld
        d.stats, d.stats_hw = stats, stats_hw
        if aux1_16 is not None:
            d.aux1_16_hi, d.aux1_16_lo, d.ld_aux1_16 = aux1_16[0].hi(aux1_16[1]), aux1_16[0].lo(aux1_16[1]), aux1_16[0].ld
        self.keep.append(d)
        M = N * Ho * Wo
        if (wt.kh == 3 and wt.kw == 3 and wt.stride == 1 and (ph, pw) == (1, 1) and c0 == 64 and c1 == 0 and wt.cout == 64 and W % 8 == 0 and
                epi == 'std' and res is None and act1 in ('none', 'relu') and act2 in ('none', 'relu') and M >= 148 * 128 and
                os.environ.get('BFLOW_SLAB', '1') != '0'):
            # layer1 of the encoders: weights resident in shared memory + halo slabs (bflow_conv2d_slab64)
            img, acc_scale = wt.tc3_image(64, c0)
            maps = (C.c_uint8 * 256)()
            src = srcs[0][0]
            for j, base in enumerate((src.hi(srcs[0][1]), src.lo(srcs[0][1]))):
                check(lib.bflow_tma_tile_map(C.addressof(maps) + 128 * j, base, N, H, W, 64, src.ld, 8, 18, 0), 'tma_tile_map')
            self.keep.append(maps)
            self._add(lib.bflow_conv2d_slab64, C.byref(d), C.addressof(maps), img.data_ptr(), acc_scale, self.eng.err.data_ptr(),
                      label=f'conv_slab64 64->64 3x3/1 M={M}', flops=2.0 * M * 64 * 9 * 64)
            self.n_tc += 1
            return Ho, Wo
        bn = choose_bn(wt.cout, (M + 127) // 128)
        img, acc_scale = wt.tc3_image(bn, c0)
        orient = {(3, 3): 1, (5, 1): 1, (1, 5): 2}.get((wt.kh, wt.kw), 0)
        if (orient and wt.stride == 1 and (ph, pw) == (wt.kh // 2, wt.kw // 2) and bn in (64, 128) and self.eng.prec == 0 and
                os.environ.get('BFLOW_TC3_SLAB', '0') == '1'):
            # halo slabs instead of im2col rows: A traffic / 2.7 (3x3) ... / 4 (1x5, 5x1).  Opt-in: measured on B200 it buys nothing here -- the
            # main loops of these layers are bound by the tensor pipe (bn 128) or by MMA issue (bn 64), not by L2 -> SM traffic
            taps = wt.kh if orient == 1 else wt.kw
            bw, bh = (8, 16 + taps - 1) if orient == 1 else (16 + taps - 1, 8)
            maps = (C.c_uint8 * 512)()
            for i, (s_, c_off, cc) in enumerate(srcs):
                for j, base in enumerate((s_.hi(c_off), s_.lo(c_off))):
                    check(lib.bflow_tma_tile_map(C.addressof(maps) + 128 * (2 * i + j), base, N, H, W, cc, s_.ld, bw, bh, 1 if orient == 2 else 0), 'tma_tile_map')
            self.keep.append(maps)
            self._add(lib.bflow_conv2d_nhwc_tc3s, C.byref(d), C.addressof(maps), img.data_ptr(), bn, acc_scale, orient, self.eng.err.data_ptr(),
                      label=f'conv_tc3s_{bn} {c0 + c1}->{wt.cout} {wt.kh}x{wt.kw}/{wt.stride} M={M}', flops=2.0 * M * wt.cout * wt.kh * wt.kw * (c0 + c1))
            self.n_tc += 1
            return Ho, Wo
        maps = self._maps(srcs, N, H, W, wt)
        if (epi == 'std' and res is None and res16 is None and stats is None and act1 in ('none', 'relu') and act2 in ('none', 'relu') and
                (M + 127) // 128 * ((wt.cout + bn - 1) // bn) <= 148 and wt.cout % 4 == 0 and os.environ.get('BFLOW_TC3_OSTORE', '1') != '0'):
            # single-tile launch with a plain epilogue: outputs leave through tensor-map stores
            omaps = (C.c_uint8 * 640)()
            if y16 is not None and y16[0].ld % 8 == 0:
                for j, base in enumerate((y16[0].hi(y16[1]), y16[0].lo(y16[1]))):
                    check(lib.bflow_tma_out_map(C.addressof(omaps) + 128 * j, base, M, wt.cout & ~7, y16[0].ld, 2), 'tma_out_map')      # a 4-channel tail is the kernel's
            if y is not None and ldy % 4 == 0:
                check(lib.bflow_tma_out_map(C.addressof(omaps) + 256, y, M, wt.cout, ldy, 4), 'tma_out_map')
            if (y16 is None or y16[0].ld % 8 == 0) and (y is None or ldy % 4 == 0):
                self.keep.append(omaps)
                self._add(lib.bflow_conv2d_nhwc_tc3o, C.byref(d), C.addressof(maps), C.addressof(omaps), img.data_ptr(), bn, acc_scale, self.eng.err.data_ptr(),
                          label=f'conv_tc3_{bn} {c0 + c1}->{wt.cout} {wt.kh}x{wt.kw}/{wt.stride} M={M}', flops=2.0 * M * wt.cout * wt.kh * wt.kw * (c0 + c1))
                self.n_tc += 1
                return Ho, Wo
        if (epi in ('gru_zr', 'gru_q') and res is not None and y is not None and wt.cout % bn == 0 and (M + 127) // 128 * (wt.cout // bn) <= 148 and
                ldy % 4 == 0 and ldr % 4 == 0 and ld_aux0 % 4 == 0):
            # GRU gate epilogues leave through tensor-map stores: {fp16 hi, fp16 lo, y fp32, res, aux0} (the last two only for BFLOW_TC3_GRU_TMA=1)
            t16 = aux1_16 if epi == 'gru_zr' else y16
            hdim = wt.cout // 2 if epi == 'gru_zr' else wt.cout
            if t16 is not None and t16[0].ld % 8 == 0 and hdim % 8 == 0:
                omaps = (C.c_uint8 * 640)()
                for j, base in enumerate((t16[0].hi(t16[1]), t16[0].lo(t16[1]))):
                    check(lib.bflow_tma_out_map(C.addressof(omaps) + 128 * j, base, M, hdim, t16[0].ld, 2), 'tma_out_map')
                check(lib.bflow_tma_out_map(C.addressof(omaps) + 256, y, M, wt.cout, ldy, 4), 'tma_out_map')
                check(lib.bflow_tma_out_map(C.addressof(omaps) + 384, res, M, wt.cout, ldr, 4), 'tma_out_map')
                check(lib.bflow_tma_out_map(C.addressof(omaps) + 512, aux0, M, hdim, ld_aux0, 4), 'tma_out_map')
                self.keep.append(omaps)
                self._add(lib.bflow_conv2d_nhwc_tc3o, C.byref(d), C.addressof(maps), C.addressof(omaps), img.data_ptr(), bn, acc_scale, self.eng.err.data_ptr(),
                          label=f'conv_tc3_{bn} {c0 + c1}->{wt.cout} {wt.kh}x{wt.kw}/{wt.stride} M={M}', flops=2.0 * M * wt.cout * wt.kh * wt.kw * (c0 + c1))
                self.n_tc += 1
                return Ho, Wo
        self._add(lib.bflow_conv2d_nhwc_tc3, C.byref(d), C.addressof(maps), img.data_ptr(), bn, acc_scale, self.eng.err.data_ptr(),
                  label=f'conv_tc3_{bn} {c0 + c1}->{wt.cout} {wt.kh}x{wt.kw}/{wt.stride} M={M}', flops=2.0 * M * wt.cout * wt.kh * wt.kw * (c0 + c1))
        self.n_tc += 1
        return Ho, Wo

    def _conv_simt16(self, wt, x0, c0, ld0, N, H, W, y=None, ldy=0, y16=None, act1='none', res=None, ldr=0, kernel='auto'):
        """CUDA-core convolution on an fp32 source (7x7 stems, convf1, the tiny Bezier head conv), fp32 and/or split output."""
        lib = self.eng.lib
        d = ConvDesc()
        d.x0, d.c0, d.ld0 = x0, c0, ld0
        d.x1, d.c1, d.ld1 = None, 0, 0
        d.w, d.ldw, d.bias = wt.w.data_ptr(), wt.ldw, wt.b.data_ptr()
        d.res, d.ldr = res, ldr
        d.y, d.ldy = y, ldy
        ph, pw = wt.pad
        Ho, Wo = (H + 2 * ph - wt.kh) // wt.stride + 1, (W + 2 * pw - wt.kw) // wt.stride + 1
        d.N, d.H, d.W, d.Ho, d.Wo, d.Cout = N, H, W, Ho, Wo, wt.cout
        d.KH, d.KW, d.stride, d.pad_h, d.pad_w = wt.kh, wt.kw, wt.stride, ph, pw
        d.act1, d.act2, d.scale, d.epi = ACT[act1], 0, 1.0, 0
        if y16 is not None:
            d.y16_hi, d.y16_lo, d.ldy16 = y16[0].hi(y16[1]), y16[0].lo(y16[1]), y16[0].ld
        self.keep.append(d)
        fn = {'small_n': lib.bflow_conv2d_small_n, 'thin7': lib.bflow_conv2d_thin7}.get(kernel, lib.bflow_conv2d_nhwc)
        name = {'small_n': 'conv_small_n', 'thin7': 'conv_thin7'}.get(kernel, 'conv_simt')
        self._add(fn, C.byref(d), label=f'{name} {c0}->{wt.cout} {wt.kh}x{wt.kw}/{wt.stride} M={N * Ho * Wo}',
                  flops=2.0 * N * Ho * Wo * wt.cout * wt.kh * wt.kw * c0)
        return Ho, Wo

    def _stem_window(self, src_nchw: torch.Tensor, C_total: int, c_off: int, cin: int, ns: int, H: int, W: int, scale: float = 1.0,
                     shift: float = 0.0, shared: Optional[dict] = None) -> dict:
        """How one input window (channels [c_off, c_off+cin) of an NCHW fp32 input) reaches the 7x7 stem."""
        from .engine import _ceil
        L, dev = self.eng.lib, self.eng.device
        shared = {} if shared is None else shared
        if 49 * cin <= 256:
            assert W % 4 == 0, 'bflow_b200: the fused stem needs W % 4 == 0 (the model itself needs H, W % 8 == 0)'
            # fused stem: input footprint -> patch matrix in shared memory -> tcgen05 (bflow_conv2d_stem7)
            return dict(kind='stem7', src=src_nchw.data_ptr(), C_total=C_total, c_off=c_off, cin=cin, ns=ns, scale=scale, shift=shift)
        if 49 * cin <= 512:
            if 'buf' not in shared:      # one patch-matrix buffer per encoder call, reused window after window (same stream)
                shared['buf'] = _S16(ns * (H // 2) * (W // 2), _ceil(49 * cin, 64), dev)
                self.keep.append(shared['buf'])
            return dict(kind='im2col', src=src_nchw.data_ptr(), C_total=C_total, c_off=c_off, cin=cin, ns=ns, scale=scale, shift=shift, buf=shared['buf'])
        if c_off % 8 == 0 and cin >= 16:
            key = ('x16', src_nchw.data_ptr())
            if key not in shared:        # NCHW fp32 -> NHWC fp32 -> split planes, once per source tensor
                ld = _ceil(C_total, 8)
                x32 = dev_zeros(ns, H, W, ld, device=dev, dtype=torch.float32)
                x16 = _S16(ns * H * W, ld, dev)
                self._add(L.bflow_nchw_to_nhwc, src_nchw.data_ptr(), x32.data_ptr(), ns, C_total, H, W, 0, C_total, ld, scale, shift)
                self._add(L.bflow_split_f16, x32.data_ptr(), ld, x16.hi(), x16.lo(), ld, ns * H * W, ld)
                self.keep += [x32, x16]
                shared[key] = x16
            return dict(kind='tma', x16=shared[key], c_off=c_off, cin=cin, ns=ns)
        key = ('x32', src_nchw.data_ptr())
        if key not in shared:
            x32 = dev_zeros(ns, H, W, C_total, device=dev, dtype=torch.float32)
            self._add(L.bflow_nchw_to_nhwc, src_nchw.data_ptr(), x32.data_ptr(), ns, C_total, H, W, 0, C_total, C_total, scale, shift)
            shared[key] = x32
            self.keep.append(x32)
        return dict(kind='simt', ptr=shared[key].data_ptr() + 4 * c_off, cin=cin, ld=C_total, ns=ns)

    # ---- encoders (models/raft_utils/extractor.py:47-55,103-125) ---------------------------------------------------------
    def _encoder16(self, E: dict, windows, Np: int, H: int, W: int, pool: List[int], final):
        """pool: device pointers of >= 5 scratch regions of Np*(H/2)*(W/2)*64*4 bytes; each holds either a raw fp32 conv output or a
        split-fp16 activation of the same element count."""
        L = self.eng.lib
        dev = self.eng.device
        inorm = E['kind'] == 'instance'
        free = list(pool)

        def s16(ptr, rows, c):
            return _S16(rows, c, dev, base=ptr)

        f16 = self.eng.prec == 1
        lo = (lambda t: None) if f16 else (lambda t: t.lo())       # BFLOW_PREC_F16: lo planes are neither written nor read

        H2, W2 = H // 2, W // 2
        rows = Np * H2 * W2
        w1 = E['conv1']
        # consecutive fused-stem windows over the same source tensor (same sample count) collapse into one launch
        merged = []
        for win in windows:
            last = merged[-1] if merged else None
            if (last is not None and win['kind'] == 'stem7' and last['kind'] == 'stem7' and last['src'] == win['src'] and last['ns'] == win['ns'] and
                    last['cin'] == win['cin'] and len(last['group']) < 8):
                last['group'].append(win)
            elif win['kind'] == 'stem7':
                merged.append(dict(win, group=[win]))
            else:
                merged.append(win)
        windows = [dict(w_, ns=sum(g_['ns'] for g_ in w_['group'])) if w_.get('group') else w_ for w_ in merged]
        # stem: 7x7 stride-2 conv (extractor.py:112), one launch group per input window.  Three forms:
        #   'im2col'  few input channels (K = 49*cin <= 512): patch matrix in split-fp16 + 1x1 tensor-core GEMM
        #   'tma'     cin >= 16 at an 8-aligned channel offset of a split-fp16 NHWC input: 7x7 im2col-TMA convolution
        #   'simt'    fp32 NHWC input on CUDA cores
        def stem(win, n0, ns, y=None, y16=None, act='none', stats=None):
            if win['kind'] == 'stem7':
                wins = win.get('group', [win])           # several channel windows of the same input tensor: one launch
                offs = (C.c_int * len(wins))(*[w_['c_off'] for w_ in wins])
                self.keep.append(offs)
                ns = sum(w_['ns'] for w_ in wins)
                wm = E['conv1_mat']
                img, acc_scale = wm.tc3_image(64, wm.cin)
                d = ConvDesc()
                d.precision = self.eng.prec
                d.max_ctas = getattr(self, '_max_ctas', 0)
                d.x0, d.c0, d.c1, d.bias = win['src'], win['cin'], 0, wm.b.data_ptr()
                d.y, d.ldy = y, 64
                d.N, d.H, d.W, d.Ho, d.Wo, d.Cout = ns, H, W, H2, W2, 64
                d.KH, d.KW, d.stride, d.pad_h, d.pad_w = 7, 7, 2, 3, 3
                d.act1, d.act2, d.scale, d.epi = ACT[act], 0, 1.0, 0
                if y16 is not None:
                    d.y16_hi, d.y16_lo, d.ldy16 = y16[0].hi(y16[1]), y16[0].lo(y16[1]), y16[0].ld
                d.stats, d.stats_hw = stats, H2 * W2
                self.keep.append(d)
                self._add(L.bflow_conv2d_stem7, C.byref(d), img.data_ptr(), win['C_total'], offs, len(wins),
                          win.get('scale', 1.0), win.get('shift', 0.0), acc_scale, self.eng.err.data_ptr(),
                          label=f'conv_stem7 {win["cin"]}->64 7x7/2 M={ns * H2 * W2}', flops=2.0 * ns * H2 * W2 * 64 * 49 * win['cin'])
                self.n_tc += 1
            elif win['kind'] == 'im2col':
                wm = E['conv1_mat']
                buf = win['buf']
                self._add(L.bflow_im2col_split16, win['src'], win['C_total'], win['c_off'], win['cin'], ns, H, W, 7, 7, 2, 3, 3, win.get('scale', 1.0),
                          win.get('shift', 0.0), buf.hi(), buf.lo(), buf.ld)
                self._conv3(wm, [(buf, 0, wm.cin)], 1, 1, ns * H2 * W2, y=y, ldy=64, y16=y16, act1=act, stats=stats, stats_hw=H2 * W2)
            elif win['kind'] == 'tma':
                self._conv3(w1, [(win['x16'], win['c_off'], win['cin'])], ns, H, W, y=y, ldy=64, y16=y16, act1=act, stats=stats)
            else:
                self._conv_simt16(w1, win['ptr'], win['cin'], win['ld'], ns, H, W, y=y, ldy=64, y16=y16, act1=act)

        if inorm:
            raw = free.pop()
            sm = self._sums(Np, 64)
            fused = all(win['kind'] != 'simt' for win in windows)      # tensor-core stems accumulate the IN statistics themselves
            n0 = 0
            for win in windows:
                stem(win, n0, win['ns'], y=raw + n0 * H2 * W2 * 64 * 4, stats=(sm + n0 * 64 * 2 * 8) if fused else None)
                n0 += win['ns']
            xptr = free.pop()
            X = s16(xptr, rows, 64)
            if not fused:
                self._add(L.bflow_plane_sums, raw, 64, sm, Np, H2 * W2, 64)
            self._add(L.bflow_instnorm_relu16, raw, 64, sm, None, 0, None, None, None, 0, None, 0, X.hi(), lo(X), 64, Np, H2 * W2, 64, 1e-5)
            free.append(raw)
        else:
            xptr = free.pop()
            X = s16(xptr, rows, 64)
            n0 = 0
            for win in windows:
                stem(win, n0, win['ns'], y16=(X.rows_from(n0 * H2 * W2, win['ns'] * H2 * W2), 0), act='relu')
                n0 += win['ns']
        Hc, Wc, Cc = H2, W2, 64
        for blk in E['blocks']:
            c1, c2, dn = blk['conv1'], blk['conv2'], blk['down']
            Co = c1.cout
            st = c1.stride
            Ho, Wo = Hc // st, Wc // st
            rin, rout = Np * Hc * Wc, Np * Ho * Wo
            if inorm:
                raw1 = free.pop()
                sm = self._sums(Np, Co)
                self._conv3(c1, [(X, 0, Cc)], Np, Hc, Wc, y=raw1, ldy=Co, stats=sm)
                y1p = free.pop()
                Y1 = s16(y1p, rout, Co)
                self._add(L.bflow_instnorm_relu16, raw1, Co, sm, None, 0, None, None, None, 0, None, 0, Y1.hi(), lo(Y1), Co, Np, Ho * Wo, Co, 1e-5)
                raw2 = raw1                                     # raw1 is dead: reuse it for conv2's output
                sm2 = self._sums(Np, Co)
                self._conv3(c2, [(Y1, 0, Co)], Np, Ho, Wo, y=raw2, ldy=Co, stats=sm2)
                outp = y1p                                      # Y1 is dead after conv2: the block output takes its place
                OUT = s16(outp, rout, Co)
                if dn is not None:
                    rawd = free.pop()
                    smd = self._sums(Np, Co)
                    self._conv3(dn, [(X, 0, Cc)], Np, Hc, Wc, y=rawd, ldy=Co, stats=smd)
                    self._add(L.bflow_instnorm_relu16, raw2, Co, sm2, rawd, Co, smd, None, None, 0, None, 0, OUT.hi(), lo(OUT), Co, Np, Ho * Wo, Co, 1e-5)
                    free.append(rawd)
                else:
                    self._add(L.bflow_instnorm_relu16, raw2, Co, sm2, None, 0, None, X.hi(), lo(X), Cc, None, 0, OUT.hi(), lo(OUT), Co, Np, Ho * Wo, Co, 1e-5)
                free.append(raw2)
                free.append(xptr)
                X, xptr = OUT, outp
            else:
                y1p = free.pop()
                Y1 = s16(y1p, rout, Co)
                self._conv3(c1, [(X, 0, Cc)], Np, Hc, Wc, y16=(Y1, 0), act1='relu')
                outp = free.pop()
                OUT = s16(outp, rout, Co)
                if dn is not None:
                    dp = free.pop()
                    D = s16(dp, rout, Co)
                    self._conv3(dn, [(X, 0, Cc)], Np, Hc, Wc, y16=(D, 0))
                    self._conv3(c2, [(Y1, 0, Co)], Np, Ho, Wo, y16=(OUT, 0), act1='relu', act2='relu', res16=(D, 0))
                    free.append(dp)
                else:
                    self._conv3(c2, [(Y1, 0, Co)], Np, Ho, Wo, y16=(OUT, 0), act1='relu', act2='relu', res16=(X, 0))
                free.append(y1p)
                free.append(xptr)
                X, xptr = OUT, outp
            Hc, Wc, Cc = Ho, Wo, Co
        final(X, Np, Hc, Wc, Cc)

    def _record_volume(self, fm_ev, fm_img, fm_ev16, fm_img16, T_ev, T):
        """Correlation volume (tensor-core GEMM, granule-tiled planes) + pyramid (corr.py:264-272, 293-305) and the lookup descriptor."""
        eng, L = self.eng, self.eng.lib
        dev = eng.device
        B, h, w, Q, R = self.B, self.h, self.w, self.Q, self.R
        f32 = dict(device=dev, dtype=torch.float32)
        deg, fd = eng.deg, eng.fdim
        hx, poff, gw = self.hx.data_ptr(), self.poff, self.gw
        self.tiled = True
        bnv = 128
        Np0 = tiled_plane_size(h, w)
        self.vol0 = torch.empty(T, R, Np0, **f32)
        img_bytes = ((Np0 + bnv - 1) // bnv) * ((fd + 63) // 64) * 2 * bnv * 128
        self.f2img = dev_zeros(T, B, img_bytes, device=dev, dtype=torch.uint8)
        srcs = [(fm_ev, (t + 1) * B, fm_ev16) for t in range(T_ev)] + ([(fm_img, B, fm_img16)] if self.use_img else [])
        for t, (fm2, n0, fm1_16) in enumerate(srcs):
            for b in range(B):
                self._add(L.bflow_pack_b_tc, fm2.data_ptr() + (n0 + b) * Q * fd * 4, fd, self.f2img[t, b].data_ptr(), Q, fd, bnv, h, w)
                # corr[bq, n] = <f1[bq, :], f2[n, :]> / sqrt(D): a 1x1 "convolution" over the Q query pixels of sample b whose weight
                # image is the packed target feature map
                d = ConvDesc()
                d.precision = self.eng.prec
                d.max_ctas = getattr(self, '_vol_max_ctas', 0)
                d.c0, d.ld0 = fd, fd
                d.y, d.ldy = self.vol0[t].data_ptr() + b * Q * Np0 * 4, Np0
                d.N, d.H, d.W, d.Ho, d.Wo, d.Cout = 1, 1, Q, 1, Q, Np0
                d.KH, d.KW, d.stride = 1, 1, 1
                d.scale = 1.0 / (fd ** 0.5)
                self.keep.append(d)
                sub = fm1_16.rows_from(b * Q, Q)
                maps = (C.c_uint8 * 512)()
                for j, base in enumerate((sub.hi(), sub.lo())):
                    check(L.bflow_tma_im2col_map(C.addressof(maps) + 128 * j, base, 1, 1, Q, fd, fd, 1, 1, 1, 0, 0), 'tma_im2col_map')
                self.keep.append(maps)
                self._add(L.bflow_conv2d_nhwc_tc3, C.byref(d), C.addressof(maps), self.f2img[t, b].data_ptr(), bnv, 1.0, eng.err.data_ptr(),
                          label=f'corr_volume_tc3 Q={Q} D={fd}', flops=2.0 * Q * Q * fd)
        pyr = [(list(range(T)), self.vol0, h, w)]
        for lvl in range(1, max(eng.levels)):
            prev_idx, prev, hp_, wp_ = pyr[-1]
            keep = [t for t in range(T) if eng.levels[t] > lvl]
            hl, wl = hp_ // 2, wp_ // 2
            cur = torch.empty(len(keep), R, tiled_plane_size(hl, wl), **f32)
            for j, t in enumerate(keep):
                self._add(L.bflow_corr_pool_tiled, prev[prev_idx.index(t)].data_ptr(), cur[j].data_ptr(), R, hp_, wp_)
            pyr.append((keep, cur, hl, wl))
        self.pyr = pyr

        # ---- lookup: centres from the fp32 Bezier params, output written as split planes for convc1 ----
        slots = [(lvl, t, pyr[lvl][1][pyr[lvl][0].index(t)], pyr[lvl][2], pyr[lvl][3]) for (lvl, t) in eng.slots]
        ld = make_lookup_desc(slots, T, B, h, w, True)
        ld.coords = None
        ld.params, ld.params_ld, ld.degree = hx + poff * 4, gw, deg
        for t in range(T):
            for k in range(deg):
                ld.coef[t][k] = float(eng.coef[t, k])
        ld.out, ld.out_nhwc, ld.out_ld = None, 1, eng.ldc
        ld.out16_hi, ld.out16_lo, ld.out16_ld = self.corr16.hi(), (None if eng.prec == 1 else self.corr16.lo()), eng.ldc
        self.keep.append(ld)
        self.lookup_desc = ld
        self.lookup_fn = L.bflow_corr_lookup


    # ---- the whole forward ---------------------------------------------------------------------------------------------------
    def _record_s16(self):
        from .engine import _ceil
        eng, L = self.eng, self.eng.lib
        cfg, dev = eng.cfg, eng.device
        B, H, W, h, w, Q, R = self.B, self.H, self.W, self.h, self.w, self.Q, self.R
        f32 = dict(device=dev, dtype=torch.float32)
        deg, hd, cd, md, fd = eng.deg, eng.hdim, eng.cdim, eng.mdim, eng.fdim
        nctx, ncorr = cfg['num_bins']['context'], cfg['num_bins']['correlation']
        T_ev = len(cfg['correlation']['ev']['target_indices']) if self.use_ev else 0
        T = len(eng.levels)
        np_max = max((T_ev + 1) * B if self.use_ev else 0, 2 * B if self.use_img else 0, B)
        pool_bytes = np_max * (H // 2) * (W // 2) * 64 * 4
        self.pool = [torch.empty(pool_bytes, device=dev, dtype=torch.uint8) for _ in range(5)]
        pool = [t.data_ptr() for t in self.pool]
        # the context encoder runs beside the feature encoder on the second stream and needs its own scratch
        self.pool_c = [torch.empty(B * (H // 2) * (W // 2) * 64 * 4, device=dev, dtype=torch.uint8) for _ in range(5)]
        pool_c = [t.data_ptr() for t in self.pool_c]
        self.sums = dev_zeros(64 * np_max * 128 * 2, device=dev, dtype=torch.float64)
        self._sums_off = 0
        self._add(L.bflow_zero, self.sums.data_ptr(), self.sums.numel() * 8)

        gw = hd + cd + md
        poff = hd + cd + md - 2 * deg
        self.hx = dev_zeros(R, gw, **f32)                 # fp32 masters: h (cols 0:hd) and the Bezier params (cols poff:)
        self.hx16 = _S16(R, gw, dev)                        # what the convolutions read
        self.poff, self.gw = poff, gw
        hx, hx16 = self.hx.data_ptr(), self.hx16

        # ---- context encoder on the second stream: net = tanh -> h (fp32 master + split copy), inp = relu -> split only
        #      (raft.py:144-147); then the iteration-invariant GRU terms conv(inp) + bias ----
        U = eng.upd
        self._fork()
        # The two encoders are the same network on 1 (context) and T+1 or 2 (features) images per sample.  Left alone, their one-CTA-per-SM
        # launches queue behind each other and each pays its own wave quantisation (a 600-tile context layer needs 5 waves of 148 for
        # 4.05 waves of work); with disjoint CTA budgets the two chains run side by side.  BFLOW_ENC_SPLIT = SMs of the context chain.
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        # measured optimum at D, batch 1: 20 of 148 SMs (f32x3), 14 (f16: the 1-image context chain gains more from the single MMA than the feature chain)
        split = int(os.environ.get('BFLOW_ENC_SPLIT', '14' if eng.prec == 1 else '20')) if eng.use_side_stream else 0
        split = max(0, min(split, n_sm // 2))
        self._max_ctas = split
        ctx_c = (nctx if self.use_ev else 0) + (3 if self.use_img else 0)
        if self.use_ev and self.use_img:
            # context = cat(voxel[:, -nctx:], image0) (raft.py:137-138): NHWC fp32 -> split planes -> 7x7 im2col-TMA stem
            ctx_ld = _ceil(ctx_c, 8)
            self.ctx = dev_zeros(B, H, W, ctx_ld, **f32)
            self.ctx16 = _S16(B * H * W, ctx_ld, dev)
            self._add(L.bflow_nchw_to_nhwc, self.voxel_in.data_ptr(), self.ctx.data_ptr(), B, self.cin_vox, H, W, self.cin_vox - nctx, nctx, ctx_ld, 1.0, 0.0)
            self._add(L.bflow_nchw_to_nhwc, self.img_in[0].data_ptr(), self.ctx.data_ptr() + nctx * 4, B, 3, H, W, 0, 3, ctx_ld, 2.0 / 255.0, -1.0)
            self._add(L.bflow_split_f16, self.ctx.data_ptr(), ctx_ld, self.ctx16.hi(), self.ctx16.lo(), ctx_ld, B * H * W, ctx_ld)
            cwin = [dict(kind='tma', x16=self.ctx16, c_off=0, cin=ctx_c, ns=B)]
        elif self.use_ev:
            cwin = [self._stem_window(self.voxel_in, self.cin_vox, self.cin_vox - nctx, nctx, B, H, W)]
        else:
            cwin = [self._stem_window(self.img_in[0], 3, 0, 3, B, H, W, 2.0 / 255.0, -1.0)]
        Ec = eng.enc['cnet']

        def cnet_final(X, n_, Hc, Wc, Cc):
            self._conv3(Ec['conv2'][0], [(X, 0, Cc)], n_, Hc, Wc, y=hx, ldy=gw, y16=(hx16, 0), act1='tanh')
            self._conv3(Ec['conv2'][1], [(X, 0, Cc)], n_, Hc, Wc, y16=(hx16, hd), act1='relu')
        self._encoder16(Ec, cwin, B, H, W, pool_c, cnet_final)
        self.pre = {k: torch.empty(R, (2 * hd if k.startswith('zr') else hd), **f32) for k in ('zr1', 'q1', 'zr2', 'q2')}
        pre = {k: v.data_ptr() for k, v in self.pre.items()}
        for k in ('zr1', 'q1', 'zr2', 'q2'):
            self._conv3(U[k + '_inp'], [(hx16, hd, cd)], B, h, w, y=pre[k], ldy=self.pre[k].shape[1])
        # initial Bezier parameters: zeros (+ flow_init) (raft.py:150-153): fp32 master + split copy
        self._add(L.bflow_nchw_to_nhwc, self.init_in.data_ptr(), hx + poff * 4, B, 2 * deg, h, w, 0, 2 * deg, gw, 1.0, 0.0)
        self._add(L.bflow_split_f16, hx + poff * 4, gw, hx16.hi(poff), hx16.lo(poff), gw, R, 2 * deg)
        self._main()
        self._max_ctas = n_sm - split if split else 0

        # ---- feature encoders: fp32 feature map for the volume GEMM's B operand, split copy for its A operand ----
        def fnet(name, windows, Np):
            fm = torch.empty(Np, h, w, fd, **f32)
            fm16 = _S16(Np * Q, fd, dev)
            E = eng.enc[name]
            self._encoder16(E, windows, Np, H, W, pool,
                            lambda X, n_, Hc, Wc, Cc: self._conv3(E['conv2'][0], [(X, 0, Cc)], n_, Hc, Wc, y=fm.data_ptr(), ldy=fd, y16=(fm16, 0)))
            return fm, fm16
        fm_ev = fm_img = fm_ev16 = fm_img16 = None
        if self.use_ev:
            idxs = [0] + list(cfg['correlation']['ev']['target_indices'])
            shared = {}
            fm_ev, fm_ev16 = fnet('fnet_ev', [self._stem_window(self.voxel_in, self.cin_vox, i, ncorr, B, H, W, shared=shared) for i in idxs], (T_ev + 1) * B)
        if self.use_img:
            shared = {}
            fm_img, fm_img16 = fnet('fnet_img', [self._stem_window(self.img_in[i], 3, 0, 3, B, H, W, 2.0 / 255.0, -1.0, shared=shared) for i in range(2)], 2 * B)
        self.fm_ev, self.fm_img, self.fm_ev16, self.fm_img16 = fm_ev, fm_img, fm_ev16, fm_img16
        # the context chain is still running when the feature encoder is done: the all-pairs GEMMs keep to the feature chain's SMs too (a
        # 148-CTA launch whose last 20 CTAs start late was measured at 58-64 us instead of 38)
        self._vol_max_ctas = self._max_ctas
        self._max_ctas = 0

        self.corr16 = _S16(R, eng.ldc, dev)                 # zero-filled: the channels padding S*81 up to ldc stay zero
        if eng.corr_mode == 'otf':
            # ---- row (f3): no volume.  Pooled TARGET FEATURE pyramid (avg_pool2d is linear: pooled features give the pooled correlation
            #      planes of corr.py:119 exactly) + on-the-fly lookup (bflow_corr_lookup_otf) ----
            self.tiled = False
            feats = [(fm_ev, 0, (t + 1) * B) for t in range(T_ev)] + ([(fm_img, 0, B)] if self.use_img else [])     # (tensor, f1 image, f2 image)
            pyr_f = []                                      # per target: [level-0 view, pooled level 1, ...]
            self.fpool = []
            for t, (fm, i1, i2) in enumerate(feats):
                lv = [fm[i2:i2 + B]]
                hh, ww = h, w
                for _ in range(1, eng.levels[t]):
                    nxt = torch.empty(B, hh // 2, ww // 2, fd, **f32)
                    self._add(L.bflow_feat_pool, lv[-1].data_ptr(), nxt.data_ptr(), B, hh, ww, fd, fd, fd)
                    lv.append(nxt)
                    hh, ww = hh // 2, ww // 2
                pyr_f.append(lv)
                self.fpool.append(lv)
            slots = [(lvl, t, feats[t][0][feats[t][1]:feats[t][1] + B], pyr_f[t][lvl]) for (lvl, t) in eng.slots]
            ld = make_lookup_otf_desc(slots, T, B, h, w, fd)
            ld.coords = None
            ld.params, ld.params_ld, ld.degree = hx + poff * 4, gw, deg
            for t in range(T):
                for k in range(deg):
                    ld.coef[t][k] = float(eng.coef[t, k])
            ld.out, ld.out_ld = None, 0
            ld.out16_hi, ld.out16_lo, ld.out16_ld = self.corr16.hi(), (None if eng.prec == 1 else self.corr16.lo()), eng.ldc
            self.keep.append(ld)
            self.lookup_desc = ld
            self.lookup_fn = L.bflow_corr_lookup_otf
        else:
            self._record_volume(fm_ev, fm_img, fm_ev16, fm_img16, T_ev, T)

        self._join()                                        # context branch done

        # ---- update-block workspace ----
        self.c1_16 = _S16(R, 256, dev)
        self.cb16 = _S16(R, 256, dev)
        self.f1_16 = _S16(R, 128, dev)
        self.rh16 = _S16(R, hd, dev)
        self.hm16 = _S16(R, 256, dev)
        self.zr = torch.empty(R, 2 * hd, **f32)
        self.hh = torch.empty(R, 256, **f32)
        self.mask = torch.empty(R, 576, **f32)
        zr, hh, mk = self.zr.data_ptr(), self.hh.data_ptr(), self.mask.data_ptr()
        c1_16, cb16, f1_16, rh16, hm16 = self.c1_16, self.cb16, self.f1_16, self.rh16, self.hm16

        def mask_head():
            self._conv3(U['mask0'], [(hx16, 0, hd)], B, h, w, y16=(hm16, 0), act1='relu')
            self._conv3(U['mask2'], [(hm16, 0, 256)], B, h, w, y=mk, ldy=576, scale=0.25)

        def upsample(out_t, mask_done=False):
            if not mask_done:
                mask_head()
            self._add(L.bflow_cvx_upsample, hx + poff * 4, gw, 0, mk, 576, 0, out_t.data_ptr(), B, 2 * deg, h, w)

        self.iter_start = len(self.launches)
        for itr in range(self.iters):
            if itr == 1:
                self.iter_len = len(self.launches) - self.iter_start
            # motion encoder (update.py:88-97): the Bezier branch (convf1 -> convf2) runs beside lookup -> convc1 -> convc2.  (Forking after
            # the lookup instead was measured: the lookup drops from 21 to 15 us, but convc1 then starts ~10 us late behind the fork and
            # convc2 shares SMs with convf2: no net gain.)
            self._fork()
            thin = U['convf1'].cout == 128 and (2 * deg) % 4 == 0 and poff % 4 == 0
            self._conv_simt16(U['convf1'], hx + poff * 4, 2 * deg, gw, B, h, w, y16=(f1_16, 0), act1='relu', kernel='thin7' if thin else 'auto')
            self._conv3(U['convf2'], [(f1_16, 0, 128)], B, h, w, y16=(cb16, 192), act1='relu')
            self._main()
            self._add(self.lookup_fn, C.byref(self.lookup_desc), label='corr_lookup' if eng.corr_mode != 'otf' else 'corr_lookup_otf')
            self._conv3(U['convc1'], [(self.corr16, 0, eng.ldc)], B, h, w, y16=(c1_16, 0), act1='relu')
            self._conv3(U['convc2'], [(c1_16, 0, 256)], B, h, w, y16=(cb16, 0), act1='relu')
            self._join()
            self._conv3(U['conv'], [(cb16, 0, 256)], B, h, w, y16=(hx16, hd + cd), act1='relu')
            # SepConvGRU (update.py:33-48) with the iteration-invariant inp part hoisted and the gate arithmetic in the epilogues
            for sfx in '12':
                self._conv3(U['zr' + sfx + '_dyn'], [(hx16, 0, hd), (hx16, hd + cd, md)], B, h, w, y=zr, ldy=2 * hd, bias=False,
                            res=pre['zr' + sfx], ldr=2 * hd, epi='gru_zr', aux0=hx, ld_aux0=gw, aux1_16=(rh16, 0))
                self._conv3(U['q' + sfx + '_dyn'], [(rh16, 0, hd), (hx16, hd + cd, md)], B, h, w, y=hx, ldy=gw, y16=(hx16, 0), bias=False,
                            res=pre['q' + sfx], ldr=hd, epi='gru_q', aux0=zr, ld_aux0=2 * hd)
            # the mask head only reads the new hidden state: in test mode (last iteration only) it runs beside the Bezier head on the second stream
            side_mask = self.test_mode and itr == self.iters - 1 and 2 * deg <= 8
            if side_mask:
                self._fork()
                mask_head()
                self._main()
            # Bezier head + delta update in place (update.py:17-18, bezier.py:137-139): fp32 master and split copy
            if 2 * deg <= 8:      # tiny head: warp-per-pixel CUDA-core kernel on the fp32 hidden tensor
                self._conv3(U['head1'], [(hx16, 0, hd)], B, h, w, y=hh, ldy=256, act1='relu')
                self._conv_simt16(U['head2'], hh, 256, 256, B, h, w, y=hx + poff * 4, ldy=gw, y16=(hx16, poff), res=hx + poff * 4, ldr=gw, kernel='small_n')
            else:
                self._conv3(U['head1'], [(hx16, 0, hd)], B, h, w, y16=(hm16, 0), act1='relu')
                self._conv3(U['head2'], [(hm16, 0, 256)], B, h, w, y=hx + poff * 4, ldy=gw, y16=(hx16, poff), res=hx + poff * 4, ldr=gw)
            if side_mask:
                self._join()
            if not self.test_mode:
                upsample(self.ups[itr])
        if self.iters == 1:
            self.iter_len = len(self.launches) - self.iter_start
        if self.test_mode:
            upsample(self.ups[0], mask_done=2 * deg <= 8)
        self._add(L.bflow_nhwc_to_nchw, hx + poff * 4, self.low.data_ptr(), B, 2 * deg, h, w, gw)
        self.n_launches = len(self.launches)
