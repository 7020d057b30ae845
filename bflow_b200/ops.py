"""Tensor-level wrappers over the C ABI: argument checks (dtype/device/contiguity → AssertionError, the
reference's error convention), output allocation, current-stream plumbing.  These are the reference-level
operator mirrors (NCHW layouts of models/raft_utils/corr.py, utils.py and bezier.py) used by :class:`BezierCurves`
and by the parity tests; the engine calls the ABI directly.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import ACT, ConvDesc, LookupDesc, LookupOtfDesc, check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def dev_zeros(*shape, device, dtype=torch.float32) -> torch.Tensor:
    """Zero-filled device tensor through cudaMemsetAsync (bflow_zero) instead of an ATen fill kernel: building a plan then launches
    nothing but this library's kernels, which is what a launch capture of a forward should show first."""
    t = torch.empty(*shape, device=device, dtype=dtype)
    if t.numel():
        with torch.cuda.device(t.device):
            check(_lib.lib().bflow_zero(t.data_ptr(), t.numel() * t.element_size(), torch.cuda.current_stream(t.device).cuda_stream), 'zero')
    return t


def _on_tensor_device(fn):
    """Runs the wrapped operator with the device of its first CUDA tensor argument current: allocations, the stream handle and the
    `<<<>>>` launches inside the library all follow the CURRENT device, which need not be the tensors' device."""
    import functools

    def first_cuda(objs):
        for o in objs:
            if isinstance(o, torch.Tensor) and o.is_cuda:
                return o
            if isinstance(o, (list, tuple)):
                t = first_cuda(o)
                if t is not None:
                    return t
        return None

    @functools.wraps(fn)
    def wrapper(*a, **k):
        t = first_cuda(list(a) + list(k.values()))
        if t is None:
            return fn(*a, **k)
        with torch.cuda.device(t.device):
            return fn(*a, **k)
    return wrapper


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    assert isinstance(t, torch.Tensor), f'{name}: tensor expected'
    assert t.is_cuda, f'{name}: CUDA tensor expected (bflow_b200 has no CPU path)'
    assert t.dtype == torch.float32, f'{name}: float32 expected'
    return t.contiguous()


# ---- layout ------------------------------------------------------------------------------------------
@_on_tensor_device
def nchw_to_nhwc(x: torch.Tensor, c_off: int = 0, c_cnt: Optional[int] = None, out: Optional[torch.Tensor] = None,
                 out_ld: Optional[int] = None, scale: float = 1.0, shift: float = 0.0) -> torch.Tensor:
    x = _f32c(x, 'x')
    N, Ct, H, W = x.shape
    c_cnt = Ct - c_off if c_cnt is None else c_cnt
    ld = c_cnt if out_ld is None else out_ld
    if out is None:
        out = torch.zeros(N, H, W, ld, device=x.device, dtype=torch.float32)
    check(_lib.lib().bflow_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), N, Ct, H, W, c_off, c_cnt, ld, scale, shift, _stream()),
          'nchw_to_nhwc')
    return out


@_on_tensor_device
def nhwc_to_nchw(x: torch.Tensor, C_: Optional[int] = None) -> torch.Tensor:
    x = _f32c(x, 'x')
    N, H, W, ld = x.shape
    C_ = ld if C_ is None else C_
    out = torch.empty(N, C_, H, W, device=x.device, dtype=torch.float32)
    check(_lib.lib().bflow_nhwc_to_nchw(x.data_ptr(), out.data_ptr(), N, C_, H, W, ld, _stream()), 'nhwc_to_nchw')
    return out


# ---- convolution ----------------------------------------------------------------------------------------
def pack_conv_weight(w: torch.Tensor, cin_pad: Optional[int] = None) -> Tuple[torch.Tensor, int]:
    """OIHW → K-major [(kh*KW+kw)*Cin + c][ldw] with Cout zero-padded to a multiple of 4 (and Cin to cin_pad)."""
    O, I, KH, KW = w.shape
    ldw = (O + 3) // 4 * 4
    ci = I if cin_pad is None else cin_pad
    p = torch.zeros(KH, KW, ci, ldw, device=w.device, dtype=torch.float32)
    p[:, :, :I, :O] = w.detach().float().permute(2, 3, 1, 0)
    return p.reshape(KH * KW * ci, ldw).contiguous(), ldw


def pack_conv_weight_tc(w: torch.Tensor, bn: int, cin_pad: Optional[int] = None, prescale: bool = True,
                        block_per_tap: bool = False, c0: Optional[int] = None) -> Tuple[torch.Tensor, float]:
    """OIHW fp32 → (tensor-core weight image of bflow_conv2d_nhwc_tc3, acc_scale).
    Image: [ceil(O/bn)][ceil(K/64)][hi | lo (fp16)][bn][64] with K = (kh*KW+kw)*Cin + c flattened and the 16-byte
    chunks of every 128-byte row XOR-swizzled by (row % 8) — byte for byte the SWIZZLE_128B shared-memory tile.
    The weights are multiplied by 2^k (largest magnitude in [0.5, 1)) so that the fp16 residuals stay normal;
    acc_scale = 2^-k is applied to the fp32 accumulator (exact)."""
    O, I, KH, KW = w.shape
    ci = I if cin_pad is None else cin_pad
    w = w.detach().float()
    if block_per_tap:
        # K order of the TMA-fed kernel: (tap, 64-channel block); each of the (up to two) concatenated sources is padded to a
        # multiple of 64 channels on its own
        c0 = ci if c0 is None else c0
        c1 = ci - c0
        p0, p1 = (c0 + 63) // 64 * 64, (c1 + 63) // 64 * 64
        wp = torch.zeros(O, KH, KW, p0 + p1, device=w.device, dtype=torch.float32)
        wsrc = torch.zeros(O, KH, KW, ci, device=w.device, dtype=torch.float32)
        wsrc[..., :I] = w.permute(0, 2, 3, 1)
        wp[..., :c0] = wsrc[..., :c0]
        if c1 > 0:
            wp[..., p0:p0 + c1] = wsrc[..., c0:]
        w = wp.permute(0, 3, 1, 2).contiguous()
        O, I, KH, KW = w.shape
        ci = I
    K = KH * KW * ci
    nkb, nt = (K + 63) // 64, (O + bn - 1) // bn
    k = 0
    amax = float(w.abs().max())
    if prescale and amax > 0:
        k = max(-16, min(16, int(math.floor(-math.log2(amax)))))
    wk = torch.zeros(nt * bn, nkb * 64, device=w.device, dtype=torch.float32)
    wp = torch.zeros(O, KH, KW, ci, device=w.device, dtype=torch.float32)
    wp[..., :I] = (w * (2.0 ** k)).permute(0, 2, 3, 1)
    wk[:O, :K] = wp.reshape(O, K)
    hi = wk.clamp(-65504.0, 65504.0).to(torch.float16)
    lo = (wk - hi.float()).clamp(-65504.0, 65504.0).to(torch.float16)
    r = torch.arange(bn, device=w.device) % 8
    src_chunk = (torch.arange(8, device=w.device)[None, :] ^ r[:, None])          # dst chunk j <- source chunk j ^ (row % 8)
    idx = src_chunk[None, None, :, :, None].expand(nt, nkb, bn, 8, 8)

    def tile(x):
        x = x.view(torch.int16).view(nt, bn, nkb, 8, 8).permute(0, 2, 1, 3, 4)     # tile, k-block, row, chunk, element
        return torch.gather(x, 3, idx)
    img = torch.stack([tile(hi), tile(lo)], dim=2).contiguous()                    # tile, k-block, hi|lo, row, chunk, element
    return img.view(-1), 2.0 ** (-k)


@_on_tensor_device
def split_f16(x_nhwc: torch.Tensor, ld16: Optional[int] = None) -> torch.Tensor:
    """(..., C) fp32 rows → (2, rows, ld16) fp16 planes [hi, lo] with x = hi + lo."""
    x = _f32c(x_nhwc, 'x')
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    ld16 = Cc if ld16 is None else ld16
    out = torch.zeros(2, rows, ld16, device=x.device, dtype=torch.float16)
    check(_lib.lib().bflow_split_f16(x.data_ptr(), Cc, out[0].data_ptr(), out[1].data_ptr(), ld16, rows, Cc, _stream()), 'split_f16')
    return out


def tma_im2col_maps(planes: torch.Tensor, N: int, H: int, W: int, Cc: int, KH: int, KW: int, stride: int, ph: int, pw: int,
                    c_off: int = 0):
    """Host buffer with the {hi, lo} im2col tensor maps of a split-fp16 activation (2, N*H*W, ld16), channels [c_off, c_off+C)."""
    ld16 = planes.shape[-1]
    buf = (C.c_uint8 * 256)()
    for i in range(2):
        base = planes[i].data_ptr() + c_off * 2
        check(_lib.lib().bflow_tma_im2col_map(C.addressof(buf) + 128 * i, base, N, H, W, Cc, ld16, KH, KW, stride, ph, pw), 'tma_im2col_map')
    return buf


@_on_tensor_device
def conv2d(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, stride: int = 1,
           padding=(0, 0), act: str = 'none', scale: float = 1.0, backend: str = 'simt', bn: int = 128) -> torch.Tensor:
    """NCHW in / NCHW out convenience form of bflow_conv2d_nhwc (used by tests and the operator mirror)."""
    x = _f32c(x, 'x')
    N, Cin, H, W = x.shape
    O, I, KH, KW = weight.shape
    assert I == Cin
    ph, pw = (padding, padding) if isinstance(padding, int) else padding
    Ho, Wo = (H + 2 * ph - KH) // stride + 1, (W + 2 * pw - KW) // stride + 1
    xh = nchw_to_nhwc(x)
    wp, ldw = pack_conv_weight(_f32c(weight, 'weight'))
    y = torch.empty(N, Ho, Wo, O, device=x.device, dtype=torch.float32)
    d = ConvDesc()
    d.x0, d.c0, d.ld0 = xh.data_ptr(), Cin, Cin
    d.x1, d.c1, d.ld1 = None, 0, 0
    d.w, d.ldw = wp.data_ptr(), ldw
    b = _f32c(bias, 'bias') if bias is not None else None
    d.bias = b.data_ptr() if b is not None else None
    d.res, d.ldr = None, 0
    d.y, d.ldy = y.data_ptr(), O
    d.N, d.H, d.W, d.Ho, d.Wo, d.Cout = N, H, W, Ho, Wo, O
    d.KH, d.KW, d.stride, d.pad_h, d.pad_w = KH, KW, stride, ph, pw
    d.act1, d.act2, d.scale = ACT[act], 0, scale
    d.precision = _lib.PREC[getattr(conv2d, 'precision', 'f32x3')]
    if backend == 'tc3s':
        orient = {(3, 3): 1, (5, 1): 1, (1, 5): 2}[(KH, KW)]
        ld16 = (Cin + 7) // 8 * 8
        x16 = split_f16(xh, ld16)
        taps = KH if orient == 1 else KW
        bw, bh = (8, 16 + taps - 1) if orient == 1 else (16 + taps - 1, 8)
        maps = (C.c_uint8 * 512)()
        L = _lib.lib()
        for j in range(2):
            check(L.bflow_tma_tile_map(C.addressof(maps) + 128 * j, x16[j].data_ptr(), N, H, W, Cin, ld16, bw, bh, 1 if orient == 2 else 0), 'tma_tile_map')
        wtc, acc_scale = pack_conv_weight_tc(weight, bn, block_per_tap=True)
        err = torch.zeros(1, device=x.device, dtype=torch.int32)
        d.x0 = None
        check(L.bflow_conv2d_nhwc_tc3s(C.byref(d), C.addressof(maps), wtc.data_ptr(), bn, acc_scale, orient, err.data_ptr(), _stream()), 'conv2d_tc3s')
        if int(err.item()) != 0:
            raise RuntimeError('bflow_conv2d_nhwc_tc3s: pipeline wait timed out inside the kernel')
    elif backend == 'tc3':
        x16 = split_f16(xh, (Cin + 7) // 8 * 8)
        maps = (C.c_uint8 * 512)()
        # `conv2d.split_c0 = c` reads the input as the concatenation of two sources, channels [0, c) and [c, Cin) (update.py:35,38,42,45: torch.cat
        # as two tensor-map pairs; c a multiple of 64)
        sc0 = getattr(conv2d, 'split_c0', None)
        if sc0:
            assert 0 < sc0 < Cin and sc0 % 64 == 0
            C.memmove(maps, tma_im2col_maps(x16, N, H, W, sc0, KH, KW, stride, ph, pw), 256)
            C.memmove(C.addressof(maps) + 256, tma_im2col_maps(x16, N, H, W, Cin - sc0, KH, KW, stride, ph, pw, c_off=sc0), 256)
            d.c0, d.c1 = sc0, Cin - sc0
            wtc, acc_scale = pack_conv_weight_tc(weight, bn, block_per_tap=True, c0=sc0)
        else:
            C.memmove(maps, tma_im2col_maps(x16, N, H, W, Cin, KH, KW, stride, ph, pw), 256)
            wtc, acc_scale = pack_conv_weight_tc(weight, bn, block_per_tap=True)
        err = torch.zeros(1, device=x.device, dtype=torch.int32)
        d.x0 = None
        if getattr(conv2d, 'tma_out', False):
            # outputs through tensor-map stores (bflow_conv2d_nhwc_tc3o): fp32 y and a split-fp16 copy that is checked against it
            L = _lib.lib()
            y16 = torch.zeros(2, N * Ho * Wo, (O + 7) // 8 * 8, device=x.device, dtype=torch.float16)
            d.y16_hi, d.y16_lo, d.ldy16 = y16[0].data_ptr(), y16[1].data_ptr(), y16.shape[-1]
            omaps = (C.c_uint8 * 640)()
            for j in range(2):
                check(L.bflow_tma_out_map(C.addressof(omaps) + 128 * j, y16[j].data_ptr(), N * Ho * Wo, O & ~7, y16.shape[-1], 2), 'tma_out_map')
            check(L.bflow_tma_out_map(C.addressof(omaps) + 256, y.data_ptr(), N * Ho * Wo, O, O, 4), 'tma_out_map')
            check(L.bflow_conv2d_nhwc_tc3o(C.byref(d), C.addressof(maps), C.addressof(omaps), wtc.data_ptr(), bn, acc_scale, err.data_ptr(), _stream()), 'conv2d_tc3o')
            conv2d.last_y16 = (y16[0].float() + y16[1].float())[:, :O].reshape(N, Ho, Wo, O).permute(0, 3, 1, 2)
            conv2d.last_y16_pad = y16[:, :, O:].clone()          # channels behind Cout must stay untouched
        else:
            check(_lib.lib().bflow_conv2d_nhwc_tc3(C.byref(d), C.addressof(maps), wtc.data_ptr(), bn, acc_scale, err.data_ptr(), _stream()), 'conv2d_tc3')
        if int(err.item()) != 0:
            raise RuntimeError('bflow_conv2d_nhwc_tc3: pipeline wait timed out inside the kernel')
    elif backend == 'stem7':
        # x stays NCHW fp32; `scale` doubles as the input map x -> in_scale * x + in_shift via the `stem_affine` attribute of this function
        in_scale, in_shift = getattr(conv2d, 'stem_affine', (1.0, 0.0))
        L = _lib.lib()
        wm = torch.zeros(O, 256, 1, 1, device=x.device, dtype=torch.float32)
        wm[:, :KH * KW * Cin, 0, 0] = _f32c(weight, 'weight').reshape(O, -1)       # K = (c, kh, kw): bflow_conv2d_stem7's order
        wtc, acc_scale = pack_conv_weight_tc(wm, 64, block_per_tap=True)
        err = torch.zeros(1, device=x.device, dtype=torch.int32)
        d.x0 = x.data_ptr()
        check(L.bflow_conv2d_stem7(C.byref(d), wtc.data_ptr(), Cin, (C.c_int * 1)(0), 1, in_scale, in_shift, acc_scale, err.data_ptr(), _stream()), 'conv2d_stem7')
        if int(err.item()) != 0:
            raise RuntimeError('bflow_conv2d_stem7: pipeline wait timed out inside the kernel')
    elif backend == 'slab64':
        x16 = split_f16(xh, 64)
        maps = (C.c_uint8 * 256)()
        L = _lib.lib()
        for j in range(2):
            check(L.bflow_tma_tile_map(C.addressof(maps) + 128 * j, x16[j].data_ptr(), N, H, W, 64, 64, 8, 18, 0), 'tma_tile_map')
        wtc, acc_scale = pack_conv_weight_tc(weight, 64, block_per_tap=True)
        err = torch.zeros(1, device=x.device, dtype=torch.int32)
        d.x0 = None
        check(L.bflow_conv2d_slab64(C.byref(d), C.addressof(maps), wtc.data_ptr(), acc_scale, err.data_ptr(), _stream()), 'conv2d_slab64')
        if int(err.item()) != 0:
            raise RuntimeError('bflow_conv2d_slab64: pipeline wait timed out inside the kernel')
    else:
        check(_lib.lib().bflow_conv2d_nhwc(C.byref(d), _stream()), 'conv2d')
    return nhwc_to_nchw(y)


@_on_tensor_device
def instance_norm_relu(x: torch.Tensor, residual: Optional[torch.Tensor] = None, residual_norm: bool = False,
                       eps: float = 1e-5) -> torch.Tensor:
    """NCHW convenience form: relu(IN(x)) or relu(relu(IN(x)) + R)."""
    L = _lib.lib()
    xh = nchw_to_nhwc(x)
    N, H, W, Cc = xh.shape
    sums = torch.zeros(N, Cc, 2, device=x.device, dtype=torch.float64)
    check(L.bflow_plane_sums(xh.data_ptr(), Cc, sums.data_ptr(), N, H * W, Cc, _stream()), 'plane_sums')
    rh = rs = None
    if residual is not None:
        rh = nchw_to_nhwc(residual)
        if residual_norm:
            rs = torch.zeros(N, Cc, 2, device=x.device, dtype=torch.float64)
            check(L.bflow_plane_sums(rh.data_ptr(), Cc, rs.data_ptr(), N, H * W, Cc, _stream()), 'plane_sums')
    out = torch.empty_like(xh)
    check(L.bflow_instnorm_relu(xh.data_ptr(), Cc, sums.data_ptr(), rh.data_ptr() if rh is not None else None, Cc,
                                rs.data_ptr() if rs is not None else None, out.data_ptr(), Cc, N, H * W, Cc, eps, _stream()),
          'instnorm_relu')
    return nhwc_to_nchw(out)


# ---- correlation -------------------------------------------------------------------------------------------
@_on_tensor_device
def corr_volume(fmap1: torch.Tensor, fmap2: torch.Tensor) -> torch.Tensor:
    """fmap1 (B,D,h,w) or (T,B,D,h,w); fmap2 (T,B,D,h,w), reference layout → (T, B*h*w, 1, h, w)
    (models/raft_utils/corr.py:264-272)."""
    fmap2 = _f32c(fmap2, 'fmap2')
    T, B, D, h, w = fmap2.shape
    fmap1 = _f32c(fmap1, 'fmap1')
    per_target = fmap1.ndim == 5
    Q = h * w
    out = torch.empty(T, B * Q, 1, h, w, device=fmap2.device, dtype=torch.float32)
    f1h = [nchw_to_nhwc(fmap1[t]) for t in range(T)] if per_target else [nchw_to_nhwc(fmap1)] * T
    for t in range(T):
        check(_lib.lib().bflow_corr_volume(f1h[t].data_ptr(), D, fmap2[t].data_ptr(), out[t].data_ptr(), B, D, Q, _stream()),
              'corr_volume')
    return out


@_on_tensor_device
def corr_pool(vol: torch.Tensor) -> torch.Tensor:
    """(..., H, W) → (..., H//2, W//2), avg_pool2d(2, 2) (corr.py:119)."""
    vol = _f32c(vol, 'vol')
    H, W = vol.shape[-2:]
    planes = vol.numel() // (H * W)
    out = torch.empty(*vol.shape[:-2], H // 2, W // 2, device=vol.device, dtype=torch.float32)
    check(_lib.lib().bflow_corr_pool(vol.data_ptr(), out.data_ptr(), planes, H, W, _stream()), 'corr_pool')
    return out


def tiled_plane_size(h: int, w: int) -> int:
    return ((h + 3) // 4) * ((w + 3) // 4) * 16


def to_tiled(planes: torch.Tensor) -> torch.Tensor:
    """(..., h, w) row-major planes → (..., ceil4(h)*ceil4(w)) in the 4x4-pixel-tiled layout of bflow_corr_lookup(tiled=1)."""
    h, w = planes.shape[-2:]
    hp, wp = (h + 3) // 4 * 4, (w + 3) // 4 * 4
    p = torch.nn.functional.pad(planes, (0, wp - w, 0, hp - h))
    lead = p.shape[:-2]
    p = p.reshape(*lead, hp // 4, 4, wp // 4, 4).transpose(-3, -2)
    return p.reshape(*lead, hp * wp).contiguous()


def from_tiled(flat: torch.Tensor, h: int, w: int) -> torch.Tensor:
    hp, wp = (h + 3) // 4 * 4, (w + 3) // 4 * 4
    lead = flat.shape[:-1]
    p = flat.reshape(*lead, hp // 4, wp // 4, 4, 4).transpose(-3, -2).reshape(*lead, hp, wp)
    return p[..., :h, :w].contiguous()


@_on_tensor_device
def corr_pool_tiled(flat: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """avg_pool2d(2,2) on tiled planes (..., tiled(h, w)) → (..., tiled(h//2, w//2))."""
    flat = _f32c(flat, 'vol')
    planes = flat.numel() // tiled_plane_size(h, w)
    out = torch.empty(*flat.shape[:-1], tiled_plane_size(h // 2, w // 2), device=flat.device, dtype=torch.float32)
    check(_lib.lib().bflow_corr_pool_tiled(flat.data_ptr(), out.data_ptr(), planes, h, w, _stream()), 'corr_pool_tiled')
    return out


def make_lookup_desc(slots: Sequence[tuple], n_targets: int, B: int, h: int, w: int, tiled: bool = False) -> LookupDesc:
    """slots: (level, base target, planes) in output order; planes is (B*Q, hl, wl) row-major, or — when ``tiled`` —
    (level, base target, planes (B*Q, tiled size), hl, wl)."""
    d = LookupDesc()
    d.n_slots, d.n_targets, d.B, d.h, d.w, d.radius = len(slots), n_targets, B, h, w, 4
    d.tiled = int(tiled)
    for s, entry in enumerate(slots):
        lvl, t, planes = entry[:3]
        assert planes.is_cuda and planes.dtype == torch.float32 and planes.is_contiguous()
        assert planes.shape[0] == B * h * w
        if tiled:
            hl, wl = entry[3], entry[4]
            assert planes.shape[1] == tiled_plane_size(hl, wl)
        else:
            hl, wl = planes.shape[-2], planes.shape[-1]
        d.vol[s] = planes.data_ptr()
        d.hl[s], d.wl[s] = hl, wl
        d.target[s] = t
        d.inv_scale[s] = 1.0 / (2 ** lvl)
    return d


@_on_tensor_device
def corr_lookup(slots: Sequence[tuple], coords: torch.Tensor, nhwc: bool = False, tiled: bool = False) -> torch.Tensor:
    """coords (T,B,2,h,w) → (B, S*81, h, w) [reference layout] or (B,h,w,S*81) when nhwc (corr.py:307-350)."""
    coords = _f32c(coords, 'coords')
    T, B, two, h, w = coords.shape
    assert two == 2
    d = make_lookup_desc(slots, T, B, h, w, tiled)
    S = len(slots)
    d.coords = coords.data_ptr()
    d.params, d.params_ld, d.degree = None, 0, 0
    if nhwc:
        out = torch.empty(B, h, w, S * 81, device=coords.device, dtype=torch.float32)
    else:
        out = torch.empty(B, S * 81, h, w, device=coords.device, dtype=torch.float32)
    d.out, d.out_nhwc, d.out_ld = out.data_ptr(), int(nhwc), S * 81
    check(_lib.lib().bflow_corr_lookup(C.byref(d), _stream()), 'corr_lookup')
    return out


@_on_tensor_device
def feat_pool(x_nhwc: torch.Tensor) -> torch.Tensor:
    """avg_pool2d(2, 2) (floor) of NHWC features (N, H, W, C) -> (N, H//2, W//2, C)."""
    x = _f32c(x_nhwc, 'x')
    N, H, W, Cc = x.shape
    out = torch.empty(N, H // 2, W // 2, Cc, device=x.device, dtype=torch.float32)
    check(_lib.lib().bflow_feat_pool(x.data_ptr(), out.data_ptr(), N, H, W, Cc, Cc, Cc, _stream()), 'feat_pool')
    return out


def make_lookup_otf_desc(slots: Sequence[tuple], n_targets: int, B: int, h: int, w: int, D: int) -> LookupOtfDesc:
    """slots: (level, base target, f1 NHWC (B,h,w,D), f2 NHWC (B,hl,wl,D)) in output order."""
    d = LookupOtfDesc()
    d.n_slots, d.n_targets, d.B, d.h, d.w, d.radius, d.D = len(slots), n_targets, B, h, w, 4, D
    d.ld1 = d.ld2 = D
    d.scale = 1.0 / math.sqrt(D)
    for s, (lvl, t, f1, f2) in enumerate(slots):
        for x in (f1, f2):
            assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.shape[0] == B and x.shape[-1] == D
        assert f1.shape[1:3] == (h, w)
        d.f1[s], d.f2[s] = f1.data_ptr(), f2.data_ptr()
        d.hl[s], d.wl[s] = f2.shape[1], f2.shape[2]
        d.target[s] = t
        d.inv_scale[s] = 1.0 / (2 ** lvl)
    return d


@_on_tensor_device
def corr_lookup_otf(fmap1: torch.Tensor, fmap2: torch.Tensor, levels: Sequence[int], coords: torch.Tensor) -> torch.Tensor:
    """On-the-fly form of CorrComputation + CorrBlockParallelMultiTarget (corr.py:128-350): fmap1 (B,D,h,w), fmap2 (T,B,D,h,w) in the
    reference layout, levels per target, coords (T,B,2,h,w) -> (B, S*81, h, w), without building the correlation volume."""
    from . import config as _cfg
    fmap1, fmap2, coords = _f32c(fmap1, 'fmap1'), _f32c(fmap2, 'fmap2'), _f32c(coords, 'coords')
    T, B, D, h, w = fmap2.shape
    f1 = nchw_to_nhwc(fmap1)
    pyr = [[nchw_to_nhwc(fmap2[t])] for t in range(T)]
    for t in range(T):
        for _ in range(1, levels[t]):
            pyr[t].append(feat_pool(pyr[t][-1]))
    slots = [(lvl, t, f1, pyr[t][lvl]) for (lvl, t) in _cfg.slot_table(list(levels))]
    d = make_lookup_otf_desc(slots, T, B, h, w, D)
    S = len(slots)
    out = torch.empty(B, h, w, S * 81, device=coords.device, dtype=torch.float32)
    d.coords, d.params, d.params_ld, d.degree = coords.data_ptr(), None, 0, 0
    d.out, d.out_ld = out.data_ptr(), S * 81
    check(_lib.lib().bflow_corr_lookup_otf(C.byref(d), _stream()), 'corr_lookup_otf')
    return out.permute(0, 3, 1, 2).contiguous()


# ---- Bezier ------------------------------------------------------------------------------------------------
@_on_tensor_device
def bezier_eval(params: torch.Tensor, coef: np.ndarray) -> torch.Tensor:
    params = _f32c(params, 'params')
    B, c2, H, W = params.shape
    deg = c2 // 2
    coef = np.ascontiguousarray(coef, dtype=np.float32)
    T = coef.shape[0]
    assert coef.shape == (T, deg)
    out = torch.empty(T, B, 2, H, W, device=params.device, dtype=torch.float32)
    for t0 in range(0, T, 32):
        t1 = min(T, t0 + 32)
        chunk = np.ascontiguousarray(coef[t0:t1])
        check(_lib.lib().bflow_bezier_eval(params.data_ptr(), chunk.ctypes.data, out[t0:t1].data_ptr(), t1 - t0, B, deg, H, W, _stream()),
              'bezier_eval')
    return out


@_on_tensor_device
def cvx_upsample(data: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """data (N,C,h,w), mask (N,576,h,w) in the reference layout → (N,C,8h,8w) (utils.py:33-48)."""
    data, mask = _f32c(data, 'data'), _f32c(mask, 'mask')
    N, Cc, h, w = data.shape
    assert mask.shape == (N, 576, h, w)
    out = torch.empty(N, Cc, 8 * h, 8 * w, device=data.device, dtype=torch.float32)
    check(_lib.lib().bflow_cvx_upsample(data.data_ptr(), 0, 1, mask.data_ptr(), 0, 1, out.data_ptr(), N, Cc, h, w, _stream()),
          'cvx_upsample')
    return out
