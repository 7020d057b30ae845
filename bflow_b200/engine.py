"""Launch planner for the RAFT-spline forward pass (the body of models/raft_spline/raft.py:101-200).

Python only does plumbing here: it packs weights once, sizes a workspace per input shape, records the
sequence of C-ABI kernel launches and captures that sequence in a CUDA graph; every arithmetic operation
runs in libbflow_b200.so.  Activations live in HBM as NHWC fp32 rows ``base[pixel*ld + channel]`` so that
channel concatenations of the reference (torch.cat in update.py) are just channel offsets into one buffer:

    hx   rows(B*Q) x [ h (hdim) | inp (cdim) | motion-encoder out (mdim-2deg) | Bezier params (2deg) ]
    cb   rows(B*Q) x [ convc2 out (192) | convf2 out (64) ]
    corr rows(B*Q) x ldc   (slot*81 + tap, zero padded to a multiple of 16 channels)

The correlation pyramid keeps the reference's layout (target, B*Q, h_l, w_l): one private plane per query.
"""
from __future__ import annotations

import ctypes as C
import gc
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from collections import OrderedDict

from . import _lib, config as _cfg
from ._lib import ACT, EPI, PREC, ConvDesc, LookupDesc, check
from .bezier import bernstein_coeffs
from .ops import pack_conv_weight, pack_conv_weight_tc, make_lookup_desc, tiled_plane_size, dev_zeros
from .engine_s16 import S16Recorder


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b * b


class _Weight:
    """One convolution's parameters.  Everything is packed on the HOST (numpy-speed torch CPU ops) and uploaded with plain copies, so
    that building an engine launches no PyTorch kernels: the first kernels a profiler sees in a forward are this library's."""
    __slots__ = ('w', 'ldw', 'b', 'cout', 'cin', 'kh', 'kw', 'stride', 'pad', 'oihw', 'cin_pad', 'tc', 'dev')

    def __init__(self, w, ldw, b, cout, cin, kh, kw, stride, pad, oihw, cin_pad, dev):
        self.w, self.ldw, self.b, self.cout, self.cin = w, ldw, b, cout, cin
        self.kh, self.kw, self.stride, self.pad = kh, kw, stride, pad
        self.oihw, self.cin_pad, self.dev = oihw, cin_pad, dev      # oihw stays on the host
        self.tc = {}                       # (bn, c0) -> tensor-core weight image on the device (packed on first use)

    def tc3_image(self, bn: int, c0: int):
        """Weight image of the TMA-fed kernel: K ordered (tap, 64-channel block), sources padded separately."""
        key = ('tc3', bn, c0)
        if key not in self.tc:
            img, acc_scale = pack_conv_weight_tc(self.oihw, bn, self.cin_pad, block_per_tap=True, c0=c0)
            self.tc[key] = (img.to(self.dev), acc_scale)
        return self.tc[key]


class Engine:
    def __init__(self, model, device: torch.device, precision: str = 'f32x3', correlation: str = 'volume'):
        self.cfg = model.model_params
        assert correlation in ('volume', 'otf'), correlation
        self.corr_mode = correlation
        self.device = device
        self.lib = _lib.lib()
        assert precision in PREC, f'precision must be one of {sorted(PREC)}'
        self.precision = precision
        self.prec = PREC[precision]
        self.use_graph = os.environ.get('BFLOW_GRAPH', '1') != '0'
        self.use_tc = os.environ.get('BFLOW_TC', '1') != '0'      # tcgen05 convolutions (0: exact-fp32 CUDA-core kernels, the numerical anchor)
        assert self.use_tc or self.prec == 0, 'BFLOW_TC=0 is the exact-fp32 anchor: it has no reduced-precision form'
        assert self.use_tc or correlation == 'volume', 'BFLOW_TC=0 is the exact-fp32 anchor of the volume path'
        self.use_side_stream = os.environ.get('BFLOW_STREAMS', '1') != '0'
        self.max_plans = max(1, int(os.environ.get('BFLOW_MAX_PLANS', '4')))
        with torch.cuda.device(device):
            self.err = dev_zeros(1, device=device, dtype=torch.int32)
            self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._err_event = None
            self.h2d_stream = torch.cuda.Stream(device=device)
            self.d2h_stream = torch.cuda.Stream(device=device)
            self._plans: 'OrderedDict[tuple, List[_Plan]]' = OrderedDict()
            self._turn: Dict[tuple, int] = {}
            self._pack(model)

    # ------------------------------------------------------------------------------------------------
    # weight preparation (once per state_dict; host-side packing, one upload per tensor)
    # ------------------------------------------------------------------------------------------------
    def _mk(self, conv, w=None, b=None, cin_pad=None) -> _Weight:
        w = conv.weight if w is None else w
        b = conv.bias if b is None else b
        w = w.detach().to('cpu', torch.float32).contiguous()
        b = b.detach().to('cpu', torch.float32).contiguous()
        O, I, KH, KW = w.shape
        wp, ldw = pack_conv_weight(w, cin_pad)
        return _Weight(wp.to(self.device), ldw, b.to(self.device), O, I if cin_pad is None else cin_pad, KH, KW, conv.stride, conv.pad, w, cin_pad,
                       self.device)

    def _mk_folded(self, conv, bn, rows: Optional[slice] = None) -> _Weight:
        """Conv followed by eval-mode BatchNorm (extractor.py:21-25) folded into weight and bias."""
        w = conv.weight.detach().to('cpu', torch.float32)
        b = conv.bias.detach().to('cpu', torch.float32)
        if bn is not None:
            g = bn.weight.detach().to('cpu', torch.float32)
            beta = bn.bias.detach().to('cpu', torch.float32)
            rm = bn.running_mean.detach().to('cpu', torch.float32)
            rv = bn.running_var.detach().to('cpu', torch.float32)
            s = g * torch.rsqrt(rv + bn.eps)
            w = w * s[:, None, None, None]
            b = (b - rm) * s + beta
        if rows is not None:
            w, b = w[rows], b[rows]
        return self._mk(conv, w.contiguous(), b.contiguous())

    def _pack_encoder(self, enc, kind: str, split: Optional[int] = None) -> dict:
        bn = kind == 'batch'
        f = (lambda c, n: self._mk_folded(c, n)) if bn else (lambda c, n: self._mk(c))
        out = {'kind': kind, 'conv1': f(enc.conv1, enc.norm1 if bn else None), 'blocks': []}
        w1 = out['conv1']
        O, I, KH, KW = w1.oihw.shape
        if KH * KW * I <= 512:      # stem as a GEMM over the materialised patch matrix (engine_s16: 'im2col' stem)
            kp = 256 if KH * KW * I <= 256 else _ceil(KH * KW * I, 64)      # bflow_conv2d_stem7 wants exactly 4 k-blocks
            wm = torch.zeros(O, kp, 1, 1, dtype=torch.float32)
            # K order: (c, kh, kw) for the fused stem (bflow_conv2d_stem7, K <= 256), (kh, kw, c) for the patch-matrix kernel (bflow_im2col_split16)
            wm[:, :KH * KW * I, 0, 0] = w1.oihw.reshape(O, -1) if kp == 256 else w1.oihw.permute(0, 2, 3, 1).reshape(O, -1)
            wp, ldw = pack_conv_weight(wm)
            out['conv1_mat'] = _Weight(wp.to(self.device), ldw, w1.b, O, kp, 1, 1, 1, (0, 0), wm, None, self.device)
        for layer in (enc.layer1, enc.layer2, enc.layer3):
            for blk in layer:
                e = {'conv1': f(blk.conv1, blk.norm1 if bn else None), 'conv2': f(blk.conv2, blk.norm2 if bn else None), 'down': None}
                if hasattr(blk, 'downsample'):
                    e['down'] = f(blk.downsample[0], blk.norm3 if bn else None)
                out['blocks'].append(e)
        if split is None:
            out['conv2'] = [self._mk(enc.conv2)]
        else:   # context encoder: hidden / context halves get different activations (raft.py:145-147)
            O = enc.conv2.weight.shape[0]
            out['conv2'] = [self._mk_folded(enc.conv2, None, slice(0, split)), self._mk_folded(enc.conv2, None, slice(split, O))]
        return out

    def _pack(self, model) -> None:
        cfg = self.cfg
        self.deg = cfg['bezier_degree']
        self.hdim, self.cdim, self.mdim = cfg['hidden']['dim'], cfg['context']['dim'], cfg['motion']['dim']
        self.fdim = cfg['feature']['dim']
        self.levels = _cfg.levels_per_target(cfg)
        self.slots = _cfg.slot_table(self.levels)
        self.ncorr = len(self.slots) * 81
        self.ldc = _ceil(self.ncorr, 16)
        assert len(self.slots) <= _lib.MAX_SLOTS and len(self.levels) <= _lib.MAX_TARGETS and self.deg <= _lib.MAX_DEGREE
        assert self.hdim % 4 == 0 and self.cdim % 4 == 0 and (self.mdim - 2 * self.deg) % 4 == 0
        self.enc = {}
        if model.fnet_ev is not None:
            self.enc['fnet_ev'] = self._pack_encoder(model.fnet_ev, cfg['feature']['norm'])
        if model.fnet_img is not None:
            self.enc['fnet_img'] = self._pack_encoder(model.fnet_img, cfg['feature']['norm'])
        self.enc['cnet'] = self._pack_encoder(model.cnet, cfg['context']['norm'], split=self.hdim)
        ub = model.update_block
        g = ub.gru
        U = {}
        U['convc1'] = self._mk(ub.encoder.convc1, cin_pad=self.ldc)
        for n in ('convc2', 'convf1', 'convf2', 'conv'):
            U[n] = self._mk(getattr(ub.encoder, n))
        hd, cd = self.hdim, self.cdim
        for sfx in '12':
            # [z | r] share one GEMM (N = 2*hdim).  The K axis of every GRU conv is [h | inp | motion] (update.py:35,38): the `inp`
            # slice is the same in every iteration, so its contribution (+ bias) is computed once per forward ("*_inp") and the
            # per-iteration convs only see [h | motion] ("*_dyn").
            z, r, q = getattr(g, 'convz' + sfx), getattr(g, 'convr' + sfx), getattr(g, 'convq' + sfx)
            cpu = lambda t: t.detach().to('cpu', torch.float32)
            wzr, bzr = torch.cat([cpu(z.weight), cpu(r.weight)], 0), torch.cat([cpu(z.bias), cpu(r.bias)], 0)
            wq = cpu(q.weight)
            dyn = lambda w: torch.cat([w[:, :hd], w[:, hd + cd:]], 1).contiguous()
            U['zr' + sfx + '_inp'] = self._mk(z, wzr[:, hd:hd + cd].contiguous(), bzr)
            U['zr' + sfx + '_dyn'] = self._mk(z, dyn(wzr), bzr)
            U['q' + sfx + '_inp'] = self._mk(q, wq[:, hd:hd + cd].contiguous(), q.bias)
            U['q' + sfx + '_dyn'] = self._mk(q, dyn(wq), q.bias)
        U['head1'], U['head2'] = self._mk(ub.bezier_head.conv1), self._mk(ub.bezier_head.conv2)
        U['mask0'], U['mask2'] = self._mk(ub.mask[0]), self._mk(ub.mask[2])
        self.upd = U
        ts = _cfg.lookup_timestamps(cfg)
        self.coef = bernstein_coeffs(ts, self.deg).astype('float32')      # (T, deg), float64 → fp32 like bezier.py:180

    # ------------------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------------------
    def _lanes(self, key) -> List['_Plan']:
        """Plans of one (B, H, W, iters, test_mode): lane 0 always, lane 1 once a pipelined (non_blocking) call needs a second set of
        input / workspace / output buffers.  Least-recently-used keys are dropped beyond BFLOW_MAX_PLANS (a D-shape plan holds ~1 GB)."""
        lanes = self._plans.get(key)
        if lanes is None:
            while len(self._plans) >= self.max_plans:
                self._plans.popitem(last=False)
            lanes = [_Plan(self, *key)]
            self._plans[key] = lanes
        else:
            self._plans.move_to_end(key)
        return lanes

    def plan(self, B, H, W, iters, test_mode) -> '_Plan':
        with torch.cuda.device(self.device):
            return self._lanes((B, H, W, iters, bool(test_mode)))[0]

    def _poll_err(self):
        """Raises if a tensor-core pipeline wait timed out in an EARLIER forward (the word travels with every forward's outputs; no sync)."""
        if self._err_event is not None and self._err_event.query() and int(self._err_host[0]) != 0:
            raise RuntimeError('bflow_b200: a tcgen05 pipeline wait timed out inside a kernel (the result of that forward is invalid)')

    def run(self, voxel, images, iters: int, init, test_mode: bool, non_blocking: bool = False):
        """One forward.  Blocking form (default, the reference's semantics): inputs are CUDA tensors, the launch list runs on the current
        stream and the outputs are fresh device tensors.  Pipelined form (non_blocking=True): inputs may be pinned host tensors; the
        H2D copy, the graph and the D2H copy of the results run on three streams over two alternating buffer sets, so that the copies of
        frames i+1 and i-1 overlap the graph of frame i.  Returns (low, ups, event): pinned host tensors that are valid once `event`
        has completed and until the second-next pipelined call reuses the lane."""
        ref = voxel if voxel is not None else images[0]
        B, _, H, W = ref.shape
        key = (B, H, W, iters, bool(test_mode))
        with torch.cuda.device(self.device):
            self._poll_err()
            lanes = self._lanes(key)
            if not non_blocking:
                plan = lanes[0]
                plan.wait_idle()
                plan.load_inputs(voxel, images, init)
                plan.execute()
                self._err_host_copy(torch.cuda.current_stream(self.device))
                return plan.low.clone(), [u.clone() for u in plan.ups]
            if len(lanes) < 2:
                lanes.append(_Plan(self, *key))
            turn = self._turn.get(key, 0)
            self._turn[key] = turn + 1
            return lanes[turn % 2].run_pipelined(voxel, images, init)

    def _err_host_copy(self, stream):
        self._err_host.copy_(self.err, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(stream)
        self._err_event = ev


class _Plan(S16Recorder):
    """Workspace + recorded launch list (+ CUDA graph) for one (B, H, W, iters, test_mode)."""

    def __init__(self, eng: Engine, B: int, H: int, W: int, iters: int, test_mode: bool):
        self.eng, self.B, self.H, self.W, self.iters, self.test_mode = eng, B, H, W, iters, test_mode
        cfg, dev = eng.cfg, eng.device
        self.h, self.w = H // 8, W // 8
        self.Q = self.h * self.w
        self.R = B * self.Q
        self.launches: List[Tuple] = []
        self.schedule: List[Tuple] = []          # ('launch', index, stream) | ('fork',) | ('join',)
        self._cur_stream = 0
        self.side_stream = None
        self.labels: List[Tuple[str, float]] = []
        self.keep: List = []           # descriptors / tensors referenced by raw pointer
        self.graph = None
        self.n_launches = 0
        self.n_tc = 0
        f32 = dict(device=dev, dtype=torch.float32)
        self.use_ev, self.use_img = cfg['use_events'], cfg['use_boundary_images']
        nctx, ncorr = cfg['num_bins']['context'], cfg['num_bins']['correlation']
        self.cin_vox = nctx + ncorr - 1
        # static inputs
        self.voxel_in = dev_zeros(B, self.cin_vox, H, W, **f32) if self.use_ev else None
        self.img_in = [dev_zeros(B, 3, H, W, **f32) for _ in range(2)] if self.use_img else None
        self.init_in = dev_zeros(B, 2 * eng.deg, self.h, self.w, **f32)
        # outputs
        self.low = torch.empty(B, 2 * eng.deg, self.h, self.w, **f32)
        n_up = 1 if test_mode else iters
        self.ups = [torch.empty(B, 2 * eng.deg, H, W, **f32) for _ in range(n_up)]
        # pipelined execution (Engine.run(non_blocking=True)): pinned result buffers and the events that order the three streams
        self.low_host = self.ups_host = None
        self.ev_done = self.ev_out = None
        if eng.use_tc:
            self._record_s16()
        else:
            self._record()

    # ---- recording helpers -------------------------------------------------------------------------------
    def _add(self, fn, *args, label: Optional[str] = None, flops: float = 0.0):
        self.schedule.append(('launch', len(self.launches), self._cur_stream))
        self.launches.append((fn, args))
        self.labels.append((label or fn.__name__.replace('bflow_', ''), flops))

    # independent branches of the forward pass run on a second stream (graph branches): every convolution of the update block
    # at batch 1 fills at most half of the SMs, so two of them side by side cost little more than one
    def _fork(self):
        self.schedule.append(('fork',))
        self._cur_stream = 1

    def _main(self):
        self._cur_stream = 0

    def _join(self):
        self.schedule.append(('join',))
        self._cur_stream = 0

    def _conv(self, wt, x0, c0, ld0, N, H, W, y, ldy, act1='none', act2='none', res=None, ldr=0,
              x1=None, c1=0, ld1=0, scale=1.0, bias=True, epi='std', aux0=None, ld_aux0=0, aux1=None, ld_aux1=0, kernel='auto'):
        d = ConvDesc()
        d.x0, d.c0, d.ld0 = x0, c0, ld0
        d.x1, d.c1, d.ld1 = x1, c1, ld1
        d.w, d.ldw, d.bias = wt.w.data_ptr(), wt.ldw, (wt.b.data_ptr() if bias else None)
        d.epi, d.aux0, d.ld_aux0, d.aux1, d.ld_aux1 = EPI[epi], aux0, ld_aux0, aux1, ld_aux1
        d.res, d.ldr = res, ldr
        d.y, d.ldy = y, ldy
        ph, pw = wt.pad
        Ho, Wo = (H + 2 * ph - wt.kh) // wt.stride + 1, (W + 2 * pw - wt.kw) // wt.stride + 1
        d.N, d.H, d.W, d.Ho, d.Wo, d.Cout = N, H, W, Ho, Wo, wt.cout
        d.KH, d.KW, d.stride, d.pad_h, d.pad_w = wt.kh, wt.kw, wt.stride, ph, pw
        d.act1, d.act2, d.scale = ACT[act1], ACT[act2], scale
        assert c0 + c1 == wt.cin, (c0, c1, wt.cin)
        self.keep.append(d)
        lib = self.eng.lib
        flops = 2.0 * N * Ho * Wo * wt.cout * wt.kh * wt.kw * (c0 + c1)
        if kernel == 'small_n':
            self._add(lib.bflow_conv2d_small_n, C.byref(d), label=f'conv_small_n {c0 + c1}->{wt.cout} {wt.kh}x{wt.kw}/{wt.stride} M={N * Ho * Wo}', flops=flops)
        else:
            self._add(lib.bflow_conv2d_nhwc, C.byref(d), label=f'conv_simt {c0 + c1}->{wt.cout} {wt.kh}x{wt.kw}/{wt.stride} M={N * Ho * Wo}', flops=flops)
        return Ho, Wo

    def _sums(self, N, Cc):
        off = self._sums_off
        self._sums_off += N * Cc * 2
        assert self._sums_off <= self.sums.numel()
        return self.sums.data_ptr() + off * 8

    # ---- encoders (models/raft_utils/extractor.py:47-55,103-125) -----------------------------------------
    def _encoder(self, E: dict, windows: Sequence[Tuple[int, int, int, int]], Np: int, H: int, W: int, bufs: List[torch.Tensor], final):
        """windows: (pointer, channels, ld, samples) per conv1 launch; sample blocks are stacked along the
        batch axis in that order (the reference's torch.cat at extractor.py:106-110)."""
        L = self.eng.lib
        kind = E['kind']
        inorm = kind == 'instance'
        free = list(bufs)

        def norm_relu(ptr, N, HW, Cc, res=None, res_sums=None):
            if not inorm:
                return
            s = self._sums(N, Cc)
            self._add(L.bflow_plane_sums, ptr, Cc, s, N, HW, Cc)
            self._add(L.bflow_instnorm_relu, ptr, Cc, s, res, Cc, res_sums, ptr, Cc, N, HW, Cc, 1e-5)

        act = 'none' if inorm else 'relu'
        # stem: 7x7 stride-2 conv, one launch per input window
        X = free.pop()
        w1 = E['conv1']
        H2, W2 = H // 2, W // 2
        n0 = 0
        for ptr, cin, ld, ns in windows:
            self._conv(w1, ptr, cin, ld, ns, H, W, X.data_ptr() + n0 * H2 * W2 * 64 * 4, 64, act1=act)
            n0 += ns
        assert n0 == Np
        norm_relu(X.data_ptr(), Np, H2 * W2, 64)
        Hc, Wc, Cc = H2, W2, 64
        for blk in E['blocks']:
            c1, c2, dn = blk['conv1'], blk['conv2'], blk['down']
            Y1 = free.pop()
            Ho, Wo = self._conv(c1, X.data_ptr(), Cc, Cc, Np, Hc, Wc, Y1.data_ptr(), c1.cout, act1=act)
            Co = c1.cout
            norm_relu(Y1.data_ptr(), Np, Ho * Wo, Co)
            Y2 = free.pop()
            if dn is not None:
                D = free.pop()
                self._conv(dn, X.data_ptr(), Cc, Cc, Np, Hc, Wc, D.data_ptr(), Co)
                res, ldr = D.data_ptr(), Co
            else:
                D = None
                res, ldr = X.data_ptr(), Cc
            if inorm:
                self._conv(c2, Y1.data_ptr(), Co, Co, Np, Ho, Wo, Y2.data_ptr(), Co)
                rs = None
                if D is not None:
                    rs = self._sums(Np, Co)
                    self._add(L.bflow_plane_sums, D.data_ptr(), Co, rs, Np, Ho * Wo, Co)
                norm_relu(Y2.data_ptr(), Np, Ho * Wo, Co, res=res, res_sums=rs)
            else:
                self._conv(c2, Y1.data_ptr(), Co, Co, Np, Ho, Wo, Y2.data_ptr(), Co, act1=act, act2='relu', res=res, ldr=ldr)
            free.append(X)
            free.append(Y1)
            if D is not None:
                free.append(D)
            X, Hc, Wc, Cc = Y2, Ho, Wo, Co
        final(X.data_ptr(), Np, Hc, Wc, Cc)

    # ---- the whole forward ----------------------------------------------------------------------------------
    def _record(self):
        """BFLOW_TC=0: every convolution on the exact-fp32 CUDA-core kernels over fp32 NHWC activations, the volume in the reference's
        row-major layout -- the numerical anchor the tensor-core path (engine_s16._record_s16) is checked against."""
        eng, L = self.eng, self.eng.lib
        cfg, dev = eng.cfg, eng.device
        B, H, W, h, w, Q, R = self.B, self.H, self.W, self.h, self.w, self.Q, self.R
        f32 = dict(device=dev, dtype=torch.float32)
        deg, hd, cd, md, fd = eng.deg, eng.hdim, eng.cdim, eng.mdim, eng.fdim
        nctx, ncorr = cfg['num_bins']['context'], cfg['num_bins']['correlation']
        T_ev = len(cfg['correlation']['ev']['target_indices']) if self.use_ev else 0
        T = len(eng.levels)
        np_max = max((T_ev + 1) * B if self.use_ev else 0, 2 * B if self.use_img else 0, B)
        bufs = [torch.empty(np_max * (H // 2) * (W // 2) * 64, **f32) for _ in range(4)]
        self.sums = dev_zeros(64 * np_max * 128 * 2, device=dev, dtype=torch.float64)
        self._sums_off = 0
        self.keep += bufs
        self._add(L.bflow_zero, self.sums.data_ptr(), self.sums.numel() * 8)

        gw = hd + cd + md                              # GRU input width = [h | inp | motion]
        poff = hd + cd + md - 2 * deg                  # Bezier params live at the tail of hx
        self.hx = dev_zeros(R, gw, **f32)
        self.poff, self.gw = poff, gw
        hx = self.hx.data_ptr()

        # ---- inputs to NHWC ----
        ctx_c = (nctx if self.use_ev else 0) + (3 if self.use_img else 0)
        self.ctx = dev_zeros(B, H, W, ctx_c, **f32)
        if self.use_ev:
            self.vox = dev_zeros(B, H, W, self.cin_vox, **f32)
            self._add(L.bflow_nchw_to_nhwc, self.voxel_in.data_ptr(), self.vox.data_ptr(), B, self.cin_vox, H, W, 0, self.cin_vox,
                      self.cin_vox, 1.0, 0.0)
        if self.use_img:
            # images -> 2*(x/255)-1 (raft.py:134); image 0 is also the tail of the context input (raft.py:137-140)
            self.imgs = dev_zeros(2 * B, H, W, 3, **f32)
            for i in range(2):
                self._add(L.bflow_nchw_to_nhwc, self.img_in[i].data_ptr(), self.imgs.data_ptr() + i * B * H * W * 3 * 4, B, 3, H, W, 0, 3, 3,
                          2.0 / 255.0, -1.0)

        # ---- feature encoders ----
        fm_ev = fm_img = None
        if self.use_ev:
            fm_ev = torch.empty((T_ev + 1) * B, h, w, fd, **f32)
            idxs = [0] + list(cfg['correlation']['ev']['target_indices'])
            wins = [(self.vox.data_ptr() + i * 4, ncorr, self.cin_vox, B) for i in idxs]
            E = eng.enc['fnet_ev']
            self._encoder(E, wins, (T_ev + 1) * B, H, W, bufs,
                          lambda x, Np, Hc, Wc, Cc: self._conv(E['conv2'][0], x, Cc, Cc, Np, Hc, Wc, fm_ev.data_ptr(), fd))
        if self.use_img:
            fm_img = torch.empty(2 * B, h, w, fd, **f32)
            E2 = eng.enc['fnet_img']
            self._encoder(E2, [(self.imgs.data_ptr(), 3, 3, 2 * B)], 2 * B, H, W, bufs,
                          lambda x, Np, Hc, Wc, Cc: self._conv(E2['conv2'][0], x, Cc, Cc, Np, Hc, Wc, fm_img.data_ptr(), fd))
        self.fm_ev, self.fm_img = fm_ev, fm_img

        # ---- context encoder → net = tanh, inp = relu straight into hx (raft.py:144-147) ----
        if self.use_ev and self.use_img:
            # context = cat(voxel[:, -nctx:], image0) (raft.py:137-138)
            self._add(L.bflow_nchw_to_nhwc, self.voxel_in.data_ptr(), self.ctx.data_ptr(), B, self.cin_vox, H, W, self.cin_vox - nctx, nctx,
                      ctx_c, 1.0, 0.0)
            self._add(L.bflow_nchw_to_nhwc, self.img_in[0].data_ptr(), self.ctx.data_ptr() + nctx * 4, B, 3, H, W, 0, 3, ctx_c, 2.0 / 255.0, -1.0)
            cwin = [(self.ctx.data_ptr(), ctx_c, ctx_c, B)]
        elif self.use_ev:
            cwin = [(self.vox.data_ptr() + (self.cin_vox - nctx) * 4, nctx, self.cin_vox, B)]
        else:
            cwin = [(self.imgs.data_ptr(), 3, 3, B)]
        Ec = eng.enc['cnet']

        def cnet_final(x, Np, Hc, Wc, Cc):
            self._conv(Ec['conv2'][0], x, Cc, Cc, Np, Hc, Wc, hx, gw, act1='tanh')
            self._conv(Ec['conv2'][1], x, Cc, Cc, Np, Hc, Wc, hx + hd * 4, gw, act1='relu')
        self._encoder(Ec, cwin, B, H, W, bufs, cnet_final)

        # ---- initial Bezier parameters: zeros (+ flow_init) (raft.py:150-153) ----
        self._add(L.bflow_nchw_to_nhwc, self.init_in.data_ptr(), hx + poff * 4, B, 2 * deg, h, w, 0, 2 * deg, gw, 1.0, 0.0)

        # ---- correlation volume + pyramid (corr.py:264-272, 293-305): reference layout, one private row-major plane per query ----
        self.tiled = False
        self.vol0 = torch.empty(T, R, h, w, **f32)
        if self.use_ev:
            self.f2_ev = torch.empty(T_ev * B, fd, Q, **f32)
            self._add(L.bflow_nhwc_to_nchw, fm_ev.data_ptr() + B * Q * fd * 4, self.f2_ev.data_ptr(), T_ev * B, fd, h, w, fd)
            for t in range(T_ev):
                self._add(L.bflow_corr_volume, fm_ev.data_ptr(), fd, self.f2_ev.data_ptr() + t * B * fd * Q * 4, self.vol0[t].data_ptr(), B, fd, Q)
        if self.use_img:
            self.f2_img = torch.empty(B, fd, Q, **f32)
            self._add(L.bflow_nhwc_to_nchw, fm_img.data_ptr() + B * Q * fd * 4, self.f2_img.data_ptr(), B, fd, h, w, fd)
            self._add(L.bflow_corr_volume, fm_img.data_ptr(), fd, self.f2_img.data_ptr(), self.vol0[T_ev].data_ptr(), B, fd, Q)
        # pyramid: level l holds the targets with more than l levels; (indices, tensor, hl, wl)
        pyr = [(list(range(T)), self.vol0, h, w)]
        for lvl in range(1, max(eng.levels)):
            prev_idx, prev, hp_, wp_ = pyr[-1]
            keep = [t for t in range(T) if eng.levels[t] > lvl]
            hl, wl = hp_ // 2, wp_ // 2
            cur = torch.empty(len(keep), R, hl, wl, **f32)
            for j, t in enumerate(keep):
                src = prev[prev_idx.index(t)]
                self._add(L.bflow_corr_pool, src.data_ptr(), cur[j].data_ptr(), R, hp_, wp_)
            pyr.append((keep, cur, hl, wl))
        self.pyr = pyr

        # ---- lookup descriptor (corr.py:307-350); centres come from the Bezier params in hx ----
        slots = [(lvl, t, pyr[lvl][1][pyr[lvl][0].index(t)]) for (lvl, t) in eng.slots]
        self.corr = dev_zeros(R, eng.ldc, **f32)
        ld = make_lookup_desc(slots, T, B, h, w, False)
        ld.coords = None
        ld.params, ld.params_ld, ld.degree = hx + poff * 4, gw, deg
        for t in range(T):
            for k in range(deg):
                ld.coef[t][k] = float(eng.coef[t, k])
        ld.out, ld.out_nhwc, ld.out_ld = self.corr.data_ptr(), 1, eng.ldc
        self.keep.append(ld)
        self.lookup_desc = ld

        # ---- update-block workspace ----
        U = eng.upd
        self.c1 = torch.empty(R, 256, **f32)
        self.cb = torch.empty(R, 256, **f32)
        self.f1 = torch.empty(R, 128, **f32)
        self.zr = torch.empty(R, 2 * hd, **f32)
        self.rh = torch.empty(R, hd, **f32)
        self.hh = torch.empty(R, 256, **f32)
        self.mask = torch.empty(R, 576, **f32)
        c1, cb, f1, zr, rh, hh, mk = (t.data_ptr() for t in (self.c1, self.cb, self.f1, self.zr, self.rh, self.hh, self.mask))
        xw = cd + md                                   # x = [inp | motion features]

        def upsample(out_t):
            self._conv(U['mask0'], hx, hd, gw, B, h, w, hh, 256, act1='relu')
            self._conv(U['mask2'], hh, 256, 256, B, h, w, mk, 576, scale=0.25)
            self._add(L.bflow_cvx_upsample, hx + poff * 4, gw, 0, mk, 576, 0, out_t.data_ptr(), B, 2 * deg, h, w)

        # iteration-invariant part of the GRU convolutions: conv(inp) + bias for z|r and q of both passes
        self.pre = {k: torch.empty(R, (2 * hd if k.startswith('zr') else hd), **f32) for k in ('zr1', 'q1', 'zr2', 'q2')}
        pre = {k: v.data_ptr() for k, v in self.pre.items()}
        for k in ('zr1', 'q1', 'zr2', 'q2'):
            self._conv(U[k + '_inp'], hx + hd * 4, cd, gw, B, h, w, pre[k], self.pre[k].shape[1])

        self.iter_start = len(self.launches)
        for itr in range(self.iters):
            if itr == 1:
                self.iter_len = len(self.launches) - self.iter_start
            # lookup around coords0 + flow(t) (raft.py:170-184)
            self._add(L.bflow_corr_lookup, C.byref(ld))
            # motion encoder (update.py:88-97)
            self._conv(U['convc1'], self.corr.data_ptr(), eng.ldc, eng.ldc, B, h, w, c1, 256, act1='relu')
            self._conv(U['convc2'], c1, 256, 256, B, h, w, cb, 256, act1='relu')
            self._conv(U['convf1'], hx + poff * 4, 2 * deg, gw, B, h, w, f1, 128, act1='relu')
            self._conv(U['convf2'], f1, 128, 128, B, h, w, cb + 192 * 4, 256, act1='relu')
            self._conv(U['conv'], cb, 256, 256, B, h, w, hx + (hd + cd) * 4, gw, act1='relu')
            # SepConvGRU (update.py:33-48): horizontal then vertical pass
            for sfx in '12':
                # z|r = sigmoid(conv([h | motion]) + inp part); epilogue also emits r*h.   q = tanh(conv([r*h | motion]) + inp part);
                # epilogue applies h = (1-z) h + z q in place.
                self._conv(U['zr' + sfx + '_dyn'], hx, hd, gw, B, h, w, zr, 2 * hd, x1=hx + (hd + cd) * 4, c1=md, ld1=gw, bias=False,
                           res=pre['zr' + sfx], ldr=2 * hd, epi='gru_zr', aux0=hx, ld_aux0=gw, aux1=rh, ld_aux1=hd)
                self._conv(U['q' + sfx + '_dyn'], rh, hd, hd, B, h, w, hx, gw, x1=hx + (hd + cd) * 4, c1=md, ld1=gw, bias=False,
                           res=pre['q' + sfx], ldr=hd, epi='gru_q', aux0=zr, ld_aux0=2 * hd)
            # Bezier head + delta update in place (update.py:17-18, bezier.py:137-139)
            self._conv(U['head1'], hx, hd, gw, B, h, w, hh, 256, act1='relu')
            self._conv(U['head2'], hh, 256, 256, B, h, w, hx + poff * 4, gw, res=hx + poff * 4, ldr=gw,
                       kernel='small_n' if 2 * deg <= 32 else 'auto')
            if not self.test_mode:
                upsample(self.ups[itr])
        if self.iters == 1:
            self.iter_len = len(self.launches) - self.iter_start
        if self.test_mode:
            upsample(self.ups[0])
        self._add(L.bflow_nhwc_to_nchw, hx + poff * 4, self.low.data_ptr(), B, 2 * deg, h, w, gw)
        self.n_launches = len(self.launches)

    # ---- execution -----------------------------------------------------------------------------------------
    def load_inputs(self, voxel, images, init, stream=None):
        """Copies the call's inputs into the plan's static buffers (on `stream`, default: the current stream of the engine's device)."""
        ctx = torch.cuda.stream(stream) if stream is not None else _NullCtx()
        with ctx:
            if self.use_ev:
                assert voxel.shape == self.voxel_in.shape, (voxel.shape, self.voxel_in.shape)
                self.voxel_in.copy_(voxel, non_blocking=True)
            if self.use_img:
                for dst, src in zip(self.img_in, images):
                    assert src.shape == dst.shape
                    dst.copy_(src, non_blocking=True)
            if init is not None:
                assert init.shape == self.init_in.shape
                self.init_in.copy_(init, non_blocking=True)
                self._init_dirty = True
            elif getattr(self, '_init_dirty', False):
                self.init_in.zero_()
                self._init_dirty = False

    def launch_all(self):
        main = torch.cuda.current_stream(self.eng.device)
        if self.side_stream is None:
            self.side_stream = torch.cuda.Stream(device=self.eng.device, priority=int(os.environ.get('BFLOW_SIDE_PRIORITY', '0')))
        side = self.side_stream
        handles = (main.cuda_stream, side.cuda_stream)
        use_side = self.eng.use_side_stream
        for item in self.schedule:
            if item[0] == 'launch':
                fn, args = self.launches[item[1]]
                rc = fn(*args, handles[item[2] if use_side else 0])
                if rc != 0:
                    check(rc, fn.__name__)
            elif use_side:
                if item[0] == 'fork':
                    side.wait_stream(main)
                else:
                    main.wait_stream(side)

    def check(self):
        """Synchronises and raises if a tensor-core pipeline wait timed out (never expected; such a kernel also traps)."""
        if int(self.eng.err.item()) != 0:
            raise RuntimeError('bflow_b200: a tcgen05 pipeline wait timed out inside a kernel')

    def execute(self):
        with torch.cuda.device(self.eng.device):
            if not self.eng.use_graph:
                self.launch_all()
                return
            if self.graph is None:
                self.launch_all()                         # eager warm-up (also surfaces contract errors outside capture)
                torch.cuda.current_stream().synchronize()
                self.check()
                g = torch.cuda.CUDAGraph()
                # A dead Python cycle that still owns CUDA objects (an earlier engine: its CUDAGraph, streams, tensors) must not be collected
                # WHILE this capture runs: destroying a graph / stream from the garbage collector is an unsafe call in global capture mode and
                # invalidates the capture (torch.cuda.graph no longer collects on entry).  Collect now, keep the collector off until capture_end.
                gc.collect()
                gc_was_enabled = gc.isenabled()
                gc.disable()
                try:
                    with torch.cuda.graph(g):
                        self.launch_all()
                finally:
                    if gc_was_enabled:
                        gc.enable()
                self.graph = g
            self.graph.replay()

    def wait_idle(self):
        """Blocking calls reuse lane 0: wait (on the current stream) for pipelined work that may still be using its buffers."""
        cur = torch.cuda.current_stream(self.eng.device)
        if self.ev_out is not None:
            cur.wait_event(self.ev_out)

    def run_pipelined(self, voxel, images, init):
        """H2D on the engine's h2d stream, the graph on the current stream, D2H on the d2h stream; see Engine.run."""
        eng = self.eng
        cur = torch.cuda.current_stream(eng.device)
        for t in ([voxel] if voxel is not None else []) + list(images or []):
            assert t.is_cuda or t.is_pinned(), 'pipelined forward: host inputs must be pinned (tensor.pin_memory())'
        if self.low_host is None:
            self.low_host = torch.empty(self.low.shape, dtype=torch.float32).pin_memory()
            self.ups_host = [torch.empty(u.shape, dtype=torch.float32).pin_memory() for u in self.ups]
        # 1. inputs: the previous graph of this lane must have finished reading them
        if self.ev_done is not None:
            eng.h2d_stream.wait_event(self.ev_done)
        else:
            eng.h2d_stream.wait_stream(cur)
        self.load_inputs(voxel, images, init, stream=eng.h2d_stream)
        ev_in = torch.cuda.Event()
        ev_in.record(eng.h2d_stream)
        # 2. compute: inputs have landed, and the previous results of this lane have left the output buffers
        cur.wait_event(ev_in)
        if self.ev_out is not None:
            cur.wait_event(self.ev_out)
        self.execute()
        self.ev_done = torch.cuda.Event()
        self.ev_done.record(cur)
        # 3. results (+ the error word) to pinned host memory
        eng.d2h_stream.wait_event(self.ev_done)
        with torch.cuda.stream(eng.d2h_stream):
            self.low_host.copy_(self.low, non_blocking=True)
            for dst, src in zip(self.ups_host, self.ups):
                dst.copy_(src, non_blocking=True)
            eng._err_host_copy(eng.d2h_stream)
        self.ev_out = torch.cuda.Event()
        self.ev_out.record(eng.d2h_stream)
        return self.low_host, self.ups_host, self.ev_out


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
