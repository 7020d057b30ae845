"""Drop-in for the reference's ``models.raft_spline.raft`` (RAFTSpline + BezierCurves).

Same constructor dict, same ``state_dict`` keys/shapes, same ``forward`` signature and return
types as models/raft_spline/raft.py:14-200, so ``modules/raft_spline.py:24,57-58`` can use it
unchanged.  The module tree below only HOLDS parameters under the reference's names; all
arithmetic runs in the hand-written sm_100a kernels of ``libbflow_b200.so`` driven by
:mod:`bflow_b200.engine`.  There is no CPU or PyTorch fallback: ``forward`` on non-CUDA inputs,
or without the built library, raises.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from .bezier import BezierCurves
from . import config as _cfg

__all__ = ['RAFTSpline', 'BezierCurves']


class _Conv(nn.Module):
    """Parameter holder with nn.Conv2d's state_dict layout (weight OIHW, bias O)."""

    def __init__(self, cin: int, cout: int, kh: int, kw: int, stride: int = 1, pad=(0, 0)):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(cout, cin, kh, kw))
        self.bias = nn.Parameter(torch.zeros(cout))
        self.stride = stride
        self.pad = pad


class _BatchNorm(nn.Module):
    """Parameter/buffer holder with nn.BatchNorm2d's state_dict layout."""

    def __init__(self, c: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer('running_mean', torch.zeros(c))
        self.register_buffer('running_var', torch.ones(c))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))
        self.eps = 1e-5


class _Holder(nn.Module):
    pass


def _norm_holder(kind: str, c: int) -> nn.Module:
    if kind == 'batch':
        return _BatchNorm(c)
    if kind in ('instance', 'none'):
        return _Holder()           # InstanceNorm2d(affine=False) and Sequential() carry no state
    raise NotImplementedError(f'norm_fn={kind!r} (the reference configs use instance/batch)')


def _res_block(cin: int, cout: int, kind: str, stride: int) -> nn.Module:
    """Parameter layout of ResidualBlock (models/raft_utils/extractor.py:5-44)."""
    b = _Holder()
    b.conv1 = _Conv(cin, cout, 3, 3, stride, (1, 1))
    b.conv2 = _Conv(cout, cout, 3, 3, 1, (1, 1))
    b.norm1 = _norm_holder(kind, cout)
    b.norm2 = _norm_holder(kind, cout)
    if stride != 1:
        b.norm3 = _norm_holder(kind, cout)
        # the reference registers norm3 a second time inside the downsample Sequential
        b.downsample = nn.Sequential(_Conv(cin, cout, 1, 1, stride, (0, 0)), b.norm3)
    return b


def _encoder(cin: int, cout: int, kind: str) -> nn.Module:
    """Parameter layout of BasicEncoder (models/raft_utils/extractor.py:58-101)."""
    e = _Holder()
    e.norm_fn = kind
    e.norm1 = _norm_holder(kind, 64)
    e.conv1 = _Conv(cin, 64, 7, 7, 2, (3, 3))
    e.layer1 = nn.Sequential(_res_block(64, 64, kind, 1), _res_block(64, 64, kind, 1))
    e.layer2 = nn.Sequential(_res_block(64, 96, kind, 2), _res_block(96, 96, kind, 1))
    e.layer3 = nn.Sequential(_res_block(96, 128, kind, 2), _res_block(128, 128, kind, 1))
    e.conv2 = _Conv(128, cout, 1, 1)
    return e


def num_cor_planes(cfg: Dict[str, Any]) -> int:
    """Channel count of the lookup output as the reference derives it from the config
    (models/raft_spline/update.py:69-86)."""
    out = 0
    if cfg['use_events']:
        ev = cfg['correlation']['ev']
        assert len(ev['levels']) == len(ev['radius']) > 0
        out += sum(l * (2 * r + 1) ** 2 for l, r in zip(ev['levels'], ev['radius']))
    if cfg['use_boundary_images']:
        im = cfg['correlation']['img']
        out += im['levels'] * (2 * im['radius'] + 1) ** 2
    return out


def _update_block(cfg: Dict[str, Any], hdim: int) -> nn.Module:
    """Parameter layout of BasicUpdateBlock (models/raft_spline/update.py:8-114)."""
    deg2 = 2 * cfg['bezier_degree']
    mdim = cfg['motion']['dim']
    cdim = cfg['context']['dim']
    u = _Holder()
    enc = _Holder()
    enc.convc1 = _Conv(num_cor_planes(cfg), 256, 1, 1)
    enc.convc2 = _Conv(256, 192, 3, 3, 1, (1, 1))
    enc.convf1 = _Conv(deg2, 128, 7, 7, 1, (3, 3))
    enc.convf2 = _Conv(128, 64, 3, 3, 1, (1, 1))
    enc.conv = _Conv(64 + 192, mdim - deg2, 3, 3, 1, (1, 1))
    u.encoder = enc
    gin = hdim + cdim + mdim
    gru = _Holder()
    for sfx, (kh, kw) in (('1', (1, 5)), ('2', (5, 1))):
        for gate in 'zrq':
            setattr(gru, f'conv{gate}{sfx}', _Conv(gin, hdim, kh, kw, 1, (kh // 2, kw // 2)))
    u.gru = gru
    head = _Holder()
    head.conv1 = _Conv(hdim, 256, 3, 3, 1, (1, 1))
    head.conv2 = _Conv(256, deg2, 3, 3, 1, (1, 1))
    u.bezier_head = head
    u.mask = nn.Sequential(_Conv(hdim, 256, 3, 3, 1, (1, 1)), _Holder(), _Conv(256, 64 * 9, 1, 1))
    return u


class RAFTSpline(nn.Module):
    """``precision``: arithmetic of the tensor-core convolutions.  ``'f32x3'`` (default): split-fp16 operands, three MMAs per product,
    fp32-equivalent (the 1e-3 px configuration of BASELINE.json).  ``'f16'``: one fp16 MMA per product on the hi planes (the reduced-
    precision configuration, 1e-2 px bar).  Default from the environment variable BFLOW_PRECISION.
    ``correlation``: ``'volume'`` (default) or ``'otf'`` (on-the-fly, no materialised volume; BFLOW_CORR)."""

    def __init__(self, model_params: Dict[str, Any], seed: Optional[int] = 0, verbose: bool = False, precision: Optional[str] = None,
                 correlation: Optional[str] = None):
        super().__init__()
        import os
        self.precision = precision if precision is not None else os.environ.get('BFLOW_PRECISION', 'f32x3')
        assert self.precision in ('f32x3', 'f16'), self.precision
        # 'volume' (default): all-pairs correlation volume + pyramid as in the reference; 'otf': on-the-fly correlation against a pooled
        # target-feature pyramid (no T x (B*Q) x Q volume: 369 MB at 480x640, the only reason the batch is capped by memory)
        self.correlation = correlation if correlation is not None else os.environ.get('BFLOW_CORR', 'volume')
        assert self.correlation in ('volume', 'otf'), self.correlation
        p = model_params
        nctx, ncorr = p['num_bins']['context'], p['num_bins']['correlation']
        self.bezier_degree = p['bezier_degree']
        self.detach_bezier = p['detach_bezier']
        assert ncorr > 0 and nctx > 0
        assert self.bezier_degree >= 1
        self.nbins_context, self.nbins_corr = nctx, ncorr
        cp = p['correlation']
        self.corr_use_cosine_sim = cp['use_cosine_sim']          # read, never used (raft.py:33)
        self.ev_corr_target_indices = list(cp['ev']['target_indices']) if p['use_events'] else []
        self.ev_corr_levels = list(cp['ev']['levels']) if p['use_events'] else []
        self.ev_corr_radius = 4                                  # hard-coded in the reference (raft.py:38-40)
        self.img_corr_params = None
        if p['use_boundary_images']:
            self.img_corr_params = cp['img']
            assert 'levels' in self.img_corr_params and 'radius' in self.img_corr_params
        self.hidden_dim = hdim = p['hidden']['dim']
        self.context_dim = cdim = p['context']['dim']
        fdim, fnorm, cnorm = p['feature']['dim'], p['feature']['norm'], p['context']['norm']

        context_in = 0
        self.fnet_img = None
        if self.img_corr_params is not None:
            self.fnet_img = _encoder(3, fdim, fnorm)
            context_in += 3
        self.fnet_ev = None
        if p['use_events']:
            assert 0 not in self.ev_corr_target_indices
            assert len(self.ev_corr_target_indices) > 0
            assert max(self.ev_corr_target_indices) < nctx
            assert len(self.ev_corr_target_indices) == len(self.ev_corr_levels)
            self.fnet_ev = _encoder(ncorr, fdim, fnorm)
            context_in += nctx
        assert self.fnet_ev is not None or self.fnet_img is not None
        self.cnet = _encoder(context_in, hdim + cdim, cnorm)
        self.update_block = _update_block(p, hdim)

        self.model_params = p
        self._engine = None
        self._engine_versions = None
        # the reference's RAFTSplineModule loads checkpoints through the PARENT module: PyTorch then copies into the parameters in place and
        # never calls this module's load_state_dict.  The post-hook fires for nested loads too; in-place updates (optimizer steps, EMA swaps,
        # param.data.copy_) are caught by the tensor version counters checked in engine().
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())
        if verbose:
            print(f'bflow_b200 RAFT-Spline: context bins {nctx}, correlation bins {ncorr}, '
                  f'degree {self.bezier_degree}, events {p["use_events"]}, images {p["use_boundary_images"]}')
        if seed is not None:
            self.reset_parameters(seed)

    # ---- parameters --------------------------------------------------------------------------
    @torch.no_grad()
    def reset_parameters(self, seed: int = 0, randomize_bn: bool = False) -> None:
        """Seeded random initialisation with the reference's distributions: fan-out Kaiming normal
        for encoder conv weights (extractor.py:91-92), PyTorch's default uniform(+-1/sqrt(fan_in))
        for everything else.  ``randomize_bn`` also draws non-trivial BatchNorm statistics so that
        BN folding is exercised."""
        g = torch.Generator().manual_seed(seed)
        for name, m in self.named_modules():
            if isinstance(m, _Conv):
                cout, cin, kh, kw = m.weight.shape
                bound = 1.0 / math.sqrt(cin * kh * kw)
                if name.startswith('update_block'):
                    m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 2 - 1) * bound)
                else:
                    m.weight.copy_(torch.randn(m.weight.shape, generator=g) * math.sqrt(2.0 / (cout * kh * kw)))
                m.bias.copy_((torch.rand(m.bias.shape, generator=g) * 2 - 1) * bound)
        seen = set()
        for m in self.modules():
            if isinstance(m, _BatchNorm) and id(m) not in seen:
                seen.add(id(m))
                c = m.weight.numel()
                if randomize_bn:
                    m.running_mean.copy_(torch.randn(c, generator=g) * 0.1)
                    m.running_var.copy_(torch.rand(c, generator=g) + 0.5)
                    m.weight.copy_(torch.rand(c, generator=g) + 0.5)
                    m.bias.copy_(torch.randn(c, generator=g) * 0.1)
                else:
                    m.running_mean.zero_(); m.running_var.fill_(1.0); m.weight.fill_(1.0); m.bias.zero_()
        self._engine = None

    def freeze_bn(self):
        """Kept for API parity (raft.py:75-78); BatchNorm here is always applied in eval form."""
        return None

    def _invalidate(self, *_):
        self._engine = None

    def _versions(self):
        """Fingerprint of the parameter / buffer storage the packed weight images were made from."""
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._engine = None
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._engine = None
        return out

    # ---- reference helpers kept for API parity ---------------------------------------------------
    def gen_voxel_grids(self, input_: torch.Tensor):
        """raft.py:88-99 (views only; the engine reads the windows in place)."""
        assert self.nbins_context + self.nbins_corr - 1 == input_.shape[-3]
        grids = [input_[:, i:i + self.nbins_corr] for i in [0] + list(self.ev_corr_target_indices)]
        return grids, input_[:, -self.nbins_context:]

    # ---- forward ---------------------------------------------------------------------------------
    def engine(self, device=None):
        device = torch.device(device) if device is not None else next(self.parameters()).device
        if device.type != 'cuda':
            raise RuntimeError('bflow_b200.RAFTSpline runs only on CUDA (sm_100a); there is no CPU path')
        ver = self._versions()
        if (self._engine is None or self._engine.device != device or self._engine.precision != self.precision or
                self._engine.corr_mode != self.correlation or self._engine_versions != ver):
            from .engine import Engine
            self._engine = Engine(self, device, self.precision, self.correlation)
            self._engine_versions = ver
        return self._engine

    @torch.no_grad()
    def forward(self,
                voxel_grid: Optional[torch.Tensor] = None,
                images: Optional[List[torch.Tensor]] = None,
                iters: int = 12,
                flow_init: Optional[BezierCurves] = None,
                test_mode: bool = False,
                non_blocking: bool = False):
        """Same contract as the reference (raft.py:101-200).  ``non_blocking=True`` (an extension) pipelines consecutive calls: inputs may be
        PINNED host tensors, the host-to-device copy of this call and the device-to-host copy of the previous results overlap the
        compute of the neighbouring calls, and the returned ``BezierCurves`` hold pinned HOST tensors that become valid when first
        accessed (``get_params`` / ``cpu`` / ``get_flow_from_reference`` wait on an event) and stay valid until the second-next such call."""
        assert voxel_grid is not None or images is not None
        assert iters > 0
        if self.fnet_ev is not None:
            assert voxel_grid is not None
            assert voxel_grid.ndim == 4
            assert self.nbins_context + self.nbins_corr - 1 == voxel_grid.shape[-3]
        if self.fnet_img is not None:
            assert images is not None and len(images) == 2
        ref = voxel_grid if voxel_grid is not None else images[0]
        assert ref.shape[-2] % 8 == 0 and ref.shape[-1] % 8 == 0
        # the coarsest pyramid level must keep at least 2 x 2 pixels: with a 1-pixel level the reference's bilinear_sampler divides by
        # (W - 1) = 0 (models/raft_utils/utils.py:13-14) and returns NaN flow, so there is nothing to be equal to
        lv = max(_cfg.levels_per_target(self.model_params))
        assert min(ref.shape[-2], ref.shape[-1]) // 8 >> (lv - 1) >= 2, \
            f'input {tuple(ref.shape[-2:])} too small for a {lv}-level correlation pyramid (the reference yields NaN here)'
        init = flow_init.get_params() if flow_init is not None else None
        if non_blocking:
            dev = next(self.parameters()).device
            if dev.type != 'cuda':
                raise RuntimeError('bflow_b200.RAFTSpline runs only on CUDA (sm_100a); move the module to a CUDA device first')
            low, ups, ev = self.engine(dev).run(voxel_grid, images, iters, init, test_mode, non_blocking=True)
            if test_mode:
                return BezierCurves(low, ready_event=ev), BezierCurves(ups[-1], ready_event=ev)
            return [BezierCurves(u, ready_event=ev) for u in ups]
        if not ref.is_cuda:
            raise RuntimeError('bflow_b200.RAFTSpline runs only on CUDA tensors (sm_100a); there is no CPU path')
        low, ups = self.engine(ref.device).run(voxel_grid, images, iters, init, test_mode)
        if test_mode:
            return BezierCurves(low), BezierCurves(ups[-1])
        return [BezierCurves(u) for u in ups]
