"""CPU fp32 oracle for bflow's RAFT-spline inference forward pass.  TEST INFRASTRUCTURE ONLY.

A pure-functional restatement (plain torch CPU tensor ops on a ``state_dict``) of what the
reference computes on the path ``RAFTSpline.forward`` (models/raft_spline/raft.py:101-200).
It is written from the behaviour documented in SURVEY.md §8a, not transcribed: e.g. the
correlation lookup is an explicit 4-corner gather with per-corner zero padding instead of
``grid_sample``, Bézier coefficients come from ``math.comb`` instead of numba/scipy, and the
convex upsampling is an explicit 9-neighbour loop instead of ``unfold``.

Parity pin: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, imported unmodified from
``/root/reference`` in the build container (``oracle/ref_loader.py``): live in
``tests/test_oracle_vs_reference.py`` whenever ``/root/reference`` exists, and through the
committed fixtures in ``tests/golden/`` (made by ``oracle/make_golden.py``) everywhere else.

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# encoders  (models/raft_utils/extractor.py:47-55, 103-125)
# --------------------------------------------------------------------------------------
def _conv(sd: SD, name: str, x: Tensor, stride=1, padding=0) -> Tensor:
    return F.conv2d(x, sd[name + '.weight'], sd[name + '.bias'], stride=stride, padding=padding)


def _norm(sd: SD, name: str, x: Tensor, kind: str) -> Tensor:
    """extractor.py:21-31 — instance: per-(n,c) biased variance, eps 1e-5, no affine;
    batch (eval): running statistics + affine."""
    if kind == 'instance':
        mu = x.mean(dim=(2, 3), keepdim=True)
        var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
        return (x - mu) * torch.rsqrt(var + 1e-5)
    if kind == 'batch':
        rm, rv = sd[name + '.running_mean'], sd[name + '.running_var']
        g, b = sd[name + '.weight'], sd[name + '.bias']
        s = g * torch.rsqrt(rv + 1e-5)
        return x * s[None, :, None, None] + (b - rm * s)[None, :, None, None]
    if kind == 'none':
        return x
    raise NotImplementedError(kind)


def _res_block(sd: SD, p: str, x: Tensor, kind: str, stride: int) -> Tensor:
    """extractor.py:47-55."""
    y = torch.relu(_norm(sd, p + '.norm1', _conv(sd, p + '.conv1', x, stride, 1), kind))
    y = torch.relu(_norm(sd, p + '.norm2', _conv(sd, p + '.conv2', y, 1, 1), kind))
    if stride != 1:
        x = _norm(sd, p + '.norm3', _conv(sd, p + '.downsample.0', x, stride, 0), kind)
    return torch.relu(x + y)


def basic_encoder(sd: SD, p: str, x: Tensor, kind: str) -> Tensor:
    """extractor.py:103-125 (the list→batch concatenation is done by the caller)."""
    x = torch.relu(_norm(sd, p + '.norm1', _conv(sd, p + '.conv1', x, 2, 3), kind))
    for layer, stride in (('layer1', 1), ('layer2', 2), ('layer3', 2)):
        x = _res_block(sd, f'{p}.{layer}.0', x, kind, stride)
        x = _res_block(sd, f'{p}.{layer}.1', x, kind, 1)
    return _conv(sd, p + '.conv2', x)


# --------------------------------------------------------------------------------------
# correlation volume, pyramid, lookup  (models/raft_utils/corr.py)
# --------------------------------------------------------------------------------------
def corr_volume(fmap1: Tensor, fmap2: Tensor) -> Tensor:
    """corr.py:264-272.  fmap1 (T,B,D,h,w) or (B,D,h,w); fmap2 (T,B,D,h,w) → (T, B*h*w, h, w)."""
    T, B, D, h, w = fmap2.shape
    if fmap1.ndim == 4:
        fmap1 = fmap1[None].expand(T, -1, -1, -1, -1)
    a = fmap1.reshape(T, B, D, h * w).transpose(-1, -2)
    b = fmap2.reshape(T, B, D, h * w)
    c = (a @ b) / math.sqrt(D)
    return c.reshape(T, B * h * w, h, w)


def corr_pyramid(vol: Tensor, levels_per_target: Sequence[int]) -> List[Tuple[List[int], Tensor]]:
    """corr.py:293-305 + 108-125.  Returns per level: (base-target indices at that level,
    tensor (n_targets_at_level, B*Q, h_l, w_l)).  Level l>=1 is avg_pool2d(2,2) (floor) of level l-1
    restricted to the targets whose level count exceeds l."""
    T = vol.shape[0]
    assert len(levels_per_target) == T
    pyr = [(list(range(T)), vol)]
    for lvl in range(2, max(levels_per_target) + 1):
        prev_idx, prev = pyr[-1]
        keep = [t for t in range(T) if levels_per_target[t] >= lvl]
        sel = torch.stack([prev[prev_idx.index(t)] for t in keep], dim=0)
        n, bq, hh, ww = sel.shape
        down = F.avg_pool2d(sel.reshape(n * bq, 1, hh, ww), 2, stride=2).reshape(n, bq, hh // 2, ww // 2)
        pyr.append((keep, down))
    return pyr


def slot_table(levels_per_target: Sequence[int]) -> List[Tuple[int, int]]:
    """Output slot order of the lookup: level-major, then ascending base-target index
    (corr.py:322-346: outer loop over pyramid levels, torch.cat over levels)."""
    out = []
    for lvl in range(max(levels_per_target)):
        for t, n in enumerate(levels_per_target):
            if n > lvl:
                out.append((lvl, t))
    return out


def corr_lookup(pyr, coords: Tensor, radius: int = 4) -> Tensor:
    """corr.py:307-350 + utils.py:5-21.  coords (T,B,2,h,w) pixel coordinates (x,y) in the level-0
    target plane → (B, S*(2r+1)^2, h, w); channel = slot*81 + iy*9 + ix, dy = iy-r, dx = ix-r.
    Bilinear with align_corners=True semantics in pixel units, zero contribution from any corner
    outside the plane."""
    T, B, _, h, w = coords.shape
    n = 2 * radius + 1
    d = torch.arange(-radius, radius + 1, dtype=coords.dtype)
    outs = []
    for lvl, (tidx, vol) in enumerate(pyr):
        hl, wl = vol.shape[-2:]
        for j, t in enumerate(tidx):
            c = coords[t].permute(0, 2, 3, 1).reshape(B * h * w, 2) / (2 ** lvl)
            # the reference maps to [-1,1] and grid_sample maps back (utils.py:13-14,19)
            gx = 2 * (c[:, 0, None, None] + d[None, None, :]) / (wl - 1) - 1      # (BQ,1,n)
            gy = 2 * (c[:, 1, None, None] + d[None, :, None]) / (hl - 1) - 1      # (BQ,n,1)
            x = ((gx + 1) / 2) * (wl - 1)
            y = ((gy + 1) / 2) * (hl - 1)
            x = x.expand(-1, n, n)
            y = y.expand(-1, n, n)
            x0 = torch.floor(x)
            y0 = torch.floor(y)
            fx = x - x0
            fy = y - y0
            x0 = x0.long()
            y0 = y0.long()
            plane = vol[j].reshape(B * h * w, hl * wl)
            acc = torch.zeros_like(x)
            for oy, wy in ((0, 1 - fy), (1, fy)):
                for ox, wx in ((0, 1 - fx), (1, fx)):
                    xi = x0 + ox
                    yi = y0 + oy
                    ok = (xi >= 0) & (xi <= wl - 1) & (yi >= 0) & (yi <= hl - 1)
                    lin = (yi.clamp(0, hl - 1) * wl + xi.clamp(0, wl - 1)).reshape(B * h * w, n * n)
                    v = torch.gather(plane, 1, lin).reshape(B * h * w, n, n)
                    acc = acc + torch.where(ok, v * wy * wx, torch.zeros_like(v))
            outs.append(acc.reshape(B, h, w, n * n))
    out = torch.stack(outs, dim=1)                       # (B, S, h, w, 81)
    return out.permute(0, 1, 4, 2, 3).reshape(B, -1, h, w).float()


# --------------------------------------------------------------------------------------
# Bézier curves  (models/raft_spline/bezier.py)
# --------------------------------------------------------------------------------------
def bezier_coeffs(timestamps: Sequence[float], degree: int) -> Tensor:
    """bezier.py:141-163,178-180: C(n,i)(1-t)^(n-i) t^i for i=1..n in float64, cast to fp32."""
    rows = []
    for t in timestamps:
        t = float(t)
        assert 0.0 <= t <= 1.0
        rows.append([math.comb(degree, i) * (1.0 - t) ** (degree - i) * t ** i for i in range(1, degree + 1)])
    return torch.tensor(rows, dtype=torch.float64).float()


def bezier_flow(params: Tensor, timestamps: Sequence[float]) -> Tensor:
    """bezier.py:134-135,165-186: params (B, 2*deg, h, w) with channel = dim*deg + (i-1) → (T,B,2,h,w)."""
    B, C, h, w = params.shape
    deg = C // 2
    coef = bezier_coeffs(timestamps, deg).to(params.dtype)
    return torch.einsum('bdphw,tp->tbdhw', params.reshape(B, 2, deg, h, w), coef)


def coords_grid(B: int, h: int, w: int) -> Tensor:
    """utils.py:24-30: channel 0 = x (column), channel 1 = y (row)."""
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    return torch.stack([xs, ys], dim=0).float()[None].repeat(B, 1, 1, 1)


def cvx_upsample(data: Tensor, mask: Tensor) -> Tensor:
    """utils.py:33-48: mask channel = k*64 + i*8 + j, k = ky*3+kx; softmax over k;
    out[n,c,8y+i,8x+j] = sum_k softmax_k * 8*data[n,c,y+ky-1,x+kx-1] (zero padded)."""
    N, C, H, W = data.shape
    m = torch.softmax(mask.reshape(N, 9, 8, 8, H, W), dim=1)
    dp = F.pad(8 * data, (1, 1, 1, 1))
    out = torch.zeros(N, C, 8, 8, H, W, dtype=data.dtype)
    for ky in range(3):
        for kx in range(3):
            nb = dp[:, :, ky:ky + H, kx:kx + W]                       # (N,C,H,W)
            out = out + m[:, ky * 3 + kx][:, None] * nb[:, :, None, None]
    return out.permute(0, 1, 4, 2, 5, 3).reshape(N, C, 8 * H, 8 * W)


# --------------------------------------------------------------------------------------
# update block  (models/raft_spline/update.py)
# --------------------------------------------------------------------------------------
def motion_encoder(sd: SD, p: str, bezier: Tensor, corr: Tensor) -> Tensor:
    """update.py:88-97."""
    cor = torch.relu(_conv(sd, p + '.convc1', corr))
    cor = torch.relu(_conv(sd, p + '.convc2', cor, 1, 1))
    bez = torch.relu(_conv(sd, p + '.convf1', bezier, 1, 3))
    bez = torch.relu(_conv(sd, p + '.convf2', bez, 1, 1))
    out = torch.relu(_conv(sd, p + '.conv', torch.cat([cor, bez], 1), 1, 1))
    return torch.cat([out, bezier], 1)


def sep_conv_gru(sd: SD, p: str, h: Tensor, x: Tensor) -> Tensor:
    """update.py:33-48."""
    for sfx, pad in (('1', (0, 2)), ('2', (2, 0))):
        hx = torch.cat([h, x], 1)
        z = torch.sigmoid(_conv(sd, f'{p}.convz{sfx}', hx, 1, pad))
        r = torch.sigmoid(_conv(sd, f'{p}.convr{sfx}', hx, 1, pad))
        q = torch.tanh(_conv(sd, f'{p}.convq{sfx}', torch.cat([r * h, x], 1), 1, pad))
        h = (1 - z) * h + z * q
    return h


def update_block(sd: SD, p: str, net: Tensor, inp: Tensor, corr: Tensor, bezier: Tensor):
    """update.py:116-126 → (net, mask, delta_bezier)."""
    mot = motion_encoder(sd, p + '.encoder', bezier, corr)
    net = sep_conv_gru(sd, p + '.gru', net, torch.cat([inp, mot], 1))
    delta = _conv(sd, p + '.bezier_head.conv2', torch.relu(_conv(sd, p + '.bezier_head.conv1', net, 1, 1)), 1, 1)
    mask = 0.25 * _conv(sd, p + '.mask.2', torch.relu(_conv(sd, p + '.mask.0', net, 1, 1)))
    return net, mask, delta


# --------------------------------------------------------------------------------------
# full forward  (models/raft_spline/raft.py:88-200)
# --------------------------------------------------------------------------------------
def forward(sd: SD, cfg: dict, voxel_grid: Optional[Tensor] = None, images: Optional[List[Tensor]] = None,
            iters: int = 12, flow_init: Optional[Tensor] = None, test_mode: bool = True,
            taps: Optional[dict] = None):
    """Returns (params_low (B,2deg,h,w), params_up (B,2deg,H,W)) in test mode, else the list of
    ``iters`` upsampled parameter tensors.  ``taps`` (a dict) receives stage outputs keyed by the
    reference's timer labels (raft.py:116-186)."""
    use_ev = cfg['use_events']
    use_img = cfg['use_boundary_images']
    nctx = cfg['num_bins']['context']
    ncorr = cfg['num_bins']['correlation']
    deg = cfg['bezier_degree']
    hdim = cfg['hidden']['dim']
    cdim = cfg['context']['dim']
    fnorm = cfg['feature']['norm']
    cnorm = cfg['context']['norm']
    tidx = list(cfg['correlation']['ev']['target_indices']) if use_ev else []
    levels: List[int] = []
    f1s, f2s = [], []
    context = None
    if use_ev:
        voxel_grid = voxel_grid.contiguous().float()
        B = voxel_grid.shape[0]
        assert voxel_grid.shape[1] == nctx + ncorr - 1
        wins = [voxel_grid[:, i:i + ncorr] for i in [0] + tidx]                    # raft.py:88-99
        fm = basic_encoder(sd, 'fnet_ev', torch.cat(wins, 0), fnorm).float()
        fm = fm.reshape(len(wins), B, *fm.shape[1:])
        f1s.append(fm[0][None].expand(len(tidx), -1, -1, -1, -1))
        f2s.append(fm[1:])
        levels += list(cfg['correlation']['ev']['levels'])
        context = voxel_grid[:, -nctx:]
        if taps is not None:
            taps['fnet_ev'] = fm
    if use_img:
        assert len(images) == 2
        imgs = [2 * (x.float().contiguous() / 255) - 1 for x in images]           # raft.py:134
        B = imgs[0].shape[0]
        fm = basic_encoder(sd, 'fnet_img', torch.cat(imgs, 0), fnorm)
        fm = fm.reshape(2, B, *fm.shape[1:])
        f1s.append(fm[0][None])
        f2s.append(fm[1][None])
        levels.append(int(cfg['correlation']['img']['levels']))
        context = imgs[0] if context is None else torch.cat([context, imgs[0]], 1)
        if taps is not None:
            taps['fnet_img'] = fm
    cn = basic_encoder(sd, 'cnet', context, cnorm)                                 # raft.py:144-147
    net = torch.tanh(cn[:, :hdim])
    inp = torch.relu(cn[:, hdim:hdim + cdim])
    B, _, H, W = context.shape
    h, w = H // 8, W // 8
    coords0 = coords_grid(B, h, w)
    params = torch.zeros(B, 2 * deg, h, w)
    if flow_init is not None:
        params = params + flow_init
    vol = corr_volume(torch.cat(f1s, 0), torch.cat(f2s, 0))
    pyr = corr_pyramid(vol, levels)
    if taps is not None:
        taps['net0'], taps['inp'] = net, inp
        taps['pyramid'] = pyr
    dt = 1.0 / (nctx - 1)
    ts = [dt * i for i in tidx] + ([1] if use_img else [])                         # raft.py:170-177
    ups = []
    up = None
    for itr in range(iters):
        coords1 = coords0[None] + bezier_flow(params, ts)
        corr = corr_lookup(pyr, coords1)
        net, mask, delta = update_block(sd, 'update_block', net, inp, corr, params)
        params = params + delta
        if taps is not None and itr == 0:
            taps['corr0'], taps['net1'], taps['delta0'], taps['mask0'] = corr, net, delta, mask
        if (not test_mode) or itr == iters - 1:
            up = cvx_upsample(params, mask)
            ups.append(up)
    if test_mode:
        return params, up
    return ups


def epe_masked(src: Tensor, tgt: Tensor, valid: Optional[Tensor] = None):
    """utils/metrics.py:196-213: per-pixel sqrt(sum_c (src-tgt)^2); returns (sum, count)."""
    e = torch.sqrt(((src - tgt) ** 2).sum(dim=1))
    if valid is not None:
        e = e[valid.reshape(e.shape).bool()]
    return e.double().sum(), e.numel()
