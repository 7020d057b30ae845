"""CPU oracle (numpy) for scope rows (f1)/(f2).  TEST INFRASTRUCTURE ONLY.

voxel_grid  : VoxelGrid.convert  (data/utils/representations.py:64-111)
norm_voxel  : norm_voxel_grid    (data/utils/representations.py:9-18)
epe_masked  : epe_masked         (utils/metrics.py:196-213)
flow_metrics: epe_masked, ae_masked (:259-296), n_pixel_error_masked (:161-193) and the linear-assumption prediction (:298-305)
Pinned against the reference's own functions through tests/golden/events.npz and tests/golden/metrics.npz
(oracle/make_golden.py: make_events, make_metrics).
"""
import numpy as np


def voxel_grid(x, y, pol, time, C, H, W, t0, t1):
    x, y, pol, time = map(np.asarray, (x, y, pol, time))
    out = np.zeros(C * H * W, dtype=np.float32)
    # torch: int64 tensor / python int -> float32 true division, then * (C-1) in float32
    tn = (time - t0).astype(np.float32) / np.float32(t1 - t0) * np.float32(C - 1)
    tf = np.floor(tn).astype(np.int32)
    value = (2 * pol.astype(np.float32) - 1).astype(np.float32)
    if np.issubdtype(x.dtype, np.integer):
        for tl in (tf, tf + 1):
            m = (tl >= 0) & (tl < C)
            wgt = value * (1 - np.abs(tl.astype(np.float32) - tn))
            idx = H * W * tl.astype(np.int64) + W * y.astype(np.int64) + x.astype(np.int64)
            np.add.at(out, idx[m], wgt[m].astype(np.float32))
    else:
        x = x.astype(np.float32)
        y = y.astype(np.float32)
        x0, y0 = np.floor(x).astype(np.int32), np.floor(y).astype(np.int32)
        for xl in (x0, x0 + 1):
            for yl in (y0, y0 + 1):
                for tl in (tf, tf + 1):
                    m = (xl < W) & (xl >= 0) & (yl < H) & (yl >= 0) & (tl >= 0) & (tl < C)
                    wgt = value * (1 - np.abs(xl - x)) * (1 - np.abs(yl - y)) * (1 - np.abs(tl.astype(np.float32) - tn))
                    idx = H * W * tl.astype(np.int64) + W * yl.astype(np.int64) + xl.astype(np.int64)
                    np.add.at(out, idx[m], wgt[m].astype(np.float32))
    return out.reshape(C, H, W)


def norm_voxel(v):
    v = np.array(v, dtype=np.float32, copy=True)
    nz = v != 0
    if nz.sum() > 0:
        vals = v[nz].astype(np.float32)
        mean = vals.mean(dtype=np.float64)
        std = vals.std(ddof=1, dtype=np.float64) if vals.size > 1 else float('nan')
        v[nz] = (vals - np.float32(mean)) / np.float32(std) if std > 0 else vals - np.float32(mean)
    return v


def epe_masked(src, tgt, valid=None):
    e = np.sqrt(((np.asarray(src, np.float32) - np.asarray(tgt, np.float32)) ** 2).sum(1))
    if valid is not None:
        e = e[np.asarray(valid).astype(bool)]
    return float(e.astype(np.float64).sum()), int(e.size)


def flow_metrics(src, tgt, valid=None, n_pixels=(), scale=1.0):
    """dict(epe, ae_deg, npe{n}: percent, count) of scale * src against tgt over the valid pixels; (N, C, *) arrays."""
    src = np.asarray(src, np.float32) * np.float32(scale)
    tgt = np.asarray(tgt, np.float32)
    m = np.ones(src.shape[:1] + src.shape[2:], bool) if valid is None else np.asarray(valid).astype(bool)
    e = np.sqrt(((src - tgt) ** 2).sum(1))
    ext_s = np.concatenate([src, np.ones_like(src[:, :1])], 1)
    ext_t = np.concatenate([tgt, np.ones_like(tgt[:, :1])], 1)
    cs = (ext_s * ext_t).sum(1) / (np.linalg.norm(ext_s, axis=1) * np.linalg.norm(ext_t, axis=1))
    ae = np.arccos(np.clip(cs, -1.0, 1.0))
    n = int(m.sum())
    out = {'count': n,
           'epe': float(e[m].astype(np.float64).sum() / n) if n else None,
           'ae_deg': float(ae[m].astype(np.float64).sum() / n / np.pi * 180) if n else float('nan')}
    rel = e / np.clip(np.sqrt((tgt ** 2).sum(1)), 1e-6, None)
    for k in n_pixels:
        bad = (e > k) & (rel >= 0.05)
        out[f'npe{k}'] = float(bad[m].sum() / n * 100) if n else float('nan')
    return out
