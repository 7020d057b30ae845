"""CPU oracle (numpy) for scope rows (f1)/(f2).  TEST INFRASTRUCTURE ONLY.

voxel_grid  : VoxelGrid.convert  (data/utils/representations.py:64-111)
norm_voxel  : norm_voxel_grid    (data/utils/representations.py:9-18)
epe_masked  : epe_masked         (utils/metrics.py:196-213)
Pinned against the reference's own functions through tests/golden/events.npz (oracle/make_golden.py: make_events).
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
