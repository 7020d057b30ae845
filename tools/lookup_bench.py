"""Correlation-lookup microbench (BASELINE config #5) — also the ncu target for the roofline kernel.

    python tools/lookup_bench.py [--batch 16] [--slots d|single] [--reps 20] [--nchw]
"""
import argparse
import ctypes
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bflow_b200 import ops, _lib, config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--slots', default='single', choices=['single', 'd', 'm'])
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--nchw', action='store_true')
    ap.add_argument('--h', type=int, default=60)
    ap.add_argument('--w', type=int, default=80)
    ap.add_argument('--noflush', action='store_true')
    ap.add_argument('--tiled', action='store_true')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    lib = _lib.lib()
    h, w, B = a.h, a.w, a.batch
    levels = {'single': [4], 'd': [1, 1, 1, 4], 'm': [1, 1, 1, 1, 4, 4]}[a.slots]
    table = config.slot_table(levels)
    T, S, R = len(levels), len(table), B * h * w
    g = torch.Generator().manual_seed(7)
    planes = {}
    for (l, t) in table:
        p = torch.randn(R, h >> l, w >> l, device=dev)
        planes[(l, t)] = ops.to_tiled(p) if a.tiled else p
    slots = [(l, t, planes[(l, t)], h >> l, w >> l) for (l, t) in table] if a.tiled else [(l, t, planes[(l, t)]) for (l, t) in table]
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    coords = (torch.stack([xs, ys], 0).float()[None, None] + 8 * torch.randn(T, B, 2, h, w, generator=g)).to(dev)
    out = torch.empty(B, S * 81, h, w, device=dev) if a.nchw else torch.empty(R, S * 81, device=dev)
    d = ops.make_lookup_desc(slots, T, B, h, w, a.tiled)
    d.coords, d.params, d.params_ld, d.degree = coords.data_ptr(), None, 0, 0
    d.out, d.out_nhwc, d.out_ld = out.data_ptr(), int(not a.nchw), S * 81
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    for _ in range(3):
        _lib.check(lib.bflow_corr_lookup(ctypes.byref(d), stream), 'lookup')
    ts = []
    for _ in range(a.reps):
        if not a.noflush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.bflow_corr_lookup(ctypes.byref(d), stream)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = statistics.median(ts)
    nbytes = R * (S * 724 + T * 8)
    print(f'lookup B={B} {h}x{w} slots={S} targets={T} layout={"nchw" if a.nchw else "nhwc"}{" tiled-volume" if a.tiled else ""}: median {t*1e3:.1f} us, min {min(ts)*1e3:.1f} us, '
          f'{nbytes/1e6:.2f} MB algorithmic -> {nbytes/(t*1e-3)/1e9:.0f} GB/s ({nbytes/(t*1e-3)/1e9/6552.6:.3f} of measured HBM peak)')


if __name__ == '__main__':
    main()
