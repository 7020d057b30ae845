"""Per-launch device times of one forward step (eager launches, CUDA events), grouped by launch label.
    python tools/step_profile.py [--preset E_LU4_BD2] [--batch 1] [--h 480] [--w 640] [--iters 12] [--all]"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bflow_b200 import RAFTSpline, config, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--preset', default='E_LU4_BD2')
    ap.add_argument('--batch', type=int, default=1)
    ap.add_argument('--h', type=int, default=480)
    ap.add_argument('--w', type=int, default=640)
    ap.add_argument('--iters', type=int, default=12)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--all', action='store_true')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    cfg = config.preset(a.preset)
    net = RAFTSpline(cfg, seed=0).to(dev)
    vg, im = synthetic.inputs(cfg, a.batch, a.h, a.w)
    vg = vg.to(dev) if vg is not None else None
    im = [t.to(dev) for t in im] if im is not None else None
    net(voxel_grid=vg, images=im, iters=a.iters, test_mode=True)
    plan = net.engine().plan(a.batch, a.h, a.w, a.iters, True)
    plan.load_inputs(vg, im, None)
    stream = torch.cuda.current_stream().cuda_stream
    acc = collections.OrderedDict()
    for rep in range(a.reps):
        evs = []
        for (fn, args), (label, flops) in zip(plan.launches, plan.labels):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(*args, stream); e1.record()
            evs.append((label, flops, e0, e1))
        torch.cuda.synchronize()
        for label, flops, e0, e1 in evs:
            s = acc.setdefault(label, [0, 0.0, flops])
            s[0] += 1; s[1] += e0.elapsed_time(e1)
    tot = sum(v[1] for v in acc.values()) / a.reps
    print(f'{a.preset} B={a.batch} {a.h}x{a.w} iters={a.iters}: sum of launches {tot:.3f} ms/step, {len(plan.launches)} launches')
    rows = sorted(acc.items(), key=lambda kv: -kv[1][1])
    for label, (n, ms, flops) in rows if a.all else rows[:40]:
        per = ms / n
        tf = f'{flops / (per * 1e-3) / 1e12:7.1f} TF/s' if flops else ''
        print(f'{label:50s} x{n // a.reps:4d}  {per * 1e3:9.1f} us each  {ms / a.reps:8.3f} ms/step {100 * ms / a.reps / tot:5.1f}%  {tf}')


if __name__ == '__main__':
    main()
