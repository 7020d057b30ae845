"""Runs N eager forwards (no CUDA graph, one stream) of config D so that ncu sees every launch of a step in order.
    BFLOW_GRAPH=0 BFLOW_STREAMS=0 ncu ... python tools/one_step.py [--n 2] [--batch 1]"""
import argparse
import os
import sys

os.environ.setdefault('BFLOW_GRAPH', '0')
os.environ.setdefault('BFLOW_STREAMS', '0')
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bflow_b200 import RAFTSpline, config, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--preset', default='E_LU4_BD2')
ap.add_argument('--n', type=int, default=2)
ap.add_argument('--batch', type=int, default=1)
ap.add_argument('--h', type=int, default=480)
ap.add_argument('--w', type=int, default=640)
ap.add_argument('--iters', type=int, default=12)
a = ap.parse_args()
dev = torch.device('cuda:0')
cfg = config.preset(a.preset)
net = RAFTSpline(cfg, seed=0).to(dev)
vg, im = synthetic.inputs(cfg, a.batch, a.h, a.w)
vg = vg.to(dev) if vg is not None else None
im = [t.to(dev) for t in im] if im is not None else None
with torch.inference_mode():
    for _ in range(a.n):
        low, up = net(voxel_grid=vg, images=im, iters=a.iters, test_mode=True)
torch.cuda.synchronize()
print('launches per forward:', net.engine().plan(a.batch, a.h, a.w, a.iters, True).n_launches)
