"""Repeats the small reference fixtures N times on one GPU and prints (8 x low-res max EPE, full-res max EPE) per run: run-to-run spread of the
parity margin (atomics order in the InstanceNorm sums) -- the margin to the 1e-3 px bar is ~10x.
    python tools/flake_probe.py [N]"""
import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import load_golden
from test_gpu_forward import build_case, run_cuda, flow_epe
for name in ['d_128_i4', 'd_128_i4_bn', 'm_128_i3_bn']:
    g = load_golden(name)
    errs = []
    for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
        cfg, net, sd, vg, im = build_case(g)
        low, up = run_cuda(net, vg, im, iters=int(g['iters']), test_mode=True)
        low, up = low.get_params().cpu(), up.get_params().cpu()
        mx, mean = flow_epe(low, torch.from_numpy(g['low']))
        mxu = flow_epe(up, torch.from_numpy(g['up']))[0] if 'up' in g.files else -1
        errs.append((round(8 * float(mx), 6), round(float(mxu), 6)))
    print(name, errs)
