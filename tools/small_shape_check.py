"""Forward vs the CPU oracle on small / odd input sizes (rows per image below one 128-row tile at the coarse levels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bflow_b200 import RAFTSpline, config, synthetic  # noqa: E402
from oracle import raft_spline_oracle as O  # noqa: E402

for preset, B, H, W, iters, seed in [('E_LU4_BD2', 1, 64, 96, 2, 1), ('E_LU4_BD2', 1, 64, 64, 2, 1), ('E_LU4_BD2', 2, 64, 96, 2, 1), ('E_LU4_BD2', 1, 128, 128, 2, 1),
                                     ('E_LU4_BD2', 1, 96, 128, 2, 1), ('E_LU4_BD2', 1, 72, 88, 2, 1), ('E_I_LU5_BD10', 1, 64, 96, 2, 1)]:
    cfg = config.preset(preset)
    net = RAFTSpline(cfg, seed=seed)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    vg, im = synthetic.inputs(cfg, B, H, W, seed=3)
    with torch.inference_mode():
        want = O.forward(sd, cfg, vg, im, iters=iters, test_mode=True)[1]
    for env in ({}, {'BFLOW_TC': '0'}):
        os.environ.pop('BFLOW_TC', None)
        os.environ.update(env)
        n2 = RAFTSpline(cfg, seed=seed).to('cuda:0')
        got = n2(voxel_grid=vg.cuda() if vg is not None else None, images=[t.cuda() for t in im] if im is not None else None, iters=iters, test_mode=True)[1]
        err = (got.get_params().cpu() - want).abs().max().item()
        print(f'{preset} B={B} {H}x{W} iters={iters} {env}: max |diff| vs oracle {err:.3e}', flush=True)
    os.environ.pop('BFLOW_TC', None)
