"""How far is the REFERENCE ITSELF, run on the GPU in PyTorch's default TF32 mode, from its own CPU fp32 result on the golden fixtures?
(context for the reduced-precision configuration, row g).   python tools/ref_tf32_on_fixture.py d_480x640_i12 m_384x512_i12"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_golden, build_case, flow_epe  # noqa: E402
from oracle import ref_loader  # noqa: E402

for name in sys.argv[1:]:
    g = load_golden(name)
    cfg, net, sd, vg, im = build_case(g)
    ref = ref_loader.build(cfg)
    ref.load_state_dict(sd, strict=True)
    ref = ref.cuda()
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.inference_mode():
            low, up = ref(voxel_grid=vg.cuda() if vg is not None else None, images=[t.cuda() for t in im] if im is not None else None,
                          iters=int(g['iters']), test_mode=True)
        mx, mean = flow_epe(low.get_params().cpu(), torch.from_numpy(g['low']))
        print(f'{name}: reference on cuda, allow_tf32={tf32}: low-res max EPE x8 = {8 * mx:.3e} px, mean x8 = {8 * mean:.3e} px', flush=True)
