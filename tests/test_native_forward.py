"""The whole-forward native entry (bflow_forward_load / run / destroy): a plan exported by the Python planner is replayed by a plain C
program -- no Python, no torch in that process -- and must reproduce RAFTSpline.forward bit for bit (same kernels, same launch list)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from bflow_b200 import RAFTSpline, config, synthetic, _lib


def build_native(tmp):
    exe = os.path.join(tmp, 'forward_main')
    libdir = os.path.dirname(_lib._build.LIB)
    subprocess.check_call(['gcc', '-O1', '-I', os.path.join(ROOT, 'include'), os.path.join(ROOT, 'tests', 'native', 'forward_main.c'), '-o', exe,
                           '-L', libdir, '-lbflow_b200', f'-Wl,-rpath,{libdir}'])
    return exe


def test_native_caller_builds_against_the_header_alone(tmp_path):
    """CPU: the C program compiles and links against include/bflow_b200.h + libbflow_b200.so only (no CUDA or torch headers)."""
    _lib.lib()
    exe = build_native(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and 'usage' in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize('preset,B,H,W,iters', [('E_LU4_BD2', 1, 128, 160, 3), ('E_I_LU5_BD10', 2, 128, 128, 2)])
def test_plan_file_replayed_by_a_c_program_equals_the_python_forward(tmp_path, preset, B, H, W, iters):
    tmp = str(tmp_path)
    exe = build_native(tmp)
    cfg = config.preset(preset)
    net = RAFTSpline(cfg, seed=None)
    net.reset_parameters(7, randomize_bn=True)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    vg, im = synthetic.inputs(cfg, B, H, W, seed=5)
    netc = net.to('cuda:0')
    low, up = netc(voxel_grid=vg.cuda() if vg is not None else None, images=[t.cuda() for t in im] if im is not None else None, iters=iters, test_mode=True)
    low, up = low.get_params().cpu().numpy(), up.get_params().cpu().numpy()
    torch.save(sd, os.path.join(tmp, 'w.pt'))
    plan = os.path.join(tmp, 'x.plan')
    # the exporter runs in its own process: the arena claims a fixed virtual address range of the process that uses it
    code = ('import sys, torch; sys.path.insert(0, %r); from bflow_b200 import config, export; '
            'print(export.export_plan(config.preset(%r), torch.load(%r), %d, %d, %d, %d, %r))' % (ROOT, preset, os.path.join(tmp, 'w.pt'), B, H, W, iters, plan))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    files = []
    for name, t in (('vox', vg), ('im0', im[0] if im else None), ('im1', im[1] if im else None)):
        if t is None:
            files.append('-')
        else:
            p = os.path.join(tmp, name + '.f32')
            t.numpy().astype('<f4').tofile(p)
            files.append(p)
    out_low, out_up = os.path.join(tmp, 'low.f32'), os.path.join(tmp, 'up.f32')
    r = subprocess.run([exe, plan] + files + [out_low, out_up, '3'], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got_low = np.fromfile(out_low, dtype='<f4').reshape(low.shape)
    got_up = np.fromfile(out_up, dtype='<f4').reshape(up.shape)
    # same kernels, same launch order; only the atomics of the InstanceNorm sums reorder
    assert np.abs(got_low - low).max() <= 1e-5 and np.abs(got_up - up).max() <= 1e-5
    assert os.path.getsize(plan) < 200 << 20
