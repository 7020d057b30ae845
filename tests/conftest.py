import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)


def state_checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def build_case(g):
    """Regenerates (cfg, model, state_dict, voxel, images) of a golden end-to-end fixture and proves
    the regenerated tensors are the ones the fixture was made from."""
    from bflow_b200 import RAFTSpline, config, synthetic
    cfg = config.preset(str(g['preset']))
    net = RAFTSpline(cfg, seed=None)
    net.reset_parameters(int(g['wseed']), randomize_bn=bool(g['bn']))
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    assert abs(state_checksum(sd) - float(g['weights_checksum'])) <= 1e-9 * float(g['weights_checksum'])
    vg, im = synthetic.inputs(cfg, int(g['B']), int(g['H']), int(g['W']), kind=str(g['kind']))
    if vg is not None:
        assert abs(float(vg.double().abs().sum()) - float(g['voxel_checksum'])) <= 1e-9 * float(g['voxel_checksum'])
    if im is not None:
        cs = float(im[0].double().abs().sum() + im[1].double().abs().sum())
        assert abs(cs - float(g['image_checksum'])) <= 1e-9 * float(g['image_checksum'])
    return cfg, net, sd, vg, im


def flow_epe(a: torch.Tensor, b: torch.Tensor):
    """max / mean over pixels of the end-point error between the final flows (last control point) of two
    (B, 2*deg, H, W) parameter tensors — the parity metric of BASELINE.json."""
    deg = a.shape[1] // 2
    fa = a.reshape(a.shape[0], 2, deg, *a.shape[2:])[:, :, -1]
    fb = b.reshape(b.shape[0], 2, deg, *b.shape[2:])[:, :, -1]
    e = torch.sqrt(((fa - fb) ** 2).sum(1))
    return float(e.max()), float(e.mean())
