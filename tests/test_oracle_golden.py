"""CPU: the oracle (oracle/raft_spline_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  This is what pins the oracle on machines without /root/reference."""
import numpy as np
import pytest
import torch

from conftest import load_golden, build_case
from oracle import raft_spline_oracle as O
from bflow_b200 import synthetic

TOL = 2e-5   # reference self-noise under kernel reordering is <= 1.5e-5 px


@pytest.mark.parametrize('name', ['d_128_i4', 'd_128_i4_bn', 'm_128_i3_bn', 'd_480x640_i12', 'm_384x512_i12'])
def test_forward_matches_reference_fixture(name):
    g = load_golden(name)
    cfg, net, sd, vg, im = build_case(g)
    with torch.inference_mode():
        low, up = O.forward(sd, cfg, vg, im, iters=int(g['iters']), test_mode=True)
    assert np.abs(low.numpy() - g['low']).max() < TOL
    if 'up' in g.files:
        assert np.abs(up.numpy() - g['up']).max() < TOL
    else:
        got = up.reshape(-1)[torch.from_numpy(g['up_index'])].numpy()
        assert np.abs(got - g['up_samples']).max() < TOL


def test_lookup_matches_reference_fixture():
    g = load_golden('lookup_16x24')
    B, h, w, D = int(g['B']), int(g['h']), int(g['w']), int(g['D'])
    levels = [int(v) for v in g['levels']]
    f1, f2, coords = synthetic.lookup_case(B, h, w, dim=D, targets=len(levels), seed=7)
    assert abs(float(f1.double().abs().sum()) - float(g['f1_checksum'])) < 1e-6
    assert abs(float(coords.double().abs().sum()) - float(g['coords_checksum'])) < 1e-6
    vol = O.corr_volume(f1, f2)
    pyr = O.corr_pyramid(vol, levels)
    idx = g['lvl0_index']
    assert np.abs(vol.reshape(-1).numpy()[idx] - g['lvl0_samples']).max() < 1e-5
    for lvl in (1, 2, 3):
        assert pyr[lvl][0] == [3]
        assert np.abs(pyr[lvl][1].numpy() - g[f'lvl{lvl}']).max() < 1e-5
    out = O.corr_lookup(pyr, coords)
    assert out.shape == g['out'].shape
    assert np.abs(out.numpy() - g['out']).max() < 2e-5
    assert O.slot_table(levels) == [(0, 0), (0, 1), (0, 2), (0, 3), (1, 3), (2, 3), (3, 3)]


@pytest.mark.parametrize('deg', [1, 2, 10])
def test_bezier_matches_reference_fixture(deg):
    g = load_golden('bezier')
    p = torch.from_numpy(g[f'deg{deg}_params'])
    ts = [float(t) for t in g['ts']]
    assert np.abs(O.bezier_flow(p, ts).numpy() - g[f'deg{deg}_flows']).max() < 1e-6
    assert np.abs(O.bezier_flow(p, [0.5])[0].numpy() - g[f'deg{deg}_scalar_half']).max() < 1e-6
    up = O.cvx_upsample(p, torch.from_numpy(g[f'deg{deg}_mask']))
    assert np.abs(up.numpy() - g[f'deg{deg}_up']).max() < 1e-5


def test_known_answers():
    # Bernstein rows for degree 2 at t = 1/4, 1/2
    c = O.bezier_coeffs([0.25, 0.5, 1.0], 2)
    assert torch.allclose(c, torch.tensor([[0.375, 0.0625], [0.5, 0.25], [0.0, 1.0]]))
    # convex upsampling of a constant field returns 8*const in the interior for any mask
    up = O.cvx_upsample(torch.full((1, 2, 5, 6), 0.5), torch.randn(1, 576, 5, 6))
    assert torch.allclose(up[:, :, 8:-8, 8:-8], torch.full((1, 2, 24, 32), 4.0), atol=1e-5)
    # a lookup centred on integer coordinates returns the volume entry at the centre tap
    vol = torch.randn(1, 12, 3, 4)
    co = O.coords_grid(1, 3, 4)[None]
    out = O.corr_lookup([([0], vol)], co)
    centre = out[0, 40].reshape(-1)
    assert torch.allclose(centre, vol[0].reshape(12, 12).diagonal(), atol=1e-6)
