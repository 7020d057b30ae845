"""Scope rows (f1)/(f2): events -> voxel grid, voxel normalisation, masked EPE.  CPU: oracle vs reference fixture; GPU: kernels vs both."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import events_oracle as E


def test_events_oracle_matches_reference_fixture():
    g = load_golden('events')
    C_, H, W = int(g['C']), int(g['H']), int(g['W'])
    assert np.abs(E.voxel_grid(g['xi'], g['yi'], g['pol'], g['t'], C_, H, W, 5000, 45000) - g['out_int']).max() < 1e-6
    assert np.abs(E.voxel_grid(g['xf'], g['yf'], g['pol'], g['t'], C_, H, W, 5000, 45000) - g['out_flt']).max() < 2e-6
    assert np.abs(E.voxel_grid(g['xi'], g['yi'], g['pol'], g['t'], C_, H, W, int(g['t'][0]), int(g['t'][-1])) - g['out_default']).max() < 1e-6
    assert np.abs(E.norm_voxel(g['out_int']) - g['normed']).max() < 1e-6
    # degenerate inputs: empty grid and a single non-zero (std undefined -> only the mean is removed)
    assert np.all(E.norm_voxel(np.zeros((2, 3, 3), np.float32)) == 0)
    one = np.zeros((2, 3, 3), np.float32)
    one[0, 1, 1] = 2.5
    assert np.all(E.norm_voxel(one) == 0)


def test_metrics_oracle_matches_reference_fixture():
    """oracle/events_oracle.flow_metrics against the reference's own utils/metrics.py outputs (tests/golden/metrics.npz)."""
    g = load_golden('metrics')
    src, tgt, valid = g['src'], g['tgt'], g['valid']
    a = E.flow_metrics(src, tgt, None, n_pixels=(1,))
    v = E.flow_metrics(src, tgt, valid, n_pixels=(2, 3))
    assert abs(a['epe'] - float(g['epe'])) < 1e-5 and abs(v['epe'] - float(g['epe_v'])) < 1e-5
    assert abs(a['ae_deg'] - float(g['ae'])) < 1e-3 and abs(v['ae_deg'] - float(g['ae_v'])) < 1e-3
    assert abs(a['npe1'] - float(g['npe1'])) < 1e-4
    assert abs(v['npe2'] - float(g['npe2_v'])) < 1e-4 and abs(v['npe3'] - float(g['npe3_v'])) < 1e-4
    lin = [E.flow_metrics(src, t, None, scale=float(ts)) for ts, t in zip(g['ts'], g['tgts'])]
    assert abs(np.mean([m['epe'] for m in lin]) - float(g['epe_multi_lin'])) < 1e-5
    assert abs(np.mean([m['ae_deg'] for m in lin]) - float(g['ae_multi_lin'])) < 1e-3


@pytest.mark.gpu
def test_flow_metrics_on_device_match_reference_fixture():
    """Row (f2): bflow_flow_metrics through the reference-named mirrors of bflow_b200.metrics."""
    from bflow_b200 import metrics as M
    g = load_golden('metrics')
    dev = 'cuda:0'
    src, tgt, valid = (torch.from_numpy(g[k]).to(dev) for k in ('src', 'tgt', 'valid'))
    assert abs(float(M.epe_masked(src, tgt)) - float(g['epe'])) < 1e-5
    assert abs(float(M.epe_masked(src, tgt, valid)) - float(g['epe_v'])) < 1e-5
    assert abs(float(M.ae_masked(src, tgt)) - float(g['ae'])) < 2e-3
    assert abs(float(M.ae_masked(src, tgt, valid)) - float(g['ae_v'])) < 2e-3
    assert abs(float(M.ae_masked(src, tgt, valid, degrees=False)) - float(g['ae_rad_v'])) < 1e-4
    assert abs(float(M.n_pixel_error_masked(src, tgt, None, 1)) - float(g['npe1'])) < 1e-3
    assert abs(float(M.n_pixel_error_masked(src, tgt, valid, 2)) - float(g['npe2_v'])) < 1e-3
    assert abs(float(M.n_pixel_error_masked(src, tgt, valid, 3)) - float(g['npe3_v'])) < 1e-3
    ts = [float(t) for t in g['ts']]
    tgts = [torch.from_numpy(t).to(dev) for t in g['tgts']]
    epe_lin, ae_lin = M.lin_assumption_metrics(src, ts, tgts)
    assert abs(float(epe_lin) - float(g['epe_multi_lin'])) < 1e-5 and abs(float(ae_lin) - float(g['ae_multi_lin'])) < 2e-3
    epe_lin_v, ae_lin_v = M.lin_assumption_metrics(src, ts, tgts, [valid] * len(ts))
    assert abs(float(epe_lin_v) - float(g['epe_multi_lin_v'])) < 1e-5 and abs(float(ae_lin_v) - float(g['ae_multi_lin_v'])) < 2e-3
    # the materialised form of the linear-assumption predictions through the *_multi mirrors gives the same numbers
    lin = [p.tensor() for p in M.predictions_from_lin_assumption(src, ts)]
    assert abs(float(M.epe_masked_multi(lin, tgts)) - float(g['epe_multi_lin'])) < 1e-5
    assert abs(float(M.ae_masked_multi(lin, tgts)) - float(g['ae_multi_lin'])) < 2e-3
    # no valid pixel: epe_masked returns None (metrics.py:210-211), and the multi form skips it
    none = torch.zeros_like(valid)
    assert M.epe_masked(src, tgt, none) is None
    assert M.epe_masked_multi([src, src], [tgt, tgt], [none, none]) is None
    # large random case against the numpy oracle
    gen = torch.Generator().manual_seed(9)
    a, b = torch.randn(2, 2, 480, 640, generator=gen) * 5, torch.randn(2, 2, 480, 640, generator=gen) * 5
    m = torch.rand(2, 480, 640, generator=gen) > 0.5
    want = E.flow_metrics(a.numpy(), b.numpy(), m.numpy(), n_pixels=(1, 3))
    assert abs(float(M.epe_masked(a.to(dev), b.to(dev), m.to(dev))) - want['epe']) < 1e-5 * want['epe']
    assert abs(float(M.ae_masked(a.to(dev), b.to(dev), m.to(dev))) - want['ae_deg']) < 2e-3
    assert abs(float(M.n_pixel_error_masked(a.to(dev), b.to(dev), m.to(dev), 3)) - want['npe3']) < 1e-3


@pytest.mark.gpu
def test_voxelize_out_of_range_integer_events_raise_and_never_write():
    """an integer event outside the sensor used to be written out of bounds.  The reference's put_ raises; so do we."""
    from bflow_b200.events import VoxelGrid
    dev = 'cuda:0'
    C_, H, W = 3, 8, 10
    guard = torch.zeros(4 * C_ * H * W, device=dev)                     # the grid is carved from the middle of a guarded buffer
    t = torch.tensor([0, 5, 10, 10], device=dev)
    pol = torch.tensor([1, 0, 1, 1], device=dev, dtype=torch.bool)
    x = torch.tensor([1, W, -1, 3], device=dev)
    y = torch.tensor([1, 2, 3, H + 5], device=dev)
    with pytest.raises(IndexError, match='3 events'):
        VoxelGrid(C_, H, W).convert(x, y, pol, t, 0, 10)
    out = VoxelGrid(C_, H, W, check_bounds=False).convert(x, y, pol, t, 0, 10)
    want = E.voxel_grid(np.array([1]), np.array([1]), np.array([True]), np.array([0]), C_, H, W, 0, 10)
    assert np.abs(out.cpu().numpy() - want).max() == 0                     # only the in-range event landed
    assert float(guard.abs().sum()) == 0


@pytest.mark.gpu
def test_voxelize_and_norm_match_reference_fixture():
    from bflow_b200.events import VoxelGrid, norm_voxel_grid
    g = load_golden('events')
    C_, H, W = int(g['C']), int(g['H']), int(g['W'])
    dev = 'cuda:0'
    t = torch.from_numpy(g['t']).to(dev)
    pol = torch.from_numpy(g['pol']).to(dev)
    vg = VoxelGrid(C_, H, W)
    out = vg.convert(torch.from_numpy(g['xi']).to(dev), torch.from_numpy(g['yi']).to(dev), pol, t, 5000, 45000)
    assert np.abs(out.cpu().numpy() - g['out_int']).max() < 2e-5          # float atomics: summation order differs from the CPU loop
    out_f = vg.convert(torch.from_numpy(g['xf']).to(dev), torch.from_numpy(g['yf']).to(dev), pol, t, 5000, 45000)
    assert np.abs(out_f.cpu().numpy() - g['out_flt']).max() < 2e-5
    out_d = vg.convert(torch.from_numpy(g['xi']).to(dev), torch.from_numpy(g['yi']).to(dev), pol, t)
    assert np.abs(out_d.cpu().numpy() - g['out_default']).max() < 2e-5
    normed = norm_voxel_grid(torch.from_numpy(g['out_int']).to(dev))
    assert np.abs(normed.cpu().numpy() - g['normed']).max() < 2e-5
    assert float(norm_voxel_grid(torch.zeros(2, 4, 4, device=dev)).abs().sum()) == 0
    one = torch.zeros(2, 4, 4, device=dev)
    one[1, 2, 2] = -3.0
    assert float(norm_voxel_grid(one).abs().sum()) == 0
    empty = vg.convert(torch.zeros(0, dtype=torch.long, device=dev), torch.zeros(0, dtype=torch.long, device=dev),
                       torch.zeros(0, dtype=torch.bool, device=dev), torch.zeros(0, dtype=torch.long, device=dev), 0, 10)
    assert float(empty.abs().sum()) == 0
    assert vg.get_extended_time_window(1000, 5000) == (0, 6000)


@pytest.mark.gpu
def test_voxelize_large_random_against_oracle():
    from bflow_b200.events import VoxelGrid
    gen = torch.Generator().manual_seed(3)
    n, C_, H, W = 300000, 9, 96, 128
    t = torch.sort(torch.randint(0, 100000, (n,), generator=gen)).values
    x = torch.randint(0, W, (n,), generator=gen)
    y = torch.randint(0, H, (n,), generator=gen)
    pol = torch.randint(0, 2, (n,), generator=gen).bool()
    want = E.voxel_grid(x.numpy(), y.numpy(), pol.numpy(), t.numpy(), C_, H, W, 10000, 90000)
    got = VoxelGrid(C_, H, W).convert(x.cuda(), y.cuda(), pol.cuda(), t.cuda(), 10000, 90000).cpu().numpy()
    assert np.abs(got - want).max() < 1e-4
    assert abs(float(got.sum()) - float(want.sum())) < 1e-2


@pytest.mark.gpu
def test_epe_masked_on_device():
    from bflow_b200.events import epe_sum_count
    gen = torch.Generator().manual_seed(5)
    src, tgt = torch.randn(3, 2, 40, 56, generator=gen), torch.randn(3, 2, 40, 56, generator=gen)
    valid = torch.rand(3, 40, 56, generator=gen) > 0.4
    s, n = epe_sum_count(src.cuda(), tgt.cuda(), valid.cuda())
    ws, wn = E.epe_masked(src.numpy(), tgt.numpy(), valid.numpy())
    assert int(n) == wn and abs(float(s) - ws) < 1e-6 * ws
    s, n = epe_sum_count(src.cuda(), tgt.cuda())
    ws, wn = E.epe_masked(src.numpy(), tgt.numpy())
    assert int(n) == wn and abs(float(s) - ws) < 1e-6 * ws
    s, n = epe_sum_count(src.cuda(), tgt.cuda(), torch.zeros(3, 40, 56, dtype=torch.bool).cuda())     # no valid pixel (metrics.py:210-211)
    assert int(n) == 0 and float(s) == 0.0
