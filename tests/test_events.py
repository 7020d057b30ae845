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
