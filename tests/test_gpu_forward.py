"""GPU parity of the whole drop-in forward against the reference fixtures and the CPU oracle.
Bar (BASELINE.json north_star): max over pixels of the final-flow EPE <= 1e-3 px."""
import numpy as np
import pytest
import torch

from conftest import load_golden, build_case, flow_epe
from oracle import raft_spline_oracle as O
from bflow_b200 import RAFTSpline, BezierCurves, config, synthetic

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
EPE_BAR = 1e-3


def run_cuda(net, vg, im, **kw):
    net = net.to(DEV)
    vgc = vg.to(DEV) if vg is not None else None
    imc = [t.to(DEV) for t in im] if im is not None else None
    return net(voxel_grid=vgc, images=imc, **kw)


@pytest.mark.parametrize('name', ['d_128_i4', 'd_128_i4_bn', 'm_128_i3_bn', 'd_480x640_i12', 'm_384x512_i12'])
def test_forward_matches_reference_fixture(name):
    g = load_golden(name)
    cfg, net, sd, vg, im = build_case(g)
    low, up = run_cuda(net, vg, im, iters=int(g['iters']), test_mode=True)
    assert isinstance(low, BezierCurves) and isinstance(up, BezierCurves)
    low, up = low.get_params().cpu(), up.get_params().cpu()
    mx, mean = flow_epe(low, torch.from_numpy(g['low']))
    assert 8 * mx <= EPE_BAR, f'low-res EPE {mx} (x8 in full-res pixels)'
    if 'up' in g.files:
        mx, mean = flow_epe(up, torch.from_numpy(g['up']))
        assert mx <= EPE_BAR, f'max EPE {mx}, mean {mean}'
        assert np.abs(up.numpy() - g['up']).max() <= EPE_BAR
    else:
        got = up.reshape(-1)[torch.from_numpy(g['up_index'])].numpy()
        assert np.abs(got - g['up_samples']).max() <= EPE_BAR / 1.4


@pytest.mark.parametrize('preset,B,H,W,iters,kind', [
    ('E_LU4_BD2', 2, 128, 160, 3, 'randn'),
    ('E_I_LU5_BD10', 1, 128, 128, 2, 'sparse_norm'),
])
def test_forward_matches_oracle_all_modes(preset, B, H, W, iters, kind):
    cfg = config.preset(preset)
    net = RAFTSpline(cfg, seed=None)
    net.reset_parameters(11, randomize_bn=True)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    vg, im = synthetic.inputs(cfg, B, H, W, seed=21, kind=kind)
    init = 0.5 * torch.randn(B, 2 * cfg['bezier_degree'], H // 8, W // 8, generator=torch.Generator().manual_seed(2))
    with torch.inference_mode():
        want_list = O.forward(sd, cfg, vg, im, iters=iters, test_mode=False)
        want_low, want_up = O.forward(sd, cfg, vg, im, iters=iters, flow_init=init, test_mode=True)
    got_list = run_cuda(net, vg, im, iters=iters, test_mode=False)
    assert isinstance(got_list, list) and len(got_list) == iters
    for a, b in zip(got_list, want_list):
        assert flow_epe(a.get_params().cpu(), b)[0] <= EPE_BAR
        assert (a.get_params().cpu() - b).abs().max() <= EPE_BAR
    low, up = run_cuda(net, vg, im, iters=iters, flow_init=BezierCurves(init.to(DEV)), test_mode=True)
    assert (low.get_params().cpu() - want_low).abs().max() * 8 <= EPE_BAR
    assert (up.get_params().cpu() - want_up).abs().max() <= EPE_BAR
    # and again without flow_init on the cached plan: the init buffer must have been cleared
    low2, up2 = run_cuda(net, vg, im, iters=iters, test_mode=True)
    assert (up2.get_params().cpu() - want_list[-1]).abs().max() <= EPE_BAR


def test_forward_cuda_core_fp32_path(monkeypatch):
    """BFLOW_TC=0: every convolution on the exact-fp32 CUDA-core kernel (the numerical anchor of the tcgen05 path)."""
    monkeypatch.setenv('BFLOW_TC', '0')
    g = load_golden('d_128_i4_bn')
    cfg, net, sd, vg, im = build_case(g)
    low, up = run_cuda(net, vg, im, iters=int(g['iters']), test_mode=True)
    assert net.engine().plan(int(g['B']), int(g['H']), int(g['W']), int(g['iters']), True).n_tc == 0
    assert np.abs(up.get_params().cpu().numpy() - g['up']).max() <= 1e-4


def test_tensor_core_path_is_the_default():
    cfg = config.preset('E_LU4_BD2')
    net = RAFTSpline(cfg, seed=1).to(DEV)
    vg, _ = synthetic.inputs(cfg, 1, 64, 96, seed=3)
    net(voxel_grid=vg.to(DEV), iters=1, test_mode=True)
    plan = net.engine().plan(1, 64, 96, 1, True)
    assert plan.n_tc >= 40, plan.n_tc           # all but the 7x7 stems and convf1 run on tcgen05
    plan.check()


def test_graph_replay_equals_eager_and_is_repeatable(monkeypatch):
    cfg = config.preset('E_LU4_BD2')
    vg, _ = synthetic.inputs(cfg, 1, 128, 160, seed=3)
    net = RAFTSpline(cfg, seed=2).to(DEV)
    a = net(voxel_grid=vg.to(DEV), iters=3, test_mode=True)[1].get_params().clone()
    b = net(voxel_grid=vg.to(DEV), iters=3, test_mode=True)[1].get_params().clone()
    monkeypatch.setenv('BFLOW_GRAPH', '0')
    net2 = RAFTSpline(cfg, seed=2).to(DEV)
    c = net2(voxel_grid=vg.to(DEV), iters=3, test_mode=True)[1].get_params()
    assert (a - b).abs().max() <= 1e-5      # atomics in the InstanceNorm sums: order noise only
    assert (a - c).abs().max() <= 1e-5


def test_batch_shards_are_independent():
    """rank r's result equals rows [r*b, (r+1)*b) of the single-GPU result."""
    cfg = config.preset('E_LU4_BD2')
    vg, _ = synthetic.inputs(cfg, 4, 64, 96, seed=9)
    net = RAFTSpline(cfg, seed=4).to(DEV)
    full = net(voxel_grid=vg.to(DEV), iters=2, test_mode=True)[1].get_params()
    for r in range(2):
        part = net(voxel_grid=vg[2 * r:2 * r + 2].to(DEV), iters=2, test_mode=True)[1].get_params()
        assert (part - full[2 * r:2 * r + 2]).abs().max() <= 1e-5


def test_host_tensor_input_is_rejected_loudly():
    net = RAFTSpline(config.preset('E_LU4_BD2')).to(DEV)
    with pytest.raises(RuntimeError, match='CUDA'):
        net(voxel_grid=torch.zeros(1, 9, 64, 64), iters=1, test_mode=True)
