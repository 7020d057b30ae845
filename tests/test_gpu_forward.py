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


@pytest.mark.parametrize('name', ['d_128_i4', 'd_128_i4_bn', 'm_128_i3_bn', 'd_480x640_i12', 'm_384x512_i12', 'd_480x640_i12_b4',
                                  'm_384x512_i12_b4'])
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
        if int(g['B']) == 1:
            # the fixture samples the full-resolution output; every pixel of it is checked against the oracle port (itself pinned to the
            # reference at 2e-5 by tests/test_oracle_vs_reference.py and the sampled values above)
            with torch.inference_mode():
                want_low, want_up = O.forward(sd, cfg, vg, im, iters=int(g['iters']), test_mode=True)
            assert np.abs(want_up.reshape(-1)[torch.from_numpy(g['up_index'])].numpy() - g['up_samples']).max() <= 5e-5
            assert flow_epe(up, want_up)[0] <= EPE_BAR
    net.engine().plan(int(g['B']), int(g['H']), int(g['W']), int(g['iters']), True).check()      # no pipeline wait timed out


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
    vg, _ = synthetic.inputs(cfg, 1, 128, 136, seed=3)
    net(voxel_grid=vg.to(DEV), iters=1, test_mode=True)
    plan = net.engine().plan(1, 128, 136, 1, True)
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
    vg, _ = synthetic.inputs(cfg, 4, 128, 136, seed=9)
    net = RAFTSpline(cfg, seed=4).to(DEV)
    full = net(voxel_grid=vg.to(DEV), iters=2, test_mode=True)[1].get_params()
    for r in range(2):
        part = net(voxel_grid=vg[2 * r:2 * r + 2].to(DEV), iters=2, test_mode=True)[1].get_params()
        assert (part - full[2 * r:2 * r + 2]).abs().max() <= 1e-5


# Deviation of the REFERENCE ITSELF from its CPU fp32 result when it runs on the B200 in PyTorch's default GPU arithmetic (cuDNN / cuBLAS
# TF32, 10-bit mantissa operands) on the same fixtures: max over pixels of 8 x low-res final-flow EPE, measured with
# tools/ref_tf32_on_fixture.py (profiles/r02_reference_tf32_on_fixtures.txt).  These fixtures use randomised BatchNorm statistics, which
# amplify operand rounding ~5x over the benchmark weights.
REF_TF32_DEVIATION = {'d_480x640_i12': 3.32e-2, 'm_384x512_i12': 4.32e-2, 'd_128_i4_bn': 2.09e-2}


@pytest.mark.parametrize('name', ['d_480x640_i12', 'm_384x512_i12', 'd_128_i4_bn'])
def test_reduced_precision_configuration_on_the_fixtures(name):
    """Row (g): precision='f16' (one fp16 MMA per product, hi planes only) against the reference fixtures.  With these stress weights
    no single-pass 16-bit-operand arithmetic reaches 1e-2 px -- the reference's own TF32 GPU run is 2-4e-2 away from its CPU result --
    so the gate here is "no worse than 1.5 x the reference's default GPU arithmetic"; the 1e-2 px bar of BASELINE.json is asserted on
    the benchmark configuration below.  The result must also differ from the fp32-equivalent one (no silent fallback)."""
    g = load_golden(name)
    cfg, net, sd, vg, im = build_case(g)
    net.precision = 'f16'
    low, up = run_cuda(net, vg, im, iters=int(g['iters']), test_mode=True)
    assert net.engine().prec == 1
    net.engine().plan(int(g['B']), int(g['H']), int(g['W']), int(g['iters']), True).check()
    low = low.get_params().cpu()
    mx, mean = flow_epe(low, torch.from_numpy(g['low']))
    assert 8 * mx <= 1.5 * REF_TF32_DEVIATION[name], f'low-res EPE {mx} (x8 in full-res pixels), mean {mean}'
    assert 8 * mx > 1e-5, 'identical to the split path: the reduced-precision kernels did not run'
    net.precision = 'f32x3'
    low3 = run_cuda(net, vg, im, iters=int(g['iters']), test_mode=True)[0].get_params().cpu()
    assert net.engine().prec == 0 and 8 * flow_epe(low3, torch.from_numpy(g['low']))[0] <= EPE_BAR


def test_reduced_precision_configuration_meets_1e_2_on_the_benchmark_workload():
    """Row (g), the bar itself: BASELINE.json configs[1] exactly as bench.py runs it (E_LU4_BD2, 640x480, 12 iterations, seed-0 weights,
    sparse_norm events seed 1234) with precision='f16' against the CPU oracle: max final-flow EPE <= 1e-2 px."""
    cfg = config.preset('E_LU4_BD2')
    net = RAFTSpline(cfg, seed=0, precision='f16')
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    vg, _ = synthetic.inputs(cfg, 1, 480, 640, seed=1234)
    with torch.inference_mode():
        want_low, want_up = O.forward(sd, cfg, vg, None, iters=12, test_mode=True)
    low, up = run_cuda(net, vg, None, iters=12, test_mode=True)
    mx, mean = flow_epe(up.get_params().cpu(), want_up)
    assert mx <= 1e-2, (mx, mean)
    assert mx > 1e-4


@pytest.mark.parametrize('name', ['d_128_i4_bn', 'm_128_i3_bn', 'd_480x640_i12'])
def test_on_the_fly_correlation_forward_matches_reference_fixture(name):
    """Row (f3): correlation='otf' -- no correlation volume, lookups computed from a pooled target-feature pyramid -- meets the same 1e-3 px
    bar against the reference fixtures, and really is the other path (no volume tensor in the plan)."""
    g = load_golden(name)
    cfg, net, sd, vg, im = build_case(g)
    net.correlation = 'otf'
    low, up = run_cuda(net, vg, im, iters=int(g['iters']), test_mode=True)
    plan = net.engine().plan(int(g['B']), int(g['H']), int(g['W']), int(g['iters']), True)
    assert net.engine().corr_mode == 'otf' and not hasattr(plan, 'vol0') and any('corr_lookup_otf' == lab for lab, _ in plan.labels)
    low, up = low.get_params().cpu(), up.get_params().cpu()
    mx, mean = flow_epe(low, torch.from_numpy(g['low']))
    assert 8 * mx <= EPE_BAR, f'low-res EPE {mx} (x8 in full-res pixels)'
    if 'up' in g.files:
        assert flow_epe(up, torch.from_numpy(g['up']))[0] <= EPE_BAR
    else:
        got = up.reshape(-1)[torch.from_numpy(g['up_index'])].numpy()
        assert np.abs(got - g['up_samples']).max() <= EPE_BAR / 1.4


def test_pipelined_forward_with_host_buffers_equals_blocking_forward():
    """forward(non_blocking=True): pinned host inputs, H2D / graph / D2H on three streams over two alternating buffer sets."""
    cfg = config.preset('E_LU4_BD2')
    net = RAFTSpline(cfg, seed=6).to(DEV)
    frames = [synthetic.inputs(cfg, 1, 128, 160, seed=30 + i, pinned=True)[0] for i in range(5)]
    want = [net(voxel_grid=f.to(DEV), iters=2, test_mode=True)[1].get_params().cpu() for f in frames]
    pending = [None, None]
    got = []
    for i, f in enumerate(frames):               # classic software pipeline: submit i, collect i-1
        cur = net(voxel_grid=f, iters=2, test_mode=True, non_blocking=True)
        assert cur[1].height == 128 and cur[1].batch_size == 1          # shape accessors do not wait
        if pending[0] is not None:
            got.append(pending[1].get_params().clone())
        pending = cur
    got.append(pending[1].get_params().clone())
    assert not got[0].is_cuda
    for a, b in zip(got, want):
        assert (a - b).abs().max() <= 1e-5
    # a blocking call afterwards still works on lane 0 and waits for the pipelined work
    again = net(voxel_grid=frames[0].to(DEV), iters=2, test_mode=True)[1].get_params().cpu()
    assert (again - want[0]).abs().max() <= 1e-5
    with pytest.raises(AssertionError, match='pinned'):
        net(voxel_grid=torch.zeros(1, 9, 128, 160), iters=2, test_mode=True, non_blocking=True)


def test_plan_cache_is_bounded(monkeypatch):
    monkeypatch.setenv('BFLOW_MAX_PLANS', '2')
    cfg = config.preset('E_LU4_BD2')
    net = RAFTSpline(cfg, seed=1).to(DEV)
    for hw in ((128, 128), (128, 136), (136, 128), (128, 128)):
        vg, _ = synthetic.inputs(cfg, 1, hw[0], hw[1], seed=3)
        net(voxel_grid=vg.to(DEV), iters=1, test_mode=True)
        assert len(net.engine()._plans) <= 2
    assert list(net.engine()._plans)[-1][:3] == (1, 128, 128)


def test_in_place_weight_update_rebuilds_the_engine():
    cfg = config.preset('E_LU4_BD2')
    net = RAFTSpline(cfg, seed=1).to(DEV)
    vg, _ = synthetic.inputs(cfg, 1, 128, 136, seed=3)
    a = net(voxel_grid=vg.to(DEV), iters=2, test_mode=True)[1].get_params().clone()
    with torch.no_grad():
        net.update_block.bezier_head.conv2.weight.mul_(2.0)             # what an optimizer step / EMA swap does
    b = net(voxel_grid=vg.to(DEV), iters=2, test_mode=True)[1].get_params().clone()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    with torch.inference_mode():
        want = O.forward({k: v.cpu() for k, v in sd.items()}, cfg, vg, None, iters=2, test_mode=True)[1]
    assert (a - b).abs().max() > 1e-4
    assert (b.cpu() - want).abs().max() <= EPE_BAR


def test_host_tensor_input_is_rejected_loudly():
    net = RAFTSpline(config.preset('E_LU4_BD2')).to(DEV)
    with pytest.raises(RuntimeError, match='CUDA'):
        net(voxel_grid=torch.zeros(1, 9, 128, 128), iters=1, test_mode=True)
    with pytest.raises(AssertionError, match='too small'):      # a 1-pixel pyramid level: the reference returns NaN there
        net(voxel_grid=torch.zeros(1, 9, 64, 96, device=DEV), iters=1, test_mode=True)


@pytest.mark.gpu
def test_capture_survives_garbage_of_an_earlier_engine():
    """A dead engine (reference cycles: only the cyclic collector frees its CUDAGraph and streams) must not be collected while the next plan is
    being captured -- destroying a graph from the collector invalidates a global-mode capture.  Collector thresholds at 1 make any unguarded
    capture hit it."""
    import gc
    cfg = config.preset('E_LU4_BD2')
    vg, _ = synthetic.inputs(cfg, 1, 128, 160)
    old = gc.get_threshold()
    try:
        net = RAFTSpline(cfg, seed=0).to(DEV)
        first = net(voxel_grid=vg.to(DEV), iters=2, test_mode=True)[1].get_params().clone()
        del net                                   # garbage now, still uncollected
        gc.set_threshold(1, 1, 1)
        net2 = RAFTSpline(cfg, seed=0).to(DEV)
        second = net2(voxel_grid=vg.to(DEV), iters=2, test_mode=True)[1].get_params()
        assert (first - second).abs().max() <= 1e-5      # same weights, same input; InstanceNorm atomics: order noise only
    finally:
        gc.set_threshold(*old)
