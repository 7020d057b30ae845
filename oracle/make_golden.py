"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_loader.py) on seeded synthetic inputs.

Run in the build container:   python -m oracle.make_golden
Weights: ``bflow_b200.RAFTSpline.reset_parameters(seed)`` (deterministic CPU generator) loaded into the
reference with ``load_state_dict(strict=True)``.  Inputs: ``bflow_b200.synthetic``.  Each fixture stores
checksums of the weights and inputs so that a test can prove it regenerated the same tensors.
Large outputs are stored at ``n_samples`` seeded pixel positions instead of in full.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bflow_b200 import RAFTSpline, config, synthetic   # noqa: E402
from oracle import ref_loader                          # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')

# name, preset, B, H, W, iters, weight seed, randomise BN, input kind, full upsampled output?
CASES = [
    ('d_128_i4',        'E_LU4_BD2',    1, 128, 128, 4, 0, False, 'sparse_norm', True),
    ('d_128_i4_bn',     'E_LU4_BD2',    2, 128, 160, 4, 1, True,  'randn',       True),
    ('m_128_i3_bn',     'E_I_LU5_BD10', 2, 128, 160, 3, 0, True,  'sparse_norm', False),
    ('d_480x640_i12',   'E_LU4_BD2',    1, 480, 640, 12, 0, True, 'sparse_norm', False),
    ('m_384x512_i12',   'E_I_LU5_BD10', 1, 384, 512, 12, 0, True, 'sparse_norm', False),
    # BASELINE configs #3 (M at batch 4) and #4 (D at batch 4 per GPU): the batch-4 code paths (24-sample fnet, 12 288-row update
    # block, per-pixel-walk lookup) at full size
    ('m_384x512_i12_b4', 'E_I_LU5_BD10', 4, 384, 512, 12, 0, True, 'sparse_norm', False),
    ('d_480x640_i12_b4', 'E_LU4_BD2',    4, 480, 640, 12, 0, True, 'sparse_norm', False),
]
N_SAMPLES = 40000


def checksum(t):
    return float(t.double().abs().sum())


def sample_index(numel, seed=99):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (min(N_SAMPLES, numel),), generator=g)


def state_checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def make_case(name, preset, B, H, W, iters, wseed, bn, kind, full):
    cfg = config.preset(preset)
    net = RAFTSpline(cfg, seed=None)
    net.reset_parameters(wseed, randomize_bn=bn)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    ref = ref_loader.build(cfg)
    ref.load_state_dict(sd, strict=True)
    vg, im = synthetic.inputs(cfg, B, H, W, kind=kind)
    with torch.inference_mode():
        low, up = ref(voxel_grid=vg, images=im, iters=iters, test_mode=True)
    low, up = low.get_params().float(), up.get_params().float()
    out = dict(preset=preset, B=B, H=H, W=W, iters=iters, wseed=wseed, bn=bn, kind=kind,
               weights_checksum=state_checksum(sd),
               voxel_checksum=checksum(vg) if vg is not None else 0.0,
               image_checksum=checksum(im[0]) + checksum(im[1]) if im is not None else 0.0,
               low=low.numpy())
    if full:
        out['up'] = up.numpy()
    else:
        idx = sample_index(up.numel())
        out['up_index'] = idx.numpy()
        out['up_samples'] = up.reshape(-1)[idx].numpy()
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)
    print(name, 'low', tuple(low.shape), 'up', tuple(up.shape), 'max|flow|', float(up.abs().max()))


def make_lookup():
    """Reference CorrComputation + CorrBlockParallelMultiTarget (models/raft_utils/corr.py) on the
    config-#5 style case at a fixture-sized feature map."""
    ref_loader.load()
    from models.raft_utils.corr import CorrComputation, CorrBlockParallelMultiTarget
    B, h, w, D = 1, 16, 24, 64
    levels = [1, 1, 1, 4]
    f1, f2, coords = synthetic.lookup_case(B, h, w, dim=D, targets=len(levels), seed=7)
    cc = CorrComputation(f1, f2, num_levels_per_target=levels)
    blk = CorrBlockParallelMultiTarget(corr_computation_events=cc)
    out = blk(coords)
    pyr = [c.corr.squeeze(2).numpy() for c in blk._corr_pyramid]
    idx = sample_index(pyr[0].size)
    np.savez_compressed(os.path.join(GOLD, 'lookup_16x24.npz'), B=B, h=h, w=w, D=D, levels=np.array(levels),
                        f1_checksum=checksum(f1), f2_checksum=checksum(f2), coords_checksum=checksum(coords),
                        out=out.numpy(), lvl0_index=idx.numpy(), lvl0_samples=pyr[0].reshape(-1)[idx.numpy()],
                        lvl0_checksum=float(np.abs(pyr[0].astype('float64')).sum()),
                        lvl1=pyr[1], lvl2=pyr[2], lvl3=pyr[3])
    print('lookup', tuple(out.shape), [p.shape for p in pyr])


def make_bezier():
    ref_loader.load()
    from models.raft_spline.bezier import BezierCurves
    g = torch.Generator().manual_seed(3)
    out = {}
    for deg in (1, 2, 10):
        p = torch.randn(2, 2 * deg, 4, 6, generator=g)
        mask = torch.randn(2, 576, 4, 6, generator=g)
        b = BezierCurves(p)
        ts = [0.0, 0.2, 0.25, 0.5, 0.75, 1.0]
        out[f'deg{deg}_params'] = p.numpy()
        out[f'deg{deg}_mask'] = mask.numpy()
        out[f'deg{deg}_flows'] = b.get_flow_from_reference(ts).numpy()
        out[f'deg{deg}_scalar_half'] = b.get_flow_from_reference(0.5).numpy()
        out[f'deg{deg}_scalar_one'] = b.get_flow_from_reference(1.0).numpy()
        out[f'deg{deg}_up'] = b.create_upsampled(mask).get_params().numpy()
    out['ts'] = np.array(ts)
    np.savez_compressed(os.path.join(GOLD, 'bezier.npz'), **out)
    print('bezier ok')


def make_events():
    """Reference VoxelGrid.convert / norm_voxel_grid (data/utils/representations.py) on seeded synthetic events."""
    import importlib
    ref_loader.load()
    rep = importlib.import_module('data.utils.representations')
    g = torch.Generator().manual_seed(11)
    C_, H, W, n = 5, 24, 32, 6000
    t = torch.sort(torch.randint(1000, 51000, (n,), generator=g)).values
    xi = torch.randint(0, W, (n,), generator=g)
    yi = torch.randint(0, H, (n,), generator=g)
    xf = torch.rand(n, generator=g) * (W + 1) - 1          # some events fall outside the sensor
    yf = torch.rand(n, generator=g) * (H + 1) - 1
    pol = torch.randint(0, 2, (n,), generator=g).bool()
    vg = rep.VoxelGrid(C_, H, W)
    out_int = vg.convert(xi, yi, pol, t, 5000, 45000)
    out_flt = vg.convert(xf, yf, pol, t, 5000, 45000)
    out_default = vg.convert(xi, yi, pol, t)
    normed = rep.norm_voxel_grid(out_int.clone())
    np.savez_compressed(os.path.join(GOLD, 'events.npz'), C=C_, H=H, W=W, t=t.numpy(), xi=xi.numpy(), yi=yi.numpy(), xf=xf.numpy(), yf=yf.numpy(),
                        pol=pol.numpy(), out_int=out_int.numpy(), out_flt=out_flt.numpy(), out_default=out_default.numpy(), normed=normed.numpy())
    print('events ok', float(out_int.abs().sum()), float(out_flt.abs().sum()))


def make_metrics():
    """Reference utils/metrics.py functions (its only missing import, torchmetrics.Metric, is a base class of the Metric wrappers
    around them and is shimmed by an empty class) on seeded flow pairs."""
    import importlib
    import types
    ref_loader.load()
    if 'torchmetrics' not in sys.modules:
        tm = types.ModuleType('torchmetrics')
        tm.Metric = type('Metric', (), {})
        sys.modules['torchmetrics'] = tm
    sys.path.insert(0, ref_loader.LIVE_ROOT)
    M = importlib.import_module('utils.metrics')
    g = torch.Generator().manual_seed(17)
    N, H, W = 3, 40, 56
    tgt = torch.randn(N, 2, H, W, generator=g) * 6
    src = tgt + torch.randn(N, 2, H, W, generator=g) * torch.tensor([0.2, 2.0, 6.0]).view(3, 1, 1, 1)
    tgt[0, :, :4] = 0                                   # zero ground truth: the relative-error clip path
    valid = torch.rand(N, H, W, generator=g) > 0.3
    ts = [0.2, 0.4, 0.6, 0.8, 1.0]
    tgts = [tgt * t + 0.3 * torch.randn(N, 2, H, W, generator=g) for t in ts]
    lin = M.predictions_from_lin_assumption(src, ts)
    out = dict(src=src.numpy(), tgt=tgt.numpy(), valid=valid.numpy(), ts=np.array(ts), tgts=np.stack([t.numpy() for t in tgts]),
               epe=float(M.epe_masked(src, tgt)), epe_v=float(M.epe_masked(src, tgt, valid)),
               ae=float(M.ae_masked(src, tgt)), ae_v=float(M.ae_masked(src, tgt, valid)), ae_rad_v=float(M.ae_masked(src, tgt, valid, degrees=False)),
               npe1=float(M.n_pixel_error_masked(src, tgt, None, 1)), npe2_v=float(M.n_pixel_error_masked(src, tgt, valid, 2)),
               npe3_v=float(M.n_pixel_error_masked(src, tgt, valid, 3)),
               epe_multi_lin=float(M.epe_masked_multi(lin, tgts)), ae_multi_lin=float(M.ae_masked_multi(lin, tgts)),
               epe_multi_lin_v=float(M.epe_masked_multi(lin, tgts, [valid] * len(ts))),
               ae_multi_lin_v=float(M.ae_masked_multi(lin, tgts, [valid] * len(ts))))
    np.savez_compressed(os.path.join(GOLD, 'metrics.npz'), **out)
    print('metrics ok', out['epe_v'], out['ae_v'], out['npe2_v'], out['epe_multi_lin'])


if __name__ == '__main__':
    assert ref_loader.live(), 'needs /root/reference (build container only)'
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    only = set(sys.argv[1:])
    want = lambda name: not only or name in only
    if want('bezier'):
        make_bezier()
    if want('lookup'):
        make_lookup()
    if want('events'):
        make_events()
    if want('metrics'):
        make_metrics()
    for c in CASES:
        if want(c[0]):
            make_case(*c)
