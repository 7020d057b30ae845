"""CPU, build container only: the oracle against the LIVE reference imported from /root/reference."""
import pytest
import torch

from oracle import ref_loader, raft_spline_oracle as O
from bflow_b200 import RAFTSpline, config, synthetic

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason='/root/reference not present (GPU box)')


@pytest.mark.parametrize('preset,B,H,W,iters', [('E_LU4_BD2', 1, 128, 128, 3), ('E_I_LU5_BD10', 1, 128, 160, 2)])
def test_oracle_equals_live_reference(preset, B, H, W, iters):
    cfg = config.preset(preset)
    net = RAFTSpline(cfg, seed=None)
    net.reset_parameters(3, randomize_bn=True)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    ref = ref_loader.build(cfg)
    ref.load_state_dict(sd, strict=True)          # proves state_dict key/shape compatibility as well
    vg, im = synthetic.inputs(cfg, B, H, W, seed=5, kind='randn')
    with torch.inference_mode():
        lo_r, up_r = ref(voxel_grid=vg, images=im, iters=iters, test_mode=True)
        lo_o, up_o = O.forward(sd, cfg, vg, im, iters=iters, test_mode=True)
        lst_r = ref(voxel_grid=vg, images=im, iters=iters, test_mode=False)
        lst_o = O.forward(sd, cfg, vg, im, iters=iters, test_mode=False)
    assert (lo_r.get_params() - lo_o).abs().max() < 2e-5
    assert (up_r.get_params() - up_o).abs().max() < 2e-5
    assert len(lst_r) == len(lst_o) == iters
    for a, b in zip(lst_r, lst_o):
        assert (a.get_params() - b).abs().max() < 2e-5


def test_bezier_class_matches_reference():
    import numpy as np
    ref_loader.load()
    from models.raft_spline.bezier import BezierCurves as RefBezier
    from bflow_b200 import BezierCurves
    p = torch.randn(2, 20, 5, 7)
    a, b = RefBezier(p), BezierCurves(p)
    for t in (0, 0.0, 0.3, 1, 1.0, [0.1, 0.9], np.array([0.25, 0.5])):
        assert torch.allclose(a.get_flow_from_reference(t), b.get_flow_from_reference(t), atol=1e-6)
    for attr in ('batch_size', 'degree', 'dim', 'height', 'width', 'requires_grad', 'n_ctrl_pts'):
        assert getattr(a, attr) == getattr(b, attr)
    assert torch.equal(a.detach(cpu=True).get_params(), b.detach(cpu=True).get_params())
