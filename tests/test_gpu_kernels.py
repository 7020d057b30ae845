"""GPU parity of every C-ABI kernel against the CPU oracle / torch CPU fp32 on seeded inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from oracle import raft_spline_oracle as O
from bflow_b200 import ops, synthetic, BezierCurves
from bflow_b200.bezier import bernstein_coeffs

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def g(seed=0):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize('N,Cin,H,W,Cout,k,s,p,act', [
    (2, 5, 20, 28, 64, (7, 7), 2, (3, 3), 'relu'),        # encoder stem, odd Cin → generic path
    (2, 64, 12, 20, 96, (3, 3), 2, (1, 1), 'none'),       # vector path, stride 2, Cout not multiple of 64
    (1, 96, 9, 11, 128, (1, 1), 2, (0, 0), 'none'),       # 1x1 downsample
    (1, 384, 8, 12, 256, (1, 5), 1, (0, 2), 'sigmoid'),   # GRU horizontal
    (1, 384, 8, 12, 128, (5, 1), 1, (2, 0), 'tanh'),      # GRU vertical
    (1, 256, 8, 12, 4, (3, 3), 1, (1, 1), 'none'),        # Bezier head: tiny Cout
    (3, 20, 8, 12, 128, (7, 7), 1, (3, 3), 'relu'),       # convf1 at degree 10
    (4, 128, 40, 60, 64, (3, 3), 1, (1, 1), 'relu'),      # big enough for the 128-row tile
    (1, 576, 6, 10, 124, (1, 1), 1, (0, 0), 'relu'),      # convc1-like, Cout 124
])
def test_conv2d_matches_torch(N, Cin, H, W, Cout, k, s, p, act):
    x = torch.randn(N, Cin, H, W, generator=g(1))
    w = torch.randn(Cout, Cin, *k, generator=g(2)) / (Cin * k[0] * k[1]) ** 0.5
    b = torch.randn(Cout, generator=g(3))
    ref = F.conv2d(x, w, b, stride=s, padding=p)
    ref = {'none': lambda t: t, 'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh}[act](ref)
    out = ops.conv2d(x.to(DEV), w.to(DEV), b.to(DEV), stride=s, padding=p, act=act).cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max() < 2e-5


@pytest.mark.parametrize('N,Cin,H,W,Cout,k,s,p,act,bn', [
    (1, 64, 8, 16, 64, (1, 1), 1, (0, 0), 'none', 64),
    (1, 576, 10, 13, 256, (1, 1), 1, (0, 0), 'relu', 128),     # ring wraps, ragged M, 2 n-tiles
    (1, 64, 8, 16, 256, (1, 1), 1, (0, 0), 'none', 256),
    (2, 64, 12, 20, 96, (3, 3), 2, (1, 1), 'none', 128),
    (1, 96, 16, 24, 128, (3, 3), 1, (1, 1), 'relu', 64),       # k-blocks straddle taps
    (1, 384, 8, 12, 256, (1, 5), 1, (0, 2), 'sigmoid', 64),
    (1, 384, 8, 12, 128, (5, 1), 1, (2, 0), 'tanh', 128),
    (4, 128, 40, 60, 64, (3, 3), 1, (1, 1), 'relu', 64),
])
def test_conv2d_tc3_small_shapes_match_torch(N, Cin, H, W, Cout, k, s, p, act, bn):
    """tcgen05 path with split-fp16 operands (hi*hi + hi*lo + lo*hi): ~22 mantissa bits per operand, fp32 accumulate."""
    x = torch.randn(N, Cin, H, W, generator=g(1))
    w = torch.randn(Cout, Cin, *k, generator=g(2)) / (Cin * k[0] * k[1]) ** 0.5
    b = torch.randn(Cout, generator=g(3))
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p).float()
    ref = {'none': lambda t: t, 'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh}[act](ref)
    out = ops.conv2d(x.to(DEV), w.to(DEV), b.to(DEV), stride=s, padding=p, act=act, backend='tc3', bn=bn).cpu()
    assert (out - ref).abs().max() < 1e-4


@pytest.mark.parametrize('backend,N,Cin,H,W,Cout,k,s,p,act,bn', [
    ('tc3', 1, 576, 60, 80, 256, (1, 1), 1, (0, 0), 'relu', 128),        # convc1 shape, tensor-map store epilogue
    ('tc3', 1, 256, 60, 80, 124, (3, 3), 1, (1, 1), 'relu', 64),         # ragged Cout
    ('tc3', 2, 64, 96, 100, 96, (3, 3), 2, (1, 1), 'none', 128),         # multi-tile, stride 2
    ('tc3', 1, 256, 60, 80, 256, (1, 5), 1, (0, 2), 'sigmoid', 128),     # GRU-shaped, bulk-copy epilogue
    ('tc3', 1, 64, 8, 16, 256, (1, 1), 1, (0, 0), 'none', 256),          # bn 256 (three-MMA form in split mode)
    ('slab64', 2, 64, 40, 48, 64, (3, 3), 1, (1, 1), 'relu', 64),
    ('stem7', 2, 5, 64, 96, 64, (7, 7), 2, (3, 3), 'relu', 64),
])
def test_conv2d_single_mma_f16_precision(backend, N, Cin, H, W, Cout, k, s, p, act, bn):
    """BFLOW_PREC_F16 (row g): ONE fp16 MMA per k-step on the hi planes, fp32 accumulate.  Checked two ways: against fp64 on the
    fp16-ROUNDED operands it is as exact as the split mode (proves that the lo planes / the lo half of the weight image are really
    ignored and nothing else changed), and against the unrounded fp32 operands it is within the fp16 rounding of the inputs."""
    x = torch.randn(N, Cin, H, W, generator=g(1))
    w = torch.randn(Cout, Cin, *k, generator=g(2)) / (Cin * k[0] * k[1]) ** 0.5
    b = torch.randn(Cout, generator=g(3))
    fn = {'none': lambda t: t, 'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh}[act]
    kexp = int(torch.floor(-torch.log2(w.abs().max())))                 # the packer's power-of-two prescale (exact)
    w16 = (w * 2.0 ** kexp).half().double() * 2.0 ** (-kexp)
    ref16 = fn(F.conv2d(x.half().double(), w16, b.double(), stride=s, padding=p).float())
    ref32 = fn(F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p).float())
    ops.conv2d.precision = 'f16'
    try:
        out = ops.conv2d(x.to(DEV), w.to(DEV), b.to(DEV), stride=s, padding=p, act=act, backend=backend, bn=bn).cpu()
    finally:
        ops.conv2d.precision = 'f32x3'
    assert (out - ref16).abs().max() < 1e-4
    assert (out - ref32).abs().max() < 2e-2
    assert (out - ref32).abs().max() > 1e-5          # ... and it is NOT the split path


@pytest.mark.parametrize('N,Cin,H,W,Cout,k,s,p,act,bn', [
    (1, 64, 160, 128, 256, (1, 1), 1, (0, 0), 'none', 128),    # 320 tiles: multi-tile CTAs, bulk row stores from the staging tile
    (1, 64, 150, 129, 200, (1, 1), 1, (0, 0), 'relu', 128),    # the same with ragged M and Cout
    (2, 64, 96, 100, 64, (3, 3), 1, (1, 1), 'relu', 64),       # multi-tile, direct stores
    (1, 256, 60, 80, 124, (3, 3), 1, (1, 1), 'relu', 64),      # single-tile, ragged Cout
    (1, 256, 60, 80, 256, (1, 5), 1, (0, 2), 'sigmoid', 128),  # single-tile bulk-copy epilogue, bn 128
    (1, 128, 60, 80, 64, (3, 3), 1, (1, 1), 'tanh', 64),
    (2, 96, 120, 160, 96, (3, 3), 1, (1, 1), 'relu', 96),      # narrow weight tile (bn = Cout = 96 on the 128-column kernel) + half k-blocks (96 = 64 + 32 channels)
    (2, 64, 120, 160, 96, (3, 3), 2, (1, 1), 'none', 96),      # the same, stride 2, multi-tile bulk row stores
    (1, 96, 60, 80, 96, (1, 1), 1, (0, 0), 'relu', 96),        # single-tile CTAs
    (1, 28, 64, 96, 64, (7, 7), 2, (3, 3), 'relu', 64),        # 28 channels: every k-block is a half block (two of four k-steps issued)
    (1, 160, 60, 80, 112, (3, 3), 1, (1, 1), 'none', 112),     # 160 = 2 x 64 + 32 channels, 112-row weight tiles
])
def test_conv2d_tc3_matches_torch(N, Cin, H, W, Cout, k, s, p, act, bn):
    """TMA-fed kernel and its epilogue variants (single-tile bulk copy, multi-tile bulk row stores, direct stores, ragged Cout)."""
    x = torch.randn(N, Cin, H, W, generator=g(1))
    w = torch.randn(Cout, Cin, *k, generator=g(2)) / (Cin * k[0] * k[1]) ** 0.5
    b = torch.randn(Cout, generator=g(3))
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p).float()
    ref = {'none': lambda t: t, 'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh}[act](ref)
    out = ops.conv2d(x.to(DEV), w.to(DEV), b.to(DEV), stride=s, padding=p, act=act, backend='tc3', bn=bn).cpu()
    assert (out - ref).abs().max() < 1e-4


@pytest.mark.parametrize('N,Cin,c0,H,W,Cout,k,p,act,bn', [
    (1, 256, 128, 60, 80, 128, (1, 5), (0, 2), 'tanh', 64),    # GRU candidate shape: cat(r*h, motion features)
    (1, 224, 128, 60, 80, 256, (5, 1), (2, 0), 'sigmoid', 128),    # second source 96 = 64 + 32 channels: its last k-block is a half block
    (2, 84, 64, 50, 44, 96, (3, 3), (1, 1), 'relu', 96),       # second source 20 channels (config M's Bezier parameters): half block, narrow tile, ragged M
])
def test_conv2d_tc3_two_sources_match_torch(N, Cin, c0, H, W, Cout, k, p, act, bn):
    """Channel concatenation of two sources as two tensor-map pairs (update.py:35-45, 94), each padded to whole 64-channel blocks on its own."""
    x = torch.randn(N, Cin, H, W, generator=g(1))
    w = torch.randn(Cout, Cin, *k, generator=g(2)) / (Cin * k[0] * k[1]) ** 0.5
    b = torch.randn(Cout, generator=g(3))
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=1, padding=p).float()
    ref = {'none': lambda t: t, 'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh}[act](ref)
    ops.conv2d.split_c0 = c0
    try:
        out = ops.conv2d(x.to(DEV), w.to(DEV), b.to(DEV), stride=1, padding=p, act=act, backend='tc3', bn=bn).cpu()
    finally:
        ops.conv2d.split_c0 = None
    assert (out - ref).abs().max() < 1e-4


@pytest.mark.parametrize('N,Cin,H,W,Cout,k,p,act,bn', [
    (1, 64, 16, 8, 64, (3, 3), (1, 1), 'none', 64),            # one tile
    (1, 256, 60, 80, 192, (3, 3), (1, 1), 'relu', 64),         # convc2 shape: 40 x 3 single-tile CTAs, ragged tile rows (60 = 3.75 x 16)
    (1, 256, 60, 80, 256, (1, 5), (0, 2), 'sigmoid', 128),     # GRU horizontal pass: transposed slabs
    (1, 256, 60, 80, 128, (5, 1), (2, 0), 'tanh', 64),         # GRU vertical pass
    (2, 96, 50, 44, 96, (3, 3), (1, 1), 'relu', 128),          # channels beyond C zero-filled, ragged both ways, 2 images
    (5, 128, 60, 80, 128, (3, 3), (1, 1), 'none', 128),        # multi-tile CTAs (200 tiles)
    (1, 128, 13, 21, 124, (3, 3), (1, 1), 'relu', 64),         # ragged Cout
])
def test_conv2d_tc3_slab_mode_matches_torch(N, Cin, H, W, Cout, k, p, act, bn):
    """Slab mode of the TMA-fed kernel (8 x 16 pixel tiles, halo slabs, both orientations) through every epilogue it can meet."""
    x = torch.randn(N, Cin, H, W, generator=g(1))
    w = torch.randn(Cout, Cin, *k, generator=g(2)) / (Cin * k[0] * k[1]) ** 0.5
    b = torch.randn(Cout, generator=g(3))
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=1, padding=p).float()
    ref = {'none': lambda t: t, 'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh}[act](ref)
    out = ops.conv2d(x.to(DEV), w.to(DEV), b.to(DEV), stride=1, padding=p, act=act, backend='tc3s', bn=bn).cpu()
    assert (out - ref).abs().max() < 1e-4


@pytest.mark.parametrize('N,Cin,H,W,Cout,k,p,act,bn', [
    (1, 576, 60, 80, 256, (1, 1), (0, 0), 'relu', 128),        # convc1 shape
    (1, 256, 60, 80, 192, (3, 3), (1, 1), 'relu', 64),         # convc2 shape
    (1, 256, 60, 80, 120, (3, 3), (1, 1), 'relu', 64),         # ragged Cout (multiple of 8): the tensor map clips the last box
    (1, 256, 60, 80, 124, (3, 3), (1, 1), 'relu', 64),         # Cout % 8 == 4: maps cover 120 channels, the kernel stores the 4-channel tail
    (1, 128, 37, 50, 256, (3, 3), (1, 1), 'none', 128),        # ragged M: rows beyond M clipped
    (1, 64, 9, 13, 40, (1, 1), (0, 0), 'relu', 64),            # one partial tile
])
def test_conv2d_tc3_tensor_map_store_epilogue(N, Cin, H, W, Cout, k, p, act, bn):
    """Single-tile launches whose outputs (fp32 and both split-fp16 planes) leave through cp.async.bulk.tensor stores of swizzled boxes."""
    x = torch.randn(N, Cin, H, W, generator=g(1))
    w = torch.randn(Cout, Cin, *k, generator=g(2)) / (Cin * k[0] * k[1]) ** 0.5
    b = torch.randn(Cout, generator=g(3))
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=1, padding=p).float()
    ref = torch.relu(ref) if act == 'relu' else ref
    ops.conv2d.tma_out = True
    try:
        out = ops.conv2d(x.to(DEV), w.to(DEV), b.to(DEV), stride=1, padding=p, act=act, backend='tc3', bn=bn).cpu()
        out16 = ops.conv2d.last_y16.cpu()
    finally:
        ops.conv2d.tma_out = False
    assert (out - ref).abs().max() < 1e-4
    assert (out16 - ref).abs().max() < 1e-4
    assert float(ops.conv2d.last_y16_pad.abs().sum()) == 0.0, 'the store epilogue wrote behind the last output channel'


@pytest.mark.parametrize('N,H,W,act', [(1, 16, 8, 'none'), (2, 40, 24, 'relu'), (3, 37, 16, 'none'), (5, 120, 160, 'relu')])
def test_conv2d_slab64_matches_torch(N, H, W, act):
    """Slab kernel (weights resident, halo slabs): 3x3/1, 64 -> 64; ragged tile rows (H % 16 != 0), several images, multi-tile CTAs."""
    x = torch.randn(N, 64, H, W, generator=g(1))
    w = torch.randn(64, 64, 3, 3, generator=g(2)) / 24.0
    b = torch.randn(64, generator=g(3))
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=1, padding=1).float()
    ref = torch.relu(ref) if act == 'relu' else ref
    out = ops.conv2d(x.to(DEV), w.to(DEV), b.to(DEV), stride=1, padding=(1, 1), act=act, backend='slab64').cpu()
    assert (out - ref).abs().max() < 1e-4


@pytest.mark.parametrize('N,Cin,H,W,affine', [(1, 5, 32, 16, None), (2, 5, 96, 160, None), (2, 3, 64, 48, (2.0 / 255.0, -1.0)), (3, 4, 40, 24, None)])
def test_conv2d_stem7_matches_torch(N, Cin, H, W, affine):
    """Fused stem (TMA footprint -> patch matrix in shared memory -> tcgen05): 7x7/2, thin input; the affine input map must leave the padding at zero."""
    x = torch.randn(N, Cin, H, W, generator=g(1)) if affine is None else torch.randint(0, 256, (N, Cin, H, W), generator=g(1)).float()
    w = torch.randn(64, Cin, 7, 7, generator=g(2)) / (49 * Cin) ** 0.5
    b = torch.randn(64, generator=g(3))
    xin = x if affine is None else x * affine[0] + affine[1]
    ref = torch.relu(F.conv2d(xin.double(), w.double(), b.double(), stride=2, padding=3).float())
    ops.conv2d.stem_affine = affine or (1.0, 0.0)
    try:
        out = ops.conv2d(x.to(DEV), w.to(DEV), b.to(DEV), stride=2, padding=(3, 3), act='relu', backend='stem7').cpu()
    finally:
        ops.conv2d.stem_affine = (1.0, 0.0)
    assert (out - ref).abs().max() < 1e-4


def test_instance_norm_relu_variants():
    x = torch.randn(3, 96, 17, 23, generator=g(1)) * 3 + 1
    r = torch.randn(3, 96, 17, 23, generator=g(2))
    inorm = lambda t: F.instance_norm(t, eps=1e-5)
    for res, rn, want in ((None, False, torch.relu(inorm(x))),
                          (r, False, torch.relu(torch.relu(inorm(x)) + r)),
                          (r, True, torch.relu(torch.relu(inorm(x)) + inorm(r)))):
        out = ops.instance_norm_relu(x.to(DEV), None if res is None else res.to(DEV), rn).cpu()
        assert (out - want).abs().max() < 1e-5


def test_layout_round_trip_and_window():
    x = torch.randn(2, 9, 16, 24, generator=g(1))
    y = ops.nchw_to_nhwc(x.to(DEV), c_off=2, c_cnt=5)
    assert torch.equal(y.cpu(), x[:, 2:7].permute(0, 2, 3, 1))
    z = ops.nchw_to_nhwc(x.to(DEV), scale=2.0 / 255.0, shift=-1.0)
    assert torch.allclose(z.cpu(), (2 * (x / 255) - 1).permute(0, 2, 3, 1), atol=1e-6)
    assert torch.equal(ops.nhwc_to_nchw(ops.nchw_to_nhwc(x.to(DEV))).cpu(), x)


def _pyramid_slots(vol, levels):
    """GPU pyramid in engine form from a level-0 volume (T, BQ, 1, h, w)."""
    T = vol.shape[0]
    lv = [vol[:, :, 0].contiguous()]
    idx = [list(range(T))]
    for lvl in range(1, max(levels)):
        keep = [t for t in range(T) if levels[t] > lvl]
        prev = torch.stack([lv[-1][idx[-1].index(t)] for t in keep], 0)
        lv.append(ops.corr_pool(prev))
        idx.append(keep)
    return [(l, t, lv[l][idx[l].index(t)].contiguous()) for (l, t) in O.slot_table(levels)], lv


def test_corr_volume_pool_lookup_match_reference_fixture():
    gd = load_golden('lookup_16x24')
    B, h, w, D = int(gd['B']), int(gd['h']), int(gd['w']), int(gd['D'])
    levels = [int(v) for v in gd['levels']]
    f1, f2, coords = synthetic.lookup_case(B, h, w, dim=D, targets=len(levels), seed=7)
    vol = ops.corr_volume(f1.to(DEV), f2.to(DEV))
    assert vol.shape == (4, B * h * w, 1, h, w)
    assert np.abs(vol.cpu().reshape(-1).numpy()[gd['lvl0_index']] - gd['lvl0_samples']).max() < 2e-5
    slots, lv = _pyramid_slots(vol, levels)
    for lvl in (1, 2, 3):
        assert np.abs(lv[lvl].cpu().numpy() - gd[f'lvl{lvl}']).max() < 2e-5
    out = ops.corr_lookup(slots, coords.to(DEV))
    assert np.abs(out.cpu().numpy() - gd['out']).max() < 3e-5
    out2 = ops.corr_lookup(slots, coords.to(DEV), nhwc=True)
    assert torch.equal(out2.permute(0, 3, 1, 2), out)


@pytest.mark.parametrize('B,h,w,levels', [(1, 60, 80, [4]), (3, 24, 40, [1, 1, 1, 1, 4, 4]), (2, 9, 13, [2, 1])])
def test_lookup_matches_oracle(B, h, w, levels):
    T = len(levels)
    f1, f2, coords = synthetic.lookup_case(B, h, w, dim=32, targets=T, seed=11)
    coords[0, 0, :, 0, 0] = torch.tensor([-30.0, 1e9])            # far outside: all taps zero-padded
    coords[0, 0, :, 0, 1] = torch.tensor([float(w - 1), float(h - 1)])   # exact integer corner
    vol = O.corr_volume(f1, f2)
    pyr = O.corr_pyramid(vol, levels)
    want = O.corr_lookup(pyr, coords)
    slots = [(l, t, pyr[l][1][pyr[l][0].index(t)].contiguous().to(DEV)) for (l, t) in O.slot_table(levels)]
    out = ops.corr_lookup(slots, coords.to(DEV)).cpu()
    assert out.shape == want.shape
    assert (out - want).abs().max() < 3e-5
    assert out[0, :81, 0, 0].abs().max() == 0


@pytest.mark.parametrize('B,h,w,levels', [(1, 60, 80, [4]), (2, 24, 40, [1, 1, 1, 1, 4, 4]), (2, 9, 13, [2, 1]), (1, 15, 22, [3])])
def test_tiled_lookup_and_pool_match_oracle(B, h, w, levels):
    """Granule-tiled volume layout (4x4-pixel tiles): pooling and lookup against the row-major oracle."""
    T = len(levels)
    f1, f2, coords = synthetic.lookup_case(B, h, w, dim=32, targets=T, seed=13)
    coords[0, 0, :, 0, 0] = torch.tensor([-30.0, 1e9])
    coords[0, 0, :, 0, 1] = torch.tensor([float(w - 1), float(h - 1)])
    coords[0, 0, :, 0, 2] = torch.tensor([-3.25, -2.5])                   # window straddles the top-left corner
    vol = O.corr_volume(f1, f2)
    pyr = O.corr_pyramid(vol, levels)
    want = O.corr_lookup(pyr, coords)
    # tiled pyramid built on the GPU from the tiled level 0
    lv = {0: ops.to_tiled(vol.to(DEV))}
    dims = {0: (h, w)}
    idx = {0: list(range(T))}
    for lvl in range(1, max(levels)):
        keep = [t for t in range(T) if levels[t] > lvl]
        prev = torch.stack([lv[lvl - 1][idx[lvl - 1].index(t)] for t in keep], 0)
        lv[lvl] = ops.corr_pool_tiled(prev, *dims[lvl - 1])
        dims[lvl] = (dims[lvl - 1][0] // 2, dims[lvl - 1][1] // 2)
        idx[lvl] = keep
        got = ops.from_tiled(lv[lvl], *dims[lvl]).cpu()
        assert (got - pyr[lvl][1]).abs().max() < 1e-6
        assert torch.equal(ops.to_tiled(got.to(DEV)), lv[lvl])          # padding of the tiled planes is exactly zero
    slots = [(l, t, lv[l][idx[l].index(t)].contiguous(), dims[l][0], dims[l][1]) for (l, t) in O.slot_table(levels)]
    out = ops.corr_lookup(slots, coords.to(DEV), nhwc=True, tiled=True).cpu().permute(0, 3, 1, 2)
    assert out.shape == want.shape
    assert (out - want).abs().max() < 3e-5
    out2 = ops.corr_lookup(slots, coords.to(DEV), nhwc=False, tiled=True).cpu()      # reference layout out of tiled planes
    assert (out2 - want).abs().max() < 3e-5


def test_on_the_fly_lookup_matches_reference_fixture_and_oracle():
    """Row (f3): bflow_corr_lookup_otf (no materialised volume, pooled target-feature pyramid) gives the reference's lookup output."""
    gd = load_golden('lookup_16x24')
    B, h, w, D = int(gd['B']), int(gd['h']), int(gd['w']), int(gd['D'])
    levels = [int(v) for v in gd['levels']]
    f1, f2, coords = synthetic.lookup_case(B, h, w, dim=D, targets=len(levels), seed=7)
    out = ops.corr_lookup_otf(f1.to(DEV), f2.to(DEV), levels, coords.to(DEV))
    assert np.abs(out.cpu().numpy() - gd['out']).max() < 3e-5
    # feature pooling == correlation pooling (avg_pool2d is linear), including the floor rule on odd sizes
    x = torch.randn(2, 15, 22, 64, generator=g(3))
    want = F.avg_pool2d(x.permute(0, 3, 1, 2), 2, stride=2).permute(0, 2, 3, 1)
    assert (ops.feat_pool(x.to(DEV)).cpu() - want).abs().max() < 1e-6
    for (B, h, w, levels, D) in [(1, 60, 80, [4], 256), (2, 24, 40, [1, 1, 1, 1, 4, 4], 128), (2, 9, 13, [2, 1], 32), (1, 15, 22, [3], 384)]:
        T = len(levels)
        f1, f2, coords = synthetic.lookup_case(B, h, w, dim=D, targets=T, seed=17)
        coords[0, 0, :, 0, 0] = torch.tensor([-30.0, 1e9])            # far outside: all taps zero-padded
        coords[0, 0, :, 0, 1] = torch.tensor([float(w - 1), float(h - 1)])
        coords[0, 0, :, 0, 2] = torch.tensor([-3.25, -2.5])           # window straddles the top-left corner
        want = O.corr_lookup(O.corr_pyramid(O.corr_volume(f1, f2), levels), coords)
        got = ops.corr_lookup_otf(f1.to(DEV), f2.to(DEV), levels, coords.to(DEV)).cpu()
        assert got.shape == want.shape
        assert (got - want).abs().max() < 1e-4 * max(1.0, float(want.abs().max()))
        assert got[0, :81, 0, 0].abs().max() == 0


def test_lookup_known_answer_centre_tap():
    h, w = 12, 20
    vol = torch.randn(1, h * w, 1, h, w, generator=g(5))
    co = O.coords_grid(1, h, w)[None]
    out = ops.corr_lookup([(0, 0, vol[0, :, 0].contiguous().to(DEV))], co.to(DEV)).cpu()
    assert torch.equal(out[0, 40].reshape(-1), vol[0, :, 0].reshape(h * w, h * w).diagonal())


@pytest.mark.parametrize('deg', [1, 2, 10])
def test_bezier_and_upsample_match_reference_fixture(deg):
    gd = load_golden('bezier')
    p = torch.from_numpy(gd[f'deg{deg}_params']).to(DEV)
    b = BezierCurves(p)
    ts = [float(t) for t in gd['ts']]
    assert np.abs(b.get_flow_from_reference(ts).cpu().numpy() - gd[f'deg{deg}_flows']).max() < 1e-6
    assert np.abs(b.get_flow_from_reference(0.5).cpu().numpy() - gd[f'deg{deg}_scalar_half']).max() < 1e-6
    assert np.abs(b.get_flow_from_reference(1.0).cpu().numpy() - gd[f'deg{deg}_scalar_one']).max() == 0
    up = b.create_upsampled(torch.from_numpy(gd[f'deg{deg}_mask']).to(DEV))
    assert np.abs(up.get_params().cpu().numpy() - gd[f'deg{deg}_up']).max() < 1e-5


def test_upsample_of_constant_is_eight_times_constant():
    up = ops.cvx_upsample(torch.full((2, 4, 7, 9), 0.25, device=DEV), torch.randn(2, 576, 7, 9, device=DEV))
    assert torch.allclose(up[:, :, 8:-8, 8:-8], torch.full((2, 4, 40, 56), 2.0, device=DEV), atol=1e-6)


def test_contract_errors_are_assertions():
    with pytest.raises(AssertionError):
        ops.conv2d(torch.zeros(1, 3, 8, 8), torch.zeros(4, 3, 3, 3, device=DEV))       # CPU tensor
    with pytest.raises(AssertionError):
        ops.corr_lookup([(0, 0, torch.zeros(5, 4, 4, device=DEV))], torch.zeros(1, 1, 2, 4, 4, device=DEV))
