"""CPU: the C-ABI library builds, loads and exports every symbol the header declares; host-side logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from bflow_b200 import RAFTSpline, BezierCurves, config, _lib, dist as bdist
from bflow_b200.bezier import bernstein_coeffs
from bflow_b200.ops import pack_conv_weight


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'bflow_b200.h')).read()
    return sorted(set(re.findall(r'\b(bflow_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    declared = header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/bflow_b200.h but not exported'
    assert sorted(_lib.exported_symbols()) == declared, 'ctypes binding and header disagree'
    assert lib.bflow_abi_version() == 2
    assert lib.bflow_built_for_sm() == 100


def test_contract_violations_return_invalid_without_touching_the_gpu():
    lib = _lib.lib()
    assert lib.bflow_conv2d_nhwc(None, None) == 1
    assert b'null descriptor' in lib.bflow_last_error()
    d = _lib.LookupDesc()
    d.n_slots = 99
    assert lib.bflow_corr_lookup(C.byref(d), None) == 1
    assert b'bad slot count' in lib.bflow_last_error()
    # a descriptor built against another header (wrong struct_size) is refused before anything behind that field is read
    d2 = _lib.LookupDesc()
    d2.struct_size -= 24
    assert lib.bflow_corr_lookup(C.byref(d2), None) == 1
    assert b'descriptor size mismatch' in lib.bflow_last_error()
    c = _lib.ConvDesc()
    c.struct_size = 0
    for fn in (lib.bflow_conv2d_nhwc, lib.bflow_conv2d_small_n, lib.bflow_conv2d_thin7):
        assert fn(C.byref(c), None) == 1 and b'descriptor size mismatch' in lib.bflow_last_error()
    with pytest.raises(AssertionError):
        _lib.check(lib.bflow_corr_pool(None, None, 1, 4, 4, None), 'corr_pool')


def test_struct_layouts_match_the_header():
    # sizes computed by the C compiler for the same declarations
    import subprocess, tempfile, textwrap
    code = textwrap.dedent('''
        #include <stdio.h>
        #include "bflow_b200.h"
        int main(void) { printf("%zu %zu\\n", sizeof(bflow_conv_desc), sizeof(bflow_lookup_desc)); return 0; }
    ''')
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, 'a.c'), os.path.join(td, 'a.out')
        open(src, 'w').write(code)
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), src, '-o', exe])
        a, b = map(int, subprocess.check_output([exe]).split())
    assert a == C.sizeof(_lib.ConvDesc)
    assert b == C.sizeof(_lib.LookupDesc)
    # ... and as the LIBRARY was compiled
    lib = _lib.lib()
    assert lib.bflow_sizeof_conv_desc() == a and lib.bflow_sizeof_lookup_desc() == b
    assert _lib.ConvDesc().struct_size == a and _lib.LookupDesc().struct_size == b


def test_library_is_built_from_the_sources_in_the_tree():
    from bflow_b200 import build
    h = _lib.lib().bflow_source_hash().decode()
    assert len(h) == 64 and h == build.source_hash()


def test_integration_md_ctypes_stub_matches_the_header():
    """The binding INTEGRATION.md tells a maintainer to paste must describe the struct the library reads."""
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    blocks = re.findall(r'```python\n(.*?)```', doc, flags=re.S)
    stub = next(b for b in blocks if 'class LookupDesc' in b)
    ns = {}
    exec(compile(stub.replace("C.CDLL('bflow_b200/libbflow_b200.so')", f"C.CDLL({_lib._build.LIB!r})"), 'INTEGRATION.md', 'exec'), ns)
    lib = _lib.lib()
    assert C.sizeof(ns['LookupDesc']) == lib.bflow_sizeof_lookup_desc() == C.sizeof(_lib.LookupDesc)
    assert [f[0] for f in ns['LookupDesc']._fields_] == [f[0] for f in _lib.LookupDesc._fields_]
    for (n1, t1), (n2, t2) in zip(ns['LookupDesc']._fields_, _lib.LookupDesc._fields_):
        assert C.sizeof(t1) == C.sizeof(t2), n1
    assert callable(ns['corr_block_call'])


def test_packed_weights_follow_in_place_parameter_updates():
    """a checkpoint loaded through a PARENT module, an optimizer step or param.data.copy_ never calls
    RAFTSpline.load_state_dict; the engine must notice through the load-state-dict post hook / the tensor version counters."""
    net = RAFTSpline(config.preset('E_LU4_BD2'), seed=0)
    v0 = net._versions()
    net._engine, net._engine_versions = 'packed', v0                       # stands for a built engine (no GPU here)
    with torch.no_grad():
        net.update_block.gru.convz1.weight.mul_(1.5)                        # in-place update
    assert net._versions() != v0
    parent = torch.nn.Module()
    parent.net = net
    net._engine = 'packed'
    parent.load_state_dict({'net.' + k: v for k, v in RAFTSpline(config.preset('E_LU4_BD2'), seed=3).state_dict().items()}, strict=True)
    assert net._engine is None                                              # the post hook fired for the nested load
    net.precision = 'f16'
    assert net.precision == 'f16'
    with pytest.raises(AssertionError):
        RAFTSpline(config.preset('E_LU4_BD2'), precision='int8')


def test_reference_copy_under_oracle_ref_is_unmodified():
    from oracle import build_ref, ref_loader
    if not os.path.isdir(build_ref.DST):
        pytest.skip('oracle/_ref not staged (python -m oracle.build_ref)')
    assert build_ref.verify()
    if ref_loader.live():
        for rel in build_ref.FILES:
            assert build_ref.sha256(os.path.join(build_ref.SRC, rel)) == build_ref.sha256(os.path.join(build_ref.DST, rel))


def test_config_tables():
    d, m = config.preset('E_LU4_BD2'), config.preset('E_I_LU5_BD10')
    assert config.input_channels(d) == 9 and config.input_channels(m) == 65
    assert config.levels_per_target(d) == [1, 1, 1, 4]
    assert config.levels_per_target(m) == [1, 1, 1, 1, 4, 4]
    assert config.slot_table([1, 1, 1, 4]) == [(0, 0), (0, 1), (0, 2), (0, 3), (1, 3), (2, 3), (3, 3)]
    assert len(config.slot_table(config.levels_per_target(m))) == 12
    assert config.lookup_timestamps(d) == [0.25, 0.5, 0.75, 1.0]
    assert np.allclose(config.lookup_timestamps(m), [0.2, 0.4, 0.6, 0.8, 1.0, 1])


def test_bernstein_known_answers():
    c = bernstein_coeffs([0.25, 0.5], 2)
    assert np.allclose(c, [[0.375, 0.0625], [0.5, 0.25]])
    assert np.allclose(bernstein_coeffs([1.0], 10)[0], [0] * 9 + [1])
    assert np.allclose(bernstein_coeffs([0.0], 3)[0], 0)


def test_module_surface_and_state_dict_layout():
    from bflow_b200.raft import num_cor_planes
    d = config.preset('E_LU4_BD2')
    net = RAFTSpline(d)
    sd = net.state_dict()
    assert len(sd) == 179
    assert sd['update_block.encoder.convc1.weight'].shape == (256, 567, 1, 1)
    assert sd['update_block.gru.convz1.weight'].shape == (128, 384, 1, 5)
    assert sd['update_block.gru.convq2.weight'].shape == (128, 384, 5, 1)
    assert sd['update_block.mask.2.weight'].shape == (576, 256, 1, 1)
    # norm3 is registered twice in the reference (extractor.py:43-44): the alias must exist and share storage
    assert sd['cnet.layer2.0.downsample.1.running_var'].data_ptr() == sd['cnet.layer2.0.norm3.running_var'].data_ptr()
    assert not any(k.startswith('fnet_ev') and 'norm' in k for k in sd)       # InstanceNorm carries no state
    m = RAFTSpline(config.preset('E_I_LU5_BD10'))
    assert len(m.state_dict()) == 211 and num_cor_planes(m.model_params) == 972
    assert sum(p.numel() for p in net.parameters()) == 5344832
    # strict round trip
    net2 = RAFTSpline(d, seed=5)
    net2.load_state_dict(sd, strict=True)
    assert all(torch.equal(a, b) for a, b in zip(net2.state_dict().values(), sd.values()))
    grids, ctx = net.gen_voxel_grids(torch.zeros(1, 9, 8, 8))
    assert len(grids) == 5 and grids[2].shape[1] == 5 and ctx.shape[1] == 5


def test_forward_refuses_cpu_tensors():
    net = RAFTSpline(config.preset('E_LU4_BD2'))
    with pytest.raises(RuntimeError, match='CUDA'):
        net(voxel_grid=torch.zeros(1, 9, 128, 128), iters=1, test_mode=True)
    with pytest.raises(AssertionError):
        net(voxel_grid=torch.zeros(1, 8, 128, 128), iters=1, test_mode=True)
    with pytest.raises(AssertionError):
        net(iters=1)


def test_bezier_curves_cpu_api():
    p = torch.randn(2, 4, 3, 5)
    b = BezierCurves(p)
    assert b.degree == 2 and b.batch_size == 2 and b.height == 3 and b.width == 5 and b.dim == 4
    assert torch.equal(b.get_flow_from_reference(1.0), p[:, [1, 3]])
    assert torch.equal(b.get_flow_from_reference(0), torch.zeros(2, 2, 3, 5))
    f = b.get_flow_from_reference([0.25, 0.5])
    assert f.shape == (2, 2, 2, 3, 5)
    assert torch.allclose(f[1, :, 0], 0.5 * p[:, 0] + 0.25 * p[:, 1])
    b.delta_update_params(torch.ones_like(p))
    assert torch.equal(b.get_params(), p + 1)
    z = BezierCurves.create_from_voxel_grid(torch.zeros(1, 9, 64, 96), bezier_degree=10)
    assert z.get_params().shape == (1, 20, 8, 12)
    assert b.cpu().get_params().device.type == 'cpu' and not b.detach().requires_grad


def test_weight_images_are_packed_on_the_host():
    """Engine packing must not launch PyTorch kernels (the driver's launch capture should start with this library's kernels)."""
    from bflow_b200.ops import pack_conv_weight_tc
    w = torch.randn(96, 70, 3, 3)
    img, acc = pack_conv_weight_tc(w, 64, block_per_tap=True)
    assert img.device.type == 'cpu' and img.dtype == torch.int16
    assert img.numel() == 2 * 9 * 2 * 2 * 64 * 64                          # n-tiles x (tap, 64-ch block) x hi|lo x rows x 64
    k = -int(np.log2(acc))
    # tile 0, k-block 0 (tap 0, channels 0..63), hi plane, row r: chunk j holds source chunk j ^ (r % 8)
    hi = img.view(2, 18, 2, 64, 8, 8)[0, 0, 0].view(torch.float16)
    want = (w[:64, :64, 0, 0] * 2.0 ** k).half()
    for r in (0, 3, 63):
        for j in (0, 5):
            assert torch.equal(hi[r, j], want[r, (j ^ (r % 8)) * 8:(j ^ (r % 8)) * 8 + 8])


def test_weight_packing_layout():
    w = torch.arange(2 * 3 * 1 * 2, dtype=torch.float32).reshape(2, 3, 1, 2)   # O=2, I=3, KH=1, KW=2
    p, ldw = pack_conv_weight(w)
    assert ldw == 4 and p.shape == (6, 4)
    for kw in range(2):
        for c in range(3):
            for o in range(2):
                assert p[kw * 3 + c, o] == w[o, c, 0, kw]
    assert torch.all(p[:, 2:] == 0)
    p2, _ = pack_conv_weight(torch.ones(5, 3, 1, 1), cin_pad=16)
    assert p2.shape == (16, 8) and p2[3:].abs().sum() == 0


def test_shard_ranges_cover_the_batch():
    for gb in (1, 7, 8, 32):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_range(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_tile_width_policy():
    """engine_s16.choose_bn: 64-column tiles for narrow layers and for the 38-row-tile update block, 128 columns once the row tiles fill the machine,
    and a narrow weight tile (bn = Cout) only where bflow_conv2d_nhwc_tc3 accepts one: a single column tile, 64 < Cout < 128, Cout % 16 == 0."""
    from bflow_b200.engine_s16 import choose_bn
    assert choose_bn(64, 3000) == 64 and choose_bn(48, 10) == 64
    assert choose_bn(96, 750) == 96 and choose_bn(112, 148) == 112 and choose_bn(80, 200) == 80
    assert choose_bn(124, 150) == 128 and choose_bn(128, 188) == 128 and choose_bn(192, 150) == 128 and choose_bn(256, 750) == 128
    assert choose_bn(96, 38) == 64 and choose_bn(192, 38) == 64 and choose_bn(124, 38) == 64
    assert choose_bn(256, 38) == 128 and choose_bn(576, 38) == 64
