"""TEST INFRASTRUCTURE ONLY.  Imports the UNMODIFIED reference model code from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  The single missing
import on the model path is ``omegaconf.ListConfig`` (models/raft_utils/corr.py:8, used in one
``isinstance``); a two-class shim is injected before import.  Nothing is written to the
reference tree (bytecode writing is disabled)."""
import os
import sys
import types

REF_ROOT = os.environ.get('BFLOW_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, 'models', 'raft_spline', 'raft.py'))


def load():
    """Returns the reference ``models.raft_spline.raft`` module."""
    if not available():
        raise RuntimeError(f'reference not found at {REF_ROOT}')
    sys.dont_write_bytecode = True
    if 'omegaconf' not in sys.modules:
        shim = types.ModuleType('omegaconf')
        shim.ListConfig = type('ListConfig', (list,), {})
        shim.DictConfig = type('DictConfig', (dict,), {})
        sys.modules['omegaconf'] = shim
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        mod = importlib.import_module('models.raft_spline.raft')
    return mod


def build(cfg: dict):
    """Constructs the reference RAFTSpline(cfg).eval() with its construction prints silenced."""
    import contextlib
    import io
    mod = load()
    with contextlib.redirect_stdout(io.StringIO()):
        net = mod.RAFTSpline(cfg)
    return net.eval()
