"""TEST / BENCH INFRASTRUCTURE ONLY.  Imports the UNMODIFIED reference model code.

Search order: the live tree (/root/reference, build container only), then the byte-for-byte copy that
``oracle/build_ref.py`` stages under ``oracle/_ref/`` (git-ignored, shipped to the GPU box with the working tree).
The single missing import on the model path is ``omegaconf.ListConfig`` (models/raft_utils/corr.py:8, used in one
``isinstance``); a two-class shim is injected before import.  Nothing is written to the reference tree (bytecode
writing is disabled)."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
LIVE_ROOT = os.environ.get('BFLOW_REFERENCE_ROOT', '/root/reference')
STAGED_ROOT = os.path.join(HERE, '_ref')


def _has(root: str) -> bool:
    return os.path.isfile(os.path.join(root, 'models', 'raft_spline', 'raft.py'))


def root() -> str:
    """Directory the reference is imported from ('' when neither the live tree nor the staged copy exists)."""
    if _has(LIVE_ROOT):
        return LIVE_ROOT
    if _has(STAGED_ROOT):
        return STAGED_ROOT
    return ''


REF_ROOT = root()


def available() -> bool:
    return root() != ''


def live() -> bool:
    """The reference checkout itself (not the staged copy) is present: build container only."""
    return _has(LIVE_ROOT)


def load():
    """Returns the reference ``models.raft_spline.raft`` module."""
    r = root()
    if not r:
        raise RuntimeError(f'reference not found at {LIVE_ROOT} nor staged under {STAGED_ROOT} (python -m oracle.build_ref)')
    sys.dont_write_bytecode = True
    if 'omegaconf' not in sys.modules:
        shim = types.ModuleType('omegaconf')
        shim.ListConfig = type('ListConfig', (list,), {})
        shim.DictConfig = type('DictConfig', (dict,), {})
        sys.modules['omegaconf'] = shim
    if r not in sys.path:
        sys.path.insert(0, r)
    import importlib
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        mod = importlib.import_module('models.raft_spline.raft')
    # utils/timers.py:71 registers an atexit printer of its (here always empty) timer table; keep a bench's JSON line the last thing printed
    import atexit
    timers = sys.modules.get('utils.timers')
    if timers is not None and hasattr(timers, 'print_timing_info'):
        atexit.unregister(timers.print_timing_info)
    return mod


def build(cfg: dict):
    """Constructs the reference RAFTSpline(cfg).eval() with its construction prints silenced."""
    import contextlib
    import io
    mod = load()
    with contextlib.redirect_stdout(io.StringIO()):
        net = mod.RAFTSpline(cfg)
    return net.eval()
