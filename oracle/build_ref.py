"""TEST / BENCH INFRASTRUCTURE ONLY.  Stages the UNMODIFIED reference model code under ``oracle/_ref/``.

The reference (uzh-rpg/bflow) is pure Python on PyTorch: there is nothing to compile and nothing to pip-install
(no setup.py / pyproject.toml), so "building" it means making its model files importable where /root/reference does
not exist (the GPU box).  This recipe copies, byte for byte, the files on the RAFT-spline inference path

    models/raft_spline/{raft,update,bezier}.py   models/raft_utils/{corr,extractor,utils}.py   utils/timers.py
    data/utils/representations.py                (row f1: VoxelGrid / norm_voxel_grid)

from /root/reference into ``oracle/_ref/`` — a directory that is git-ignored (reference SOURCES never enter the history)
but not gpurun-ignored, so it travels to the GPU box exactly like the built ``.so``.  ``oracle/ref_loader.py`` imports the
live tree when it exists and this copy otherwise; ``bench.py --impl reference`` and the ``cpu_baseline`` / ``pytorch_gpu``
legs then time the reference ITSELF (``kind: "reference"``).  A manifest with the sha256 of every staged file is written
next to them so that a test can prove the copy is unmodified.

    python -m oracle.build_ref        (build container; a no-op with a message where /root/reference is absent)
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
SRC = os.environ.get('BFLOW_REFERENCE_ROOT', '/root/reference')
FILES = [
    'models/raft_spline/raft.py', 'models/raft_spline/update.py', 'models/raft_spline/bezier.py',
    'models/raft_utils/corr.py', 'models/raft_utils/extractor.py', 'models/raft_utils/utils.py',
    'utils/timers.py', 'data/utils/representations.py', 'LICENSE',
]


def sha256(path: str) -> str:
    with open(path, 'rb') as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose: bool = True) -> bool:
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        if verbose:
            print(f'oracle/_ref: {SRC} not present, nothing staged (using what is already in {DST})')
        return os.path.isfile(os.path.join(DST, FILES[0]))
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        manifest[rel] = sha256(dst)
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': 'uzh-rpg/bflow (unmodified copies)', 'sha256': manifest}, f, indent=1)
    if verbose:
        print(f'oracle/_ref: staged {len(FILES)} reference files from {SRC}')
    return True


def verify() -> bool:
    """True when every staged file still has the hash recorded when it was copied."""
    try:
        with open(os.path.join(DST, 'MANIFEST.json')) as f:
            man = json.load(f)['sha256']
    except OSError:
        return False
    return all(os.path.isfile(os.path.join(DST, rel)) and sha256(os.path.join(DST, rel)) == h for rel, h in man.items())


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
