"""In-tree build of libbflow_b200.so (nvcc, sm_100a only).  `python -m bflow_b200.build [--force]`."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libbflow_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-Wall', '--shared']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(HERE, '..', 'include', 'bflow_b200.h')]
    return any(os.path.getmtime(p) > t for p in deps)


def nvcc_path() -> str:
    p = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.isfile(p):
        raise RuntimeError('nvcc not found; libbflow_b200.so must be built where CUDA 12.9 is installed')
    return p


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    objs = []
    bdir = os.path.join(HERE, 'build')
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc_path()] + [f for f in NVCC_FLAGS if f != '--shared'] + ['-c', src, '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [nvcc_path(), '--shared', '-o', LIB] + objs + ['-lcudart']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
