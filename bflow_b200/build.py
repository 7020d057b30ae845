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


def source_hash() -> str:
    """sha256 over csrc/*.cu, csrc/*.cuh and include/bflow_b200.h (sorted, name + content): baked into the library as
    bflow_source_hash(), so that source <-> binary identity is checked by content at load time, not by mtime."""
    import hashlib
    h = hashlib.sha256()
    deps = sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + [os.path.join(HERE, '..', 'include', 'bflow_b200.h')]
    for p in deps:
        h.update(os.path.basename(p).encode())
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def built_hash() -> str:
    """Source hash the library on disk was built from, read from the side-car file build() writes next to it ('' when missing).
    (Not through dlopen: a handle to a stale library would stay cached in this process and shadow the rebuilt one.  The loaded
    library's own bflow_source_hash() is checked against the sources in _lib.lib().)"""
    try:
        with open(LIB + '.sha256') as f:
            return f.read().strip() if os.path.isfile(LIB) else ''
    except OSError:
        return ''


def _stale() -> bool:
    return built_hash() != source_hash()


def nvcc_path() -> str:
    p = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.isfile(p):
        raise RuntimeError('nvcc not found; libbflow_b200.so must be built where CUDA 12.9 is installed')
    return p


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    objs = []
    digest = source_hash()
    bdir = os.path.join(HERE, 'build')
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc_path()] + [f for f in NVCC_FLAGS if f != '--shared'] + [f'-DBFLOW_SOURCE_HASH="{digest}"', '-c', src, '-o', obj]
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
    with open(LIB + '.sha256', 'w') as f:
        f.write(digest + '\n')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
