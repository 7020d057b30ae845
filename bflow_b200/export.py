"""Exports the launch plan of one (batch, height, width, iterations) as a PLAN FILE for the native entry points
``bflow_forward_load / bflow_forward_run / bflow_forward_destroy`` (include/bflow_b200.h; csrc/forward.cu).

    python -m bflow_b200.export --preset E_LU4_BD2 --batch 1 --height 480 --width 640 --iters 12 --out d_480x640.plan [--weights ckpt.pt]

The planner is the one ``RAFTSpline.forward`` uses (engine.py / engine_s16.py): this module only makes it allocate every device buffer
from the library's fixed-address arena (a torch pluggable allocator over ``bflow_arena_alloc``), then writes the arena's non-zero chunks
(packed weights, biases) and the recorded launch list — entry-point name plus typed arguments, descriptors and tensor maps as byte
blobs — to the file.  All pointers are absolute: the loader maps fresh memory at the same virtual address.
"""
from __future__ import annotations

import argparse
import ctypes as C
import struct
from typing import Optional

import torch

from . import _lib, config as _cfg
from ._lib import PREC

ARENA_BASE = 0x600000000000          # 96 TiB: far from the heap and from the mmap / CUDA regions of a 47-bit address space
CHUNK = 64 * 1024


def _blob(obj) -> bytes:
    return bytes(memoryview(obj).cast('B')) if not isinstance(obj, C.Structure) else bytes(obj)


def export_plan(model_params: dict, state_dict: Optional[dict], batch: int, height: int, width: int, iters: int, path: str, precision: str = 'f32x3',
                correlation: str = 'volume', device: str = 'cuda:0', reserve_bytes: int = 32 << 30, seed: int = 0) -> dict:
    """Writes the plan of forward(voxel_grid[, images], iters, test_mode=True) for the given shape; returns a summary dict."""
    from .raft import RAFTSpline
    from .engine import Engine
    lib = _lib.lib()
    dev = torch.device(device)
    torch.cuda.set_device(dev)
    net = RAFTSpline(model_params, seed=seed if state_dict is None else None, precision=precision, correlation=correlation)
    if state_dict is not None:
        net.load_state_dict(state_dict, strict=True)
    _lib.check(lib.bflow_arena_open(ARENA_BASE, reserve_bytes), 'arena_open')
    try:
        alloc = torch.cuda.memory.CUDAPluggableAllocator(_lib._build.LIB, 'bflow_arena_alloc', 'bflow_arena_free')
        pool = torch.cuda.MemPool(alloc.allocator())
        with torch.cuda.use_mem_pool(pool, device=dev):
            eng = Engine(net, dev, precision, correlation)          # packs the weights on the host, uploads them into the arena
            plan = eng.plan(batch, height, width, iters, True)      # allocates workspace + I/O buffers, records the launch list
        torch.cuda.synchronize(dev)
        used = int(lib.bflow_arena_used())
        lo, hi = ARENA_BASE, ARENA_BASE + used

        def inside(p):
            return p is not None and lo <= int(p) < hi

        # host-side objects the recorded arguments point to (descriptors, tensor-map buffers, channel-offset arrays), by address
        host = {}
        for obj in plan.keep:
            if isinstance(obj, (C.Structure, C.Array)):
                host[C.addressof(obj)] = obj
        io = [plan.voxel_in, plan.img_in[0] if plan.img_in else None, plan.img_in[1] if plan.img_in else None, plan.init_in, plan.low, plan.ups[0]]
        for t in io:
            assert t is None or inside(t.data_ptr()), 'an I/O buffer was allocated outside the arena'

        with open(path, 'wb') as f:
            f.write(b'BFLOWPLN')
            f.write(struct.pack('<I', 1))
            f.write(struct.pack('<QQQ', ARENA_BASE, reserve_bytes, used))
            meta = [batch, plan.cin_vox if plan.use_ev else 0, height, width, plan.h, plan.w, 2 * eng.deg, iters, int(plan.use_ev), int(plan.use_img),
                    PREC[precision], 0 if correlation == 'volume' else 1, 0, 0, 0, _lib.ABI_VERSION]
            f.write(struct.pack('<16i', *meta))
            f.write(struct.pack('<6Q', *[(t.data_ptr() - lo) if t is not None else 0xFFFFFFFFFFFFFFFF for t in io]))
            f.write(struct.pack('<6Q', *[(t.numel() * 4) if t is not None else 0 for t in io]))
            # non-zero 64 KB chunks of the arena (fresh arena memory is zero-filled, torch.empty regions were never written)
            chunks = []
            step = 256 << 20
            for off in range(0, used, step):
                n = min(step, used - off)
                raw = (C.c_uint8 * n)()
                _lib.check(lib.bflow_arena_read(off, C.addressof(raw), n), 'arena_read')
                host_bytes = torch.frombuffer(raw, dtype=torch.uint8)
                nz = host_bytes.view(-1)[: n // CHUNK * CHUNK].view(-1, CHUNK).ne(0).any(dim=1).nonzero().flatten().tolist() if n >= CHUNK else []
                for c in nz:
                    chunks.append((off + c * CHUNK, bytes(raw[c * CHUNK:(c + 1) * CHUNK])))
                tail = n // CHUNK * CHUNK
                if tail < n and any(raw[tail:n]):
                    chunks.append((off + tail, bytes(raw[tail:n])))
            f.write(struct.pack('<I', len(chunks)))
            for off, data in chunks:
                f.write(struct.pack('<QI', off, len(data)))
                f.write(data)
            # the schedule
            f.write(struct.pack('<I', len(plan.schedule)))
            n_launch = 0
            for item in plan.schedule:
                if item[0] == 'fork':
                    f.write(struct.pack('<B', 1))
                    continue
                if item[0] == 'join':
                    f.write(struct.pack('<B', 2))
                    continue
                fn, args = plan.launches[item[1]]
                n_launch += 1
                name = fn.__name__.encode()
                f.write(struct.pack('<BBH', 0, item[2] if eng.use_side_stream else 0, len(name)))
                f.write(name)
                types = fn.argtypes[:-1]                     # the trailing argument is the stream
                assert len(types) == len(args), (fn.__name__, len(types), len(args))
                f.write(struct.pack('<H', len(args)))
                for ty, a in zip(types, args):
                    if ty in (C.c_int, C.c_longlong, C.c_ulonglong, C.c_long):
                        f.write(struct.pack('<Bq', 0, int(a)))
                    elif ty is C.c_float:
                        f.write(struct.pack('<Bd', 1, float(a)))
                    else:                                    # pointer-typed
                        obj = getattr(a, '_obj', None)       # ctypes.byref(struct)
                        if obj is None and isinstance(a, (C.Structure, C.Array)):
                            obj = a
                        if obj is None and isinstance(a, int) and a in host:
                            obj = host[a]
                        if obj is not None:
                            data = _blob(obj)
                            f.write(struct.pack('<BI', 3, len(data)))
                            f.write(data)
                        else:
                            v = 0 if a is None else int(a)
                            assert v == 0 or inside(v), f'{fn.__name__}: pointer argument {v:#x} is neither inside the arena nor a recorded host object'
                            f.write(struct.pack('<BQ', 2, v))
        return {'path': path, 'arena_used_bytes': used, 'weight_chunks': len(chunks), 'weight_bytes': sum(len(d) for _, d in chunks), 'launches': n_launch,
                'schedule_items': len(plan.schedule)}
    finally:
        try:
            del plan, eng
        except NameError:
            pass
        torch.cuda.synchronize(dev)
        torch.cuda.empty_cache()
        lib.bflow_arena_close()


def main():
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('--preset', default='E_LU4_BD2', choices=sorted(_cfg.PRESETS))
    ap.add_argument('--batch', type=int, default=1)
    ap.add_argument('--height', type=int, default=480)
    ap.add_argument('--width', type=int, default=640)
    ap.add_argument('--iters', type=int, default=12)
    ap.add_argument('--precision', default='f32x3', choices=sorted(PREC))
    ap.add_argument('--correlation', default='volume', choices=['volume', 'otf'])
    ap.add_argument('--weights', default=None, help='torch.save()d state_dict of RAFTSpline (default: seed-0 random initialisation)')
    ap.add_argument('--out', required=True)
    a = ap.parse_args()
    sd = torch.load(a.weights, map_location='cpu') if a.weights else None
    print(export_plan(_cfg.preset(a.preset), sd, a.batch, a.height, a.width, a.iters, a.out, a.precision, a.correlation))


if __name__ == '__main__':
    main()
