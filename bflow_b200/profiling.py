"""Development / benchmark instrumentation: the in-graph schedule of one forward.

Every instrumented kernel of the library takes a "timeline slot" at LAUNCH time (``bflow_timeline``): two 64-bit words that its
CTAs update with ``atomicMin(start, %globaltimer)`` / ``atomicMax(end, %globaltimer)``.  Capturing a plan's launch list into a
CUDA graph while the timeline is armed bakes one slot into every launch, so a replay of that graph leaves the true in-graph
{first CTA start, last CTA end} of every kernel behind — with both graph branches running, which per-kernel CUDA events on an
eagerly launched step cannot show.  ``bench.py`` takes the lookup roofline and the kernel time shares from here;
``tools/timeline.py`` prints the whole schedule.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch

from . import _lib

# entry points whose kernels take a timeline slot, in the order bflow_timeline_name reports them
INSTRUMENTED = ('conv2d_nhwc_tc3', 'conv2d_nhwc_tc3o', 'conv2d_nhwc_tc3s', 'conv2d_slab64', 'conv2d_stem7', 'corr_lookup', 'corr_lookup_otf', 'conv2d_small_n',
                'conv2d_thin7', 'conv2d_nhwc', 'instnorm_relu16', 'im2col_split16')


def _dev_lib() -> C.CDLL:
    L = C.CDLL(_lib._build.LIB)
    L.bflow_timeline.argtypes = [C.c_void_p, C.c_int]
    L.bflow_timeline_name.restype = C.c_char_p
    L.bflow_tc3_cta_trace.argtypes = [C.c_void_p, C.c_int]
    return L


def graph_timeline(plan, replays: int = 3) -> Tuple[List[Dict], float]:
    """Captures `plan`'s launch list in a fresh CUDA graph with the timeline armed, replays it and returns
    ([{label, stream, start_us, end_us, flops}], graph replay time in ms).  The plan's inputs must be loaded."""
    dev = plan.eng.device
    L = _dev_lib()
    cap = 4096
    with torch.cuda.device(dev):
        buf = torch.zeros(cap, 2, device=dev, dtype=torch.int64)
        plan.launch_all()                      # eager warm-up, timeline off
        torch.cuda.synchronize()
        L.bflow_timeline(buf.data_ptr(), cap)  # armed: the capture below bakes slot i into launch i
        g = torch.cuda.CUDAGraph()
        import gc
        gc.collect()                           # see _Plan.execute: no collection of CUDA-owning garbage while the capture runs
        gc_was_enabled = gc.isenabled()
        gc.disable()
        try:
            try:
                with torch.cuda.graph(g):
                    plan.launch_all()
            finally:
                if gc_was_enabled:
                    gc.enable()
            n = L.bflow_timeline_used()
            names = [L.bflow_timeline_name(i).decode() for i in range(n)]
        finally:
            L.bflow_timeline(None, 0)
        for _ in range(replays):
            g.replay()
        torch.cuda.synchronize()
        buf[:, 0] = torch.iinfo(torch.int64).max
        buf[:, 1] = 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        t = buf[:n].cpu()
        del g
    inst = [(lab, fl, item[2]) for (fn, _), (lab, fl), item in zip(plan.launches, plan.labels, [it for it in plan.schedule if it[0] == 'launch'])
            if fn.__name__.replace('bflow_', '') in INSTRUMENTED]
    if len(inst) != n:                         # a kernel was added without registering it above: fall back to the library's own names
        inst = [(nm, 0.0, 0) for nm in names]
    t0 = int(t[:, 0].min())
    rows = [dict(label=lab, stream=st, start_us=(int(t[i, 0]) - t0) / 1e3, end_us=(int(t[i, 1]) - t0) / 1e3, flops=fl)
            for i, (lab, fl, st) in enumerate(inst)]
    return rows, e0.elapsed_time(e1)
