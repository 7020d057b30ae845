"""In-graph schedule of one forward step: every instrumented launch writes {first CTA start, last CTA end} (globaltimer ns)
into its own slot, baked in at record time, so the numbers come from a CUDA-graph replay -- not from eager launches.
    python tools/timeline.py [--preset E_LU4_BD2] [--batch 1] [--iters 12] [--raw]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bflow_b200 import RAFTSpline, config, synthetic, _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--preset', default='E_LU4_BD2')
    ap.add_argument('--batch', type=int, default=1)
    ap.add_argument('--h', type=int, default=480)
    ap.add_argument('--w', type=int, default=640)
    ap.add_argument('--iters', type=int, default=12)
    ap.add_argument('--raw', action='store_true')
    ap.add_argument('--cta', type=int, default=-1, help='per-CTA stamps of the nth tc3 launch of the step')
    ap.add_argument('--cta-label', default='', help='... or of the launch whose label contains this text')
    ap.add_argument('--cta-occ', type=int, default=6, help='which occurrence of --cta-label')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    L = C.CDLL(_lib._build.LIB)
    L.bflow_timeline.argtypes = [C.c_void_p, C.c_int]
    L.bflow_timeline_name.restype = C.c_char_p
    cap = 4096
    buf = torch.zeros(cap, 2, device=dev, dtype=torch.int64)

    def reset():
        buf[:, 0] = torch.iinfo(torch.int64).max
        buf[:, 1] = 0

    cfg = config.preset(a.preset)
    net = RAFTSpline(cfg, seed=0).to(dev)
    vg, im = synthetic.inputs(cfg, a.batch, a.h, a.w)
    vg = vg.to(dev) if vg is not None else None
    im = [t.to(dev) for t in im] if im is not None else None
    with torch.inference_mode():
        # every launch call takes the next slot; the eager warm-up and the capture call the same launch list, so reset the slot counter
        # before the capture by re-arming the timeline inside execute(): simplest is to record eagerly first with the timeline off
        os.environ['BFLOW_GRAPH'] = '1'
        plan = net.engine(dev).plan(a.batch, a.h, a.w, a.iters, True)
        plan.load_inputs(vg, im, None)
        plan.launch_all()                      # eager warm-up, timeline off
        torch.cuda.synchronize()
        cta = torch.zeros(148, 16, device=dev, dtype=torch.int64)
        if a.cta_label:
            tc3l = [lab for (fn, _), (lab, _) in zip(plan.launches, plan.labels) if fn.__name__ in ('bflow_conv2d_nhwc_tc3', 'bflow_conv2d_nhwc_tc3o', 'bflow_conv2d_nhwc_tc3s')]
            hits = [i for i, lab in enumerate(tc3l) if a.cta_label in lab]
            a.cta = hits[min(a.cta_occ, len(hits) - 1)]
        if a.cta >= 0:
            L.bflow_tc3_cta_trace.argtypes = [C.c_void_p, C.c_int]
            L.bflow_tc3_cta_trace(cta.data_ptr(), a.cta)
        L.bflow_timeline(buf.data_ptr(), cap)  # armed: the capture below bakes slot i into launch i
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            plan.launch_all()
        n = L.bflow_timeline_used()
        names = [L.bflow_timeline_name(i).decode() for i in range(n)]
        L.bflow_timeline(None, 0)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        reset()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
    t = buf[:n].cpu()
    t0 = int(t[:, 0].min())
    print(f'{a.preset} B={a.batch} {a.h}x{a.w} iters={a.iters}: graph replay {e0.elapsed_time(e1):.3f} ms; {n} instrumented launches of {plan.n_launches}; '
          f'span of instrumented launches {(int(t[:, 1].max()) - t0) / 1e6:.3f} ms')
    # label = instrumented launches in plan order (the plan's labels for those kernels)
    inst = ('conv2d_nhwc_tc3', 'conv2d_nhwc_tc3o', 'conv2d_nhwc_tc3s', 'conv2d_slab64', 'conv2d_stem7', 'corr_lookup', 'conv2d_small_n', 'conv2d_thin7', 'conv2d_nhwc', 'instnorm_relu16', 'im2col_split16')
    labels = [lab for (fn, _), (lab, _) in zip(plan.launches, plan.labels) if fn.__name__.replace('bflow_', '') in inst]
    streams = [item[2] for item in plan.schedule if item[0] == 'launch' and plan.launches[item[1]][0].__name__.replace('bflow_', '') in inst]
    if len(labels) != n:
        labels = names
        streams = [0] * n
    rows = []
    for i in range(n):
        rows.append((int(t[i, 0]) - t0, int(t[i, 1]) - t0, labels[i], streams[i]))
    if a.raw:
        for s, e, lab, st in rows:
            print(f'{s / 1e3:9.2f} {e / 1e3:9.2f}  {(e - s) / 1e3:7.2f} us  s{st} {lab}')
    if a.cta >= 0:
        tc3 = [lab for (fn, _), (lab, _) in zip(plan.launches, plan.labels) if fn.__name__ in ('bflow_conv2d_nhwc_tc3', 'bflow_conv2d_nhwc_tc3o', 'bflow_conv2d_nhwc_tc3s')]
        c = cta.cpu()
        c = c[c[:, 0] > 0]
        c0 = int(c[:, 0].min())
        print(f'per-CTA stamps of tc3 launch #{a.cta} ({tc3[a.cta] if a.cta < len(tc3) else "?"}), {c.shape[0]} CTAs, us from the first CTA start:')
        print('  cta   start prolog  1stTMA 1stfull lastMMA accrdy epidone    end')
        for i in range(c.shape[0]):
            print(f'  {i:3d} ' + ' '.join(f'{(int(v) - c0) / 1e3:7.2f}' for v in c[i] if int(v) > 0))
        cols = ['start', 'prolog', '1stTMA', '1stfull', 'lastMMA', 'accrdy', 'epidone', 'end', 'tmem->reg', 'staged', 'batch0', 'batch1', 'batch2', 'batch3', 'batch4', '-']
        for j, nme in enumerate(cols):
            if int(c[:, j].max()) == 0:
                continue
            col = (c[:, j] - c0).double() / 1e3
            print(f'  {nme:8s} min {col.min():7.2f} median {col.median():7.2f} max {col.max():7.2f}')
    # per-label aggregate: in-kernel duration and the gap to the previous launch's end on the same stream
    agg = {}
    last_end = {}
    for s, e, lab, st in rows:
        gap = s - last_end[st] if st in last_end else 0
        last_end[st] = e
        x = agg.setdefault((lab, st), [0, 0.0, 0.0])
        x[0] += 1; x[1] += (e - s) / 1e3; x[2] += gap / 1e3
    print(f'{"label":52s} {"n":>4s} {"in-kernel us":>13s} {"gap before us":>14s}')
    for (lab, st), (cnt, dur, gap) in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
        print(f's{st} {lab:50s} {cnt:4d} {dur / cnt:13.2f} {gap / cnt:14.2f}   total {(dur + gap) / 1e3:7.3f} ms')


if __name__ == '__main__':
    main()
