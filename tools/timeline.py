"""In-graph schedule of one forward step: every instrumented launch writes {first CTA start, last CTA end} (globaltimer ns)
into its own slot, baked in at record time, so the numbers come from a CUDA-graph replay -- not from eager launches.
    python tools/timeline.py [--preset E_LU4_BD2] [--batch 1] [--iters 12] [--raw] [--precision f16]
    python tools/timeline.py --cta-iter 6        per-CTA phase stamps of EVERY tensor-core launch of update-block iteration 6"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bflow_b200 import RAFTSpline, config, synthetic, profiling  # noqa: E402

TC3 = ('bflow_conv2d_nhwc_tc3', 'bflow_conv2d_nhwc_tc3o', 'bflow_conv2d_nhwc_tc3s')
COLS = ['start', 'prolog', '1stTMA', '1stfull', 'lastMMA', 'accrdy', 'epidone', 'end', 'tmem->reg', 'staged', 'batch0', 'batch1', 'batch2', 'batch3', 'batch4', '-']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--preset', default='E_LU4_BD2')
    ap.add_argument('--batch', type=int, default=1)
    ap.add_argument('--h', type=int, default=480)
    ap.add_argument('--w', type=int, default=640)
    ap.add_argument('--iters', type=int, default=12)
    ap.add_argument('--precision', default='f32x3')
    ap.add_argument('--correlation', default='volume')
    ap.add_argument('--raw', action='store_true')
    ap.add_argument('--cta', type=int, default=-1, help='per-CTA stamps of the nth tc3 launch of the step')
    ap.add_argument('--cta-label', default='', help='... or of the launch whose label contains this text')
    ap.add_argument('--cta-occ', type=int, default=6, help='which occurrence of --cta-label')
    ap.add_argument('--cta-iter', type=int, default=-1, help='per-CTA phase medians of every tc3 launch of this update-block iteration')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    L = profiling._dev_lib()
    L.bflow_tc3_cta_trace_range.argtypes = [C.c_void_p, C.c_int, C.c_int]
    cfg = config.preset(a.preset)
    net = RAFTSpline(cfg, seed=0, precision=a.precision, correlation=a.correlation).to(dev)
    vg, im = synthetic.inputs(cfg, a.batch, a.h, a.w)
    vg = vg.to(dev) if vg is not None else None
    im = [t.to(dev) for t in im] if im is not None else None
    with torch.inference_mode():
        plan = net.engine(dev).plan(a.batch, a.h, a.w, a.iters, True)
        plan.load_inputs(vg, im, None)
        tc3_idx = [i for i, (fn, _) in enumerate(plan.launches) if fn.__name__ in TC3]       # launch index of every tc3 launch
        tc3l = [plan.labels[i][0] for i in tc3_idx]
        span = 1
        if a.cta_label:
            hits = [i for i, lab in enumerate(tc3l) if a.cta_label in lab]
            a.cta = hits[min(a.cta_occ, len(hits) - 1)]
        if a.cta_iter >= 0:
            lo = plan.iter_start + a.cta_iter * plan.iter_len
            sel = [j for j, i in enumerate(tc3_idx) if lo <= i < lo + plan.iter_len]
            a.cta, span = sel[0], len(sel)
        cta = torch.zeros(span, 148, 16, device=dev, dtype=torch.int64)
        if a.cta >= 0:
            # the eager warm-up inside graph_timeline launches the list once before the capture: skip one whole pass of tc3 launches
            L.bflow_tc3_cta_trace_range(cta.data_ptr(), a.cta + len(tc3_idx), span)
        rows, ms = profiling.graph_timeline(plan)
        L.bflow_tc3_cta_trace_range(None, -1, 1)
    n = len(rows)
    print(f'{a.preset} B={a.batch} {a.h}x{a.w} iters={a.iters} precision={a.precision}: graph replay {ms:.3f} ms; {n} instrumented launches of {plan.n_launches}; '
          f'span of instrumented launches {max(r["end_us"] for r in rows) / 1e3:.3f} ms')
    if a.raw:
        for r in rows:
            print(f'{r["start_us"]:9.2f} {r["end_us"]:9.2f}  {r["end_us"] - r["start_us"]:7.2f} us  s{r["stream"]} {r["label"]}')
    if a.cta >= 0:
        call = cta.cpu()
        for k in range(span):
            c = call[k]
            c = c[c[:, 0] > 0]
            if c.shape[0] == 0:
                continue
            c0 = int(c[:, 0].min())
            lab = tc3l[a.cta + k] if a.cta + k < len(tc3l) else '?'
            print(f'per-CTA stamps of tc3 launch #{a.cta + k} ({lab}), {c.shape[0]} CTAs, us from the first CTA start:')
            if span == 1:
                print('  cta   ' + ' '.join(f'{nme:>7s}' for nme in COLS[:8]))
                for i in range(c.shape[0]):
                    print(f'  {i:3d} ' + ' '.join(f'{(int(v) - c0) / 1e3:7.2f}' for v in c[i] if int(v) > 0))
            for j, nme in enumerate(COLS):
                if int(c[:, j].max()) == 0:
                    continue
                col = (c[:, j][c[:, j] > 0] - c0).double() / 1e3
                print(f'  {nme:9s} min {col.min():7.2f} median {col.median():7.2f} max {col.max():7.2f}')
    # per-label aggregate: in-kernel duration and the gap to the previous launch's end on the same stream
    agg, last_end = {}, {}
    for r in rows:
        s, e, lab, st = r['start_us'], r['end_us'], r['label'], r['stream']
        gap = s - last_end[st] if st in last_end else 0
        last_end[st] = e
        x = agg.setdefault((lab, st), [0, 0.0, 0.0])
        x[0] += 1; x[1] += e - s; x[2] += gap
    print(f'{"label":52s} {"n":>4s} {"in-kernel us":>13s} {"gap before us":>14s}')
    for (lab, st), (cnt, dur, gap) in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
        print(f's{st} {lab:50s} {cnt:4d} {dur / cnt:13.2f} {gap / cnt:14.2f}   total {(dur + gap) / 1e3:7.3f} ms')
    it = [r for r in rows if r['label'].startswith('corr_lookup')]
    if len(it) >= 3:
        print(f'update-block iteration (lookup start to lookup start): {(it[-1]["start_us"] - it[1]["start_us"]) / (len(it) - 2):.2f} us')


if __name__ == '__main__':
    main()
