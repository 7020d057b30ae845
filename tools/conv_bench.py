"""Single-convolution microbench / ncu target.
    python tools/conv_bench.py --n 1 --cin 384 --h 60 --w 80 --cout 128 --kh 1 --kw 5 --backend tc3 --bn 64"""
import argparse
import ctypes as C
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bflow_b200 import ops, _lib  # noqa: E402
from bflow_b200._lib import ConvDesc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    for k, v in dict(n=1, cin=384, h=60, w=80, cout=128, kh=1, kw=5, stride=1, bn=64, reps=10).items():
        ap.add_argument('--' + k, type=int, default=v)
    ap.add_argument('--backend', default='tc3', choices=['tc3', 'simt'])
    ap.add_argument('--dbg', type=int, default=0)
    ap.add_argument('--trace', action='store_true')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    lib = _lib.lib()
    ph, pw = a.kh // 2, a.kw // 2
    Ho, Wo = (a.h + 2 * ph - a.kh) // a.stride + 1, (a.w + 2 * pw - a.kw) // a.stride + 1
    x = torch.randn(a.n, a.h, a.w, a.cin, device=dev)
    w = torch.randn(a.cout, a.cin, a.kh, a.kw, device=dev) / (a.cin * a.kh * a.kw) ** 0.5
    b = torch.randn(a.cout, device=dev)
    y = torch.empty(a.n, Ho, Wo, a.cout, device=dev)
    wp, ldw = ops.pack_conv_weight(w)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    x16 = ops.split_f16(x, (a.cin + 7) // 8 * 8)
    m = ops.tma_im2col_maps(x16, a.n, a.h, a.w, a.cin, a.kh, a.kw, a.stride, ph, pw)
    maps = (C.c_uint8 * 512)()
    C.memmove(maps, m, 256)
    wtc3, acc3 = ops.pack_conv_weight_tc(w, a.bn, block_per_tap=True)
    d = ConvDesc()
    d.x0, d.c0, d.ld0 = x.data_ptr(), a.cin, a.cin
    d.x1, d.c1, d.ld1 = None, 0, 0
    d.w, d.ldw, d.bias = wp.data_ptr(), ldw, b.data_ptr()
    d.res, d.ldr = None, 0
    d.y, d.ldy = y.data_ptr(), a.cout
    d.N, d.H, d.W, d.Ho, d.Wo, d.Cout = a.n, a.h, a.w, Ho, Wo, a.cout
    d.KH, d.KW, d.stride, d.pad_h, d.pad_w = a.kh, a.kw, a.stride, ph, pw
    d.act1, d.act2, d.scale = 1, 0, 1.0
    st = torch.cuda.current_stream().cuda_stream
    C.CDLL(_lib._build.LIB).bflow_tc3_debug(a.dbg)

    def run():
        if a.backend == 'tc3':
            _lib.check(lib.bflow_conv2d_nhwc_tc3(C.byref(d), C.addressof(maps), wtc3.data_ptr(), a.bn, acc3, err.data_ptr(), st), 'tc3')
        else:
            _lib.check(lib.bflow_conv2d_nhwc(C.byref(d), st), 'simt')
    for _ in range(3):
        run()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = statistics.median(ts)
    if a.trace:
        tr = torch.zeros(6, 256, device=dev, dtype=torch.int64)
        L = C.CDLL(_lib._build.LIB)
        L.bflow_tc3_trace.argtypes = [C.c_void_p]
        L.bflow_tc3_trace(tr.data_ptr())
        run(); torch.cuda.synchronize()
        L.bflow_tc3_trace(None)
        tr = tr.cpu()
        t0 = int(tr[tr > 0].min())
        names = ['prod:empty-ok', 'mma:full-ok', 'mma:committed', 'epi:tfull-ok', 'epi:tmem-released', 'epi:tile-done']
        print('cta start', int(tr[5, 255]) - t0, 'prologue done', int(tr[3, 255]) - t0, 'cta end', int(tr[4, 255]) - t0, '(SM cycles)')
        tr[5, 255] = 0; tr[3, 255] = 0; tr[4, 255] = 0
        for nm, r in (('mma:before-wait', 3), ('mma:issued', 4)):
            print(f'{nm:18s}', [int(v) - t0 for v in tr[r, 100:140] if int(v) > 0])
        tr[3, 100:] = 0; tr[4, 100:] = 0
        for r in range(6):
            vals = [int(v) - t0 for v in tr[r] if int(v) > 0][:40]
            print(f'{names[r]:18s}', vals)
    fl = 2.0 * a.n * Ho * Wo * a.cout * a.kh * a.kw * a.cin
    print(f'{a.backend} bn={a.bn} {a.cin}->{a.cout} {a.kh}x{a.kw}/{a.stride} M={a.n * Ho * Wo}: {t * 1e3:.1f} us  {fl / (t * 1e-3) / 1e12:.1f} TF/s useful; err flag {int(err.item())} dbg {a.dbg}')


if __name__ == '__main__':
    main()
