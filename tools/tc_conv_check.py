"""Diagnostic for the tcgen05 convolution kernel: compares bflow_conv2d_nhwc_tc with the fp32 CUDA-core kernel and
torch CPU on a ladder of shapes and, on mismatch, prints where the error sits (row % 8, channel chunk, output column)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bflow_b200 import ops  # noqa: E402

DEV = 'cuda:0'
CASES = [
    # N, Cin, H, W, Cout, k, stride, pad, bn
    (1, 64, 8, 16, 64, (1, 1), 1, (0, 0), 64),      # single k-block, single tile: pure GEMM 128x64x64
    (1, 64, 8, 16, 128, (1, 1), 1, (0, 0), 128),
    (1, 64, 8, 16, 256, (1, 1), 1, (0, 0), 256),
    (1, 256, 8, 16, 64, (1, 1), 1, (0, 0), 64),     # 4 k-blocks: pipeline wraps nothing yet
    (1, 576, 10, 13, 256, (1, 1), 1, (0, 0), 128),  # 9 k-blocks: ring wraps, ragged M, 2 n-tiles
    (2, 64, 12, 20, 96, (3, 3), 2, (1, 1), 128),    # 3x3 stride 2, Cout 96
    (1, 96, 16, 24, 128, (3, 3), 1, (1, 1), 128),   # Cin 96: k-blocks straddle taps
    (1, 384, 8, 12, 256, (1, 5), 1, (0, 2), 64),    # GRU horizontal
    (1, 384, 8, 12, 128, (5, 1), 1, (2, 0), 128),   # GRU vertical
    (1, 256, 8, 12, 4, (3, 3), 1, (1, 1), 64),      # tiny Cout
    (4, 128, 40, 60, 64, (3, 3), 1, (1, 1), 64),    # 75 M-tiles
]


def main():
    backend = sys.argv[1] if len(sys.argv) > 1 else 'tc'
    torch.manual_seed(0)
    worst = 0.0
    for (N, Cin, H, W, Cout, k, s, p, bn) in CASES:
        x = torch.randn(N, Cin, H, W)
        w = torch.randn(Cout, Cin, *k) / (Cin * k[0] * k[1]) ** 0.5
        b = torch.randn(Cout)
        ref = F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p).float()
        try:
            tc = ops.conv2d(x.to(DEV), w.to(DEV), b.to(DEV), stride=s, padding=p, backend=backend, bn=bn).cpu()
        except Exception as e:   # noqa: BLE001
            print(f'case {(N, Cin, H, W, Cout, k, s, p, bn)}: EXCEPTION {e}')
            worst = float('inf')
            continue
        simt = ops.conv2d(x.to(DEV), w.to(DEV), b.to(DEV), stride=s, padding=p).cpu()
        e_tc, e_simt = (tc - ref).abs().max().item(), (simt - ref).abs().max().item()
        worst = max(worst, e_tc)
        print(f'case N={N} Cin={Cin} {H}x{W} Cout={Cout} k={k} s={s} bn={bn}: tc err {e_tc:.3e}  simt err {e_simt:.3e}  |ref| {ref.abs().max():.2f}')
        if e_tc > 1e-3:
            d = (tc - ref).abs()                      # (N, Cout, Ho, Wo)
            flat = d.permute(0, 2, 3, 1).reshape(-1, Cout)   # rows = output pixels m, cols = channels
            rows = flat.max(1).values
            print('   err by (m % 8):     ', [f'{rows[i::8].max().item():.2e}' for i in range(8)])
            print('   err by (m // 32)%4: ', [f'{rows[torch.arange(len(rows)) // 32 % 4 == i].max().item():.2e}' for i in range(4)])
            cols = flat.max(0).values
            print('   err by (n % 16):    ', [f'{cols[i::16].max().item():.2e}' for i in range(min(16, Cout))])
            print('   first rows of tc vs ref:', tc.permute(0, 2, 3, 1).reshape(-1, Cout)[0, :6].tolist(), ref.permute(0, 2, 3, 1).reshape(-1, Cout)[0, :6].tolist())
    print('WORST', worst)
    return 0 if worst < 1e-3 else 1


if __name__ == '__main__':
    sys.exit(main())
