#!/usr/bin/env python
"""Benchmark of the RAFT-spline inference hot path (BASELINE.json: frames/s at 640x480, 12 iterations,
E_LU4_BD2, batch 1 per GPU; correlation-lookup HBM GB/s as the roofline kernel).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--preset E_LU4_BD2|E_I_LU5_BD10] [--batch-per-gpu B] [--precision f32x3|f16]

One process per GPU (torchrun for N > 1).  A step = one forward(voxel_grid, iters=12, test_mode=True) over one batch of
synthetic events of the preset's dataset shape.  Rank 0 prints ONE JSON line.

Default arm (bflow_b200): `value` = device-resident frames/s (CUDA events, max over ranks); `e2e` = the same through
RAFTSpline.forward(non_blocking=True) with pinned HOST buffers, H2D and D2H inside the timed region; `roofline` = the lookup kernel's
in-graph duration against the measured HBM peak; `tensor_roofline` = the tensor-core convolutions against the measured bf16 peak;
at N = 1 also: `cpu_baseline` (the reference itself on the host cores), `pytorch_gpu` (the reference itself, eager, on the same
B200), `parity`, `reduced_precision` (precision='f16') and the other BASELINE configs that fit one GPU (`configs`).
--impl reference: the UNMODIFIED reference model (oracle/_ref, staged by oracle/build_ref.py) on the host CPU cores, same
metric / unit / config; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPES = {'E_LU4_BD2': (480, 640), 'E_I_LU5_BD10': (384, 512)}       # dataset frame sizes
ITERS = 12
METRIC = 'frames/sec at 640x480x12-iter RAFT-spline'
SEED0 = 1234                                                           # rank r draws its shard with seed SEED0 + r


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), float(p.get('bf16_tflops_sustained', p.get('bf16_tflops', 1400.0))), 'measured (MEASURED_PEAKS.json hbm_gbs / bf16_tflops_sustained)'
    except Exception:
        return 6650.0, 1400.0, 'fallback (B200_PROFILING.md: 6.65 TB/s, 1.4 PFLOP/s sustained)'


def workload_config(preset: str, H: int, W: int, Bp: int, world: int) -> dict:
    """The workload, described identically by both arms (what the driver compares)."""
    kind = 'DSEC' if preset == 'E_LU4_BD2' else 'MultiFlow'
    extra = ' + 2 boundary images' if preset == 'E_I_LU5_BD10' else ''
    return {'workload': f'{preset} {W}x{H} synthetic {kind} events (sparse_norm voxel grid){extra}, {ITERS} iters, batch {Bp} per GPU, '
                        f'random-init weights seed 0, inputs seed {SEED0}+rank',
            'global_batch': Bp * world, 'parallelism': f'batch-sharded x{world}'}


class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(',')]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            hi = [v for v in sm if v >= 0.5 * max(sm)]        # "under load" = samples in the upper half of the observed range
            out = {'sm_mhz': statistics.median(hi), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}
        return out


def lookup_traffic(batch: int):
    """DRAM bytes (read + write) of ONE in-step lookup launch from the committed ncu --set full capture (profiles/lookup_traffic.json,
    written from the .ncu-rep by tools/ncu_summary.py); None when the capture is for another batch size."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'lookup_traffic.json')) as f:
            t = json.load(f)
        return (float(t['dram_bytes_per_launch']), t.get('source', 'profiles/lookup_traffic.json')) if int(t.get('batch', 1)) == batch else (None, None)
    except Exception:
        return None, None


def lookup_bytes(B: int, h: int, w: int, slots: int, targets: int) -> int:
    """Algorithmic bytes of one lookup launch: 400 B read + 324 B written per (pixel, slot),
    + 8 B of centre coordinates per (pixel, target)."""
    return B * h * w * (slots * 724 + targets * 8)


def flow_epe(a: torch.Tensor, b: torch.Tensor, deg: int):
    d = (a - b)
    fl = d.reshape(d.shape[0], 2, deg, *d.shape[2:])[:, :, -1]
    epe = torch.sqrt((fl ** 2).sum(1))
    return float(epe.max()), float(epe.mean()), float(d.abs().max())


# ----------------------------------------------------------------------------------------------------------------------
# the reference itself (oracle/_ref or /root/reference) -- CPU baseline, --impl reference, and eager on the GPU
# ----------------------------------------------------------------------------------------------------------------------
def reference_model(cfg, sd):
    """(callable forward(vg, im) -> (low, up) parameter tensors, kind).  The unmodified reference when importable, else the oracle port."""
    from oracle import ref_loader
    if ref_loader.available():
        ref = ref_loader.build(cfg)
        ref.load_state_dict(sd, strict=True)

        def fwd(vg, im, model=ref):
            low, up = model(voxel_grid=vg, images=im, iters=ITERS, test_mode=True)
            return low.get_params(), up.get_params()
        return ref, fwd, 'reference'
    from oracle import raft_spline_oracle as O
    return None, (lambda vg, im: O.forward(sd, cfg, vg, im, iters=ITERS, test_mode=True)), 'port'


def cpu_baseline(cfg, sd, vg, im, H, W, budget_s: float = 25.0):
    """The reference's CPU path (torch CPU fp32, all host threads) timed on a bounded sample of the same workload."""
    torch.set_num_threads(os.cpu_count())
    _, fwd, kind = reference_model(cfg, sd)
    with torch.inference_mode():
        t0 = time.perf_counter()
        fwd(vg, im)                                        # warm-up (oneDNN primitive creation, numba JIT)
        warm = time.perf_counter() - t0
        times = []
        up = None
        while len(times) < 3 and (sum(times) + warm) < budget_s:
            t0 = time.perf_counter()
            low, up = fwd(vg, im)
            times.append(time.perf_counter() - t0)
        if up is None:
            low, up = fwd(vg, im)
    t = statistics.median(times) if times else warm
    frames = vg.shape[0] if vg is not None else im[0].shape[0]
    what = 'the unmodified reference (oracle/_ref: models/raft_spline/raft.py)' if kind == 'reference' else 'oracle/raft_spline_oracle.py (port)'
    return {'value': frames / t, 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'kind': kind,
            'sample': f'{len(times) or 1} full forward passes of the same workload ({frames}x{H}x{W}, {ITERS} iters) after 1 warm-up; '
                      f'{what} on torch CPU fp32, {os.cpu_count()} host cpus'}, up


def pytorch_gpu(cfg, sd, vg_dev, im_dev, ours_up, deg, steps: int = 10):
    """The comparison SURVEY.md §2.1 calls the bar: the reference model itself in PyTorch eager fp32 on the same B200."""
    ref, fwd, kind = reference_model(cfg, sd)
    if ref is None:
        return {'unavailable': 'reference not importable here (oracle/_ref not staged)'}
    dev = vg_dev.device if vg_dev is not None else im_dev[0].device
    ref = ref.to(dev)
    out = {'impl': 'reference RAFTSpline.forward, PyTorch eager fp32 (cuDNN / cuBLAS / grid_sample) on the same GPU', 'torch': torch.__version__}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.inference_mode():
                for _ in range(3):
                    low, up = fwd(vg_dev, im_dev, ref)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    low, up = fwd(vg_dev, im_dev, ref)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            frames = up.shape[0]
            mx, mean, _ = flow_epe(ours_up, up, deg)
            out['tf32' if tf32 else 'fp32'] = {'frames_s': frames / (ms * 1e-3), 'ms_per_step': ms, 'allow_tf32': tf32,
                                               'max_epe_px_bflow_b200_vs_this': mx, 'mean_epe_px': mean}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return out


def run_reference(args, rank, world, preset, H, W):
    """--impl reference: the unmodified reference on the host CPU cores (rank 0 only; the global batch of the N-GPU workload)."""
    if rank != 0:
        return
    from bflow_b200 import RAFTSpline, config, synthetic
    cfg = config.preset(preset)
    sd = {k: v.clone() for k, v in RAFTSpline(cfg, seed=0).state_dict().items()}
    shards = [synthetic.inputs(cfg, args.batch_per_gpu, H, W, seed=SEED0 + r) for r in range(args.gpus)]
    vg = torch.cat([s[0] for s in shards]) if shards[0][0] is not None else None
    im = [torch.cat([s[1][i] for s in shards]) for i in range(2)] if shards[0][1] is not None else None
    torch.set_num_threads(os.cpu_count())
    _, fwd, kind = reference_model(cfg, sd)
    steps, warm = max(1, min(args.steps, 5 if args.gpus <= 2 else 2)), max(1, min(args.warmup, 1))
    with torch.inference_mode():
        for _ in range(warm):
            fwd(vg, im)
        t0 = time.perf_counter()
        for _ in range(steps):
            fwd(vg, im)
        dt = (time.perf_counter() - t0) / steps
    frames = args.batch_per_gpu * args.gpus
    v = frames / dt
    what = 'unmodified reference (oracle/_ref)' if kind == 'reference' else 'oracle port (reference not staged)'
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': warm,
            'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(preset, H, W, args.batch_per_gpu, args.gpus),
            'implementation': f'{what}, torch CPU fp32 (oneDNN), {torch.get_num_threads()} threads, one process; each step = the global batch of {frames} frame(s)',
            'cpu_baseline': {'value': v, 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'kind': kind,
                             'sample': f'{steps} forward passes of {frames} frame(s) (bounded from --steps {args.steps}) after {warm} warm-up'},
            'e2e': {'value': v, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# timed loops of the bflow_b200 arm
# ----------------------------------------------------------------------------------------------------------------------
def time_device(net, vg_dev, im_dev, K, Wm, barrier):
    with torch.inference_mode():
        for _ in range(Wm):
            low, up = net(voxel_grid=vg_dev, images=im_dev, iters=ITERS, test_mode=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            low, up = net(voxel_grid=vg_dev, images=im_dev, iters=ITERS, test_mode=True)
        e1.record()
        barrier()
    return e0.elapsed_time(e1), low, up


def time_e2e(net, vg_host, im_host, K, Wm, barrier):
    """Public API with HOST buffers: every step copies its inputs from pinned host memory and brings both results back to the host.
    The calls are software-pipelined (submit step i, collect step i-1), which is how a data-loader loop uses forward(non_blocking=True)."""
    def loop(n):
        prev = None
        for _ in range(n):
            cur = net(voxel_grid=vg_host, images=im_host, iters=ITERS, test_mode=True, non_blocking=True)
            if prev is not None:
                prev[0].get_params(); prev[1].get_params()        # waits for the D2H of the previous step
            prev = cur
        low_h, up_h = prev[0].get_params(), prev[1].get_params()
        return low_h, up_h
    with torch.inference_mode():
        loop(max(2, Wm))
        barrier()
        t0 = time.perf_counter()
        low_h, up_h = loop(K)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
    return dt, low_h, up_h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='bflow_b200', choices=['bflow_b200', 'reference'])
    ap.add_argument('--preset', default='E_LU4_BD2', choices=sorted(SHAPES))
    ap.add_argument('--batch-per-gpu', type=int, default=1)
    ap.add_argument('--precision', default='f32x3', choices=['f32x3', 'f16'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-sweep', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip pytorch_gpu / reduced_precision / the other BASELINE configs')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != 'reference' else args.warmup
    preset = args.preset
    H, W = SHAPES[preset]

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world, preset, H, W)
        return

    import torch.distributed as dist
    from bflow_b200 import RAFTSpline, config, synthetic, profiling, dist as bdist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: bflow_b200 has no CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = config.preset(preset)
    deg = cfg['bezier_degree']
    net = RAFTSpline(cfg, seed=0, precision=args.precision).to(dev)
    Bp = args.batch_per_gpu
    # every rank owns its shard of the global batch (weak scaling: Bp samples per GPU), seeded per rank
    vg_host, im_host = synthetic.inputs(cfg, Bp, H, W, seed=SEED0 + rank, pinned=True)
    vg_dev = vg_host.to(dev) if vg_host is not None else None
    im_dev = [t.to(dev) for t in im_host] if im_host is not None else None
    K, Wm = args.steps, args.warmup

    sampler = ClockSampler(local) if rank == 0 else None
    ms, low, up = time_device(net, vg_dev, im_dev, K, Wm, barrier)               # ---- device-resident throughput ("value")
    e2e_s, low_h, up_h = time_e2e(net, vg_host, im_host, K, Wm, barrier)         # ---- end to end with host buffers
    # the sampler (nvidia-smi at 100 ms) ran across both timed regions: at ~4 ms per step the first one alone is shorter than one period
    clocks = sampler.stop() if sampler else None
    e2e_dev_diff = float((up_h - up.get_params().cpu()).abs().max())
    t = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    frames = Bp * world * K
    h2d = sum(x.numel() * 4 for x in ([vg_host] if vg_host is not None else []) + list(im_host or []))
    d2h = (up_h.numel() + low_h.numel()) * 4

    # ---- per-rank EPE state gathered over NCCL (vs a zero-flow target): the one collective of the path ----
    flow = up.get_flow_from_reference(1.0)
    s, n = bdist.epe_sum_count(flow, torch.zeros_like(flow))
    epe_mean, epe_n, _ = bdist.gather_epe(s, n)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- in-graph schedule of one step: lookup roofline + tensor-core roofline + kernel time shares ----
    eng = net.engine(dev)
    plan = eng.plan(Bp, H, W, ITERS, True)
    with torch.inference_mode():
        plan.load_inputs(vg_dev, im_dev, None)
        rows, graph_ms = profiling.graph_timeline(plan)
        # the roofline kernel once more as a burst: the step's 12 lookup launches back to back inside ONE event pair, same coordinates
        # (L2-warm after the first) -- a footnote, not the headline
        lk_launch = [(fn, a) for fn, a in plan.launches if fn.__name__ == 'bflow_corr_lookup']
        stream = torch.cuda.current_stream().cuda_stream
        burst = []
        for rep in range(5):
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(400000)
            b0.record()
            for fn, a in lk_launch:
                fn(*a, stream)
            b1.record()
            torch.cuda.synchronize()
            burst.append(b0.elapsed_time(b1) / max(1, len(lk_launch)))
    hbm_peak, tc_peak, peak_src = peaks()
    per = {}
    for r in rows:
        per.setdefault(r['label'].split(' ')[0], []).append(r['end_us'] - r['start_us'])
    busy = sum(sum(v) for v in per.values())
    shares = {k: round(sum(v) / busy, 4) for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1]))}
    lk = per.get('corr_lookup', [0.0])
    lk_us = sum(lk) / len(lk)
    S, T = len(eng.slots), len(eng.levels)
    lk_bytes = lookup_bytes(Bp, H // 8, W // 8, S, T)
    achieved = lk_bytes / (lk_us * 1e-6) / 1e9
    traffic, traffic_src = lookup_traffic(Bp)
    # tensor-core convolutions: algorithmic FLOPs of the launches on tcgen05 (x3 executed in the split mode) over the step time
    tc_rows = [r for r in rows if r['label'].startswith(('conv_tc3', 'conv_slab64', 'conv_stem7', 'corr_volume_tc3'))]
    tc_flops = sum(r['flops'] for r in tc_rows)
    tc_busy_us = sum(r['end_us'] - r['start_us'] for r in tc_rows)
    mma_per_product = 3 if eng.prec == 0 else 1
    step_s = ms / K * 1e-3
    tensor_roofline = {'kernel': 'conv_tc3_kernel / conv_slab64_kernel / conv_stem7_kernel (all tcgen05 convolutions + the all-pairs correlation GEMM)',
                       'bound': 'tensor', 'unit': 'TFLOP/s', 'peak': tc_peak,
                       'achieved': tc_flops * mma_per_product / step_s / 1e12, 'frac': tc_flops * mma_per_product / step_s / 1e12 / tc_peak,
                       'algorithmic_tflops': tc_flops / step_s / 1e12, 'algorithmic_frac': tc_flops / step_s / 1e12 / tc_peak,
                       'algorithmic_gflop_per_step': tc_flops / 1e9, 'mma_per_product': mma_per_product,
                       'share_of_step_time_in_these_kernels': tc_busy_us / (graph_ms * 1e3),
                       'note': 'achieved = FLOPs the tensor pipe executes per step (algorithmic FLOPs of every tcgen05 launch x MMAs per product) / '
                               'whole step time (so launch gaps, epilogues and the non-tensor kernels count against it); peak = measured sustained '
                               'cuBLAS bf16 rate; algorithmic_* counts each product once'}

    line = {
        'metric': METRIC, 'value': frames / (ms * 1e-3), 'unit': 'frames/s', 'n_gpus': world, 'steps': K, 'warmup': Wm,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': ('f16x2-split (fp32-equivalent, 3 MMAs per product)' if eng.prec == 0 else 'f16 (single MMA per product, fp32 accumulate)') if eng.use_tc else 'f32',
        'data': 'synthetic',
        'config': workload_config(preset, H, W, Bp, world),
        'implementation': {
            'l2': 'no explicit flush: one step streams a 369 MB correlation volume and ~0.5 GB of encoder activations (> 126 MB L2)',
            'cuda_graph': eng.use_graph, 'graph_branches': eng.use_side_stream, 'precision': eng.precision,
            'arithmetic': ('split-fp16 operands (x = hi + lo; hi*hi + hi*lo + lo*hi) on tcgen05 with fp32 TMEM accumulation; activations stored as '
                           'fp16 hi/lo planes, state/volume/outputs fp32' if eng.prec == 0 else
                           'fp16 operands (hi planes only), one tcgen05 MMA per product, fp32 TMEM accumulation; state/volume/outputs fp32') +
                          f'; {plan.n_tc} of {sum(1 for f, _ in plan.launches if "conv2d" in f.__name__)} convolution launches on tensor cores '
                          '(convf1 and the 4-channel Bezier head on fp32 CUDA cores)'},
        'e2e': {'value': frames / (e2e_ms * 1e-3), 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_ms / K, 'max_abs_diff_vs_device_resident_result': e2e_dev_diff,
                'api': 'RAFTSpline.forward(voxel_grid=<pinned host tensor>, non_blocking=True) -> BezierCurves on pinned host memory; consecutive '
                       'calls are pipelined (H2D of step i+1 and D2H of step i-1 overlap the graph of step i on separate streams, two buffer sets)'},
        'gpu_launches': plan.n_launches * K,
        'roofline': {'kernel': 'corr_lookup_tiled_kernel (bflow_corr_lookup, granule-tiled volume)', 'bound': 'hbm', 'achieved': achieved, 'peak': hbm_peak,
                     'unit': 'GB/s', 'frac': achieved / hbm_peak, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                     'bytes_per_launch': lk_bytes, 'us_per_launch': lk_us, 'launches_timed': len(lk),
                     'us_per_launch_burst_of_12_same_coords': statistics.median(burst) * 1e3,
                     'note': 'achieved = algorithmic bytes / mean IN-GRAPH duration (first CTA start to last CTA end, %globaltimer stamps baked into the '
                             'captured launches: bflow_b200/profiling.py) of the step\'s lookup launches, i.e. with the convf1 branch running beside '
                             'it as in production. At batch 1 a launch moves 24.5 MB and is latency-bound; see lookup_sweep for the bandwidth '
                             'regime (L2 flushed, batch up to 32)'},
        'tensor_roofline': tensor_roofline,
        'kernel_time_shares': shares,
        'graph_replay_ms_instrumented': graph_ms,
        'clocks': clocks,
        'epe_allgather': {'mean_flow_px': epe_mean, 'pixels': epe_n, 'backend': 'nccl' if world > 1 else 'none'},
    }
    if not args.no_sweep and preset == 'E_LU4_BD2':
        line['lookup_sweep'] = lookup_sweep(dev, hbm_peak, H, W)

    # ---------------- N = 1 only: CPU baseline, parity, the reference on the GPU, reduced precision, other BASELINE configs ----------------
    if world == 1:
        sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
        up_dev = up.get_params()
        if not args.no_cpu_baseline:
            cpu, up_cpu = cpu_baseline(cfg, sd, vg_host.clone() if vg_host is not None else None, [t.clone() for t in im_host] if im_host else None, H, W)
            mx, mean, mabs = flow_epe(up_dev.cpu(), up_cpu, deg)
            line['cpu_baseline'] = cpu
            line['parity'] = {'max_epe_px': mx, 'mean_epe_px': mean, 'max_abs_ctrl': mabs, 'bar_px': 1e-3 if eng.prec == 0 else 1e-2,
                              'against': ('the unmodified reference' if cpu['kind'] == 'reference' else 'the oracle port') +
                                         ' (CPU fp32) on the same inputs, final upsampled flow, every pixel'}
        if not args.no_extras:
            try:
                line['pytorch_gpu'] = pytorch_gpu(cfg, sd, vg_dev, im_dev, up_dev, deg)
            except Exception as e:       # reported, never fatal for the bench line
                line['pytorch_gpu'] = {'error': repr(e)[:300]}
            if eng.prec == 0 and eng.use_tc:
                line['reduced_precision'] = reduced_precision(cfg, sd, vg_dev, vg_host, im_dev, im_host, up_dev, deg, K, Wm, barrier, dev)
            if preset == 'E_LU4_BD2' and Bp == 1:
                line['configs'] = other_configs(dev, barrier, min(K, 10), net)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def reduced_precision(cfg, sd, vg_dev, vg_host, im_dev, im_host, up_full, deg, K, Wm, barrier, dev):
    """Row (g): the same workload with precision='f16' (one fp16 MMA per product), reported beside the fp32-equivalent headline."""
    from bflow_b200 import RAFTSpline
    net = RAFTSpline(cfg, seed=0, precision='f16').to(dev)
    ms, low, up = time_device(net, vg_dev, im_dev, K, Wm, barrier)
    e2e_s, _, _ = time_e2e(net, vg_host, im_host, K, Wm, barrier)
    mx, mean, _ = flow_epe(up.get_params(), up_full, deg)
    frames = (vg_dev.shape[0] if vg_dev is not None else im_dev[0].shape[0]) * K
    out = {'precision': 'f16', 'value': frames / (ms * 1e-3), 'unit': 'frames/s', 'ms_per_step': ms / K, 'e2e': frames / e2e_s,
           'parity': {'max_epe_px_vs_fp32_equivalent_mode': mx, 'mean_epe_px': mean, 'bar_px': 1e-2},
           'arithmetic': 'one tcgen05 kind::f16 MMA per product on the hi planes; lo planes neither read nor written on the hot paths'}
    del net
    torch.cuda.empty_cache()
    return out


def other_configs(dev, barrier, K, net_d):
    """The BASELINE.json configs that fit one GPU besides the headline: #3 (E_I_LU5_BD10, 384x512, batch 4) and the per-GPU share of #4
    (E_LU4_BD2 at batch 4 per GPU, global 32 on 8 GPUs).  Device-resident frames/s, bounded to K steps each."""
    from bflow_b200 import RAFTSpline, config, synthetic
    out = {}
    try:
        cfg = config.preset('E_LU4_BD2')
        vg, _ = synthetic.inputs(cfg, 4, 480, 640, seed=SEED0)
        ms, _, _ = time_device(net_d, vg.to(dev), None, K, 3, barrier)
        out['config4_share_E_LU4_BD2_batch4_per_gpu'] = {'value': 4 * K / (ms * 1e-3), 'unit': 'frames/s', 'ms_per_step': ms / K, 'batch': 4, 'steps': K}
        net_d.engine(dev)._plans.clear()
        torch.cuda.empty_cache()
        cfg = config.preset('E_I_LU5_BD10')
        net = RAFTSpline(cfg, seed=0).to(dev)
        vg, im = synthetic.inputs(cfg, 4, 384, 512, seed=SEED0)
        ms, _, _ = time_device(net, vg.to(dev), [t.to(dev) for t in im], K, 3, barrier)
        out['config3_E_I_LU5_BD10_384x512_batch4'] = {'value': 4 * K / (ms * 1e-3), 'unit': 'frames/s', 'ms_per_step': ms / K, 'batch': 4, 'steps': K}
        del net
        torch.cuda.empty_cache()
    except Exception as e:
        out['error'] = repr(e)[:300]
    return out


def lookup_sweep(dev, peak, H, W):
    """BASELINE config #5: 640x480 input -> 80x60 feature map, radius 4, one target with a 4-level pyramid;
    algorithmic GB/s of the lookup kernel against batch size (CUDA events, L2 flushed between launches)."""
    import ctypes
    from bflow_b200 import ops, _lib
    lib = _lib.lib()
    h, w = H // 8, W // 8
    out = []
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for B in (1, 4, 8, 16, 32):
        g = torch.Generator(device='cpu').manual_seed(7)
        R = B * h * w
        lv = [ops.to_tiled(torch.randn(R, h >> l, w >> l, device=dev)) for l in range(4)]       # granule-tiled planes (production layout)
        slots = [(l, 0, lv[l], h >> l, w >> l) for l in range(4)]
        ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
        coords = (torch.stack([xs, ys], 0).float()[None, None] + 8 * torch.randn(1, B, 2, h, w, generator=g)).to(dev)
        res = torch.empty(R, 4 * 81, device=dev)
        d = ops.make_lookup_desc(slots, 1, B, h, w, True)
        d.coords, d.params, d.params_ld, d.degree = coords.data_ptr(), None, 0, 0
        d.out, d.out_nhwc, d.out_ld = res.data_ptr(), 1, 4 * 81
        for _ in range(3):
            _lib.check(lib.bflow_corr_lookup(ctypes.byref(d), stream), 'lookup')
        ts = []
        for _ in range(10):
            flush.zero_()                       # evicts L2 and keeps the GPU busy while the next three calls are enqueued
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            lib.bflow_corr_lookup(ctypes.byref(d), stream)
            a1.record()
            torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1))
        t = statistics.median(ts)
        nbytes = lookup_bytes(B, h, w, 4, 1)
        out.append({'batch': B, 'us': t * 1e3, 'gbs': nbytes / (t * 1e-3) / 1e9, 'frac': nbytes / (t * 1e-3) / 1e9 / peak})
        del lv, slots, coords, res
    return out


if __name__ == '__main__':
    main()
