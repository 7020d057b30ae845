#!/bin/bash
# the driver's multi-GPU launch form at N = 2: both arms
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_bench_2gpu_ref.json 2> gpurun_out/r02_bench_2gpu_ref.err; echo "ref rc=$?"
python -c "
import json
for f in ('r02_bench_2gpu','r02_bench_2gpu_ref'):
    lines=[l for l in open('gpurun_out/'+f+'.json') if l.startswith('{')]
    d=json.loads(lines[-1]); print(f, d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config'])
"
