mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
BFLOW_TC3_SLAB=1 CUDA_VISIBLE_DEVICES=0 timeout 120 python tools/timeline.py > gpurun_out/timeline_slabmode.txt 2>&1
tail -3 gpurun_out/bench_2gpu.err
