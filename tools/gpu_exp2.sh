mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
timeout 120 python tools/timeline.py --raw > gpurun_out/timeline.txt 2>&1
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_flat.json 2> gpurun_out/bench_flat.err
BFLOW_LK_FLAT=0 timeout 200 python bench.py --no-cpu-baseline --no-sweep > gpurun_out/bench_noflat.json 2> gpurun_out/bench_noflat.err
