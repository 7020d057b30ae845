mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
timeout 120 python tools/timeline.py --raw > gpurun_out/timeline3.txt 2>&1
timeout 200 python bench.py --no-sweep > gpurun_out/bench3.json 2> gpurun_out/bench3.err
BFLOW_TC3_STAGED=0 timeout 200 python bench.py --no-cpu-baseline --no-sweep > gpurun_out/bench3_nostaged.json 2> gpurun_out/bench3_nostaged.err
