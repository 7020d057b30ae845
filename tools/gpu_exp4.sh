mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
timeout 120 python tools/timeline.py --raw > gpurun_out/timeline4.txt 2>&1
BFLOW_TC3_STAGED=2 timeout 120 python tools/timeline.py --raw > gpurun_out/timeline4_staged_all.txt 2>&1
timeout 200 python bench.py --no-sweep > gpurun_out/bench4.json 2> gpurun_out/bench4.err
