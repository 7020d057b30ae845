mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
timeout 120 python tools/timeline.py --raw > gpurun_out/timeline24.txt 2>&1
timeout 200 python bench.py --no-sweep > gpurun_out/bench24.json 2> gpurun_out/bench24.err
