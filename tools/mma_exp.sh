mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
timeout 120 python tools/timeline.py --raw > gpurun_out/timeline23.txt 2>&1
timeout 200 python bench.py --no-sweep > gpurun_out/bench23.json 2> gpurun_out/bench23.err
