mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench20.json 2> gpurun_out/bench20.err
