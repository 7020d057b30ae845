mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "stem7" 2>&1 | tail -30 > gpurun_out/pytest_stem7.txt
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
timeout 120 python tools/timeline.py --raw > gpurun_out/timeline10.txt 2>&1
timeout 200 python bench.py --no-sweep > gpurun_out/bench10.json 2> gpurun_out/bench10.err
