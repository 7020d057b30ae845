mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
timeout 120 python tools/timeline.py --raw > gpurun_out/timeline15.txt 2>&1
BFLOW_TC3_BULK_Q=0 timeout 120 python tools/timeline.py --raw > gpurun_out/timeline15_q1.txt 2>&1
timeout 200 python bench.py --no-sweep --no-cpu-baseline > gpurun_out/bench15.json 2> gpurun_out/bench15.err
