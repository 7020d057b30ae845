mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
timeout 120 python tools/timeline.py --raw --cta-label "256->256 1x5" > gpurun_out/timeline8_zr.txt 2>&1
BFLOW_TC3_BULK=0 timeout 120 python tools/timeline.py --raw --cta-label "256->256 1x5" > gpurun_out/timeline8_zr_nobulk.txt 2>&1
timeout 200 python bench.py --no-sweep > gpurun_out/bench8.json 2> gpurun_out/bench8.err
BFLOW_TC3_BULK=0 timeout 200 python bench.py --no-sweep --no-cpu-baseline > gpurun_out/bench8_nobulk.json 2> gpurun_out/bench8_nobulk.err
