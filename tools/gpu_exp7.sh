mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
timeout 120 python tools/timeline.py --raw --cta-label "256->256 1x5" > gpurun_out/timeline7_zr.txt 2>&1
timeout 120 python tools/timeline.py --cta-label "576->256 1x1" > gpurun_out/timeline7_c1.txt 2>&1
timeout 120 python tools/timeline.py --cta-label "256->128 1x5" > gpurun_out/timeline7_q.txt 2>&1
timeout 200 python bench.py --no-sweep > gpurun_out/bench7.json 2> gpurun_out/bench7.err
