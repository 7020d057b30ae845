mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
timeout 120 python tools/timeline.py --raw --cta-label "256->256 1x5" > gpurun_out/timeline5_zr.txt 2>&1
timeout 120 python tools/timeline.py --cta-label "256->192 3x3" > gpurun_out/timeline5_c2.txt 2>&1
timeout 120 python tools/timeline.py --cta-label "576->256 1x1" > gpurun_out/timeline5_c1.txt 2>&1
timeout 120 python tools/timeline.py --cta-label "64->64 3x3/1 M=384000" --cta-occ 1 > gpurun_out/timeline5_enc.txt 2>&1
timeout 200 python bench.py --no-sweep --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err
