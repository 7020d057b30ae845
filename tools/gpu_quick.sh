#!/bin/bash
# quick GPU pass: forward parity + timeline (+ per-CTA phases of iteration 6)
mkdir -p gpurun_out
tag=${1:-q}
timeout 900 python -m pytest tests/test_gpu_forward.py -x -q > gpurun_out/${tag}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.txt
timeout 300 python tools/timeline.py --raw --cta-iter 6 > gpurun_out/${tag}_timeline.txt 2>&1
grep "graph replay\|update-block" gpurun_out/${tag}_timeline.txt
timeout 300 python tools/timeline.py --precision f16 > gpurun_out/${tag}_timeline_f16.txt 2>&1
grep "graph replay\|update-block" gpurun_out/${tag}_timeline_f16.txt
