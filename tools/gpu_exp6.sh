mkdir -p gpurun_out
timeout 120 python tools/timeline.py --cta-label "256->256 1x5" > gpurun_out/timeline6_zr.txt 2>&1
timeout 120 python tools/timeline.py --cta-label "576->256 1x1" > gpurun_out/timeline6_c1.txt 2>&1
timeout 120 python tools/timeline.py --cta-label "256->128 1x5" > gpurun_out/timeline6_q.txt 2>&1
