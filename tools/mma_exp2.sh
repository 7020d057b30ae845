mkdir -p gpurun_out
timeout 120 python tools/timeline.py --cta-label "256->256 1x5" > gpurun_out/cta_zr5.txt 2>&1
timeout 120 python tools/timeline.py --cta-label "256->128 1x5" > gpurun_out/cta_q5.txt 2>&1
