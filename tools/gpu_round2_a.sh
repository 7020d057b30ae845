#!/bin/bash
# first GPU pass of round 2: parity suite, bench line, in-graph timeline + per-CTA phases of one update-block iteration
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.txt
tail -5 gpurun_out/r2a_pytest.txt
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.err
timeout 300 python tools/timeline.py --raw --cta-iter 6 > gpurun_out/r2a_timeline.txt 2>&1; echo "timeline rc=$?"
timeout 300 python tools/timeline.py --precision f16 --raw > gpurun_out/r2a_timeline_f16.txt 2>&1; echo "timeline f16 rc=$?"
head -c 3000 gpurun_out/r2a_bench.json
