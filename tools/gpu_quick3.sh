#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "on_the_fly or otf" > gpurun_out/${tag}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.txt
timeout 300 python tools/timeline.py --correlation otf > gpurun_out/${tag}_timeline_otf.txt 2>&1
grep "graph replay\|update-block\|corr_lookup" gpurun_out/${tag}_timeline_otf.txt
timeout 300 python tools/timeline.py > gpurun_out/${tag}_timeline.txt 2>&1
grep "graph replay\|update-block\|corr_lookup\|corr_volume" gpurun_out/${tag}_timeline.txt
