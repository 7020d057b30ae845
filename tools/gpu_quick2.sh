#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_kernels.py -x -q -k "lookup or forward or pool" > gpurun_out/${tag}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.txt
timeout 300 python tools/timeline.py > gpurun_out/${tag}_timeline.txt 2>&1
grep "graph replay\|update-block\|corr_lookup" gpurun_out/${tag}_timeline.txt
BFLOW_PDL=1 timeout 300 python tools/timeline.py > gpurun_out/${tag}_timeline_pdl.txt 2>&1
grep "graph replay\|update-block\|corr_lookup" gpurun_out/${tag}_timeline_pdl.txt
timeout 300 python tools/lookup_bench.py > gpurun_out/${tag}_lookup_bench.txt 2>&1; tail -8 gpurun_out/${tag}_lookup_bench.txt
