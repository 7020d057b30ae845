#!/bin/bash
for v in 0 1; do
  BFLOW_TC3_ACC3=$v timeout 300 python -m pytest tests/test_gpu_forward.py -x -q -k "d_128_i4_bn or d_480x640_i12-" 2>&1 | tail -1
  BFLOW_TC3_ACC3=$v timeout 300 python tools/timeline.py > gpurun_out/r2s_timeline_acc3_$v.txt 2>&1
  grep "graph replay\|update-block\|256->192\|256->124\|128->256 3x3\|576->256" gpurun_out/r2s_timeline_acc3_$v.txt
done
