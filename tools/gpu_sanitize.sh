#!/bin/bash
# compute-sanitizer memcheck + racecheck over the kernel-level and the small forward tests of what ships
mkdir -p gpurun_out
SEL='tc3_small or single_mma or slab64 or stem7 or tensor_map_store or lookup or events or metrics or d_128_i4 or m_128_i3 or oracle_all_modes or pipelined'
for tool in memcheck racecheck; do
  timeout 1700 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/r02_sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02_sanitizer_$tool.txt | tail -4
done
