#!/bin/bash
# tools/ab.sh REPS dir1 dir2 ...   : graph replay time of the default workload in every tree, interleaved
reps=$1; shift
for i in $(seq $reps); do
  for d in "$@"; do
    ( cd $d && PYTHONPATH=$PWD python tools/timeline.py 2>&1 | grep "graph replay\|update-block" | sed "s|^|$d |" | cut -c1-110 )
  done
done
