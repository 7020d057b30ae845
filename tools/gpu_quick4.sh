#!/bin/bash
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_kernels.py -x -q -k "lookup or forward or pool" > gpurun_out/${tag}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.txt
timeout 300 python tools/timeline.py > gpurun_out/${tag}_timeline.txt 2>&1
grep "graph replay\|update-block\|corr_lookup" gpurun_out/${tag}_timeline.txt
timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/${tag}_bench.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'],d['roofline']['us_per_launch'])
print('sweep',[(x['batch'],round(x['us'],1),round(x['frac'],3)) for x in d.get('lookup_sweep',[])])
"
