#!/bin/bash
# full GPU pass: whole parity suite, bench line, timeline
tag=${1:-full}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.txt
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/${tag}_bench.err
timeout 300 python tools/timeline.py --raw > gpurun_out/${tag}_timeline.txt 2>&1
grep "graph replay\|update-block" gpurun_out/${tag}_timeline.txt
python - <<'PY'
import json,sys
d=json.load(open('gpurun_out/'+sys.argv[1]+'_bench.json')) if len(sys.argv)>1 else None
PY
python -c "
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('roofline',d['roofline']['frac'],d['roofline']['us_per_launch'],'tensor',d['tensor_roofline']['frac'])
for k in ('parity','cpu_baseline','reduced_precision','configs','pytorch_gpu'):
    print(k, json.dumps(d.get(k))[:400])
print('sweep',[(x['batch'],round(x['frac'],3)) for x in d.get('lookup_sweep',[])])
"
