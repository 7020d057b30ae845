#!/bin/bash
# what the driver does at round end: smoke() on cuda:0, and the list of the first kernels it launches
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 120 --csv --log-file gpurun_out/r02_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_smoke_launches.csv')) if len(r)>5 and r[0].isdigit()]
names=[r[4] for r in rows]
first=next((i for i,n in enumerate(names) if 'bflow::' in n), None)
print('launches captured', len(names), '; non-library launches before the first bflow:: kernel:', first)
print('first 12:', [n[:50] for n in names[:12]])
PY
