#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r01d_pytest_gpu.log
PROBE_ROWS=1000000 PROBE_BATCHES=4,16,64,256 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | tee gpurun_out/k3_1m_probe.log
PROBE_ROWS=1000000 PROBE_BATCHES=16,256 PROBE_PATHS=mma ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_mma_topk|k_merge_keys|k_prep_queries" -c 40 --csv --log-file gpurun_out/k3_1m_launches.csv python tests/quick_probe.py > /dev/null 2>&1
python - <<'PY'
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open('gpurun_out/k3_1m_launches.csv')) if len(r)>10 and r[0].isdigit()]
d=defaultdict(list)
for r in rows: d[r[4][:30]+' grid '+r[8]].append(int(r[-1]))
for k,v in d.items(): print(k, len(v), round(sum(v)/len(v)/1e3,2),'us')
PY
