#!/bin/bash
set -x
mkdir -p gpurun_out
T=r02f
O=gpurun_out
OLD=$PWD/ab_libs/2988303.so
for i in 1 2; do
  for which in new old; do
    if [ $which = old ]; then export CODDVEC_LIB=$OLD; else unset CODDVEC_LIB; fi
    PROBE_BATCHES=1,2,4,16,64,128,256,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$which /" | tee -a $O/${T}_ab_1m.log
    PROBE_ROWS=100000 PROBE_BATCHES=4,16,256 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$which /" | tee -a $O/${T}_ab_100k.log
  done
done
unset CODDVEC_LIB
for ew in 4 8; do
  CODDVEC_K3_EW=$ew PROBE_BATCHES=4,16,64,256 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/ew=$ew /" | tee -a $O/${T}_ew.log
done
timeout 900 python -m pytest tests/test_striped_gpu.py tests/test_mma_gpu.py tests/test_client.py tests/test_store_scenarios.py -m gpu -x -q 2>&1 | tail -8 | tee $O/${T}_pytest.log
timeout 900 python bench.py --rows 12500000 --steps 10 --warmup 3 > $O/${T}_bench_12m.json 2> $O/${T}_bench_12m.err; echo "bench rc=$?"
tail -c 1500 $O/${T}_bench_12m.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench_12m.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_check','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])
for n,l in d['legs'].items(): print(n, {k:l[k] for k in ('value','ms_per_step','parity_check','launches_per_step')}, l['roofline'], l['e2e']['value'])
for c in d.get('configs',[]): print(c.get('workload'), c.get('ms_per_batch'), c.get('path'), c.get('frac_of_roofline'))
print(d.get('cpu_baseline'))
PY
