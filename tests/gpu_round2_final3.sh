#!/bin/bash
# Final K3 sources (lockstep landed): GPU suite, default bench, K3 ncu captures for the traffic stamp.
mkdir -p gpurun_out
O=gpurun_out
T=r02h
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/${T}_smoke.log
for b in 64 256; do
  P="env PROBE_BATCHES=$b PROBE_PATHS=mma python tests/quick_probe.py"
  ncu --set full --clock-control none --import-source on -k regex:k_mma_topk -s 6 -c 1 -f -o $O/${T}_prof_k3_1m_b$b $P > $O/${T}_ncu_k3_b$b.log 2>&1
done
P="env PROBE_ROWS=12500000 PROBE_BATCHES=1024 PROBE_PATHS=mma python tests/quick_probe.py"
ncu --set full --clock-control none --import-source on -k regex:k_mma_topk -s 3 -c 1 -f -o $O/${T}_prof_k3_12m_b1024 $P > $O/${T}_ncu_k3_12m_b1024.log 2>&1
python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err
tail -c 300 $O/${T}_bench_default.json
