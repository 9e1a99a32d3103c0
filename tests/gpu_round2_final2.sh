#!/bin/bash
# Round-2 evidence on the round's last kernel sources (one B200): GPU suite, smoke, both bench arms, host-side probe,
# K3 ncu captures (the K2 capture of tests/gpu_round2_final.sh stays valid: its sources did not change), K3 timeline.
mkdir -p gpurun_out
O=gpurun_out
T=r02f
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/${T}_smoke.log
python bench.py --impl reference --steps 5 --warmup 3 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err
python tests/host_scaling_probe.py 2>&1 | tail -4 > $O/${T}_host_scaling.jsonl
PROBE_BATCHES=1,2,4,16,64,128,256,512,1024 python tests/quick_probe.py > $O/${T}_probe_1m.log 2>&1
PROBE_ROWS=100000 PROBE_BATCHES=1,4,16,64,256 python tests/quick_probe.py > $O/${T}_probe_100k.log 2>&1
CODDVEC_LIB=$PWD/codd_query_engine_b200/libcoddvec_timeline.so CODDVEC_K3_TRACE=1 CODDVEC_K3_TIMELINE=1 PROBE_BATCHES=16,64,256 PROBE_PATHS=mma python tests/quick_probe.py > $O/${T}_timeline_raw.log 2>&1
for b in 16 64 256; do grep -A78 "timeline\] B=$b " $O/${T}_timeline_raw.log | tail -79; done > $O/${T}_k3_timeline.log
for b in 64 256; do
  P="env PROBE_BATCHES=$b PROBE_PATHS=mma python tests/quick_probe.py"
  ncu --set full --clock-control none --import-source on -k regex:k_mma_topk -s 6 -c 1 -f -o $O/${T}_prof_k3_1m_b$b $P > $O/${T}_ncu_k3_b$b.log 2>&1
done
P="env PROBE_ROWS=12500000 PROBE_BATCHES=1024 PROBE_PATHS=mma python tests/quick_probe.py"
ncu --set full --clock-control none --import-source on -k regex:k_mma_topk -s 3 -c 1 -f -o $O/${T}_prof_k3_12m_b1024 $P > $O/${T}_ncu_k3_12m_b1024.log 2>&1
P="env PROBE_BATCHES=64 PROBE_PATHS=mma python tests/quick_probe.py"
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_mma_topk|k_prep_queries|k_merge_keys" -c 40 --csv --log-file $O/${T}_launches_k3_b64.csv $P > /dev/null 2>&1
tail -c 600 $O/${T}_bench_default.json
