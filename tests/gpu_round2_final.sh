#!/bin/bash
# Round-2 evidence, one B200: GPU suite, smoke, reference arm, default bench, secondary configs, then ncu:
# the launch list of the default bench command and --set full captures of K2 (default bench), K3 at B = 64 and
# B = 256 on the 1M-row corpus (cfg2) and B = 1024 on a 12.5M-row shard (cfg4's per-GPU share).  Outputs: gpurun_out/r02z_*.
mkdir -p gpurun_out
O=gpurun_out
T=r02z
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/${T}_smoke.log
python bench.py --impl reference --steps 5 --warmup 3 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err
python tests/bench_configs.py cfg1 cfg3 cfg5 > $O/${T}_configs.jsonl 2> $O/${T}_configs.err
PROBE_BATCHES=1,2,4,16,64,128,256,512,1024 python tests/quick_probe.py > $O/${T}_probe_1m.log 2>&1
PROBE_ROWS=12500000 PROBE_BATCHES=1,16,64,256,1024 python tests/quick_probe.py > $O/${T}_probe_12m.log 2>&1
CODDVEC_LIB=$PWD/codd_query_engine_b200/libcoddvec_timeline.so CODDVEC_K3_TRACE=1 CODDVEC_K3_TIMELINE=1 PROBE_BATCHES=16,64,256 PROBE_PATHS=mma python tests/quick_probe.py > $O/${T}_timeline_raw.log 2>&1
for b in 16 64 256; do grep -A78 "timeline\] B=$b " $O/${T}_timeline_raw.log | tail -79; done > $O/${T}_k3_timeline.log
CMD="python bench.py --steps 5 --warmup 3 --no-extras"
$CMD > $O/${T}_plain_default.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_scan_topk|k_mma_topk|k_merge_keys|k_prep_queries|k_exchange_merge" -c 60 --csv \
    --log-file $O/${T}_launches_default.csv $CMD > $O/${T}_ncu_default_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scan_topk -s 4 -c 1 -f -o $O/${T}_prof_k2_default $CMD > $O/${T}_ncu_k2.log 2>&1
for b in 64 256; do
  P="env PROBE_BATCHES=$b PROBE_PATHS=mma python tests/quick_probe.py"
  $P > $O/${T}_plain_k3_b$b.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_mma_topk -s 6 -c 1 -f -o $O/${T}_prof_k3_1m_b$b $P > $O/${T}_ncu_k3_b$b.log 2>&1
done
P="env PROBE_ROWS=12500000 PROBE_BATCHES=1024 PROBE_PATHS=mma python tests/quick_probe.py"
$P > $O/${T}_plain_k3_12m_b1024.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_mma_topk -s 3 -c 1 -f -o $O/${T}_prof_k3_12m_b1024 $P > $O/${T}_ncu_k3_12m_b1024.log 2>&1
P="env PROBE_BATCHES=64 PROBE_PATHS=mma python tests/quick_probe.py"
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_mma_topk|k_prep_queries|k_merge_keys" -c 40 --csv --log-file $O/${T}_launches_k3_b64.csv $P > /dev/null 2>&1
tail -c 800 $O/${T}_bench_default.json
