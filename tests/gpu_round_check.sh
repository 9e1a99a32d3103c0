#!/bin/bash
# One-call verification of the whole path on a B200: GPU parity tests, smoke, default bench, reference arm,
# secondary configs, then the ncu evidence for the default bench command (launch list + one --set full
# capture of the dominant kernel) and for the predicate scan (cfg3).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
T=${TAG:-r01b}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${T}_smoke.log
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
python tests/bench_configs.py cfg1 cfg3 cfg5 > gpurun_out/${T}_configs.jsonl 2> gpurun_out/${T}_configs.err
CMD="python bench.py --steps 5 --warmup 3 --no-extras"
$CMD > gpurun_out/${T}_plain_default.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_scan_topk|k_mma_topk|k_merge_keys|k_prep_queries|k_exchange_merge" -c 50 --csv \
    --log-file gpurun_out/${T}_launches_default.csv $CMD > gpurun_out/${T}_ncu_default_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scan_topk -s 4 -c 1 -f -o gpurun_out/${T}_prof_default $CMD > gpurun_out/${T}_ncu_default_full.log 2>&1
CMD3="python tests/bench_configs.py cfg3"
ncu --set full --clock-control none --import-source on -k regex:k_scan_topk -s 20 -c 1 -f -o gpurun_out/${T}_prof_cfg3 $CMD3 > gpurun_out/${T}_ncu_cfg3.log 2>&1
tail -c 600 gpurun_out/${T}_bench_default.json
tail -n 3 gpurun_out/${T}_configs.jsonl | cut -c1-700
