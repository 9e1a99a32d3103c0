#!/bin/bash
# lockstep of sibling query super-tiles (three or more, long scans): DRAM bytes and time against the previous commit's
# library; parity with the lockstep forced on for every multi-tile batch
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02_lockstep_ab.log
: > $L
run() { local label=$1; shift; env "$@" PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$label /" | tee -a $L; }
PREV="CODDVEC_LIB=$PWD/ab_libs/HEAD.so CODDVEC_LIB_ALLOW_MISSING=1"
for rep in 1 2 3; do
  run prev_12m PROBE_ROWS=12500000 PROBE_BATCHES=768,1024,2048 $PREV
  run new_12m PROBE_ROWS=12500000 PROBE_BATCHES=768,1024,2048
done
for rep in 1 2; do
  run prev_1m PROBE_BATCHES=4,64,256,512,1024 $PREV
  run new_1m PROBE_BATCHES=4,64,256,512,1024
done
: > $O/r02_lockstep_traffic.log
for w in 0 4; do
  P="env CODDVEC_K3_LOCKSTEP=$w PROBE_ROWS=12500000 PROBE_BATCHES=1024 PROBE_PATHS=mma python tests/quick_probe.py"
  for s in 3 5; do
  ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:k_mma_topk -s $s -c 1 --csv $P 2>/dev/null | grep -E "dram__bytes_read|hit_rate|time_duration" | awk -F'","' -v r="window=$w" '{print r, $(NF-2), $(NF-1), $NF}' | tee -a $O/r02_lockstep_traffic.log
  done
done
CODDVEC_K3_LOCKSTEP_MIN_TILES=0 timeout 1200 python -m pytest tests/test_mma_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r02_lockstep_pytest_forced.log
CODDVEC_K3_LOCKSTEP_MIN_TILES=0 CODDVEC_LIB=$PWD/codd_query_engine_b200/libcoddvec_checked.so timeout 600 python -m pytest tests/test_mma_gpu.py -m gpu -x -q -k "parity or large_batch" 2>&1 | tail -3 | tee $O/r02_lockstep_pytest_checked.log
