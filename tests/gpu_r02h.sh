#!/bin/bash
set -x
mkdir -p gpurun_out
T=r02h
O=gpurun_out
OLD=$PWD/ab_libs/2988303.so
timeout 300 python -m pytest tests/test_striped_gpu.py -m gpu -x -q 2>&1 | tail -12 | tee $O/${T}_pytest_striped.log
CODDVEC_K3_NO_PREFILTER=1 PROBE_BATCHES=16,64,256 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/noprefilter /" | tee -a $O/${T}_bisect.log
PROBE_BATCHES=16,64,256 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/new /" | tee -a $O/${T}_bisect.log
CODDVEC_LIB=$OLD CODDVEC_LIB_ALLOW_MISSING=1 PROBE_BATCHES=16,64,256 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/old /" | tee -a $O/${T}_bisect.log
for which in new old; do
  if [ $which = old ]; then export CODDVEC_LIB=$OLD CODDVEC_LIB_ALLOW_MISSING=1; else unset CODDVEC_LIB CODDVEC_LIB_ALLOW_MISSING; fi
  for B in 64 256; do
    timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k "regex:k_mma_topk|k_merge_keys|k_prep_queries" -s 9 -c 6 --csv --log-file $O/${T}_launches_${which}_b$B.csv env PROBE_BATCHES=$B PROBE_PATHS=mma python tests/quick_probe.py > /dev/null 2>&1
  done
done
unset CODDVEC_LIB CODDVEC_LIB_ALLOW_MISSING
for f in $O/${T}_launches_*.csv; do echo $f; grep -E "k_mma|k_merge|k_prep" $f | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-160 | head -12; done
