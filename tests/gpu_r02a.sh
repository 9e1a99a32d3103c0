#!/bin/bash
# Round-2 first GPU call: same-box A/B of the two branches that never ran (K2 queries in shared memory, predicate
# code prefetch), their parity tests, ncu --set full of K3 in the mid-batch regime on cfg2, and a first
# compute-sanitizer pass over the reduced subset.  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
T=r02a
O=gpurun_out
SMEMQ=$PWD/ab_libs/wip_k2-smem-queries.so
PREF=$PWD/ab_libs/wip_pred-code-prefetch.so
: > $O/${T}_ab_k2.log
for i in 1 2; do
  for which in main smemq; do
    if [ $which = smemq ]; then export CODDVEC_LIB=$SMEMQ; else unset CODDVEC_LIB; fi
    PROBE_BATCHES=1,2,3,4,8 PROBE_PATHS=scan python tests/quick_probe.py 2>&1 | sed "s/^/$which /" | tee -a $O/${T}_ab_k2.log
    PROBE_DIM=768 PROBE_BATCHES=1,2,4 PROBE_PATHS=scan python tests/quick_probe.py 2>&1 | sed "s/^/$which /" | tee -a $O/${T}_ab_k2.log
  done
done
: > $O/${T}_ab_pred.log
for which in main pref main pref; do
  if [ $which = pref ]; then export CODDVEC_LIB=$PREF; else unset CODDVEC_LIB; fi
  PROBE_DIMS=768 python tests/pred_probe.py 2>&1 | sed "s/^/$which /" | tee -a $O/${T}_ab_pred.log
done
for lib in $SMEMQ $PREF; do
  CODDVEC_LIB=$lib timeout 900 python -m pytest tests/test_native_gpu.py tests/test_properties_gpu.py -m gpu -x -q 2>&1 | tail -4 | sed "s|^|$(basename $lib) |" | tee -a $O/${T}_ab_parity.log
done
unset CODDVEC_LIB
# K3 mid-batch on cfg2: where do the 40 us above the HBM floor go?
for B in 16 64 256; do
  CMD="env PROBE_BATCHES=$B PROBE_PATHS=mma python tests/quick_probe.py"
  $CMD > $O/${T}_plain_k3_b$B.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mma_topk -s 6 -c 1 -f -o $O/${T}_prof_k3_b$B $CMD > $O/${T}_ncu_k3_b$B.log 2>&1
done
CMD="env PROBE_BATCHES=64 PROBE_PATHS=mma python tests/quick_probe.py"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${T}_launches_k3_b64.csv $CMD > /dev/null 2>&1
# sanitizer, first look (each tool bounded)
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitizer_subset.py > $O/${T}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $O/${T}_sanitizer_rc.log
  tail -5 $O/${T}_sanitizer_$tool.log
done
tail -30 $O/${T}_ab_k2.log
tail -40 $O/${T}_ab_pred.log
cat $O/${T}_ab_parity.log
