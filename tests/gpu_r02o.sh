#!/bin/bash
# A/B on one box: 4 accumulator stages (single CTAs) + early final refresh + PDL after k_prep_queries, against the
# previous commit's library and against each switch turned off.
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02o_ab.log
: > $L
run() {  # label, env...
  local label=$1; shift
  env "$@" PROBE_BATCHES=4,16,32,64,128,256,1024 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$label /" | tee -a $L
}
for rep in 1 2; do
  run new
  run old CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1
  run lead0 CODDVEC_K3_LAST_LEAD=0
  run nopdl CODDVEC_K3_PDL=0
  run prep0 CODDVEC_K3_FUSED_PREP_MAX_B=0
done
run new12m PROBE_ROWS=12500000
run old12m PROBE_ROWS=12500000 CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1
CODDVEC_K3_TIMELINE=1 PROBE_BATCHES=16,64,256 PROBE_PATHS=mma python tests/quick_probe.py > $O/r02o_timeline_raw.log 2>&1
for b in 16 64 256; do grep -A10 "timeline\] B=$b " $O/r02o_timeline_raw.log | tail -11; done | tee $O/r02o_timeline.log
CODDVEC_K3_FUSED_PREP_MAX_B=0 CODDVEC_K3_TIMELINE=1 PROBE_BATCHES=16 PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | grep -A10 "timeline\] B=16 " | tail -11 | tee -a $O/r02o_timeline.log
timeout 900 python -m pytest tests/test_mma_gpu.py tests/test_cfg2_gpu.py tests/test_striped_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r02o_pytest.log
