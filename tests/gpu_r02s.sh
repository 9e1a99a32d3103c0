#!/bin/bash
# ablation: heads-bound refresh by every thread / one duty CTA per query / none  x  publications per insertion / per tile
mkdir -p gpurun_out
O=gpurun_out
L=$O/r02s_ab.log
: > $L
run() { local label=$1; shift; env "$@" PROBE_PATHS=mma python tests/quick_probe.py 2>&1 | sed "s/^/$label /" | tee -a $L; }
for rows in 1000000 100000; do
for rep in 1 2; do
  run old_$rows PROBE_ROWS=$rows PROBE_BATCHES=4,16,64,128,256,1024 CODDVEC_LIB=$PWD/ab_libs/d43732d.so CODDVEC_LIB_ALLOW_MISSING=1
  run r1p1_$rows PROBE_ROWS=$rows PROBE_BATCHES=4,16,64,128,256,1024
  run r1p0_$rows PROBE_ROWS=$rows PROBE_BATCHES=4,16,64,128,256,1024 CODDVEC_K3_PUB_TILE=0
  run r0p1_$rows PROBE_ROWS=$rows PROBE_BATCHES=4,16,64,128,256,1024 CODDVEC_K3_REFRESH=0
  run r0p0_$rows PROBE_ROWS=$rows PROBE_BATCHES=4,16,64,128,256,1024 CODDVEC_K3_REFRESH=0 CODDVEC_K3_PUB_TILE=0
  run r2p1_$rows PROBE_ROWS=$rows PROBE_BATCHES=4,16,64,128,256,1024 CODDVEC_K3_REFRESH=2
done
done
CODDVEC_K3_TRACE=1 CODDVEC_K3_TIMELINE=1 PROBE_BATCHES=4,64,256 PROBE_PATHS=mma python tests/quick_probe.py > $O/r02s_trace_raw.log 2>&1
for b in 4 64 256; do grep -A78 "timeline\] B=$b " $O/r02s_trace_raw.log | tail -79; done > $O/r02s_trace.log
CODDVEC_K3_REFRESH=2 CODDVEC_K3_TRACE=1 CODDVEC_K3_TIMELINE=1 PROBE_BATCHES=64 PROBE_PATHS=mma python tests/quick_probe.py > $O/r02s_trace_r2_raw.log 2>&1
grep -A78 "timeline\] B=64 " $O/r02s_trace_r2_raw.log | tail -79 > $O/r02s_trace_r2.log
timeout 900 python -m pytest tests/test_mma_gpu.py tests/test_cfg2_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r02s_pytest.log
