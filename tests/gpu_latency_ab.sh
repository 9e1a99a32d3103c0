#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_native_gpu.py tests/test_properties_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/latency_ab.log
for which in new old; do
  if [ $which = old ]; then export CODDVEC_LIB=$PWD/build_old_libcoddvec.so; else unset CODDVEC_LIB; fi
  python tests/latency_probe.py 2>&1 | sed "s/^/$which /" | tee -a gpurun_out/latency_ab.log
done
unset CODDVEC_LIB
for rows in 1000 10000 100000 1000000; do
PROBE_ROWS=$rows PROBE_BATCHES=1 PROBE_PATHS=scan python tests/quick_probe.py 2>&1 | tee -a gpurun_out/latency_ab.log
done
