#!/bin/bash
# same-box A/B of the headline scan (100M x 384, batch 1): previous library vs current, alternating;
# then K2 parity tests and the dynamic instruction count of one scan launch
mkdir -p gpurun_out
: > gpurun_out/ab_headline.log
for i in 1 2; do
  for which in new old; do
    if [ $which = old ]; then export CODDVEC_LIB=$PWD/build_old_libcoddvec.so; else unset CODDVEC_LIB; fi
    python bench.py --steps 60 --warmup 5 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$which', round(d['ms_per_step'],4), 'ms/step', round(d['roofline']['achieved'],1), 'GB/s kernel', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],2), d['clocks'])
" | tee -a gpurun_out/ab_headline.log
  done
done
unset CODDVEC_LIB
python -m pytest tests/test_native_gpu.py tests/test_properties_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee -a gpurun_out/ab_headline.log
for which in new old; do
  if [ $which = old ]; then export CODDVEC_LIB=$PWD/build_old_libcoddvec.so; else unset CODDVEC_LIB; fi
  ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_scan_topk -s 4 -c 1 python bench.py --rows 12500000 --steps 3 --warmup 3 --no-extras 2>&1 | grep -E "smsp__inst|gpu__time" | sed "s/^/$which /" | tee -a gpurun_out/ab_headline.log
done
