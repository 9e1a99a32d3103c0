#!/bin/bash
# K3 per-CTA phase timeline (where do the ~45 us above the HBM floor go at 1M rows?), then the whole GPU suite and
# the default bench of the same code.
mkdir -p gpurun_out
O=gpurun_out
CODDVEC_K3_TIMELINE=1 PROBE_BATCHES=4,16,64,128,256 PROBE_PATHS=mma python tests/quick_probe.py > $O/r02n_timeline_raw.log 2>&1
grep -A9 "k3 timeline\] B=" $O/r02n_timeline_raw.log | awk '/B=4 /{c4++} /B=16 /{c16++} /B=64 /{c64++} /B=128 /{c128++} /B=256 /{c256++} {print}' | tail -120 > $O/r02n_timeline.log
PROBE_BATCHES=1,2,4,16,64,128,256,1024 python tests/quick_probe.py 2>&1 | tee $O/r02n_probe.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/r02n_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/r02n_bench_default.json 2> $O/r02n_bench_default.err
tail -c 1500 $O/r02n_bench_default.json
