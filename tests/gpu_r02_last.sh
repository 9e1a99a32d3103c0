#!/bin/bash
mkdir -p gpurun_out
python tests/bench_configs.py cfg1 cfg3 cfg5 > gpurun_out/r02_final_configs.jsonl 2> gpurun_out/r02_final_configs.err
python tests/host_scaling_probe.py 2>&1 | tail -3 > gpurun_out/r02_final_host_scaling.jsonl
tail -c 700 gpurun_out/r02_final_host_scaling.jsonl
