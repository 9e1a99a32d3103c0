#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/r02g_bench_default.json 2> gpurun_out/r02g_bench_default.err
tail -c 400 gpurun_out/r02g_bench_default.json; tail -3 gpurun_out/r02g_bench_default.err
