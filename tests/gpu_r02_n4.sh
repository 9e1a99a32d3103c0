#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/r02_final_bench_n4.json 2> gpurun_out/r02_final_bench_n4.err
tail -c 300 gpurun_out/r02_final_bench_n4.json
