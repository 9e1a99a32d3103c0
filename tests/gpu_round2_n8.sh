#!/bin/bash
# Round-2 evidence, eight B200s: multi-process parity, the striped collection on 4 devices, both bench legs at N = 8.
mkdir -p gpurun_out
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_check.py 2>&1 | tail -8 | tee $O/r02_n8_multi_gpu_check.log
python -m pytest tests/test_striped_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $O/r02_n8_pytest_striped.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 30 --warmup 5 > $O/r02_bench_n8.json 2> $O/r02_bench_n8.err
tail -c 500 $O/r02_bench_n8.json
