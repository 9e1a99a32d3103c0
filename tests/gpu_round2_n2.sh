#!/bin/bash
# Round-2 evidence, two B200s: striped collection and multi-process parity, then both bench arms at N = 2.
mkdir -p gpurun_out
O=gpurun_out
T=r02z
python -m pytest tests/test_striped_gpu.py -m gpu -x -q 2>&1 | tail -4 | tee $O/${T}_n2_pytest_striped.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | tail -6 | tee $O/${T}_n2_multi_gpu_check.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > $O/${T}_bench_reference_n2.json 2> $O/${T}_bench_reference_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 5 > $O/${T}_bench_n2.json 2> $O/${T}_bench_n2.err
tail -c 600 $O/${T}_bench_n2.json
