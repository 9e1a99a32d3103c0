"""Multi-GPU parity check, run under torchrun on N real GPUs:

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Shards a small corpus N ways (NCCL all-gather + K4 merge) and requires the answer to be bit-identical
to the 1-GPU answer of the same library, for the scan path (B=1), the tensor path (B=64) and a
namespace predicate.  Prints one OK line per rank 0.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from codd_query_engine_b200 import native  # noqa: E402
from codd_query_engine_b200.sharded import ShardedIndex  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, D, k = 300_001, 384, 10
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, D)).astype(np.float32)
    x[n - 7] = x[11]  # cross-shard exact tie
    ns = rng.integers(1, 6, n).astype(np.uint32)
    whole = native.NativeIndex(D, device=local)
    whole.upsert(x, ns_codes=ns)
    ok = True
    modes = [m for m in os.environ.get("EXCHANGE_MODES", "nccl,fused").split(",") if m]
    for mode in modes:
        ok &= check_mode(mode, rank, world, dev, x, ns, n, D, k, whole, rng)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if int(flag.item()) == 1 else "FAILED")
    whole.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def check_mode(mode, rank, world, dev, x, ns, n, D, k, whole, rng):
    idx = ShardedIndex(D, n, device=dev, exchange=mode, max_queries=256, max_k=32)
    idx.append_local(x[idx.lo:idx.hi], ns_codes=ns[idx.lo:idx.hi])
    assert idx.count() == n
    if rank == 0:
        print(f"exchange requested={mode} active={idx.exchange}")
    ok = idx.exchange == mode
    rng = np.random.default_rng(1)
    for B, path in ((1, native.PATH_SCAN), (3, native.PATH_SCAN), (7, native.PATH_SCAN), (64, native.PATH_MMA),
                    (200, native.PATH_MMA)):
        q = np.concatenate([x[11:12], rng.standard_normal((B - 1, D)).astype(np.float32)]) if B > 1 else x[11:12].copy()
        s, r, c = idx.query(q, k, path=path)
        ws, wr, wc = whole.query(q, k, path=path)
        same = np.array_equal(r, wr) and np.array_equal(s, ws) and np.array_equal(c, wc)
        tie = r[0, 0] == 11 and r[0, 1] == n - 7
        ok &= same and tie
        if rank == 0:
            print(f"world={world} {mode} B={B} path={path}: identical_to_1gpu={same} tie_by_global_row={tie}")
        s, r, c = idx.query(q, k, path=path, ns_code=3)
        ws, wr, wc = whole.query(q, k, path=path, ns_code=3)
        same = np.array_equal(r, wr) and np.array_equal(s, ws)
        ok &= same
        if rank == 0:
            print(f"world={world} {mode} B={B} path={path} ns=3: identical_to_1gpu={same}")
    # repeated batches exercise the two key slots / sequence flags of the fused exchange
    q = rng.standard_normal((2, D)).astype(np.float32)
    ws, wr, wc = whole.query(q, k, path=native.PATH_SCAN)
    for _ in range(50):
        s, r, c = idx.query(q, k, path=native.PATH_SCAN)
        ok &= bool(np.array_equal(r, wr) and np.array_equal(s, ws))
    idx.close()
    return ok


if __name__ == "__main__":
    main()
