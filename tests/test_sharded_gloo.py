"""Row-sharded search (sharded.ShardedIndex) with world_size 2 over gloo on the CPU: the host-side
logic of the multi-GPU path -- block partition, global row ids, all-gather of per-shard top-k,
merge -- must give exactly the single-shard answer.  Compute is the oracle-backed double."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from codd_query_engine_b200.sharded import shard_bounds


def test_shard_bounds_cover_rows_exactly():
    for total in (0, 1, 7, 8, 9, 1000, 100_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(lo <= hi for lo, hi in spans)
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans[:-1] or spans) <= max(1, -(-total // world))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, D, B, k, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from codd_query_engine_b200 import sharded
        from tests import doubles

        sharded.NativeIndex = doubles.OracleIndex            # test doubles: the product has no CPU path
        sharded.merge_keys_device = doubles.oracle_merge_keys_device
        rng = np.random.default_rng(0)
        x = rng.standard_normal((n, D)).astype(np.float32)
        x[n // 2 + 3] = x[5]                                 # a cross-shard exact tie
        q = np.concatenate([x[5:6], rng.standard_normal((B - 1, D)).astype(np.float32)])
        idx = sharded.ShardedIndex(D, n, device=torch.device("cpu"))
        assert (idx.lo, idx.hi) == sharded.shard_bounds(n, world, rank)
        got = idx.append_local(x[idx.lo:idx.hi])
        assert got.tolist() == list(range(idx.lo, idx.hi))   # global row ids
        assert idx.count() == n
        s, r, c = idx.query(q, k)
        keys = idx.query_device(torch.from_numpy(q), k)      # device-resident API, same answer
        np.save(os.path.join(tmp, f"rows{rank}.npy"), r)
        np.save(os.path.join(tmp, f"scores{rank}.npy"), s)
        np.save(os.path.join(tmp, f"keys{rank}.npy"), keys.numpy().view(np.uint64))
        if rank == 0:
            np.save(os.path.join(tmp, "x.npy"), x)
            np.save(os.path.join(tmp, "q.npy"), q)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_single_shard(tmp_path, world):
    from oracle import brute_force as bf

    n, D, B, k = 1001, 64, 4, 10
    mp.spawn(_worker, args=(world, _free_port(), n, D, B, k, str(tmp_path)), nprocs=world, join=True)
    x, q = np.load(tmp_path / "x.npy"), np.load(tmp_path / "q.npy")
    stored, scale = bf.normalise_store(x)
    es, er, ec = bf.query(stored, scale, q, k)
    for rank in range(world):
        r = np.load(tmp_path / f"rows{rank}.npy")
        s = np.load(tmp_path / f"scores{rank}.npy")
        assert np.array_equal(r, er), f"rank {rank}"       # bit-identical to the one-shard answer
        assert np.array_equal(s, es)
        ks, kr = bf.unpack_keys(np.load(tmp_path / f"keys{rank}.npy"))
        assert np.array_equal(kr, er)
    # the planted duplicate lives in the other shard: the tie is broken by global row id
    assert er[0, 0] == 5 and er[0, 1] == n // 2 + 3
