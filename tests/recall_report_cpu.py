#!/usr/bin/env python
"""CPU-only recall / timing report of the HNSW restatement (oracle/hnsw.c, the reference's collection parameters)
against the exact brute-force oracle -- which IS this package's answer (GPU parity tests) -- at sizes beyond cfg1.

    python tests/recall_report_cpu.py [--rows N] [--kind clustered|gaussian] [--queries Q] [--threads T]

One JSON line per run.  A single-threaded graph build (one upsert after another) takes minutes at 1M rows;
--threads T inserts concurrently, as the library's add_items(num_threads) does."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (C oracle loader)
from oracle import brute_force as bf  # noqa: E402
from oracle import hnsw  # noqa: E402


def arg(name, default, cast=int):
    return cast(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    n, nq, D, k = arg("--rows", 1_000_000), arg("--queries", 100), 384, 10
    kind = arg("--kind", "clustered", str)
    rng = np.random.default_rng(0)
    if kind == "clustered":  # 256 Gaussian centroids, sigma 0.3
        cent = rng.standard_normal((256, D), dtype=np.float32)
        x = cent[rng.integers(0, 256, n)] + 0.3 * rng.standard_normal((n, D), dtype=np.float32)
        q = cent[rng.integers(0, 256, nq)] + 0.3 * rng.standard_normal((nq, D), dtype=np.float32)
    else:
        x = rng.standard_normal((n, D), dtype=np.float32)
        q = rng.standard_normal((nq, D), dtype=np.float32)
    lib = bench.load_c_oracle()
    stored, scale = bf.normalise_store(x)
    t_bf, threads, keys = bench.cpu_time_queries(lib, stored, scale, q[:1], k, 5)
    _, exact = bf.unpack_keys(bench.cpu_time_queries(lib, stored, scale, q, k, 1)[2])
    threads = arg("--threads", 1)
    t0 = time.perf_counter()
    g = hnsw.HnswIndex(D, n)
    g.add(x, threads=threads)
    t_build = time.perf_counter() - t0
    g.search(q[:10], k)
    g.dist_evals(reset=True)
    t0 = time.perf_counter()
    approx = np.concatenate([g.search(q[i:i + 1], k)[0] for i in range(nq)])
    t_q = (time.perf_counter() - t0) / nq
    evals = g.dist_evals() / nq
    print(json.dumps({"rows": n, "dim": D, "kind": kind, "k": k, "queries": nq,
                      "hnsw_params": "cosine, M=16, ef_construction=200, ef_search=100",
                      "hnsw_build_s": t_build, "hnsw_build_threads": threads, "hnsw_query_ms_1_thread": t_q * 1e3, "hnsw_dist_evals_per_query": evals,
                      "hnsw_max_level": g.max_level, "recall_at_k_hnsw_vs_exact": hnsw.recall_at_k(approx, exact),
                      "exact_cpu_oracle_ms_per_query": float(np.median(t_bf)) * 1e3, "exact_cpu_oracle_threads": threads}), flush=True)
    g.close()


if __name__ == "__main__":
    main()
