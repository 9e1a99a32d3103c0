"""Recall@k report (north-star: "Recall@k against ChromaDB's own approximate HNSW answer is reported
separately"): the GPU path's exact answer against the HNSW restatement with the reference's collection
parameters (oracle/hnsw.c; store.py:60-69) and against the exact CPU oracle.  Run: pytest -m gpu."""
import numpy as np
import pytest

from oracle import brute_force as bf
from oracle import hnsw
from tests.parity import check_topk, oracle_full_scores

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("clustered", [False, True])
def test_recall_report(native_lib, clustered):
    from codd_query_engine_b200 import native

    rng = np.random.default_rng(11)
    n, D, k, nq = 10_000, 384, 10, 64
    if clustered:
        cent = rng.standard_normal((256, D)).astype(np.float32)
        x = (cent[rng.integers(0, 256, n)] + 0.3 * rng.standard_normal((n, D))).astype(np.float32)
        q = (cent[rng.integers(0, 256, nq)] + 0.3 * rng.standard_normal((nq, D))).astype(np.float32)
    else:
        x = rng.standard_normal((n, D)).astype(np.float32)
        q = rng.standard_normal((nq, D)).astype(np.float32)
    idx = native.NativeIndex(D)
    idx.upsert(x)
    stored, scale, _, _ = idx.get_rows(np.arange(n))
    _, exact, _ = bf.query(stored, scale, q, k)
    rows_scan = np.concatenate([idx.query(q[i:i + 1], k, path=native.PATH_SCAN)[1] for i in range(nq)])
    sc_mma, rows_mma, cnt_mma = idx.query(q, k, path=native.PATH_MMA)
    idx.close()
    graph = hnsw.HnswIndex(D, n)
    graph.add(x)
    approx, _ = graph.search(q, k)
    graph.close()
    # the exact scan IS the brute-force answer; the tensor path rounds the query to bf16 and may swap near-ties
    # (< 1e-3, the north-star tolerance; on iid Gaussian rows neighbouring scores at rank 10 are ~1e-4 apart, so
    # about 1 % of the slots differ -- its score parity is what tests/test_mma_gpu.py checks)
    assert hnsw.recall_at_k(rows_scan, exact) == 1.0
    # the tensor path: north-star parity (scores within 1e-3, row sets identical except near-ties inside that tolerance)
    check_topk(sc_mma, rows_mma, cnt_mma, oracle_full_scores(stored, scale, q), None, k, 1e-3)
    print(f"recall@{k} of the tensor path against the exact answer: {hnsw.recall_at_k(rows_mma, exact):.3f}")
    r = hnsw.recall_at_k(approx, rows_scan)
    print(f"recall@{k} of the HNSW answer against the GPU answer ({'clustered' if clustered else 'iid gaussian'}): {r:.3f}")
    assert r >= (0.97 if clustered else 0.3)
