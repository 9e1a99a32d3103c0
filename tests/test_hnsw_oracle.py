"""oracle/hnsw.c -- the restatement of the approximate index the reference configures
(codd_dal/metrics/metrics_semantic_metadata_store.py:60-69) -- checked against the exact brute-force oracle.
CPU only; the GPU leg (recall of the graph's answer against this package's exact answer) is
tests/bench_configs.py cfg1 and test_full_size_gpu.py::test_recall_report."""
import numpy as np
import pytest

from oracle import brute_force as bf
from oracle import hnsw


def exact_rows(x, q, k):
    xn = x / np.linalg.norm(x, axis=1, keepdims=True)
    qn = q / np.linalg.norm(q, axis=1, keepdims=True)
    s = qn @ xn.T
    order = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return order, 1.0 - np.take_along_axis(s, order, axis=1)


def test_reference_parameters():
    # the collection metadata the reference passes (store.py:63-68)
    assert (hnsw.REF_M, hnsw.REF_EF_CONSTRUCTION, hnsw.REF_EF_SEARCH) == (16, 200, 100)


def test_empty_index_pads():
    g = hnsw.HnswIndex(8, 4)
    ids, dist = g.search(np.ones((2, 8), np.float32), 3)
    assert (ids == -1).all() and np.isinf(dist).all() and len(g) == 0


@pytest.mark.parametrize("n", [1, 5, 17, 100])
def test_small_collections_are_exact(n):
    """With n <= ef_search every reachable element is kept, so the walk is exhaustive -- which is why the
    reference's <=5-document tests see exact answers from an approximate index."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal((n, 24)).astype(np.float32)
    q = rng.standard_normal((7, 24)).astype(np.float32)
    g = hnsw.HnswIndex(24, n)
    g.add(x)
    k = min(10, n)
    ids, dist = g.search(q, k)
    er, ed = exact_rows(x, q, k)
    assert np.array_equal(ids, er)
    assert np.abs(dist - ed).max() <= 1e-6


def test_distances_are_one_minus_cosine_ascending():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3000, 48)).astype(np.float32) * rng.uniform(0.1, 9.0, (3000, 1)).astype(np.float32)
    q = rng.standard_normal((20, 48)).astype(np.float32)
    g = hnsw.HnswIndex(48, 3000)
    g.add(x)
    ids, dist = g.search(q, 10)
    assert (ids >= 0).all() and (np.diff(dist, axis=1) >= 0).all()
    xn = x / np.linalg.norm(x, axis=1, keepdims=True)
    qn = q / np.linalg.norm(q, axis=1, keepdims=True)
    want = 1.0 - np.einsum("bkd,bd->bk", xn[ids], qn)
    assert np.abs(want - dist).max() <= 2e-6
    assert all(len(set(r)) == len(r) for r in ids.tolist())


def test_recall_grows_with_ef_and_reaches_one():
    rng = np.random.default_rng(4)
    x = rng.standard_normal((4000, 64)).astype(np.float32)
    q = rng.standard_normal((50, 64)).astype(np.float32)
    g = hnsw.HnswIndex(64, 4000)
    g.add(x)
    er, _ = exact_rows(x, q, 10)
    rec = [hnsw.recall_at_k(g.search(q, 10, ef_search=ef)[0], er) for ef in (10, 100, 4000)]
    assert rec[0] <= rec[1] <= rec[2] and rec[1] >= 0.8 and rec[2] == 1.0


def test_clustered_data_recall_and_matches_bf16_oracle_rows():
    """On embeddings with structure the reference's parameters recover the exact top-10, and the exact oracle
    over the bf16-stored rows (what the GPU path computes) agrees with the fp32 graph on those rows."""
    rng = np.random.default_rng(5)
    D, n = 96, 5000
    cent = rng.standard_normal((64, D)).astype(np.float32)
    x = (cent[rng.integers(0, 64, n)] + 0.3 * rng.standard_normal((n, D))).astype(np.float32)
    q = (cent[rng.integers(0, 64, 40)] + 0.3 * rng.standard_normal((40, D))).astype(np.float32)
    g = hnsw.HnswIndex(D, n)
    g.add(x)
    ids, dist = g.search(q, 10)
    stored, scale = bf.normalise_store(x)
    s, r, c = bf.query(stored, scale, q, 10)
    assert hnsw.recall_at_k(ids, r) >= 0.97
    # same rows -> same distances within the bf16 storage error (north-star tolerance 1e-3)
    same = ids == r
    assert same.mean() > 0.9 and np.abs((1.0 - s)[same] - dist[same]).max() <= 1e-3


def test_threads_do_not_change_answers():
    rng = np.random.default_rng(6)
    x = rng.standard_normal((2000, 32)).astype(np.float32)
    q = rng.standard_normal((64, 32)).astype(np.float32)
    g = hnsw.HnswIndex(32, 2000)
    g.add(x)
    a = g.search(q, 10, threads=1)
    b = g.search(q, 10, threads=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert g.dist_evals(reset=True) > 0 and g.dist_evals() == 0


def test_concurrent_insertion_builds_an_equally_good_graph():
    """hnsw_add_parallel (the library's add_items(num_threads) analogue): same recall as the sequential build, every
    element inserted exactly once, incremental use after a sequential prefix."""
    rng = np.random.default_rng(8)
    D, n = 48, 12000
    cent = rng.standard_normal((32, D)).astype(np.float32)
    x = (cent[rng.integers(0, 32, n)] + 0.3 * rng.standard_normal((n, D))).astype(np.float32)
    q = (cent[rng.integers(0, 32, 60)] + 0.3 * rng.standard_normal((60, D))).astype(np.float32)
    er, _ = exact_rows(x, q, 10)
    seq = hnsw.HnswIndex(D, n)
    seq.add(x)
    par = hnsw.HnswIndex(D, n)
    par.add(x[:500])
    par.add(x[500:], threads=4)
    assert len(par) == len(seq) == n
    r_seq = hnsw.recall_at_k(seq.search(q, 10)[0], er)
    ids, dist = par.search(q, 10)
    r_par = hnsw.recall_at_k(ids, er)
    assert r_par >= r_seq - 0.02 and r_par >= 0.95
    assert (ids >= 0).all() and (np.diff(dist, axis=1) >= 0).all()
    assert all(len(set(r)) == len(r) for r in ids.tolist())
    # every element is findable as its own nearest neighbour (nothing was lost or left unlinked)
    probe = rng.integers(0, n, 200)
    own, _ = par.search(x[probe], 1, ef_search=200)
    assert (own[:, 0] == probe).mean() >= 0.99
