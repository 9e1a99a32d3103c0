"""MicroBatcher: concurrent single-query callers are coalesced into one Collection.query and each gets
exactly the answer it would have got alone."""
import threading

import numpy as np
import pytest

from codd_query_engine_b200 import client as client_mod
from codd_query_engine_b200.batching import MicroBatcher
from tests.doubles import OracleIndex


@pytest.fixture(params=["double", pytest.param("native", marks=pytest.mark.gpu)])
def coll(request, monkeypatch):
    if request.param == "double":
        monkeypatch.setattr(client_mod, "NativeIndex", OracleIndex)
    c = client_mod.Client(endpoint=("batch", request.node.name))
    coll = c.get_or_create_collection("c", metadata={"hnsw:space": "cosine"})
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3000, 384)).astype(np.float32)
    coll.upsert(ids=[str(i) for i in range(3000)], embeddings=x,
                metadatas=[{"namespace": "a" if i % 2 else "b"} for i in range(3000)])
    yield coll, x
    c.reset()


def test_concurrent_requests_are_coalesced_and_answered_individually(coll):
    coll, x = coll
    mb = MicroBatcher(coll, max_batch=64, max_wait_ms=50)
    n_threads = 24
    out = [None] * n_threads
    start = threading.Barrier(n_threads)

    def worker(i):
        start.wait()
        where = {"namespace": "a"} if i % 3 == 0 else None
        out[i] = mb.query(query_embeddings=x[i * 7], n_results=5, where=where)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(n_threads)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    mb.close()
    assert mb.requests == n_threads and mb.batches < n_threads  # fewer launches than requests
    for i in range(n_threads):
        where = {"namespace": "a"} if i % 3 == 0 else None
        assert out[i]["ids"][0][0] == str(i * 7) or where is not None  # a vector's nearest neighbour is itself
        assert len(out[i]["ids"]) == 1 and len(out[i]["ids"][0]) == 5
        if where:
            assert all(int(j) % 2 == 1 for j in out[i]["ids"][0])
        d = out[i]["distances"][0]
        assert all(d[j] <= d[j + 1] for j in range(4))
        # the batched answer is the same top-5 (modulo the 1e-3 tensor-path tolerance) as a lone query
        alone = coll.query(query_embeddings=x[i * 7: i * 7 + 1], n_results=5, where=where)
        assert len(set(alone["ids"][0]) & set(out[i]["ids"][0])) >= 4
        assert np.allclose(sorted(alone["distances"][0]), sorted(d), atol=1e-3)


def test_text_queries_and_errors(coll):
    coll, x = coll
    mb = MicroBatcher(coll, max_batch=8, max_wait_ms=1)
    res = mb.query(query_texts=["cpu utilization"], n_results=3)
    assert len(res["ids"][0]) == 3 and res["included"]
    with pytest.raises(ValueError):
        mb.query(query_embeddings=x[0], n_results=0)      # the collection's error reaches the caller
    with pytest.raises(ValueError):
        mb.query(n_results=3)
    big = mb.query(query_embeddings=x[:4], n_results=2)    # explicit batches pass straight through
    assert len(big["ids"]) == 4
    mb.close()
    with pytest.raises(RuntimeError):
        mb.query(query_embeddings=x[0], n_results=3)


def test_env_wiring_routes_single_queries_through_a_batcher(coll, monkeypatch):
    """CODDVEC_MICROBATCH=1: unmodified callers (store.py:314-316 calls collection.query with one text) are
    coalesced without source edits; answers equal the direct search; batches of one do not recurse."""
    coll, x = coll
    want = [coll.query(query_embeddings=x[i], n_results=4) for i in range(12)]
    monkeypatch.setenv("CODDVEC_MICROBATCH", "1")
    monkeypatch.setenv("CODDVEC_MICROBATCH_WAIT_MS", "30")
    got = [None] * 12
    start = threading.Barrier(12)

    def worker(i):
        start.wait()
        got[i] = coll.query(query_embeddings=x[i], n_results=4)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(12)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert coll._batcher is not None and coll._batcher.requests == 12 and coll._batcher.batches < 12
    for g, w in zip(got, want):
        assert g["ids"] == w["ids"]
        assert np.allclose(g["distances"][0], w["distances"][0], atol=1e-3)
    lone = coll.query(query_embeddings=x[100], n_results=3)   # a batch of one
    assert lone["ids"][0][0] == "100"
    multi = coll.query(query_embeddings=x[:5], n_results=3)    # already a batch: straight through
    assert [r[0] for r in multi["ids"]] == ["0", "1", "2", "3", "4"]
    assert coll._batcher.requests == 13


def test_env_device_and_snapshot_dir(monkeypatch, tmp_path):
    monkeypatch.setattr(client_mod, "NativeIndex", OracleIndex)
    monkeypatch.setenv("CODDVEC_DEVICE", "0")
    monkeypatch.setenv("CODDVEC_SNAPSHOT_DIR", str(tmp_path))
    c = client_mod.HttpClient(host="snap-host", port=8123)
    assert c._device == 0
    coll = c.get_or_create_collection("metrics_semantic_metadata", metadata={"hnsw:space": "cosine"})
    rng = np.random.default_rng(1)
    x = rng.standard_normal((50, 384)).astype(np.float32)
    coll.upsert(ids=[f"ns#m{i}" for i in range(50)], embeddings=x, documents=[f"doc {i}" for i in range(50)])
    before = coll.query(query_embeddings=x[7], n_results=3)
    root = c.persist()
    assert root.startswith(str(tmp_path))
    c.reset()                                                   # "service restart"
    c2 = client_mod.HttpClient(host="snap-host", port=8123)
    again = c2.get_or_create_collection("metrics_semantic_metadata")
    assert again.count() == 50
    after = again.query(query_embeddings=x[7], n_results=3)
    assert after["ids"] == before["ids"] and after["documents"] == before["documents"]
    c2.reset()
