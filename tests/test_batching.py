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
