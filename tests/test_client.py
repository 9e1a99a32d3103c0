"""chromadb-compatible Client / Collection: conventions of SURVEY.md section 8(b) and the extensions
(embeddings=, query_embeddings=, where=, include=, multi-query, count, delete).  Host logic runs on
the CPU with the oracle-backed double; the same tests run on the GPU through libcoddvec.so."""
import numpy as np
import pytest

from codd_query_engine_b200 import client as client_mod
from codd_query_engine_b200.where import compile_where, evaluate, pack_mask, validate_where
from oracle import brute_force as bf
from tests.doubles import OracleIndex

COSINE = {"hnsw:space": "cosine", "hnsw:construction_ef": 200, "hnsw:search_ef": 100, "hnsw:M": 16}


@pytest.fixture(params=["double", pytest.param("native", marks=pytest.mark.gpu)])
def client(request, monkeypatch):
    if request.param == "double":
        monkeypatch.setattr(client_mod, "NativeIndex", OracleIndex)
    c = client_mod.Client(endpoint=("test", request.node.name))
    yield c
    c.reset()


def vecs(n, d=384, seed=0):
    return np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)


def test_get_or_create_is_idempotent_and_isolated(client):
    a = client.get_or_create_collection(name="a", metadata=COSINE)
    assert client.get_or_create_collection(name="a", metadata=COSINE) is a
    b = client.get_or_create_collection(name="b")
    a.upsert(ids=["x"], embeddings=vecs(1))
    assert a.count() == 1 and b.count() == 0
    assert a.metadata == COSINE  # hnsw:* keys are accepted and echoed back
    assert {c.name for c in client.list_collections()} == {"a", "b"}
    with pytest.raises(ValueError):
        client.create_collection("a")
    client.delete_collection("b")
    with pytest.raises(client_mod.NotFoundError):
        client.get_collection("b")
    assert isinstance(client.heartbeat(), int)


def test_clients_naming_one_endpoint_share_collections(monkeypatch):
    monkeypatch.setattr(client_mod, "NativeIndex", OracleIndex)
    c1 = client_mod.HttpClient(host="localhost", port=8000)
    c2 = client_mod.HttpClient(host="localhost", port=8000)
    c3 = client_mod.HttpClient(host="localhost", port=8001)
    c1.get_or_create_collection("shared").upsert(ids=["a"], embeddings=vecs(1))
    assert c2.get_collection("shared").count() == 1
    with pytest.raises(client_mod.NotFoundError):
        c3.get_collection("shared")
    c1.reset()


def test_query_shapes_and_conventions(client):
    coll = client.get_or_create_collection("c", metadata=COSINE)
    empty = coll.query(query_embeddings=vecs(2, seed=9), n_results=3)
    assert empty["ids"] == [[], []] and empty["distances"] == [[], []]
    x = vecs(50)
    ids = [f"ns#{i}" for i in range(50)]
    coll.upsert(ids=ids, embeddings=x, metadatas=[{"namespace": "ns", "i": i} for i in range(50)],
                documents=[f"doc {i}" for i in range(50)])
    res = coll.query(query_embeddings=x[7:9], n_results=5)
    assert [len(r) for r in res["ids"]] == [5, 5]
    assert res["ids"][0][0] == "ns#7" and res["ids"][1][0] == "ns#8"
    for b in range(2):
        d = res["distances"][b]
        assert all(d[i] <= d[i + 1] for i in range(4)) and abs(d[0]) < 1e-3  # ascending distance, self ~ 0
        assert res["documents"][b][0] == f"doc {7 + b}" and res["metadatas"][b][0]["i"] == 7 + b
    assert res["embeddings"] is None
    # fewer rows than n_results -> what exists
    assert len(coll.query(query_embeddings=x[:1], n_results=100)["ids"][0]) == 50
    # include
    only = coll.query(query_embeddings=x[:1], n_results=2, include=["distances"])
    assert only["metadatas"] is None and only["documents"] is None and len(only["distances"][0]) == 2
    emb = coll.query(query_embeddings=x[:1], n_results=2, include=["embeddings"])["embeddings"][0]
    assert emb.shape == (2, 384) and abs(np.linalg.norm(emb[0]) - 1) < 1e-2
    with pytest.raises(ValueError):
        coll.query(query_embeddings=x[:1], n_results=0)
    with pytest.raises(ValueError):
        coll.query(query_embeddings=x[:1], query_texts=["both"])
    with pytest.raises(ValueError):
        coll.query(query_embeddings=vecs(1, d=100), n_results=1)


def test_distances_match_oracle(client):
    coll = client.get_or_create_collection("c", metadata=COSINE)
    x, q = vecs(2000, seed=1), vecs(3, seed=2)
    coll.upsert(ids=[str(i) for i in range(2000)], embeddings=x)
    res = coll.query(query_embeddings=q, n_results=10)
    stored, scale = bf.normalise_store(x)
    s, r, c = bf.query(stored, scale, q, 10)
    for b in range(3):
        # batches > 2 run on the tensor path (bf16 queries): north-star tolerance 1e-3, near-ties may swap
        assert np.allclose(sorted(res["distances"][b]), sorted(1 - s[b]), atol=1e-3)
        assert len(set(res["ids"][b]) & {str(i) for i in r[b]}) >= 9


def test_upsert_replaces_vector_document_and_metadata(client):
    coll = client.get_or_create_collection("c", metadata=COSINE)
    x = vecs(10)
    coll.upsert(ids=[str(i) for i in range(10)], embeddings=x, documents=["old"] * 10, metadatas=[{"v": 1}] * 10)
    coll.upsert(ids=["3"], embeddings=-x[3:4], documents=["new"], metadatas=[{"v": 2}])
    assert coll.count() == 10
    got = coll.get(ids=["3"])
    assert got["ids"] == ["3"] and got["documents"] == ["new"] and got["metadatas"] == [{"v": 2}]
    res = coll.query(query_embeddings=x[3:4], n_results=10)
    assert res["ids"][0][-1] == "3" and res["distances"][0][-1] > 1.99  # now opposite to its old self
    assert coll.get(ids=["nope"])["ids"] == []
    with pytest.raises(client_mod.DuplicateIDError):
        coll.upsert(ids=["a", "a"], embeddings=x[:2])
    with pytest.raises(ValueError):
        coll.upsert(ids=["a"], embeddings=vecs(1, d=8))


def test_text_path_uses_embedding_function(client):
    calls = []

    def embed(texts):
        calls.append(list(texts))
        out = np.zeros((len(texts), 16), np.float32)
        for i, t in enumerate(texts):
            out[i, len(t) % 16] = 1.0
        return out

    coll = client.get_or_create_collection("c", metadata=COSINE, embedding_function=embed)
    coll.upsert(ids=["a", "b"], documents=["xx", "yyy"])
    res = coll.query(query_texts=["zz"], n_results=2)
    assert res["ids"][0][0] == "a" and abs(res["distances"][0][0]) < 1e-6 and abs(res["distances"][0][1] - 1) < 1e-6
    assert calls == [["xx", "yyy"], ["zz"]]
    with pytest.raises(ValueError):
        coll.upsert(ids=["c"])


def test_where_filters(client):
    coll = client.get_or_create_collection("c", metadata=COSINE)
    n = 300
    x = vecs(n, seed=3)
    rng = np.random.default_rng(4)
    ns = [f"ns{rng.integers(0, 3)}" for _ in range(n)]
    cat = [["system", "app", "db"][rng.integers(0, 3)] for _ in range(n)]
    lvl = rng.integers(0, 10, n).tolist()
    coll.upsert(ids=[str(i) for i in range(n)], embeddings=x,
                metadatas=[{"namespace": ns[i], "category": cat[i], "lvl": lvl[i]} for i in range(n)])
    stored, scale = bf.normalise_store(x)
    q = vecs(1, seed=5)
    full = bf.scores_fp32(stored, scale, bf.normalise_query(q))[0]

    def expect(pred, k=10):
        el = np.array([pred(i) for i in range(n)])
        order = [i for i in np.argsort(-full, kind="stable") if el[i]][:k]
        return [str(i) for i in order]

    cases = [
        ({"namespace": "ns1"}, lambda i: ns[i] == "ns1"),
        ({"namespace": {"$eq": "ns2"}}, lambda i: ns[i] == "ns2"),
        ({"$and": [{"namespace": "ns0"}, {"category": "db"}]}, lambda i: ns[i] == "ns0" and cat[i] == "db"),
        ({"lvl": {"$gte": 7}}, lambda i: lvl[i] >= 7),
        ({"$and": [{"namespace": "ns1"}, {"lvl": {"$lt": 3}}]}, lambda i: ns[i] == "ns1" and lvl[i] < 3),
        ({"$or": [{"category": "app"}, {"lvl": 0}]}, lambda i: cat[i] == "app" or lvl[i] == 0),
        ({"category": {"$in": ["db", "app"]}}, lambda i: cat[i] in ("db", "app")),
        ({"namespace": {"$ne": "ns0"}}, lambda i: ns[i] != "ns0"),
    ]
    for where, pred in cases:
        got = coll.query(query_embeddings=q, n_results=10, where=where)["ids"][0]
        assert got == expect(pred), where
    assert coll.query(query_embeddings=q, n_results=10, where={"namespace": "missing"})["ids"] == [[]]
    assert coll.query(query_embeddings=q, n_results=10, where={"lvl": {"$gt": 99}})["ids"] == [[]]
    with pytest.raises(ValueError):
        coll.query(query_embeddings=q, n_results=10, where={"$bogus": []})
    assert len(coll.get(where={"namespace": "ns1"})["ids"]) == sum(1 for v in ns if v == "ns1")


def test_delete_add_update(client):
    coll = client.get_or_create_collection("c", metadata=COSINE)
    x = vecs(20, seed=6)
    coll.add(ids=[str(i) for i in range(20)], embeddings=x, metadatas=[{"i": i} for i in range(20)])
    coll.add(ids=["5"], embeddings=-x[5:6])  # existing id: ignored
    assert coll.query(query_embeddings=x[5:6], n_results=1)["ids"][0] == ["5"]
    coll.delete(ids=["5", "6"])
    assert coll.count() == 18 and coll.get(ids=["5"])["ids"] == []
    top = coll.query(query_embeddings=x[5:6], n_results=20)["ids"][0]
    assert len(top) == 18 and "5" not in top and "6" not in top
    coll.upsert(ids=["new"], embeddings=x[5:6])  # reuses a tombstoned row
    assert coll.count() == 19 and coll.query(query_embeddings=x[5:6], n_results=1)["ids"][0] == ["new"]
    coll.update(ids=["7", "ghost"], metadatas=[{"tag": "t"}, {"tag": "u"}])
    assert coll.get(ids=["7"])["metadatas"] == [{"i": 7, "tag": "t"}] and coll.count() == 19
    coll.delete(where={"i": {"$lt": 3}})
    assert coll.count() == 16
    assert len(coll.peek(3)["ids"]) == 3


def test_ip_space(client):
    coll = client.get_or_create_collection("c", metadata={"hnsw:space": "ip"})
    x = np.array([[1, 0, 0, 0], [2, 0, 0, 0], [0, 3, 0, 0]], np.float32)
    coll.upsert(ids=["a", "b", "c"], embeddings=x)
    res = coll.query(query_embeddings=[[1, 0, 0, 0]], n_results=3)
    assert res["ids"][0] == ["b", "a", "c"] and res["distances"][0] == [-1.0, 0.0, 1.0]
    with pytest.raises(ValueError):
        client.get_or_create_collection("l2", metadata={"hnsw:space": "l2"})


def test_snapshot_roundtrip(client, tmp_path, request, monkeypatch):
    coll = client.get_or_create_collection("metrics semantic/1", metadata=COSINE)
    x = vecs(200, d=100, seed=8)
    coll.upsert(ids=[f"ns{i % 3}#m{i}" for i in range(200)], embeddings=x,
                metadatas=[{"namespace": f"ns{i % 3}", "category": "c%d" % (i % 2), "i": i} for i in range(200)],
                documents=[f"doc {i} \u00fc" for i in range(200)])
    coll.delete(ids=["ns1#m1", "ns2#m2"])
    q = vecs(2, d=100, seed=9)
    before = coll.query(query_embeddings=q, n_results=7, where={"namespace": "ns0"})
    before_all = coll.query(query_embeddings=q, n_results=7)
    root = client.persist(str(tmp_path / "snap"))
    names = sorted(p.name for p in (tmp_path / "snap").iterdir())
    assert len(names) == 1 and names[0].startswith("metrics_semantic_1-")  # sanitised name + hash of the real one
    # a new persistent client pointed at the directory sees the same collection and answers identically
    c2 = client_mod.PersistentClient(root)
    try:
        restored = c2.get_collection("metrics semantic/1")
        assert restored.count() == 198 and restored.metadata == COSINE
        after = restored.query(query_embeddings=q, n_results=7, where={"namespace": "ns0"})
        after_all = restored.query(query_embeddings=q, n_results=7)
        assert after["ids"] == before["ids"] and after["distances"] == before["distances"]
        assert after_all["ids"] == before_all["ids"] and after_all["metadatas"] == before_all["metadatas"]
        assert restored.get(ids=["ns0#m3"])["documents"] == ["doc 3 \u00fc"]
        assert restored.get(ids=["ns1#m1"])["ids"] == []
        restored.upsert(ids=["fresh"], embeddings=q[:1])          # tombstoned slot is reused after restore
        assert restored.count() == 199 and restored.query(query_embeddings=q[:1], n_results=1)["ids"] == [["fresh"]]
    finally:
        c2.reset()


# ---------------------------------------------------------------- where compiler (pure host logic)
def test_where_compiler_splits_fused_and_residual_predicates():
    mds = [{"namespace": "a", "category": "x", "n": 1}, {"namespace": "b", "category": "x", "n": 2},
           {"namespace": "a", "category": "y", "n": 3}, None]
    ns_codes, lb_codes = {"a": 1, "b": 2}, {"x": 1, "y": 2}
    c = compile_where({"namespace": "a"}, mds, ns_codes, "category", lb_codes)
    assert (c.ns_code, c.label_code, c.mask_words, c.empty) == (1, 0xFFFFFFFF, None, False)
    c = compile_where({"$and": [{"namespace": "b"}, {"category": {"$eq": "x"}}]}, mds, ns_codes, "category", lb_codes)
    assert (c.ns_code, c.label_code, c.mask_words) == (2, 1, None)
    c = compile_where({"$and": [{"namespace": "a"}, {"n": {"$gte": 2}}]}, mds, ns_codes, "category", lb_codes)
    assert c.ns_code == 1 and c.mask_words.tolist() == [0b0110]
    c = compile_where({"$or": [{"namespace": "a"}, {"n": 2}]}, mds, ns_codes, "category", lb_codes)
    assert c.ns_code == 0xFFFFFFFF and c.mask_words.tolist() == [0b0111]
    assert compile_where({"namespace": "zzz"}, mds, ns_codes, "category", lb_codes).empty
    assert compile_where({"n": 99}, mds, ns_codes, "category", lb_codes).empty
    alive = np.array([True, False, True, True])
    c = compile_where(None, mds, ns_codes, "category", lb_codes, alive=alive)
    assert c.mask_words.tolist() == [0b1101]
    assert compile_where(None, mds, ns_codes, "category", lb_codes).mask_words is None
    assert pack_mask(np.ones(33, bool)).tolist() == [0xFFFFFFFF, 1]
    assert evaluate({"k": {"$nin": [1]}}, {}) and not evaluate({"k": 1}, None)
    for bad in ({}, {"a": 1, "b": 2}, {"$and": [{"a": 1}]}, {"a": {"$regex": "x"}}, {"a": {"$in": 3}}, {"a": [1]}, "x"):
        with pytest.raises(ValueError):
            validate_where(bad)


def test_default_embedder_is_deterministic_unit_and_nonnegative():
    from codd_query_engine_b200.embedding import HashedBagOfWordsEmbedding

    e = HashedBagOfWordsEmbedding()
    a = e(["CPU utilization percentage", "cpu   UTILIZATION percentage!", "", "Memory in bytes"])
    assert a.shape == (4, 384) and a.dtype == np.float32 and (a >= 0).all()
    assert np.array_equal(a[0], a[1]) and np.allclose(np.linalg.norm(a[[0, 3]], axis=1), 1, atol=1e-6)
    assert not a[2].any()
    assert a[0] @ e(["cpu utilization"])[0] > a[3] @ e(["cpu utilization"])[0]
    assert e([]).shape == (0, 384)


def test_filtered_batches_choose_the_scan_when_the_filter_is_selective(monkeypatch):
    """Collection.query hands a FILTERED batch to the scan path (K2: eligibility first, only selected rows are read)
    when batch size x selectivity says it beats one tensor-path pass over every row; answers are the same either
    way.  The per-namespace row counts behind the estimate follow upserts, replacements, deletes and snapshots."""
    from codd_query_engine_b200 import client as client_mod
    from codd_query_engine_b200 import native
    from tests.doubles import OracleIndex

    monkeypatch.setattr(client_mod, "NativeIndex", OracleIndex)
    c = client_mod.Client(endpoint=("paths", "t"))
    coll = c.get_or_create_collection("c", metadata={"hnsw:space": "cosine"})
    rng = np.random.default_rng(0)
    n = 2000
    x = rng.standard_normal((n, 384)).astype(np.float32)
    ns = ["rare" if i % 50 == 0 else ("mid" if i % 5 == 1 else "big") for i in range(n)]   # 2 % / 20 % / 78 %
    coll.upsert(ids=[str(i) for i in range(n)], embeddings=x, metadatas=[{"namespace": s, "k": i % 7} for i, s in enumerate(ns)])
    assert coll._ns_rows[coll._ns_codes["rare"]] == 40 and coll._ns_rows[coll._ns_codes["mid"]] == 400

    def path_of(B, where):
        coll.query(query_embeddings=x[:B], n_results=3, where=where)
        return coll._index.last_requested_path

    assert path_of(1, {"namespace": "rare"}) == native.PATH_AUTO          # one query: AUTO already scans
    assert path_of(8, {"namespace": "rare"}) == native.PATH_SCAN          # 8 x (0.02 + 0.05) < 1
    assert path_of(3, {"namespace": "mid"}) == native.PATH_SCAN           # 3 x 0.25 < 1
    assert path_of(8, {"namespace": "mid"}) == native.PATH_AUTO           # 8 x 0.25 > 1: one tensor pass is cheaper
    assert path_of(2, {"namespace": "big"}) == native.PATH_AUTO
    assert path_of(64, None) == native.PATH_AUTO                          # unfiltered
    assert path_of(4, {"k": {"$in": [3]}}) == native.PATH_SCAN            # bitmask filter, ~14 % of the rows: 4 x 0.19 < 1
    assert path_of(6, {"k": {"$ne": 3}}) == native.PATH_AUTO              # bitmask filter, ~86 %
    # the counts follow the collection
    coll.upsert(ids=["0"], embeddings=x[:1], metadatas=[{"namespace": "mid"}])   # row 0 moves rare -> mid
    coll.delete(ids=["50"])                                                         # another rare row goes
    assert coll._ns_rows[coll._ns_codes["rare"]] == 38 and coll._ns_rows[coll._ns_codes["mid"]] == 401
    # and both paths give the same hits (the double is exact on either)
    a = coll.query(query_embeddings=x[5:9], n_results=4, where={"namespace": "rare"})
    b = [coll.query(query_embeddings=x[5 + i], n_results=4, where={"namespace": "rare"})["ids"][0] for i in range(4)]
    assert a["ids"] == b
    c.reset()


def test_snapshot_is_replaced_atomically_and_validated(client, tmp_path):
    """A re-save never leaves old metadata next to new data, names that sanitise alike keep separate directories,
    deleted collections do not come back, and a truncated data file is refused."""
    a = client.get_or_create_collection("a/b", metadata=COSINE)
    b = client.get_or_create_collection("a_b", metadata=COSINE)
    a.upsert(ids=["x", "y"], embeddings=vecs(2, d=16, seed=1))
    b.upsert(ids=["z"], embeddings=vecs(1, d=16, seed=2))
    root = client.persist(str(tmp_path / "snap"))
    assert len(list((tmp_path / "snap").iterdir())) == 2
    a.upsert(ids=["w"], embeddings=vecs(1, d=16, seed=3))
    client.delete_collection("a_b")
    client.persist(root)
    dirs = sorted(p.name for p in (tmp_path / "snap").iterdir())
    assert len(dirs) == 1 and not any(".tmp-" in d or ".old-" in d for d in dirs)
    c2 = client_mod.PersistentClient(root)
    try:
        assert [c.name for c in c2.list_collections()] == ["a/b"] and c2.get_collection("a/b").count() == 3
    finally:
        c2.reset()
    data = tmp_path / "snap" / dirs[0] / "rows.bf16"
    data.write_bytes(data.read_bytes()[:-8])
    with pytest.raises(ValueError, match="corrupt"):
        client_mod.Collection.load(str(tmp_path / "snap" / dirs[0]))


def test_failed_upsert_keeps_tombstones_and_metadata_updates_reach_the_device(client):
    coll = client.get_or_create_collection("adv", metadata=COSINE)
    x = vecs(6, d=16, seed=4)
    coll.upsert(ids=list("abcdef"), embeddings=x, metadatas=[{"namespace": "n1", "category": "k"}] * 6)
    coll.delete(ids=["b", "c"])
    bad = vecs(2, d=16, seed=5)
    bad[1, 3] = np.nan
    with pytest.raises(ValueError):
        coll.upsert(ids=["g", "h"], embeddings=bad)           # rejected: the two free slots must still be free
    assert sorted(coll._free_rows) == [1, 2] and coll.count() == 4
    res = coll.query(query_embeddings=x[1:2], n_results=6)      # deleted rows never come back, with or without a mask
    assert None not in res["ids"][0] and "b" not in res["ids"][0] and len(res["ids"][0]) == 4
    coll.upsert(ids=["g"], embeddings=x[1:2])                   # reuses a tombstoned slot
    assert coll.count() == 5 and coll.query(query_embeddings=x[1:2], n_results=1)["ids"] == [["g"]]
    # metadata-only update that moves a row to another namespace: query(where=) and get(where=) must agree
    coll.update(ids=["a"], metadatas=[{"namespace": "n2"}])
    assert coll.get(ids=["a"])["metadatas"] == [{"namespace": "n2", "category": "k"}]      # merged, not replaced
    assert coll.query(query_embeddings=x[:1], n_results=3, where={"namespace": "n2"})["ids"] == [["a"]]
    assert "a" not in coll.query(query_embeddings=x[:1], n_results=6, where={"namespace": "n1"})["ids"][0]
    assert coll.get(where={"namespace": "n2"})["ids"] == ["a"]
    coll.update(ids=["d"], embeddings=x[3:4] * 2, metadatas=[{"extra": 1}])                  # with embeddings: merged too
    assert coll.get(ids=["d"])["metadatas"] == [{"namespace": "n1", "category": "k", "extra": 1}]


def test_snapshot_hand_over_between_processes(client, tmp_path, monkeypatch):
    """The indexer job and the service are two processes with one snapshot directory between them (the stand-in for
    the Chroma server's volume): what the 'job' persists, a 'service' collection that was loaded from the older
    snapshot adopts IN PLACE (its store keeps the same Collection object) -- unless it has local writes of its own."""
    monkeypatch.setenv("CODDVEC_SNAPSHOT_RELOAD_S", "0.000001")
    x = vecs(60, d=64, seed=3)
    job = client.get_or_create_collection("metrics", metadata=COSINE)
    job.upsert(ids=[f"m{i}" for i in range(40)], embeddings=x[:40], metadatas=[{"namespace": "a"}] * 40)
    root = client.persist(str(tmp_path / "vol"))
    d = str(next((tmp_path / "vol").iterdir()))
    service = client_mod.Collection.load(d, device=client._device)  # "another process": its own object, its own index
    service._bind_snapshot(d)
    try:
        assert service.count() == 40
        assert service.query(query_embeddings=x[50:51], n_results=1)["ids"][0][0] != "m50"
        # the job indexes more and persists again; the service sees it at its next read, same object
        job.upsert(ids=[f"m{i}" for i in range(40, 60)], embeddings=x[40:], metadatas=[{"namespace": "b"}] * 20)
        assert client_mod._persist_modified(client) == ["metrics"] and client_mod._persist_modified(client) == []
        assert service.query(query_embeddings=x[50:51], n_results=1)["ids"] == [["m50"]]
        assert service.count() == 60 and service.get(ids=["m59"])["metadatas"] == [{"namespace": "b"}]
        assert service.query(query_embeddings=x[50:51], n_results=1, where={"namespace": "a"})["ids"][0][0] != "m50"
        # local writes win: a service that has written itself does not adopt the job's next snapshot
        service.upsert(ids=["local"], embeddings=x[:1] * 2)
        job.delete(ids=["m0"])
        client.persist(root)
        assert service.count() == 61 and service.get(ids=["m0"])["ids"] == ["m0"]
    finally:
        service.close()


def test_http_client_persists_at_exit_and_reloads(monkeypatch, tmp_path):
    """CODDVEC_SNAPSHOT_DIR wiring for unmodified callers: HttpClient registers the exit hook once per endpoint; the
    hook writes what changed; the next process's first client loads it."""
    monkeypatch.setattr(client_mod, "NativeIndex", OracleIndex)
    monkeypatch.setenv("CODDVEC_SNAPSHOT_DIR", str(tmp_path))
    registered = []
    monkeypatch.setattr("atexit.register", lambda fn, *a, **k: registered.append(fn))
    c1 = client_mod.HttpClient(host="chroma", port=8123)
    c1b = client_mod.HttpClient(host="chroma", port=8123)
    assert len(registered) == 1  # once per endpoint
    c1.get_or_create_collection("m").upsert(ids=["a", "b"], embeddings=vecs(2, d=32))
    registered[0]()  # "the job exits"
    assert (tmp_path / "chroma_8123" / "m" / "meta.json").is_file()
    c1.reset()
    c1b.reset()
    client_mod._EXIT_HOOKS.discard(c1.endpoint)
    c2 = client_mod.HttpClient(host="chroma", port=8123)  # "the service starts"
    try:
        assert c2.get_collection("m").count() == 2
    finally:
        c2.reset()
        client_mod._EXIT_HOOKS.discard(c2.endpoint)


def test_metadata_columns_agree_with_row_by_row_evaluation_on_mixed_types():
    """The dictionary-encoded column store must answer exactly like `evaluate` row by row, also where Python's
    equality is loose (1 == 1.0 == True) and where rows lack the field, are empty or deleted."""
    from codd_query_engine_b200.where import MetadataColumns

    mds = [{"a": 1}, {"a": 1.0}, {"a": True}, {"a": "1"}, None, {}, {"b": 2}, {"a": 0}, {"a": False}, {"a": 2.5}] * 7
    cols = MetadataColumns()
    clauses = [{"a": 1}, {"a": {"$in": [1, "1"]}}, {"a": {"$nin": [0]}}, {"a": {"$ne": 1}}, {"a": True}, {"a": {"$gt": 0}},
               {"a": {"$lte": 1}}, {"b": {"$gte": 2}}, {"$or": [{"a": "1"}, {"b": 2}]}, {"$and": [{"a": {"$gt": 0}}, {"a": {"$lt": 2}}]}]
    for where in clauses:
        want = np.array([evaluate(where, md) for md in mds])
        assert np.array_equal(cols.evaluate_rows(where, mds), want), where
    cols.set_row(3, {"a": 1})   # writes after the first touch keep the column in step
    cols.set_row(0, None)
    mds[3], mds[0] = {"a": 1}, None
    for where in clauses:
        assert np.array_equal(cols.evaluate_rows(where, mds), np.array([evaluate(where, md) for md in mds])), where
