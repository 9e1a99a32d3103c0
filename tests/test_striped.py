"""One process, several devices: the striped collection's host logic (row routing r -> shard r % G, dense global
numbering, per-shard masks, deletes / code rewrites) on CPU doubles -- native.StripedIndex's own code over
OracleIndex shards, the device group query replaced by per-shard oracle queries + the oracle's key merge."""
import numpy as np
import pytest

from codd_query_engine_b200 import client as client_mod
from codd_query_engine_b200.client import Client
from tests.doubles import OracleIndex, make_oracle_striped_index


@pytest.fixture()
def Striped():
    return make_oracle_striped_index()


def same(r1, r2, s1, s2):
    """Rows identical; scores equal up to the BLAS summation order of the CPU oracle, which depends on where a row
    sits in its matrix (on the GPU the comparison is bit-exact: tests/test_striped_gpu.py)."""
    return np.array_equal(r1, r2) and np.abs(np.asarray(s1) - np.asarray(s2)).max() <= 2e-6


@pytest.mark.parametrize("G", [2, 3, 8])
def test_striped_index_equals_one_index(Striped, G):
    rng = np.random.default_rng(G)
    n, D, k = 1003, 48, 10
    x = rng.standard_normal((n, D)).astype(np.float32)
    x[700] = x[5]  # exact tie across shards: the lower global row must come first
    ns = rng.integers(1, 4, n).astype(np.uint32)
    one, many = OracleIndex(D), Striped(D, devices=list(range(G)))
    for lo in range(0, n, 97):  # ragged batches: global rows are handed out densely, in order
        a = one.upsert(x[lo:lo + 97], ns_codes=ns[lo:lo + 97])
        b = many.upsert(x[lo:lo + 97], ns_codes=ns[lo:lo + 97])
        assert np.array_equal(a, b)
    assert len(many) == n and abs(len(many.shards[0]) - len(many.shards[-1])) <= 1
    q = np.concatenate([x[5:6], rng.standard_normal((6, D)).astype(np.float32)])
    for kw in ({}, {"ns_code": 2}):
        s1, r1, c1 = one.query(q, k, **kw)
        s2, r2, c2 = many.query(q, k, **kw)
        assert same(r1, r2, s1, s2) and np.array_equal(c1, c2)
        if not kw:
            assert r2[0, 0] == 5 and r2[0, 1] == 700
    bits = rng.random(n) < 0.1
    words = np.packbits(np.concatenate([bits, np.zeros((-n) % 32, bool)]), bitorder="little").view(np.uint32)
    s1, r1, c1 = one.query(q, k, mask=words)
    for kw in ({"mask": words}, {"mask_bits": bits}):
        s2, r2, c2 = many.query(q, k, **kw)
        assert same(r1, r2, s1, s2)
    # overwrite, delete, code rewrite, readback
    rows = np.array([3, 700, 41, 1002])
    y = rng.standard_normal((4, D)).astype(np.float32)
    assert np.array_equal(one.upsert(y, rows=rows), many.upsert(y, rows=rows))
    mixed = np.array([7, -1, -1, 8])
    assert np.array_equal(one.upsert(y, rows=mixed), many.upsert(y, rows=mixed))
    one.delete_rows([5, 6, 1004])
    many.delete_rows([5, 6, 1004])
    one.set_codes([10, 11], ns_codes=[2, 2])
    many.set_codes([10, 11], ns_codes=[2, 2])
    for kw in ({}, {"ns_code": 2}):
        s1, r1, c1 = one.query(q, k, **kw)
        s2, r2, c2 = many.query(q, k, **kw)
        assert same(r1, r2, s1, s2)
        assert 5 not in r2 and 6 not in r2
    probe = np.array([0, 1, 2, 699, 700, 1003, 1004])
    for a, b in zip(one.get_rows(probe), many.get_rows(probe)):
        assert np.array_equal(a, b, equal_nan=True)
    with pytest.raises(ValueError):
        many.upsert(y, rows=np.array([0, 1, 2, 5000]))
    bad = y.copy()
    bad[2, 3] = np.nan
    before = len(many)
    with pytest.raises(ValueError):
        many.upsert(bad)
    assert len(many) == before  # nothing committed on any shard


def test_striped_snapshot_roundtrip(Striped):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((501, 32)).astype(np.float32)
    a = Striped(32, devices=[0, 1, 2])
    a.upsert(x, ns_codes=np.arange(501, dtype=np.uint32) % 5)
    a.delete_rows([17])
    vec, scale, ns, lb = a.get_rows(np.arange(501))
    b = Striped(32, devices=[0, 1])  # restore onto a different number of devices
    b.load_rows(vec, scale, ns, lb)
    q = rng.standard_normal((3, 32)).astype(np.float32)
    for kw in ({}, {"ns_code": 3}):
        sa, ra, ca = a.query(q, 7, **kw)
        sb, rb, cb = b.query(q, 7, **kw)
        assert same(ra, rb, sa, sb)


def test_collection_on_several_devices_behaves_like_one(monkeypatch, Striped):
    monkeypatch.setattr(client_mod, "NativeIndex", OracleIndex)
    monkeypatch.setattr(client_mod, "StripedIndex", Striped)
    rng = np.random.default_rng(3)
    n, D = 240, 24
    x = rng.standard_normal((n, D)).astype(np.float32)
    mds = [{"namespace": f"ns{i % 3}", "category": "c%d" % (i % 2), "n": i} for i in range(n)]
    ids = [f"ns{i % 3}#m{i}" for i in range(n)]
    one = Client("one-device").get_or_create_collection("c", metadata={"hnsw:space": "cosine"})
    many = Client("three-devices").get_or_create_collection("c", metadata={"hnsw:space": "cosine", "coddvec:devices": "0,1,2"})
    assert isinstance(many._ensure_index(D), Striped) and not isinstance(one._ensure_index(D), Striped)
    for coll in (one, many):
        for i in range(0, n, 1):  # the reference upserts one row per call
            coll.upsert(ids=[ids[i]], embeddings=[x[i]], metadatas=[mds[i]], documents=[f"doc {i}"])
        coll.delete(ids=[ids[7], ids[8]])
        coll.upsert(ids=["new#a", "new#b"], embeddings=x[:2] * 2, metadatas=[{"namespace": "new"}, {"namespace": "new"}])
        coll.update(ids=[ids[20]], metadatas=[{"namespace": "ns9"}])  # metadata-only: the device codes must follow
    q = rng.standard_normal((4, D)).astype(np.float32)
    for where in (None, {"namespace": "ns1"}, {"namespace": "ns9"}, {"namespace": "new"}, {"n": {"$in": [3, 4, 5, 100]}},
                  {"$and": [{"namespace": "ns2"}, {"category": "c0"}]}, {"$and": [{"namespace": "ns0"}, {"n": {"$gte": 100}}]}):
        a = one.query(query_embeddings=q, n_results=6, where=where)
        b = many.query(query_embeddings=q, n_results=6, where=where)
        assert a["ids"] == b["ids"] and a["metadatas"] == b["metadatas"], where
        assert all(np.allclose(da, db, atol=2e-6) for da, db in zip(a["distances"], b["distances"])), where
    assert many.query(query_embeddings=q[:1], n_results=3, where={"namespace": "ns9"})["ids"] == [[ids[20]]]
    assert one.count() == many.count() == n
    assert one.get(where={"n": {"$lt": 5}})["ids"] == many.get(where={"n": {"$lt": 5}})["ids"]


def test_devices_from_the_environment(monkeypatch, Striped):
    monkeypatch.setattr(client_mod, "NativeIndex", OracleIndex)
    monkeypatch.setattr(client_mod, "StripedIndex", Striped)
    monkeypatch.setenv("CODDVEC_DEVICES", "0,1")
    coll = Client("env-devices").get_or_create_collection("c")  # an unmodified caller: no coddvec keys anywhere
    coll.upsert(ids=["a", "b", "c"], embeddings=np.eye(3, 8, dtype=np.float32))
    assert isinstance(coll._index, Striped) and coll._index.G == 2
    assert coll.query(query_embeddings=np.eye(3, 8, dtype=np.float32)[1:2], n_results=1)["ids"] == [["b"]]
