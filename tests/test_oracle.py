"""The oracle itself: known answers computed by hand, numpy <-> C twin agreement, key algebra,
and the committed golden vectors (tests/golden/topk_small.json, made by tests/golden/make_topk_golden.py)."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import brute_force as bf

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def coracle():
    so = os.path.join(ROOT, "oracle", "liboracle_bf.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    lib.cvo_query.restype = ctypes.c_int32
    return lib


def p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def c_query(lib, stored, scale, q, k, metric=0, mask=None, ns=None, ns_code=bf.ANY_CODE, label=None,
            label_code=bf.ANY_CODE, row_offset=0, threads=0):
    n, D = stored.shape
    qhat = np.empty_like(q, dtype=np.float32)
    lib.cvo_normalise_query(p(np.ascontiguousarray(q, np.float32)), ctypes.c_int32(q.shape[0]), ctypes.c_int32(D),
                            ctypes.c_int32(metric), p(qhat))
    out = np.zeros((q.shape[0], k), np.uint64)
    lib.cvo_query(p(stored), p(scale), ctypes.c_int64(n), ctypes.c_int32(D), p(qhat), ctypes.c_int32(q.shape[0]),
                  ctypes.c_int32(k), p(mask), p(ns), ctypes.c_uint32(ns_code), p(label), ctypes.c_uint32(label_code),
                  ctypes.c_int64(row_offset), p(out), ctypes.c_int32(threads))
    return out


def test_bf16_rounding_known_values():
    x = np.array([1.0, 1.00390625, 1.005859375, 1.001953125, -2.5, 0.0, 3.0e38, 1e-40], np.float32)
    bits = bf.f32_to_bf16_bits(x)
    # 1.00390625 = 1 + 2^-8 is exactly half way between bf16 1.0 (0x3F80) and 1.0078125 (0x3F81): ties to even
    assert bits[0] == 0x3F80 and bits[1] == 0x3F80
    assert bits[2] == 0x3F81  # 1 + 1.5 * 2^-8 rounds up
    assert bits[3] == 0x3F80  # 1 + 2^-9 rounds down
    assert bits[4] == 0xC020 and bits[5] == 0
    back = bf.bf16_bits_to_f32(bits)
    assert back[4] == -2.5 and back[0] == 1.0
    assert bf.f32_to_bf16_bits(np.array([np.nan], np.float32))[0] == 0x7FFF


def test_known_answer_cosine():
    # 3-4-5 triangle: rows are exactly representable after normalisation? (0.6, 0.8) are not in bf16,
    # so check against the definition directly.
    x = np.array([[3, 4, 0, 0], [0, 0, 1, 0], [4, 3, 0, 0], [-3, -4, 0, 0], [0, 0, 0, 0]], np.float32)
    stored, scale = bf.normalise_store(x)
    e = bf.bf16_bits_to_f32(stored)
    assert np.allclose(e[0], [0.6, 0.8, 0, 0], atol=2e-3) and scale[4] == 0.0
    assert np.array_equal(e[1], [0, 0, 1, 0]) and scale[1] == 1.0
    q = np.array([[6, 8, 0, 0]], np.float32)
    s, r, c = bf.query(stored, scale, q, 5)
    assert r[0].tolist() == [0, 2, 1, 4, 3]  # cos = 1, .96, 0, 0 (zero row; tie with row 1 -> row id), -1
    assert abs(s[0, 0] - 1.0) < 1e-6 and abs(s[0, 1] - 0.96) < 1e-3 and s[0, 2] == 0 and s[0, 3] == 0
    assert abs(s[0, 4] + 1.0) < 1e-6 and c[0] == 5
    # chromadb convention (store.py:336): distance = 1 - score ascending == score descending
    assert (np.diff(1 - s[0]) >= 0).all()


def test_known_answer_ip():
    x = np.array([[1, 2], [3, 4], [-1, 0.5]], np.float32)
    stored, scale = bf.normalise_store(x, bf.METRIC_IP)
    s, r, c = bf.query(stored, scale, np.array([[2, 1]], np.float32), 3, metric=bf.METRIC_IP)
    assert r[0].tolist() == [1, 0, 2] and s[0].tolist() == [10.0, 4.0, -1.5]


def test_key_order_and_roundtrip():
    sc = np.array([1.0, 1.0, 0.5, 0.0, -0.0, -0.5, -np.inf, np.nan, np.inf, 1e-45, -1e-45], np.float32)
    rows = np.arange(len(sc))
    keys = bf.pack_keys(sc, rows)
    assert keys[0] > keys[1]  # equal score: smaller row wins
    assert keys[3] > keys[4] and (keys[3] >> np.uint64(32)) == (keys[4] >> np.uint64(32))  # -0 == +0
    assert (keys[7] >> np.uint64(32)) == (keys[6] >> np.uint64(32))  # NaN treated as -inf
    order = np.argsort(keys)[::-1].tolist()
    assert order == [8, 0, 1, 2, 9, 3, 4, 10, 5, 6, 7]
    s2, r2 = bf.unpack_keys(keys)
    assert np.array_equal(r2, rows)
    ok = ~np.isnan(sc)
    assert np.array_equal(s2[ok] + 0.0, sc[ok] + 0.0)
    assert (keys != 0).all()
    es, er = bf.unpack_keys(np.zeros(3, np.uint64))
    assert (er == -1).all()


def test_topk_edge_cases():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((10, 16)).astype(np.float32)
    stored, scale = bf.normalise_store(x)
    q = rng.standard_normal((2, 16)).astype(np.float32)
    s, r, c = bf.query(stored, scale, q, 25)  # k > n
    assert (c == 10).all() and (r[:, 10:] == -1).all() and (r[:, :10] >= 0).all()
    s, r, c = bf.query(stored[:0], scale[:0], q, 5)  # empty corpus
    assert (c == 0).all() and (r == -1).all()
    mask = bf.eligible_rows(10, mask_words=np.array([0b1000000101], np.uint32))
    assert mask.tolist() == [True, False, True, False, False, False, False, False, False, True]
    s, r, c = bf.query(stored, scale, q, 5, mask_words=np.array([0b1000000101], np.uint32))
    assert (c == 3).all() and set(r[0, :3].tolist()) == {0, 2, 9}
    s, r, c = bf.query(stored, scale, q, 5, mask_words=np.array([0], np.uint32))
    assert (c == 0).all()
    s, r, c = bf.query(stored, scale, np.zeros((1, 16), np.float32), 3)  # zero query: all ties -> row order
    assert r[0].tolist() == [0, 1, 2] and (s[0] == 0).all()


@pytest.mark.parametrize("D,B,k,metric", [(384, 3, 10, 0), (768, 1, 50, 0), (100, 2, 7, 1), (8, 4, 128, 0)])
def test_numpy_and_c_twins_agree(coracle, D, B, k, metric):
    rng = np.random.default_rng(D + B)
    n = 5000
    x = rng.standard_normal((n, D)).astype(np.float32)
    x[17] = 0
    x[18] = x[19]
    stored_np, scale_np = bf.normalise_store(x, metric)
    stored_c = np.empty((n, D), np.uint16)
    scale_c = np.empty(n, np.float32)
    coracle.cvo_normalise_store(p(x), ctypes.c_int64(n), ctypes.c_int32(D), ctypes.c_int32(metric), p(stored_c), p(scale_c))
    assert np.array_equal(stored_np, stored_c) and np.array_equal(scale_np, scale_c)
    q = rng.standard_normal((B, D)).astype(np.float32)
    ns = rng.integers(0, 4, n).astype(np.uint32)
    bits = rng.random(n) < 0.5
    words = np.packbits(np.concatenate([bits, np.zeros((-n) % 32, bool)]), bitorder="little").view(np.uint32)
    for kw in ({}, {"mask": words}, {"ns": ns, "ns_code": 2}, {"mask": words, "ns": ns, "ns_code": 1, "row_offset": 1000}):
        keys_c = c_query(coracle, stored_np, scale_np, q, k, metric, **kw)
        s, r, c = bf.query(stored_np, scale_np, q, k, metric=metric, mask_words=kw.get("mask"), ns=kw.get("ns"),
                           ns_code=kw.get("ns_code", bf.ANY_CODE), row_offset=kw.get("row_offset", 0))
        sc, rc = bf.unpack_keys(keys_c)
        assert np.array_equal((keys_c != 0).sum(1), c)
        # fp32 accumulation order differs (BLAS vs 16-lane tree): scores agree to a few ulp, rows agree
        # wherever the oracle's neighbouring scores are further apart than that
        assert np.abs(sc - s).max() <= 3e-6 * max(1.0, np.abs(s).max())
        gaps = np.abs(np.diff(s, axis=1))
        stable = np.ones_like(r, bool)
        stable[:, 1:] &= gaps > 1e-5
        stable[:, :-1] &= gaps > 1e-5
        stable &= r >= 0
        assert np.array_equal(rc[stable], r[stable])


def test_c_merge_twin(coracle):
    rng = np.random.default_rng(5)
    G, B, kin, k = 8, 3, 10, 10
    keys = bf.pack_keys(rng.standard_normal(G * B * kin).astype(np.float32), rng.permutation(G * B * kin)).reshape(G, B, kin)
    keys = np.sort(keys, axis=2)[:, :, ::-1].copy()
    keys[2, 1, 4:] = 0
    out = np.zeros((B, k), np.uint64)
    coracle.cvo_merge_keys(p(keys), ctypes.c_int32(G), ctypes.c_int32(B), ctypes.c_int32(kin), ctypes.c_int32(k), p(out))
    assert np.array_equal(out, bf.merge_keys(keys, k))
    # top-k of a union == top-k of per-shard top-k's (the property the multi-GPU path relies on)
    flat = np.sort(keys.transpose(1, 0, 2).reshape(B, -1), axis=1)[:, ::-1][:, :k]
    assert np.array_equal(out, flat)


def test_golden_vectors():
    path = os.path.join(HERE, "golden", "topk_small.json")
    g = json.load(open(path))
    for case in g["cases"]:
        x = np.array(case["x"], np.float32)
        q = np.array(case["q"], np.float32)
        stored, scale = bf.normalise_store(x, case["metric"])
        assert stored.tolist() == case["stored_bits"]
        assert scale.view(np.uint32).tolist() == case["scale_bits"]
        mask = None if case["mask_words"] is None else np.array(case["mask_words"], np.uint32)
        ns = None if case["ns"] is None else np.array(case["ns"], np.uint32)
        s, r, c = bf.query(stored, scale, q, case["k"], metric=case["metric"], mask_words=mask, ns=ns,
                           ns_code=case["ns_code"], row_offset=case["row_offset"])
        assert r.tolist() == case["rows"] and c.tolist() == case["counts"]
        assert np.abs(s - np.array(case["scores"], np.float32)).max() <= 1e-6


def test_chroma_probe_script_runs_and_reports():
    """tests/golden/make_chroma_golden.py must always be runnable: it probes sys.path and baseline/_ref for
    chromadb / hnswlib and says what it found (today: nothing -- numeric parity of the oracle is unpinned)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "golden", "make_chroma_golden.py"), "--probe"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("probe:"), out.stderr


def test_oracle_matches_chroma_golden():
    """The numeric pin of the oracle, active the day tests/golden/chroma_topk.json exists
    (written by make_chroma_golden.py from a real chromadb / hnswlib install): every distance the reference's index
    reported for a row must be 1 - (the oracle's score of that row) up to the bf16 storage rounding, and where the
    approximate index returned the exact top-k the row ids must agree."""
    import json

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chroma_topk.json")
    if not os.path.exists(path):
        pytest.skip("no chromadb / hnswlib install has been available to generate tests/golden/chroma_topk.json (parity unpinned)")
    doc = json.load(open(path))
    sys_path_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(sys_path_root, "tests", "golden"))
    from make_chroma_golden import corpus

    x, q = corpus(doc["seed"], doc["n"], doc["d"], doc["nq"])
    stored, scale = bf.normalise_store(x)
    full = bf.scores_fp32(stored, scale, bf.normalise_query(q))
    s, r, c = bf.query(stored, scale, q, doc["k"])
    for name in ("chromadb", "hnswlib"):
        if name not in doc:
            continue
        rows = np.asarray(doc[name]["rows"])
        dist = np.asarray(doc[name]["distances"], dtype=np.float32)
        ours = 1.0 - np.take_along_axis(full, rows, axis=1)
        assert np.abs(ours - dist).max() <= 1e-3, f"{name}: distances differ by {np.abs(ours - dist).max()}"  # bf16 storage
        exact = np.asarray(doc["exact_fp32_rows"])
        hit = [b for b in range(len(rows)) if set(rows[b]) == set(exact[b])]
        for b in hit:  # the graph walk found the exact answer: ours must be the same rows (near-ties aside)
            gap = np.diff(np.sort(np.asarray(doc["exact_fp32_distances"][b])))
            if gap.min() > 2e-3:
                assert list(r[b]) == list(exact[b]), f"{name}: query {b} rows differ"


@pytest.mark.parametrize("D", [64, 384, 768])
def test_oracle_agrees_with_independent_cosine_implementations(D):
    """Not the reference's dependency (chromadb / hnswlib are not installable here, so the oracle stays "unpinned" in
    the sense of DESIGN.md section 2) -- but two independent third-party implementations of the SAME definition
    hnswlib's cosine space uses, distance = 1 - x.q / (|x| |q|): scikit-learn's brute-force NearestNeighbors and
    scipy's cdist, fed the vectors the oracle actually stores (bf16 rows widened to fp32).  Distances agree to fp32
    rounding and the neighbour lists are identical wherever the k-th and (k+1)-th distances are further apart than that."""
    sk = pytest.importorskip("sklearn.neighbors")
    sp = pytest.importorskip("scipy.spatial.distance")
    rng = np.random.default_rng(D)
    n, B, k = 5000, 20, 10
    x = rng.standard_normal((n, D)).astype(np.float32)
    x[n // 2] = x[7]  # one exact duplicate: a tie, broken by row id in the oracle
    q = np.concatenate([x[7:8], rng.standard_normal((B - 1, D)).astype(np.float32)])
    stored, scale = bf.normalise_store(x)
    scores, rows, counts = bf.query(stored, scale, q, k)
    wide = bf.bf16_bits_to_f32(stored).astype(np.float64)  # what the index holds
    dist_sk, idx_sk = sk.NearestNeighbors(n_neighbors=k + 1, metric="cosine", algorithm="brute").fit(wide).kneighbors(q.astype(np.float64))
    full = sp.cdist(q.astype(np.float64), wide, metric="cosine")  # [B][n]
    assert (counts == k).all()
    for b in range(B):
        ours = 1.0 - scores[b].astype(np.float64)
        assert np.abs(ours - full[b, rows[b]]).max() < 5e-6            # every reported distance is scipy's for that row
        assert np.abs(ours - dist_sk[b, :k]).max() < 5e-6              # and the k smallest are scikit-learn's k smallest
        gap = dist_sk[b, k] - dist_sk[b, k - 1]
        if gap > 1e-5 and np.diff(dist_sk[b, :k]).min() > 1e-5:         # well separated: identical lists, same order
            assert rows[b].tolist() == idx_sk[b, :k].tolist()
    assert rows[0, 0] == 7 and rows[0, 1] == n // 2                     # the duplicate pair: lower row id first
