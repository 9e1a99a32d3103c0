"""GPU parity tests of the C-ABI (libcoddvec.so) against the oracle.  Run: pytest -m gpu."""
import numpy as np
import pytest

from oracle import brute_force as bf
from tests.parity import check_topk, oracle_full_scores

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def native(native_lib):
    from codd_query_engine_b200 import native as nat

    return nat


def make_index(native, x, metric=bf.METRIC_COSINE, ns=None, label=None, row_offset=0, chunk=None):
    idx = native.NativeIndex(x.shape[1], metric=metric, row_offset=row_offset)
    if x.shape[0]:
        if chunk is None:
            idx.upsert(x, ns_codes=ns, label_codes=label)
        else:
            for lo in range(0, x.shape[0], chunk):
                idx.upsert(x[lo:lo + chunk], ns_codes=None if ns is None else ns[lo:lo + chunk],
                           label_codes=None if label is None else label[lo:lo + chunk])
    return idx


def stored_of(idx):
    n = len(idx)
    vec, scale, ns, lb = idx.get_rows(np.arange(n))
    return vec, scale, ns, lb


@pytest.mark.parametrize("D", [8, 64, 100, 384, 768, 1000, 1024])
def test_k1_normalise_bit_exact(native, D):
    rng = np.random.default_rng(D)
    x = rng.standard_normal((3001, D)).astype(np.float32)
    x[5] = 0.0  # zero row
    x[6] *= 1e-20
    x[7] *= 1e18
    idx = make_index(native, x, chunk=1000)
    vec, scale, _, _ = stored_of(idx)
    evec, escale = bf.normalise_store(x)
    assert np.array_equal(vec, evec)
    assert np.array_equal(scale, escale)
    assert scale[5] == 0.0
    idx.close()


def test_k1_ip_stores_raw(native):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((100, 384)).astype(np.float32) * 3
    idx = make_index(native, x, metric=bf.METRIC_IP)
    vec, scale, _, _ = stored_of(idx)
    assert np.array_equal(vec, bf.f32_to_bf16_bits(x))
    assert (scale == 1.0).all()
    idx.close()


@pytest.mark.parametrize("D,B,k", [(384, 1, 10), (384, 2, 10), (384, 4, 10), (384, 3, 5), (768, 1, 50), (768, 2, 100),
                                   (64, 4, 1), (128, 1, 128), (1024, 1, 10), (200, 7, 10), (520, 5, 3)])
def test_scan_parity(native, D, B, k):
    rng = np.random.default_rng(D * 1000 + B * 10 + k)
    n = 70001
    x = rng.standard_normal((n, D)).astype(np.float32)
    q = rng.standard_normal((B, D)).astype(np.float32)
    idx = make_index(native, x)
    stored, scale, _, _ = stored_of(idx)
    scores, rows, counts = idx.query(q, k, path=native.PATH_SCAN)
    full = oracle_full_scores(stored, scale, q)
    check_topk(scores, rows, counts, full, None, k, tol=2e-6)
    # and against the oracle's own top-k through its public entry point
    es, er, ec = bf.query(stored, scale, q, k)
    assert np.array_equal(ec, counts)
    assert np.abs(es - scores).max() <= 2e-6
    idx.close()


@pytest.mark.parametrize("D,B,k", [(1100, 1, 10), (1536, 2, 10), (2048, 1, 50), (3072, 3, 5), (4096, 1, 128)])
def test_scan_wide_rows_parity(native, D, B, k):
    """Rows longer than 1024 elements take the wide kernel (query in shared memory)."""
    rng = np.random.default_rng(D + B + k)
    n = 20011
    x = rng.standard_normal((n, D)).astype(np.float32)
    x[100] = x[7]
    ns = rng.integers(1, 5, n).astype(np.uint32)
    q = rng.standard_normal((B, D)).astype(np.float32)
    idx = native.NativeIndex(D)
    idx.upsert(x, ns_codes=ns)
    stored, scale, _, _ = stored_of(idx)
    est, esc = bf.normalise_store(x)
    assert np.array_equal(stored, est) and np.array_equal(scale, esc)
    full = oracle_full_scores(stored, scale, q)
    s, r, c = idx.query(q, k, path=native.PATH_SCAN)
    assert idx.stats().last_path == native.PATH_SCAN
    check_topk(s, r, c, full, None, k, tol=4e-6)
    s, r, c = idx.query(q, k, ns_code=2, path=native.PATH_SCAN)
    check_topk(s, r, c, full, ns == 2, k, tol=4e-6)
    bits = rng.random(n) < 0.02
    words = np.packbits(np.concatenate([bits, np.zeros((-n) % 32, bool)]), bitorder="little").view(np.uint32)
    s, r, c = idx.query(q, k, mask=words, path=native.PATH_SCAN)
    check_topk(s, r, c, full, bits, k, tol=4e-6)
    idx.close()


@pytest.mark.parametrize("n", [0, 1, 2, 5, 9, 33, 127, 128, 129, 1000])
def test_scan_small_and_ragged(native, n):
    rng = np.random.default_rng(n)
    D, k = 384, 10
    x = rng.standard_normal((n, D)).astype(np.float32)
    q = rng.standard_normal((2, D)).astype(np.float32)
    idx = make_index(native, x)
    scores, rows, counts = idx.query(q, k, path=native.PATH_SCAN)
    if n == 0:
        assert (counts == 0).all() and (rows == -1).all()
    else:
        stored, scale, _, _ = stored_of(idx)
        check_topk(scores, rows, counts, oracle_full_scores(stored, scale, q), None, k, tol=2e-6)
    idx.close()


def test_scan_ties_broken_by_row_id(native):
    rng = np.random.default_rng(7)
    D, k = 384, 10
    base = rng.standard_normal((50, D)).astype(np.float32)
    x = np.concatenate([base] * 40)  # every vector appears 40 times -> exact score ties
    q = base[3:5] + 0.01 * rng.standard_normal((2, D)).astype(np.float32)
    idx = make_index(native, x)
    stored, scale, _, _ = stored_of(idx)
    scores, rows, counts = idx.query(q, k, path=native.PATH_SCAN)
    # duplicates give bit-identical scores, so the order is fully determined: rows 3, 53, 103, ...
    assert rows[0].tolist() == [3 + 50 * i for i in range(k)]
    assert rows[1].tolist() == [4 + 50 * i for i in range(k)]
    assert len(set(scores[0].tolist())) == 1
    check_topk(scores, rows, counts, oracle_full_scores(stored, scale, q), None, k, tol=2e-6)
    idx.close()


def test_scan_predicates(native):
    rng = np.random.default_rng(11)
    n, D, k = 50000, 384, 20
    x = rng.standard_normal((n, D)).astype(np.float32)
    ns = rng.integers(0, 10, n).astype(np.uint32)
    label = rng.integers(0, 3, n).astype(np.uint32)
    q = rng.standard_normal((2, D)).astype(np.float32)
    idx = make_index(native, x, ns=ns, label=label)
    stored, scale, gns, glabel = stored_of(idx)
    assert np.array_equal(gns, ns) and np.array_equal(glabel, label)
    full = oracle_full_scores(stored, scale, q)
    # namespace only
    s, r, c = idx.query(q, k, ns_code=3, path=native.PATH_SCAN)
    check_topk(s, r, c, full, ns == 3, k, tol=2e-6)
    # namespace + label
    s, r, c = idx.query(q, k, ns_code=3, label_code=1, path=native.PATH_SCAN)
    check_topk(s, r, c, full, (ns == 3) & (label == 1), k, tol=2e-6)
    # bitmask (1% selectivity) + namespace
    bits = rng.random(n) < 0.01
    words = np.packbits(np.concatenate([bits, np.zeros((-n) % 32, bool)]), bitorder="little").view(np.uint32)
    s, r, c = idx.query(q, k, mask=words, path=native.PATH_SCAN)
    check_topk(s, r, c, full, bits, k, tol=2e-6)
    s, r, c = idx.query(q, k, mask=words, ns_code=5, path=native.PATH_SCAN)
    check_topk(s, r, c, full, bits & (ns == 5), k, tol=2e-6)
    # oracle's own predicate path agrees (counts, scores, and rows wherever scores are not within rounding of a tie)
    es, er, ec = bf.query(stored, scale, q, k, mask_words=words, ns=ns, ns_code=5)
    assert np.array_equal(ec, c) and np.abs(es - s).max() <= 2e-6
    differ = er != r
    assert not differ.any() or (np.abs(np.diff(es, axis=1)).min(axis=1)[differ.any(axis=1)] <= 4e-6).all()
    # nothing eligible
    s, r, c = idx.query(q, k, ns_code=77, path=native.PATH_SCAN)
    assert (c == 0).all() and (r == -1).all()
    idx.close()


@pytest.mark.parametrize("D", [64, 384, 768, 1024])
def test_scan_predicate_ragged_sizes_and_selectivities(native, D):
    """The predicate scan compacts eligible rows 128 at a time into per-warp rings: sizes around the 128-row
    step / 4-row quad boundaries, selectivities from all rows to a single row, namespace + label + bitmask."""
    rng = np.random.default_rng(D)
    k = 7
    for n in (1, 3, 4, 5, 127, 128, 129, 131, 255, 1000, 4099, 70001):
        x = rng.standard_normal((n, D)).astype(np.float32)
        ns = rng.integers(1, 3, n).astype(np.uint32)
        label = rng.integers(1, 4, n).astype(np.uint32)
        q = rng.standard_normal((1, D)).astype(np.float32)
        idx = make_index(native, x, ns=ns, label=label)
        stored, scale, _, _ = stored_of(idx)
        full = oracle_full_scores(stored, scale, q)
        for sel in (1.0, 0.5, 0.03):
            bits = rng.random(n) < sel
            bits[rng.integers(0, n)] = True
            words = np.packbits(np.concatenate([bits, np.zeros((-n) % 32, bool)]), bitorder="little").view(np.uint32)
            s, r, c = idx.query(q, k, mask=words, path=native.PATH_SCAN)
            check_topk(s, r, c, full, bits, k, tol=2e-6)
            s, r, c = idx.query(q, k, mask=words, ns_code=1, label_code=2, path=native.PATH_SCAN)
            check_topk(s, r, c, full, bits & (ns == 1) & (label == 2), k, tol=2e-6)
        s, r, c = idx.query(q, k, ns_code=2, path=native.PATH_SCAN)
        check_topk(s, r, c, full, ns == 2, k, tol=2e-6)
        one = np.zeros(n, bool)
        one[n - 1] = True  # only the very last row
        words = np.packbits(np.concatenate([one, np.zeros((-n) % 32, bool)]), bitorder="little").view(np.uint32)
        s, r, c = idx.query(q, k, mask=words, path=native.PATH_SCAN)
        assert c[0] == 1 and r[0, 0] == n - 1
        idx.close()


def _ascending_corpus(n, D, seed):
    """Rows whose cosine to q grows with the row index: every row a warp meets beats all it has seen, so the
    per-warp candidate buffers fill and are pruned again and again (on random data the cross-CTA heads bound
    keeps them nearly empty and the prune path would go untested)."""
    rng = np.random.default_rng(seed)
    q = rng.standard_normal(D).astype(np.float32)
    qh = q / np.linalg.norm(q)
    u = rng.standard_normal((n, D)).astype(np.float32)
    u -= (u @ qh)[:, None] * qh[None, :]
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    c = np.linspace(0.0, 0.98, n, dtype=np.float32)
    x = c[:, None] * qh[None, :] + np.sqrt(1.0 - c * c)[:, None] * u
    return x.astype(np.float32), q


@pytest.mark.parametrize("k", [10, 50, 128])
def test_scan_ascending_scores_exercise_buffer_prunes(native, k):
    n, D = 1_500_000, 16
    x, q = _ascending_corpus(n, D, seed=k)
    ns = (np.arange(n) % 2 + 1).astype(np.uint32)
    idx = make_index(native, x, ns=ns)
    stored, scale, _, _ = stored_of(idx)
    for B in (1, 2, 4):
        qs = np.stack([q * (1.0 + 0.25 * b) for b in range(B)]).astype(np.float32)   # same direction: all ascending
        full = oracle_full_scores(stored, scale, qs)
        s, r, c = idx.query(qs, k, path=native.PATH_SCAN)
        check_topk(s, r, c, full, None, k, tol=2e-6)
        if B == 1:
            s, r, c = idx.query(qs, k, ns_code=2, path=native.PATH_SCAN)            # predicate kernel, 50 % selectivity
            check_topk(s, r, c, full, ns == 2, k, tol=2e-6)
    idx.close()


def test_scan_wide_ascending_scores_exercise_buffer_prunes(native):
    n, D, k = 200_000, 1040, 10
    x, q = _ascending_corpus(n, D, seed=3)
    idx = make_index(native, x)
    stored, scale, _, _ = stored_of(idx)
    full = oracle_full_scores(stored, scale, q[None, :])
    s, r, c = idx.query(q[None, :], k, path=native.PATH_SCAN)
    check_topk(s, r, c, full, None, k, tol=4e-6)
    idx.close()


def test_ip_metric(native):
    rng = np.random.default_rng(13)
    n, D, k = 20000, 384, 10
    x = rng.standard_normal((n, D)).astype(np.float32)
    q = rng.standard_normal((1, D)).astype(np.float32)
    idx = make_index(native, x, metric=bf.METRIC_IP)
    stored, scale, _, _ = stored_of(idx)
    s, r, c = idx.query(q, k, path=native.PATH_SCAN)
    check_topk(s, r, c, oracle_full_scores(stored, scale, q, bf.METRIC_IP), None, k, tol=2e-4)
    idx.close()


def test_overwrite_and_row_offset(native):
    rng = np.random.default_rng(17)
    n, D, k = 5000, 384, 5
    x = rng.standard_normal((n, D)).astype(np.float32)
    idx = make_index(native, x, row_offset=1_000_000)
    q = rng.standard_normal((1, D)).astype(np.float32)
    # overwrite row 1234 with the query itself -> it must become the top hit with score ~1
    out = idx.upsert(q, rows=np.array([1234]))
    assert out.tolist() == [1234] and len(idx) == n
    s, r, c = idx.query(q, k)
    assert r[0, 0] == 1_000_000 + 1234 and abs(s[0, 0] - 1.0) < 1e-3
    stored, scale, _, _ = stored_of(idx)
    check_topk(s, r, c, oracle_full_scores(stored, scale, q), None, k, tol=2e-6, row_offset=1_000_000)
    # append returns the next row
    out = idx.upsert(q)
    assert out.tolist() == [n] and len(idx) == n + 1
    idx.close()


def test_errors(native):
    idx = native.NativeIndex(384)
    x = np.zeros((2, 384), np.float32)
    x[1, 5] = np.nan
    with pytest.raises(ValueError):
        idx.upsert(x)
    assert len(idx) == 0
    with pytest.raises(ValueError):
        idx.upsert(np.zeros((2, 100), np.float32))
    idx.upsert(np.ones((3, 384), np.float32))
    with pytest.raises(ValueError):
        idx.query(np.ones((1, 384), np.float32), 0)
    with pytest.raises(ValueError):
        idx.query(np.ones((1, 384), np.float32), 1000)
    with pytest.raises(ValueError):
        idx.upsert(np.ones((1, 384), np.float32), rows=np.array([17]))
    q = np.ones((1, 384), np.float32)
    q[0, 0] = np.inf
    with pytest.raises(ValueError):
        idx.query(q, 3)
    idx.close()


def test_device_api_and_merge(native):
    import torch

    rng = np.random.default_rng(19)
    n, D, k, B = 40000, 384, 10, 3
    x = rng.standard_normal((n, D)).astype(np.float32)
    q = rng.standard_normal((B, D)).astype(np.float32)
    # two shards with global row offsets, device-resident query + K4 merge == single index answer
    half = n // 2
    a = make_index(native, x[:half], row_offset=0)
    b = make_index(native, x[half:], row_offset=half)
    whole = make_index(native, x)
    dq = torch.from_numpy(q).cuda()
    keys = torch.zeros((2, B, k), dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    a.query_device(dq.data_ptr(), B, k, keys[0].data_ptr(), stream=stream)
    b.query_device(dq.data_ptr(), B, k, keys[1].data_ptr(), stream=stream)
    merged = torch.zeros((B, k), dtype=torch.int64, device="cuda")
    native.merge_keys_device(0, keys.data_ptr(), 2, B, k, k, merged.data_ptr(), stream)
    torch.cuda.synchronize()
    ms, mr = native.decode_keys(merged.cpu().numpy().view(np.uint64))
    ws, wr, wc = whole.query(q, k)
    assert np.array_equal(mr, wr)
    assert np.array_equal(ms, ws)
    # K4 twin in the oracle
    om = bf.merge_keys(keys.cpu().numpy().view(np.uint64), k)
    assert np.array_equal(om, merged.cpu().numpy().view(np.uint64))
    for i in (a, b, whole):
        i.close()


def test_merge_many_lists_large_k(native):
    import torch

    rng = np.random.default_rng(23)
    G, B, kin, kout = 37, 5, 100, 100
    sc = rng.standard_normal((G, B, kin)).astype(np.float32)
    rows = rng.permutation(G * B * kin).reshape(G, B, kin).astype(np.uint32)
    keys = bf.pack_keys(sc.ravel(), rows.ravel()).reshape(G, B, kin)
    keys = np.sort(keys, axis=2)[:, :, ::-1].copy()
    keys[3, :, 50:] = 0  # a short list
    d_in = torch.from_numpy(keys.view(np.int64)).cuda()
    d_out = torch.zeros((B, kout), dtype=torch.int64, device="cuda")
    native.merge_keys_device(0, d_in.data_ptr(), G, B, kin, kout, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), bf.merge_keys(keys, kout))


def test_stats_and_profiling(native):
    rng = np.random.default_rng(29)
    x = rng.standard_normal((10000, 384)).astype(np.float32)
    idx = make_index(native, x)
    idx.set_profiling(True)
    idx.query(x[:1], 10)
    st = idx.stats()
    assert st.rows == 10000 and st.queries == 1 and st.kernel_launches >= 3
    assert st.last_query_ms > 0 and st.last_path == native.PATH_SCAN
    assert st.last_query_bytes == 10000 * (2 * 384 + 4)
    idx.close()


@pytest.mark.parametrize("B", [1, 2, 4])
def test_scan_k_sequence_on_one_index(native, B):
    """One long-lived index, n_results 100 -> 10 -> 100 -> 128 -> 1: the scan kernels' dynamic shared-memory limit is
    a per-function attribute that must never be lowered under a cached launch plan."""
    rng = np.random.default_rng(B)
    n, D = 50000, 384
    x = rng.standard_normal((n, D)).astype(np.float32)
    idx = make_index(native, x)
    stored, scale, _, _ = stored_of(idx)
    q = rng.standard_normal((B, D)).astype(np.float32)
    full = oracle_full_scores(stored, scale, q)
    for k in (100, 10, 100, 128, 1, 50, 100):
        s, r, c = idx.query(q, k, path=native.PATH_SCAN)
        check_topk(s, r, c, full, None, k, tol=2e-6)
    idx.close()
