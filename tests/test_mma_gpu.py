"""GPU parity tests of the tensor-core path (K3: TMA + tcgen05 + fused top-k) against the oracle."""
import numpy as np
import pytest

from oracle import brute_force as bf
from tests.parity import check_topk, oracle_full_scores

pytestmark = pytest.mark.gpu

TOL = 1e-3  # BASELINE.json north_star: scores within 1e-3 absolute (queries are rounded to bf16 on this path)


@pytest.fixture(scope="module")
def native(native_lib):
    from codd_query_engine_b200 import native as nat

    return nat


def build(native, n, D, seed, metric=bf.METRIC_COSINE, **kw):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, D)).astype(np.float32)
    idx = native.NativeIndex(D, metric=metric, row_offset=kw.pop("row_offset", 0))
    idx.upsert(x, **kw)
    vec, scale, _, _ = idx.get_rows(np.arange(n))
    return idx, vec, scale, rng


def bf16_query_scores(stored, scale, q, metric=bf.METRIC_COSINE, split=None):
    """What K3 computes exactly: the normalised query rounded to bf16, fp32 accumulate -- plus, on rows of up to
    128 padded elements (CV_PRECISION_AUTO) or when asked for (CV_PRECISION_SPLIT), the bf16 residual of the query
    through a second MMA."""
    qhat = bf.normalise_query(q, metric)
    hi = bf.bf16_bits_to_f32(bf.f32_to_bf16_bits(qhat))
    out = bf.scores_fp32(stored, scale, hi)
    if split is None:
        split = -(-stored.shape[1] // 64) * 64 <= 128
    if split:
        lo = bf.bf16_bits_to_f32(bf.f32_to_bf16_bits(qhat - hi))
        out = out + bf.scores_fp32(stored, scale, lo)
    return out


VARIANTS = ["1cta", "2cta"]


def variant_path(native, variant):
    return {"1cta": native.PATH_MMA_1CTA, "2cta": native.PATH_MMA_2CTA, "auto": native.PATH_MMA}[variant]


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("D,B,n", [(384, 3, 1000), (384, 128, 4096), (64, 1, 130), (128, 130, 777), (512, 17, 2049),
                                   (100, 5, 500), (384, 260, 3000), (384, 513, 70000), (768, 33, 3000)])
def test_full_score_matrix(native, D, B, n, variant):
    """Every score the tcgen05 pipeline produces (all tiles, all K blocks, tails) vs the oracle."""
    idx, stored, scale, rng = build(native, n, D, seed=D + B + n)
    q = rng.standard_normal((B, D)).astype(np.float32)
    got = idx.debug_mma_scores(q, path=variant_path(native, variant))
    want = bf16_query_scores(stored, scale, q)
    assert got.shape == want.shape
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= 3e-6, f"max err {np.abs(got - want).max()}"
    assert np.abs(got - oracle_full_scores(stored, scale, q)).max() <= TOL
    idx.close()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("D,B,k,n", [(384, 1, 10, 70001), (384, 16, 10, 70001), (384, 128, 10, 50000),
                                     (384, 129, 10, 50000), (384, 300, 5, 30000), (384, 64, 100, 20000),
                                     (384, 8, 128, 5000), (384, 5, 10, 7), (384, 5, 10, 128), (384, 5, 10, 129),
                                     (384, 5, 10, 257), (64, 32, 1, 10000), (512, 40, 20, 10000), (200, 9, 10, 10000),
                                     (384, 600, 10, 200000), (384, 1024, 30, 60000),
                                     (768, 20, 50, 30000), (768, 300, 10, 30000), (700, 7, 10, 5000),
                                     (384, 40, 50, 40000), (384, 130, 100, 40000), (256, 16, 128, 20000)])
def test_mma_topk_parity(native, D, B, k, n, variant):
    idx, stored, scale, rng = build(native, n, D, seed=D * 7 + B * 3 + k + n)
    q = rng.standard_normal((B, D)).astype(np.float32)
    scores, rows, counts = idx.query(q, k, path=variant_path(native, variant))
    assert idx.stats().last_path == native.PATH_MMA
    # exactness against what the tensor path computes ...
    check_topk(scores, rows, counts, bf16_query_scores(stored, scale, q), None, k, tol=3e-6)
    # ... and the acceptance bar against the fp32 oracle
    check_topk(scores, rows, counts, oracle_full_scores(stored, scale, q), None, k, tol=TOL)
    idx.close()


def test_mma_and_scan_paths_agree(native):
    idx, stored, scale, rng = build(native, 40000, 384, seed=5)
    q = rng.standard_normal((4, 384)).astype(np.float32)
    s1, r1, c1 = idx.query(q, 10, path=native.PATH_SCAN)
    s2, r2, c2 = idx.query(q, 10, path=native.PATH_MMA)
    full = oracle_full_scores(stored, scale, q)
    check_topk(s1, r1, c1, full, None, 10, tol=2e-6)
    check_topk(s2, r2, c2, full, None, 10, tol=TOL)  # same rows except near-ties inside the tolerance
    assert np.array_equal(c1, c2)
    idx.close()


@pytest.mark.parametrize("variant", VARIANTS)
def test_mma_ties_and_duplicates(native, variant):
    rng = np.random.default_rng(7)
    D, k = 384, 10
    base = rng.standard_normal((50, D)).astype(np.float32)
    x = np.concatenate([base] * 40)
    idx = native.NativeIndex(D)
    idx.upsert(x)
    q = np.repeat(base[3:4], 6, axis=0) + 0.0
    scores, rows, counts = idx.query(q, k, path=variant_path(native, variant))
    for b in range(6):
        assert rows[b].tolist() == [3 + 50 * i for i in range(k)]  # exact ties -> ascending row id
        assert len(set(scores[b].tolist())) == 1
    idx.close()


@pytest.mark.parametrize("variant", VARIANTS)
def test_mma_predicates_and_row_offset(native, variant):
    rng = np.random.default_rng(11)
    n, D, k, B = 30000, 384, 20, 7
    ns = rng.integers(0, 10, n).astype(np.uint32)
    label = rng.integers(0, 3, n).astype(np.uint32)
    idx, stored, scale, rng = build(native, n, D, seed=12, ns_codes=ns, label_codes=label, row_offset=5000)
    q = rng.standard_normal((B, D)).astype(np.float32)
    full = bf16_query_scores(stored, scale, q)
    bits = rng.random(n) < 0.05
    words = np.packbits(np.concatenate([bits, np.zeros((-n) % 32, bool)]), bitorder="little").view(np.uint32)
    for kw, el in (({"ns_code": 3}, ns == 3), ({"ns_code": 3, "label_code": 1}, (ns == 3) & (label == 1)),
                   ({"mask": words}, bits), ({"mask": words, "ns_code": 5}, bits & (ns == 5))):
        s, r, c = idx.query(q, k, path=variant_path(native, variant), **kw)
        check_topk(s, r, c, full, el, k, tol=3e-6, row_offset=5000)
    s, r, c = idx.query(q, k, ns_code=77, path=variant_path(native, variant))
    assert (c == 0).all() and (r == -1).all()
    idx.close()


@pytest.mark.parametrize("variant", VARIANTS)
def test_mma_exact_zero_ties_across_ctas(native, variant):
    """Scores that are exactly 0 (zero query, zero rows) tie across CTAs: the answer must still be the k
    smallest row ids, whichever CTA publishes its threshold first (regression: the predecessor of +0.0 in
    the ordered-float encoding is -0.0)."""
    rng = np.random.default_rng(3)
    n, D, k = 40000, 384, 10
    x = rng.standard_normal((n, D)).astype(np.float32)
    x[::3] = 0.0                                   # a third of the rows are all-zero (scale 0 -> score 0)
    idx = native.NativeIndex(D)
    idx.upsert(x)
    q = rng.standard_normal((140, D)).astype(np.float32)
    q[::2] = 0.0                                   # zero queries: every row scores exactly 0
    for _ in range(10):
        s, r, c = idx.query(q, k, path=variant_path(native, variant))
        assert (r[::2] == np.arange(k)[None, :]).all() and (s[::2] == 0).all()
    stored, scale, _, _ = idx.get_rows(np.arange(n))
    check_topk(s, r, c, bf16_query_scores(stored, scale, q), None, k, tol=3e-6)
    # a query orthogonal-ish to everything positive: make all non-zero rows score negative -> zeros win
    neg = -np.abs(x[1:2])
    xx = np.abs(x)
    idx2 = native.NativeIndex(D)
    idx2.upsert(xx)
    for _ in range(5):
        s2, r2, c2 = idx2.query(np.repeat(neg, 3, axis=0), k, path=variant_path(native, variant))
        assert (r2 == np.arange(0, 3 * k, 3)[None, :]).all() and (s2 == 0).all()
    idx.close()
    idx2.close()


def test_mma_ip_metric(native):
    idx, stored, scale, rng = build(native, 20000, 384, seed=13, metric=bf.METRIC_IP)
    q = rng.standard_normal((9, 384)).astype(np.float32)
    s, r, c = idx.query(q, 10, path=native.PATH_MMA)
    check_topk(s, r, c, bf16_query_scores(stored, scale, q, bf.METRIC_IP), None, 10, tol=5e-4)  # |score| ~ 60: a few fp32 ulp
    idx.close()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("D,B,k", [(768, 40, 10), (1024, 8, 10), (1536, 150, 10), (2048, 17, 20), (4096, 5, 10)])
def test_mma_long_rows_stream_the_query_tile(native, D, B, k, variant):
    """Above 512-d the query tile no longer stays resident: its K blocks ride through the TMA ring."""
    idx, stored, scale, rng = build(native, 9000, D, seed=D + B)
    q = rng.standard_normal((B, D)).astype(np.float32)
    got = idx.debug_mma_scores(q, path=variant_path(native, variant))
    assert np.abs(got - bf16_query_scores(stored, scale, q)).max() <= 5e-6
    s, r, c = idx.query(q, k, path=variant_path(native, variant))
    check_topk(s, r, c, bf16_query_scores(stored, scale, q), None, k, tol=5e-6)
    check_topk(s, r, c, oracle_full_scores(stored, scale, q), None, k, tol=TOL)
    s2, r2, c2 = idx.query(q, k)  # AUTO takes the tensor path for batches at every dimension
    assert idx.stats().last_path == native.PATH_MMA
    idx.close()


def test_auto_path_selection(native):
    idx, stored, scale, rng = build(native, 5000, 384, seed=15)
    idx.query(rng.standard_normal((1, 384)).astype(np.float32), 5)
    assert idx.stats().last_path == native.PATH_SCAN
    idx.query(rng.standard_normal((32, 384)).astype(np.float32), 5)
    assert idx.stats().last_path == native.PATH_MMA
    idx.close()


def test_large_batch_and_large_corpus_properties(native):
    """BASELINE-sized shapes through size-independent properties: planted neighbours are found."""
    import torch

    n, D, B, k = 2_000_000, 384, 1024, 10
    idx = native.NativeIndex(D, capacity_rows=n)
    g = torch.Generator(device="cuda").manual_seed(3)
    for lo in range(0, n, 500_000):
        x = torch.randn((500_000, D), device="cuda", generator=g)
        idx.upsert_device(x.data_ptr(), 500_000)
    del x
    rng = np.random.default_rng(21)
    planted = rng.choice(n, B, replace=False)
    vec, scale, _, _ = idx.get_rows(np.sort(planted))
    order = np.argsort(np.argsort(planted))
    q = bf.bf16_bits_to_f32(vec)[order]  # query b == stored row planted[b]
    s, r, c = idx.query(q, k, path=native.PATH_MMA)
    assert (c == k).all()
    assert np.array_equal(r[:, 0], planted)           # each query's own row is its top hit ...
    assert np.abs(s[:, 0] - 1.0).max() < 2e-3          # ... with cosine ~ 1
    assert (np.diff(s, axis=1) <= 0).all()             # descending
    assert (s[:, 1] < 0.5).all()                       # random 384-d vectors are nearly orthogonal
    s1, r1, c1 = idx.query(q[:4], k, path=native.PATH_SCAN)
    assert np.array_equal(r1[:, 0], planted[:4]) and np.abs(s1 - s[:4]).max() <= TOL
    idx.close()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("D", [8, 64, 100, 128])
def test_short_rows_meet_the_tolerance_through_the_query_split(native, D, variant):
    """Sparse / short vectors are the worst case of a bf16 query (error up to 2^-9 * sum|q_i c_i|): the split keeps
    K3 within 2e-5 of the fp32 oracle -- far inside the north star's 1e-3 -- at every dimension AUTO routes to it."""
    rng = np.random.default_rng(D)
    n, B, k = 30000, 40, 10
    x = rng.standard_normal((n, D)).astype(np.float32)
    x[:, 3:] *= (rng.random((n, D - 3)) < 0.1)          # mostly-zero rows: a handful of non-zeros each
    q = rng.standard_normal((B, D)).astype(np.float32)
    q[:, 3:] *= (rng.random((B, D - 3)) < 0.2)
    idx = native.NativeIndex(D)
    idx.upsert(x)
    stored, scale, _, _ = idx.get_rows(np.arange(n))
    full = oracle_full_scores(stored, scale, q)
    got = idx.debug_mma_scores(q, path=variant_path(native, variant))
    assert np.abs(got - full).max() <= 2e-5, np.abs(got - full).max()
    s, r, c = idx.query(q, k, path=variant_path(native, variant))
    check_topk(s, r, c, full, None, k, tol=2e-5)
    s, r, c = idx.query(q, k)  # AUTO
    check_topk(s, r, c, full, None, k, tol=2e-5 if idx.stats().last_path == native.PATH_MMA else 2e-6)
    # the plain bf16 tile, for contrast, is only good to its analytic bound
    one = idx.debug_mma_scores(q, path=variant_path(native, variant), precision=native.PRECISION_BF16)
    assert np.abs(one - bf16_query_scores(stored, scale, q, split=False)).max() <= 3e-6
    idx.close()


@pytest.mark.parametrize("variant", VARIANTS)
def test_forced_split_at_384(native, variant):
    idx, stored, scale, rng = build(native, 50000, 384, seed=77)
    q = rng.standard_normal((70, 384)).astype(np.float32)
    full = oracle_full_scores(stored, scale, q)
    got = idx.debug_mma_scores(q, path=variant_path(native, variant), precision=native.PRECISION_SPLIT)
    assert np.abs(got - full).max() <= 1e-5
    s, r, c = idx.query(q, 10, path=variant_path(native, variant), precision=native.PRECISION_SPLIT)
    check_topk(s, r, c, full, None, 10, tol=1e-5)
    with pytest.raises(Exception):
        idx2 = native.NativeIndex(768)
        idx2.upsert(rng.standard_normal((100, 768)).astype(np.float32))
        idx2.query(q[:, :1].repeat(768, axis=1), 5, path=native.PATH_MMA, precision=native.PRECISION_SPLIT)
    idx.close()


def test_back_to_back_batches_share_self_resetting_state(native):
    """The fused tail leaves the cross-CTA thresholds / counters zero for the next launch: batches of changing size,
    k, path and predicate on ONE index (one query context) must each be exact."""
    rng = np.random.default_rng(5)
    n, D = 120_000, 384
    ns = rng.integers(1, 4, n).astype(np.uint32)
    idx, stored, scale, rng = build(native, n, D, seed=6, ns_codes=ns)
    plan = [(200, 10, native.PATH_MMA, None), (5, 50, native.PATH_MMA_1CTA, None), (300, 10, native.PATH_MMA_2CTA, 2),
            (1, 10, native.PATH_MMA, None), (129, 64, native.PATH_MMA, None), (16, 100, native.PATH_MMA, None),
            (64, 10, native.PATH_MMA, 3), (1024, 10, native.PATH_MMA, None), (3, 1, native.PATH_MMA_2CTA, None),
            (200, 10, native.PATH_MMA, None)]
    for rep in range(2):
        for B, k, path, code in plan:
            q = rng.standard_normal((B, D)).astype(np.float32)
            kw = {} if code is None else {"ns_code": code}
            s, r, c = idx.query(q, k, path=path, **kw)
            check_topk(s, r, c, bf16_query_scores(stored, scale, q), None if code is None else ns == code, k, tol=3e-6)
    idx.close()
