"""Randomised property tests (hypothesis) of the CUDA paths against the oracle, and a concurrency test of
the thread-safety contract of include/coddvec.h (appends never disturb running queries; overwrites are
ordered against them)."""
import os
import threading

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import brute_force as bf
from tests.parity import check_topk, oracle_full_scores

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def native(native_lib):
    from codd_query_engine_b200 import native as nat

    return nat


@settings(max_examples=int(os.environ.get("CODDVEC_HYPOTHESIS_EXAMPLES", "40")), deadline=None,
          suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(data=st.data())
def test_random_shapes_match_oracle(native, data):
    D = data.draw(st.sampled_from([8, 24, 64, 100, 128, 384, 520, 768, 1500]))
    n = data.draw(st.integers(1, 3000))
    B = data.draw(st.integers(1, 9))
    k = data.draw(st.sampled_from([1, 2, 5, 10, 33, 100, 128]))
    metric = data.draw(st.sampled_from([bf.METRIC_COSINE, bf.METRIC_IP]))
    seed = data.draw(st.integers(0, 2**31 - 1))
    dup = data.draw(st.booleans())
    zero_rows = data.draw(st.booleans())
    pred = data.draw(st.sampled_from(["none", "ns", "mask", "all_masked"]))
    path = data.draw(st.sampled_from(["scan", "auto", "mma"]))
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, D)).astype(np.float32)
    if dup and n > 4:
        x[n // 2:] = x[: n - n // 2]          # exact duplicates -> exact score ties
    if zero_rows:
        x[rng.integers(0, n, max(1, n // 10))] = 0.0
    ns = rng.integers(1, 4, n).astype(np.uint32)
    q = rng.standard_normal((B, D)).astype(np.float32)
    if data.draw(st.booleans()):
        q[0] = 0.0                             # zero query: every score ties at 0 -> pure row order
    idx = native.NativeIndex(D, metric=metric)
    idx.upsert(x, ns_codes=ns)
    stored, scale, _, _ = idx.get_rows(np.arange(n))
    est, esc = bf.normalise_store(x, metric)
    assert np.array_equal(stored, est) and np.array_equal(scale, esc)
    kw, elig = {}, None
    if pred == "ns":
        kw, elig = {"ns_code": 2}, ns == 2
    elif pred in ("mask", "all_masked"):
        bits = rng.random(n) < (0.3 if pred == "mask" else 0.0)
        kw = {"mask": np.packbits(np.concatenate([bits, np.zeros((-n) % 32, bool)]), bitorder="little").view(np.uint32)}
        elig = bits
    use_mma = path == "mma"
    p = {"scan": native.PATH_SCAN, "auto": native.PATH_AUTO, "mma": native.PATH_MMA if use_mma else native.PATH_SCAN}[path]
    s, r, c = idx.query(q, k, path=p, **kw)
    took_mma = idx.stats().last_path == native.PATH_MMA
    scale_mag = max(1.0, float(np.abs(x).max()) ** 2 * D) if metric == bf.METRIC_IP else 1.0
    if took_mma:
        if D <= 128:
            # rows of up to 128 padded elements take the hi/lo query split: fp32-accurate (~2^-17 relative)
            check_topk(s, r, c, oracle_full_scores(stored, scale, q, metric), elig, k, tol=2e-5 * scale_mag)
        else:
            qb = bf.bf16_bits_to_f32(bf.f32_to_bf16_bits(bf.normalise_query(q, metric)))
            full = bf.scores_fp32(stored, scale, qb)
            check_topk(s, r, c, full, elig, k, tol=4e-6 * scale_mag)
        # the north-star tolerance, at EVERY dimension the tensor path is routed to
        check_topk(s, r, c, oracle_full_scores(stored, scale, q, metric), elig, k, tol=1e-3 * scale_mag)
    else:
        check_topk(s, r, c, oracle_full_scores(stored, scale, q, metric), elig, k, tol=4e-6 * scale_mag)
    idx.close()


def test_concurrent_upserts_and_queries(native):
    """Writers append and overwrite while readers query: every answer must be exact for SOME committed
    prefix (rows never go backwards, scores of returned rows are real, no torn state crashes)."""
    rng = np.random.default_rng(0)
    D, k = 384, 5
    base = rng.standard_normal((20000, D)).astype(np.float32)
    idx = native.NativeIndex(D, capacity_rows=0)
    idx.upsert(base[:2000])
    q = base[100:101].copy()           # row 100 is its own nearest neighbour from the start
    errors, stop = [], threading.Event()

    def reader():
        last_count = 0
        while not stop.is_set():
            try:
                s, r, c = idx.query(q, k)
                n_now = len(idx)
                assert c[0] == k and r[0, 0] == 100 and abs(s[0, 0] - 1.0) < 2e-3
                assert (r[0] < n_now).all() and (np.diff(s[0]) <= 0).all()
                assert n_now >= last_count
                last_count = n_now
                s2, r2, c2 = idx.query(base[:32], k)   # tensor path concurrently
                assert (c2 == k).all() and (r2[:, 0] == np.arange(32)).all()
            except Exception as exc:  # noqa: BLE001
                errors.append(repr(exc))
                return

    def appender():
        for lo in range(2000, 20000, 500):
            idx.upsert(base[lo:lo + 500])

    def overwriter():
        orng = np.random.default_rng(1)
        for _ in range(60):
            rows = orng.integers(1000, 2000, 8)
            idx.upsert(base[rows[:1]] * 2.0, rows=rows[:1])   # same direction, new bits: cosine unchanged

    readers = [threading.Thread(target=reader) for _ in range(3)]
    writers = [threading.Thread(target=appender), threading.Thread(target=overwriter)]
    [t.start() for t in readers + writers]
    [t.join() for t in writers]
    stop.set()
    [t.join() for t in readers]
    assert not errors, errors[:3]
    assert len(idx) == 20000
    stored, scale, _, _ = idx.get_rows(np.arange(20000))
    s, r, c = idx.query(q, k, path=native.PATH_SCAN)
    check_topk(s, r, c, oracle_full_scores(stored, scale, q), None, k, tol=2e-6)
    idx.close()
