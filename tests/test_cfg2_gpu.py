"""BASELINE.json configs[1] at its stated size: 1M x 384-d bf16 corpus, query batch 1..256, top-10, one B200 --
EXACT comparison with the fp32 oracle over every row (numpy scores the whole 1M x 384 x 256 problem in seconds).

Three corpora: iid Gaussian rows, a clustered corpus (256 centroids, sigma 0.3: well
separated top-k) and a duplicate-rows corpus (every row appears 4 times: exact ties, broken by row id).  Every
path that can serve the batch is checked: K2 scan (fp32 query: 2e-6, exact row order), K3 on single CTAs and on
CTA pairs (bf16 query: 3e-6 against the oracle evaluated with the rounded query, 1e-3 against the fp32 oracle,
row sets identical modulo near-ties inside the tolerance).
"""
import numpy as np
import pytest

from oracle import brute_force as bf
from tests.parity import check_topk

pytestmark = pytest.mark.gpu

N, D, K = 1_000_000, 384, 10
TOL = 1e-3


@pytest.fixture(scope="module")
def native(native_lib):
    from codd_query_engine_b200 import native as nat

    return nat


def corpus(kind: str) -> np.ndarray:
    rng = np.random.default_rng({"iid": 0, "clustered": 1, "duplicates": 2}[kind])
    if kind == "iid":
        return rng.standard_normal((N, D), dtype=np.float32)
    if kind == "clustered":
        centroids = rng.standard_normal((256, D), dtype=np.float32)
        which = rng.integers(0, 256, N)
        return centroids[which] + np.float32(0.3) * rng.standard_normal((N, D), dtype=np.float32)
    base = rng.standard_normal((N // 4, D), dtype=np.float32)
    return np.concatenate([base, base, base, base])  # row r == row r + N/4 == ...: exact ties across the scan


def queries(kind: str, x: np.ndarray, B: int) -> np.ndarray:
    rng = np.random.default_rng(100 + B)
    if kind == "iid":
        return rng.standard_normal((B, D), dtype=np.float32)
    # near stored rows: the top-k are real neighbours (clustered) / 4-way exact ties (duplicates)
    picks = rng.integers(0, x.shape[0], B)
    return x[picks] + np.float32(0.05) * rng.standard_normal((B, D), dtype=np.float32)


@pytest.fixture(scope="module", params=["iid", "clustered", "duplicates"])
def loaded(request, native):
    kind = request.param
    x = corpus(kind)
    idx = native.NativeIndex(D, capacity_rows=N, max_rows=N)
    for lo in range(0, N, 250_000):
        idx.upsert(x[lo:lo + 250_000])
    stored, scale, _, _ = idx.get_rows(np.arange(N))
    e32 = bf.bf16_bits_to_f32(stored)
    del stored
    yield kind, x, idx, e32, scale
    idx.close()


def full_scores(e32, scale, qhat):
    s = (e32 @ np.ascontiguousarray(qhat, dtype=np.float32).T).T
    return (s * scale[None, :]).astype(np.float32)


@pytest.mark.parametrize("B", [1, 2, 4, 16, 256])
def test_cfg2_exact(native, loaded, B):
    kind, x, idx, e32, scale = loaded
    q = queries(kind, x, B)
    qhat = bf.normalise_query(q)
    want32 = full_scores(e32, scale, qhat)
    want16 = full_scores(e32, scale, bf.bf16_bits_to_f32(bf.f32_to_bf16_bits(qhat)))
    # AUTO: whatever the library picks must meet the acceptance bar
    s, r, c = idx.query(q, K)
    check_topk(s, r, c, want32, None, K, tol=TOL)
    if B <= 16:  # K2: fp32 query (B = 16 is four passes of four queries)
        s, r, c = idx.query(q, K, path=native.PATH_SCAN)
        check_topk(s, r, c, want32, None, K, tol=2e-6)
    for path in (native.PATH_MMA_1CTA, native.PATH_MMA_2CTA):
        s, r, c = idx.query(q, K, path=path)
        assert idx.stats().last_path == native.PATH_MMA
        check_topk(s, r, c, want16, None, K, tol=3e-6)   # exactly what the tensor path computes ...
        check_topk(s, r, c, want32, None, K, tol=TOL)    # ... and the north-star bar against the fp32 oracle
    if kind == "duplicates":  # 4-way exact ties: ascending row id within equal scores, on every path
        for path in (native.PATH_SCAN if B <= 16 else native.PATH_MMA, native.PATH_MMA_1CTA, native.PATH_MMA_2CTA):
            s, r, c = idx.query(q, K, path=path)
            for b in range(B):
                groups = {}
                for sc, row in zip(s[b], r[b]):
                    groups.setdefault(float(sc), []).append(int(row))
                for rows in groups.values():
                    base = rows[0] % (N // 4)
                    # copies of ONE row, the lowest row ids of the family first (also where k cuts the family)
                    assert rows == [base + i * (N // 4) for i in range(len(rows))], rows


def test_cfg2_split_precision_is_fp32_accurate(native, loaded):
    """CV_PRECISION_SPLIT (hi/lo query tiles, two MMAs per K step) at 384-d: scores within 2e-5 of the fp32 oracle
    and the exact row order of the fp32 scan wherever the gap to the next score exceeds that."""
    kind, x, idx, e32, scale = loaded
    q = queries(kind, x, 64)
    want32 = full_scores(e32, scale, bf.normalise_query(q))
    for path in (native.PATH_MMA_1CTA, native.PATH_MMA_2CTA):
        s, r, c = idx.query(q, K, path=path, precision=native.PRECISION_SPLIT)
        check_topk(s, r, c, want32, None, K, tol=2e-5)
