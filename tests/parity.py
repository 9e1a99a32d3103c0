"""Parity checker: CUDA results vs the exact fp32 oracle on the same stored embeddings.

Acceptance (BASELINE.json north_star / SURVEY.md section 8c): every returned score within `tol`
of the oracle's score for that row; returned row set identical to the oracle's except where scores
are within `tol` of the oracle's k-th score (near-ties); output ordered by (score desc, row asc).
"""
import numpy as np

from oracle import brute_force as bf


def oracle_full_scores(stored, scale, q, metric=bf.METRIC_COSINE):
    qhat = bf.normalise_query(q, metric)
    return bf.scores_fp32(stored, scale, qhat)  # [B][n]


def check_topk(got_scores, got_rows, got_counts, full_scores, eligible, k, tol, row_offset=0, exact_rows=False):
    """full_scores: oracle [B][n] scores; eligible: bool [n] or None."""
    B, n = full_scores.shape
    for b in range(B):
        sc = full_scores[b].astype(np.float32)
        elig = np.ones(n, dtype=bool) if eligible is None else eligible
        n_el = int(elig.sum())
        want = min(k, n_el)
        assert int(got_counts[b]) == want, f"query {b}: count {got_counts[b]} != {want}"
        rows = got_rows[b, :want].astype(np.int64) - row_offset
        assert (got_rows[b, want:] == -1).all(), f"query {b}: padding rows must be -1"
        if want == 0:
            continue
        assert len(set(rows.tolist())) == want, f"query {b}: duplicate rows {rows}"
        assert ((rows >= 0) & (rows < n)).all()
        assert elig[rows].all(), f"query {b}: ineligible row returned"
        gs = got_scores[b, :want]
        # every score within tol of the oracle's score for that row
        err = np.abs(gs - sc[rows])
        assert err.max() <= tol, f"query {b}: score error {err.max()} > {tol}"
        # ordering: score desc, ties by row asc
        keys = bf.pack_keys(gs, rows + row_offset)
        assert (keys[:-1] > keys[1:]).all(), f"query {b}: not ordered by (score desc, row asc)"
        # set equality modulo near-ties around the k-th oracle score
        okeys = bf.pack_keys(np.where(elig, sc, -np.inf), np.arange(n))
        top = np.argpartition(okeys, n - want)[n - want:]  # the `want` largest keys (ineligible rows sit at -inf) ...
        order = top[np.argsort(okeys[top])[::-1]]           # ... in (score desc, row asc) order
        if exact_rows:
            assert rows.tolist() == order.tolist(), f"query {b}: rows {rows} != oracle {order}"
        kth = sc[order[-1]]
        must = set(np.nonzero(elig & (sc > kth + tol))[0].tolist())
        assert must.issubset(set(rows.tolist())), f"query {b}: missing rows {must - set(rows.tolist())}"
        assert (sc[rows] >= kth - tol).all(), f"query {b}: returned a row below the k-th score"
