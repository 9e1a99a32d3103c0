"""TEST INFRASTRUCTURE ONLY -- exact fp32 brute-force cosine / inner-product top-k.

CPU restatement (numpy) of what happens behind the reference's vector-DB
boundary for the ``search_relevant_metrics`` path:

* call sites being restated:
    codd_dal/metrics/metrics_semantic_metadata_store.py:60-69   (collection, ``hnsw:space=cosine``)
    codd_dal/metrics/metrics_semantic_metadata_store.py:236-238 (``collection.upsert``)
    codd_dal/metrics/metrics_semantic_metadata_store.py:314-316 (``collection.query``)
    codd_dal/metrics/metrics_semantic_metadata_store.py:336     (``similarity = 1 - distance``)
* the algorithm itself lives in the un-vendored third-party dependency
  ``chromadb`` (>=0.4.0 / >=1.4.0, bundled hnswlib).  Its published cosine
  space is: L2-normalise vectors at insert and at query time, distance =
  ``1 - sum_i a_i * b_i`` in fp32, results ascending by distance.  hnswlib's
  graph walk is *approximate*; this oracle is the exact answer that walk
  approximates (BASELINE.json north_star: "exact fp32 brute-force top-k").

PARITY UNPINNED: no golden vectors or known-answer distances exist in the
reference for this path and chromadb cannot be installed here, so this oracle
is definitional (north_star), pinned only by its own C twin (oracle/brute_force.c),
by hand-computed known answers in tests/test_oracle.py and by the committed
fixtures under tests/golden/ that were generated from it.

Storage format: rows are stored as bf16 of the
L2-normalised fp32 input; ``scale = 1/||stored row||`` (fp32) makes the cosine
exact over the stored values; queries are normalised in fp32 and NOT rounded.

    stored_i = bf16_rn(x_i / ||x_i||)         scale_i = 1 / ||stored_i||
    qhat     = q / ||q||
    score_i  = (sum_d stored_i[d] * qhat[d]) * scale_i
    top-k    = k largest score among eligible rows, ties -> smaller row id
    distance = 1 - score

Norms are accumulated in float64 and rounded once to fp32 so that the GPU
path (which does the same) is bit-identical in what it stores.
"""

from __future__ import annotations

import numpy as np

METRIC_COSINE = 0
METRIC_IP = 1
ANY_CODE = 0xFFFFFFFF  # "no predicate" namespace / label code
EMPTY_KEY = np.uint64(0)


# --------------------------------------------------------------------------- bf16
def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 bit pattern (uint16). NaN -> 0x7FFF."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    lsb = (u >> np.uint32(16)) & np.uint32(1)
    rounded = ((u + np.uint32(0x7FFF) + lsb) >> np.uint32(16)).astype(np.uint16)
    nan = np.isnan(x)
    if nan.any():
        rounded = rounded.copy()
        rounded[nan] = np.uint16(0x7FFF)
    return rounded


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(b, dtype=np.uint16)
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


# --------------------------------------------------------------------------- K1 twin
def normalise_store(x: np.ndarray, metric: int = METRIC_COSINE):
    """fp32 [n][D] -> (stored bf16 bits uint16 [n][D], scale fp32 [n]).

    Restates what `collection.upsert` does to a vector in cosine space
    (codd_dal/metrics/metrics_semantic_metadata_store.py:236-238 with `hnsw:space=cosine`, :64;
    hnswlib normalises at insert).

    cosine: stored = bf16(x / ||x||), scale = 1/||stored||  (0 for an all-zero row)
    ip:     stored = bf16(x),         scale = 1
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim != 2:
        raise ValueError("embeddings must be [n][D]")
    n = x.shape[0]
    if metric == METRIC_IP:
        return f32_to_bf16_bits(x), np.ones(n, dtype=np.float32)
    ss = np.einsum("ij,ij->i", x.astype(np.float64), x.astype(np.float64))
    norm = np.sqrt(ss).astype(np.float32)
    safe = np.where(norm > 0, norm, np.float32(1)).astype(np.float32)
    unit = (x / safe[:, None]).astype(np.float32)
    stored = f32_to_bf16_bits(unit)
    s64 = bf16_bits_to_f32(stored).astype(np.float64)
    ss2 = np.einsum("ij,ij->i", s64, s64)
    with np.errstate(divide="ignore"):
        scale = np.where(ss2 > 0, 1.0 / np.sqrt(ss2), 0.0).astype(np.float32)
    return stored, scale


def normalise_query(q: np.ndarray, metric: int = METRIC_COSINE) -> np.ndarray:
    """fp32 [B][D] -> qhat fp32 [B][D] (cosine: q/||q||, zero query stays zero; ip: q).

    Query-side half of `collection.query` (codd_dal/metrics/metrics_semantic_metadata_store.py:314-316;
    hnswlib cosine space normalises the query as well)."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    if q.ndim == 1:
        q = q[None, :]
    if metric == METRIC_IP:
        return q.copy()
    ss = np.einsum("ij,ij->i", q.astype(np.float64), q.astype(np.float64))
    norm = np.sqrt(ss).astype(np.float32)
    safe = np.where(norm > 0, norm, np.float32(1)).astype(np.float32)
    return (q / safe[:, None]).astype(np.float32)


# --------------------------------------------------------------------------- keys
def pack_keys(scores: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """(score fp32, row u32) -> u64 key whose unsigned order is (score desc, row asc).

    key = ordered(score) << 32 | (0xFFFFFFFF - row);  NaN scores are treated as -inf,
    -0.0 as +0.0.  key 0 is reserved for "empty slot".
    """
    s = np.ascontiguousarray(scores, dtype=np.float32).copy()
    s[np.isnan(s)] = -np.inf
    s = s + np.float32(0.0)
    u = s.view(np.uint32)
    neg = (u >> np.uint32(31)).astype(bool)
    ordered = np.where(neg, ~u, u | np.uint32(0x80000000)).astype(np.uint64)
    r = np.asarray(rows).astype(np.uint64)
    return (ordered << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - r)


def unpack_keys(keys: np.ndarray):
    """u64 keys -> (scores fp32, rows int64 with -1 for empty slots)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    ordered = (keys >> np.uint64(32)).astype(np.uint32)
    neg = (ordered >> np.uint32(31)) == 0
    u = np.where(neg, ~ordered, ordered & np.uint32(0x7FFFFFFF)).astype(np.uint32)
    scores = u.view(np.float32).copy()
    rows = (np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF))).astype(np.int64)
    empty = keys == EMPTY_KEY
    rows[empty] = -1
    scores[empty] = np.float32(0)
    return scores, rows


# --------------------------------------------------------------------------- scoring
def eligible_rows(n, mask_words=None, ns=None, ns_code=ANY_CODE, label=None, label_code=ANY_CODE):
    """Boolean [n]: bitmask bit (LSB-first in each u32 word) AND namespace AND label predicate."""
    ok = np.ones(n, dtype=bool)
    if mask_words is not None:
        w = np.ascontiguousarray(mask_words, dtype=np.uint32)
        bits = np.unpackbits(w.view(np.uint8), bitorder="little")[:n].astype(bool)
        ok &= bits
    if ns_code != ANY_CODE:
        ok &= np.asarray(ns, dtype=np.uint32)[:n] == np.uint32(ns_code)
    if label_code != ANY_CODE:
        ok &= np.asarray(label, dtype=np.uint32)[:n] == np.uint32(label_code)
    return ok


def scores_fp32(stored_bits: np.ndarray, scale: np.ndarray, qhat: np.ndarray) -> np.ndarray:
    """[B][n] fp32 scores = (stored @ qhat^T) * scale, fp32 accumulate.

    score = 1 - distance of codd_dal/metrics/metrics_semantic_metadata_store.py:336 (`similarity_score`)."""
    e = bf16_bits_to_f32(stored_bits)
    s = (e @ np.ascontiguousarray(qhat, dtype=np.float32).T).T  # [B][n]
    return (s * np.asarray(scale, dtype=np.float32)[None, :]).astype(np.float32)


def topk_from_scores(scores: np.ndarray, k: int, eligible=None, row_offset: int = 0):
    """Per query: k best keys (desc).  Returns keys u64 [B][k], EMPTY_KEY padded.

    `n_results` semantics of codd_dal/metrics/metrics_semantic_metadata_store.py:314-316: ascending distance,
    fewer rows than asked -> what exists; ties (unspecified upstream) are broken by row id here."""
    scores = np.atleast_2d(scores)
    B, n = scores.shape
    out = np.zeros((B, k), dtype=np.uint64)
    rows = np.arange(n, dtype=np.uint64) + np.uint64(row_offset)
    for b in range(B):
        keys = pack_keys(scores[b], rows)
        if eligible is not None:
            keys = keys[eligible]
        m = min(k, keys.shape[0])
        if m == 0:
            continue
        if keys.shape[0] > 4 * m:
            keys = np.partition(keys, keys.shape[0] - m)[-m:]
        out[b, :m] = np.sort(keys)[::-1][:m]
    return out


def merge_keys(per_shard_keys: np.ndarray, k: int) -> np.ndarray:
    """K4 twin: [G][B][k'] keys -> [B][k] best keys (desc), EMPTY_KEY padded."""
    a = np.ascontiguousarray(per_shard_keys, dtype=np.uint64)
    G, B, kk = a.shape
    flat = np.transpose(a, (1, 0, 2)).reshape(B, G * kk)
    srt = np.sort(flat, axis=1)[:, ::-1]
    out = np.zeros((B, k), dtype=np.uint64)
    m = min(k, srt.shape[1])
    out[:, :m] = srt[:, :m]
    return out


def query(stored_bits, scale, q, k, metric=METRIC_COSINE, mask_words=None, ns=None,
          ns_code=ANY_CODE, label=None, label_code=ANY_CODE, row_offset=0, chunk=1 << 18):
    """Full oracle query: returns (scores fp32 [B][k], rows int64 [B][k] (-1 pad), counts [B]).

    Processes the corpus in chunks so that 1M+ rows do not need a [B][n] matrix.
    """
    stored_bits = np.asarray(stored_bits)
    n = stored_bits.shape[0]
    qhat = normalise_query(q, metric)
    B = qhat.shape[0]
    elig = eligible_rows(n, mask_words, ns, ns_code, label, label_code)
    elig = elig & ~np.isnan(np.asarray(scale, dtype=np.float32))  # scale = NaN marks a deleted row (cv_delete_rows)
    best = np.zeros((B, k), dtype=np.uint64)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        sc = scores_fp32(stored_bits[lo:hi], scale[lo:hi], qhat)
        part = topk_from_scores(sc, k, elig[lo:hi], row_offset=row_offset + lo)
        best = merge_keys(np.stack([best, part]), k)
    scores, rows = unpack_keys(best)
    counts = (best != EMPTY_KEY).sum(axis=1).astype(np.int32)
    return scores, rows, counts
