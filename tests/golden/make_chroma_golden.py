"""Pin the oracle against the REAL thing the day it is installable: probe for chromadb / hnswlib (on sys.path or
under baseline/_ref/, the directory reserved for a driver-provided reference install -- SURVEY.md section 8c) and,
when one imports, run the reference's own collection configuration
(codd_dal/metrics/metrics_semantic_metadata_store.py:60-69: cosine, M = 16, ef_construction = 200, ef_search = 100)
over a seeded corpus and write what it returns to tests/golden/chroma_topk.json:

  * chromadb:  EphemeralClient().get_or_create_collection(metadata=<hnsw keys>).upsert(embeddings=...) and
               .query(query_embeddings=..., n_results=k) -- the ids and DISTANCES the reference's store would see;
  * hnswlib :  Index(space="cosine") with the same parameters -- labels and distances of knn_query.

tests/test_oracle.py::test_oracle_matches_chroma_golden then holds oracle/brute_force.py to those numbers: every
distance chromadb reports must equal 1 - (the oracle's score for that row) within 1e-5 (fp32 vs fp32 of the
unrounded vectors; the oracle stores bf16, so the script also records the fp32 brute-force distances to separate
the two effects), and wherever the approximate index returned the exact top-k the ids must agree.

Nothing here runs in the product.  Today neither package is installable (no network, no wheel): the script says so
and exits 0 without touching the fixture, and the parity of the oracle stays "unpinned" (DESIGN.md section 2).

    python tests/golden/make_chroma_golden.py            # probe + (re)generate if possible
    python tests/golden/make_chroma_golden.py --probe    # only say what is available
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "baseline", "_ref")
OUT = os.path.join(HERE, "chroma_topk.json")
HNSW = {"hnsw:space": "cosine", "hnsw:construction_ef": 200, "hnsw:search_ef": 100, "hnsw:M": 16}


def probe():
    """-> {"chromadb": module or None, "hnswlib": module or None, "why": {...}}"""
    if os.path.isdir(REF) and REF not in sys.path:
        sys.path.insert(0, REF)
    found, why = {}, {}
    for name in ("chromadb", "hnswlib"):
        try:
            found[name] = __import__(name)
        except Exception as exc:  # noqa: BLE001 - any import failure means "not available here"
            found[name] = None
            why[name] = f"{type(exc).__name__}: {exc}"
    found["why"] = why
    return found


def corpus(seed=0, n=2000, d=384, nq=16):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[n // 2] = x[3]  # one exact duplicate: a tie
    q = np.concatenate([x[3:4], rng.standard_normal((nq - 1, d)).astype(np.float32)])
    return x, q


def run_chromadb(chromadb, x, q, k):
    client = chromadb.EphemeralClient()
    coll = client.get_or_create_collection(name="golden", metadata=dict(HNSW))
    ids = [f"r{i}" for i in range(len(x))]
    for lo in range(0, len(x), 1000):
        coll.upsert(ids=ids[lo:lo + 1000], embeddings=x[lo:lo + 1000].tolist())
    res = coll.query(query_embeddings=q.tolist(), n_results=k, include=["distances"])
    return {"version": getattr(chromadb, "__version__", "?"), "rows": [[int(i[1:]) for i in row] for row in res["ids"]],
            "distances": [[float(v) for v in row] for row in res["distances"]]}


def run_hnswlib(hnswlib, x, q, k):
    idx = hnswlib.Index(space="cosine", dim=x.shape[1])
    idx.init_index(max_elements=len(x), ef_construction=200, M=16, random_seed=100)
    idx.add_items(x, np.arange(len(x)), num_threads=1)
    idx.set_ef(100)
    labels, dist = idx.knn_query(q, k=k, num_threads=1)
    return {"version": getattr(hnswlib, "__version__", "?"), "rows": labels.astype(int).tolist(), "distances": dist.astype(float).tolist()}


def main():
    found = probe()
    avail = {k: (v is not None) for k, v in found.items() if k != "why"}
    print("probe:", json.dumps({"baseline/_ref exists": os.path.isdir(REF), **avail, "why": found["why"]}))
    if "--probe" in sys.argv or not any(avail.values()):
        if not any(avail.values()):
            print("neither chromadb nor hnswlib is importable: chroma_topk.json not written, oracle parity stays unpinned")
        return 0
    k = 10
    x, q = corpus()
    xn = x / np.linalg.norm(x, axis=1, keepdims=True)
    qn = q / np.linalg.norm(q, axis=1, keepdims=True)
    exact = 1.0 - (qn.astype(np.float32) @ xn.astype(np.float32).T)  # fp32 brute force over the UNROUNDED vectors
    doc = {"seed": 0, "n": len(x), "d": x.shape[1], "nq": len(q), "k": k, "hnsw": HNSW,
           "exact_fp32_rows": np.argsort(exact, axis=1, kind="stable")[:, :k].tolist(),
           "exact_fp32_distances": np.sort(exact, axis=1)[:, :k].astype(float).tolist()}
    if found["chromadb"] is not None:
        doc["chromadb"] = run_chromadb(found["chromadb"], x, q, k)
    if found["hnswlib"] is not None:
        doc["hnswlib"] = run_hnswlib(found["hnswlib"], x, q, k)
    with open(OUT, "w") as f:
        json.dump(doc, f)
    print("wrote", OUT, "with", [name for name in ("chromadb", "hnswlib") if name in doc])
    return 0


if __name__ == "__main__":
    sys.exit(main())
