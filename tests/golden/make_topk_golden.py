"""Generate tests/golden/topk_small.json: small seeded input/output vectors of the oracle
(oracle/brute_force.py).  PARITY UNPINNED at the numeric level (the reference has no golden vectors
for this path and chromadb cannot run here); this fixture freezes the oracle's own answers so that
oracle, C twin and CUDA path are all held to the same committed numbers.

    python tests/golden/make_topk_golden.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import brute_force as bf  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case(seed, n, D, B, k, metric=0, dup=False, mask_p=None, ns_code=bf.ANY_CODE, row_offset=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, D)).astype(np.float32)
    if dup:
        x[n // 2:] = x[: n - n // 2]
    q = rng.standard_normal((B, D)).astype(np.float32)
    ns = rng.integers(0, 3, n).astype(np.uint32) if ns_code != bf.ANY_CODE else None
    mask = None
    if mask_p is not None:
        bits = rng.random(n) < mask_p
        mask = np.packbits(np.concatenate([bits, np.zeros((-n) % 32, bool)]), bitorder="little").view(np.uint32)
    stored, scale = bf.normalise_store(x, metric)
    s, r, c = bf.query(stored, scale, q, k, metric=metric, mask_words=mask, ns=ns, ns_code=ns_code, row_offset=row_offset)
    return {"seed": seed, "metric": metric, "k": k, "row_offset": row_offset, "ns_code": ns_code,
            "x": x.tolist(), "q": q.tolist(), "ns": None if ns is None else ns.tolist(),
            "mask_words": None if mask is None else mask.tolist(),
            "stored_bits": stored.tolist(), "scale_bits": scale.view(np.uint32).tolist(),
            "scores": s.tolist(), "rows": r.tolist(), "counts": c.tolist()}


def main():
    cases = [
        case(1, 64, 16, 2, 5),
        case(2, 50, 24, 1, 10, dup=True),
        case(3, 40, 8, 3, 64),                 # k > n
        case(4, 96, 32, 2, 8, mask_p=0.3),
        case(5, 96, 32, 2, 8, ns_code=1, row_offset=1000),
        case(6, 33, 40, 1, 4, metric=1),
    ]
    path = os.path.join(HERE, "topk_small.json")
    json.dump({"generator": "tests/golden/make_topk_golden.py", "cases": cases}, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
