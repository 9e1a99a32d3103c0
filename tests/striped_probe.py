"""One process, G GPUs: throughput of a striped collection (native.StripedIndex -> cv_group_query, host buffers in and
out -- the call Collection.query makes) against the same rows on one device.  The reference's service is one process
with one client (codd_lib/codd_lib/client/provider/promql_module.py:59-67); this is what CODDVEC_DEVICES buys it.

    python tests/striped_probe.py            # PROBE_ROWS (default 8M), PROBE_DEVICES (default: all visible)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codd_query_engine_b200 import native  # noqa: E402


def timed(fn, min_s=0.5):
    for _ in range(5):
        fn()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < min_s:
        fn()
        n += 1
    return (time.perf_counter() - t0) / n


def main():
    rows = int(os.environ.get("PROBE_ROWS", 8_000_000))
    D, k = 384, 10
    G = native.device_count()
    devices = [int(d) for d in os.environ.get("PROBE_DEVICES", ",".join(str(g) for g in range(G))).split(",")]
    rng = np.random.default_rng(0)
    chunk = rng.standard_normal((500_000, D), dtype=np.float32)
    one = native.NativeIndex(D, capacity_rows=rows)
    many = native.StripedIndex(D, devices=devices, capacity_rows=rows, max_queries=1024, max_k=64)
    for lo in range(0, rows, chunk.shape[0]):
        m = min(chunk.shape[0], rows - lo)
        x = chunk[:m] + np.float32(lo % 97) * np.float32(1e-3)  # (distinct rows per chunk, cheaply)
        one.upsert(x)
        many.upsert(x)
    out = {"probe": "striped collection in ONE process (cv_group_query, host buffers) vs one device", "rows": rows,
           "dim": D, "devices": devices}
    for B in (1, 64, 1024):
        q = rng.standard_normal((B, D), dtype=np.float32)
        s1, r1, c1 = one.query(q, k)
        s2, r2, c2 = many.query(q, k)
        same = bool(np.array_equal(r1, r2) and np.array_equal(s1, s2))
        t1 = timed(lambda: one.query(q, k))
        t2 = timed(lambda: many.query(q, k))
        out[f"B{B}"] = {"one_device_ms": round(t1 * 1e3, 4), "striped_ms": round(t2 * 1e3, 4),
                        "one_device_qps": round(B / t1), "striped_qps": round(B / t2), "speedup": round(t1 / t2, 2),
                        "identical": same}
    print(json.dumps(out))
    many.close()
    one.close()


if __name__ == "__main__":
    main()
