"""End-to-end latency of one interactive query (host buffers in and out through cv_query), the
reference's own usage pattern (one query per call, store.py:314-316), by collection size."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from codd_query_engine_b200 import native  # noqa: E402


def main():
    D, k = 384, 10
    rng = np.random.default_rng(0)
    for rows in (1_000, 10_000, 100_000, 1_000_000):
        x = rng.standard_normal((rows, D), dtype=np.float32)
        idx = native.NativeIndex(D, capacity_rows=rows)
        idx.upsert(x)
        for B in (1, 4):
            q = rng.standard_normal((B, D), dtype=np.float32)
            for _ in range(200):
                idx.query(q, k)
            lat = []
            for _ in range(2000):
                t0 = time.perf_counter()
                idx.query(q, k)
                lat.append(time.perf_counter() - t0)
            lat = np.array(lat) * 1e6
            print(f"rows={rows:8d} B={B} cv_query e2e: p50 {np.percentile(lat, 50):7.1f} us  p99 {np.percentile(lat, 99):7.1f} us  "
                  f"mean {lat.mean():7.1f} us", flush=True)
        idx.close()


if __name__ == "__main__":
    main()
