"""Device-timing probe of the predicate scan (K2 PRED): selectivity x k x dimension, GB/s over the
selected bytes (rows read + the namespace column), beside the unfiltered scan of the same corpus."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from codd_query_engine_b200 import native  # noqa: E402


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    rows = int(os.environ.get("PROBE_ROWS", 10_000_000))
    dims = [int(d) for d in os.environ.get("PROBE_DIMS", "768,384").split(",")]
    for D in dims:
        idx = native.NativeIndex(D, capacity_rows=rows, max_rows=rows)
        g = torch.Generator(device="cuda").manual_seed(0)
        counts = {1: 0, 2: 0, 3: 0}
        for lo in range(0, rows, 500_000):
            m = min(500_000, rows - lo)
            x = torch.randn((m, D), device="cuda", generator=g)
            u = torch.rand((m,), device="cuda", generator=g)
            ns = torch.full((m,), 4, dtype=torch.int32, device="cuda")
            ns[u < 0.61] = 3   # 1 %
            ns[u < 0.60] = 2   # 10 %
            ns[u < 0.50] = 1   # 50 %
            for c in counts:
                counts[c] += int((ns == c).sum())
            idx.upsert_device(x.data_ptr(), m, ns_ptr=ns.contiguous().data_ptr())
        del x
        s = torch.cuda.current_stream().cuda_stream
        q = torch.randn((1, D), device="cuda", generator=g)
        dp = idx.stats().dim_padded
        for k in (10, 50):
            out = torch.zeros((1, k), dtype=torch.int64, device="cuda")
            ms = timed(lambda: idx.query_device(q.data_ptr(), 1, k, out.data_ptr(), stream=s), iters=10)
            print(f"D={D} rows={rows} k={k:3d} unfiltered        : {ms:8.4f} ms  {rows * (2 * dp + 4) / ms / 1e6:7.0f} GB/s", flush=True)
            for code in (1, 2, 3):
                ms = timed(lambda: idx.query_device(q.data_ptr(), 1, k, out.data_ptr(), ns_code=code, stream=s))
                sel = counts[code]
                b = sel * (2 * dp + 4) + rows * 4
                print(f"D={D} rows={rows} k={k:3d} ns={code} sel={sel / rows:6.3f}: {ms:8.4f} ms  {b / ms / 1e6:7.0f} GB/s over selected bytes "
                      f"(ideal {b / 6550.4e6:6.4f} ms)", flush=True)
        idx.close()


if __name__ == "__main__":
    main()
