"""Quick device-timing probe (not the bench): ms, GB/s and TFLOP/s per path and batch size."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from codd_query_engine_b200 import native  # noqa: E402


def main():
    rows = int(os.environ.get("PROBE_ROWS", 1_000_000))
    D = int(os.environ.get("PROBE_DIM", 384))
    batches = [int(b) for b in os.environ.get("PROBE_BATCHES", "1,2,4,8,16,64,128,256,512,1024").split(",")]
    paths = os.environ.get("PROBE_PATHS", "scan,mma").split(",")
    k = int(os.environ.get("PROBE_K", 10))
    idx = native.NativeIndex(D, capacity_rows=rows)
    g = torch.Generator(device="cuda").manual_seed(0)
    for lo in range(0, rows, 500_000):
        m = min(500_000, rows - lo)
        x = torch.randn((m, D), device="cuda", generator=g)
        idx.upsert_device(x.data_ptr(), m)
    del x
    s = torch.cuda.current_stream().cuda_stream
    for B in batches:
        q = torch.randn((B, D), device="cuda", generator=g)
        out = torch.zeros((B, k), dtype=torch.int64, device="cuda")
        for pname in paths:
            if pname == "scan" and B > 16:
                continue
            path = {"scan": native.PATH_SCAN, "mma": native.PATH_MMA, "mma1": native.PATH_MMA_1CTA,
                    "mma2": native.PATH_MMA_2CTA}[pname]
            fn = lambda: idx.query_device(q.data_ptr(), B, k, out.data_ptr(), path=path, stream=s)  # noqa: E731
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            iters = 20 if B <= 256 else 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            gb = rows * (2 * D + 4) / 1e9
            tf = 2.0 * B * rows * D / 1e12
            print(f"D={D} rows={rows} k={k} B={B:5d} {pname:4s}: {ms:9.4f} ms  {B / ms * 1e3:12.0f} qps  "
                  f"{gb / ms * 1e3:7.0f} GB/s/pass  {tf / ms * 1e3:8.1f} TFLOP/s")
    idx.close()


if __name__ == "__main__":
    main()
