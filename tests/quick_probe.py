"""Quick device-timing probe of the scan path (not the bench): GB/s at a few sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from codd_query_engine_b200 import native

def main():
    rows = int(os.environ.get("PROBE_ROWS", 4_000_000))
    for D in (384, 768):
        idx = native.NativeIndex(D, capacity_rows=rows)
        chunk = 500_000
        g = torch.Generator(device="cuda").manual_seed(0)
        for lo in range(0, rows, chunk):
            m = min(chunk, rows - lo)
            x = torch.randn((m, D), device="cuda", generator=g)
            idx.upsert_device(x.data_ptr(), m)
        torch.cuda.synchronize()
        for B in (1, 2, 4):
            if D == 768 and B == 4:
                continue
            q = torch.randn((B, D), device="cuda", generator=g)
            out = torch.zeros((B, 10), dtype=torch.int64, device="cuda")
            s = torch.cuda.current_stream().cuda_stream
            for _ in range(5):
                idx.query_device(q.data_ptr(), B, 10, out.data_ptr(), stream=s)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 50
            e0.record()
            for _ in range(iters):
                idx.query_device(q.data_ptr(), B, 10, out.data_ptr(), stream=s)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            gb = rows * (2 * D + 4) / 1e9
            print(f"D={D} rows={rows} B={B}: {ms:.4f} ms  {gb / ms * 1e3:.0f} GB/s  frac_of_6550={gb / ms * 1e3 / 6550.4:.3f}")
        idx.close()

if __name__ == "__main__":
    main()
