#!/usr/bin/env python
"""Write profiles/traffic.json -- the per-launch DRAM bytes bench.py reports as `roofline.traffic` -- from an
`ncu --set full` summary (profiles/summarise_ncu.py output), stamped with the hash of that kernel's sources in the
working tree (bench.KERNEL_SOURCES).  bench.py only reports an entry whose stamp equals the hash of the sources it runs, so this must be run
on the tree the capture was taken from, with no kernel edits in between.

    python profiles/make_traffic.py SUMMARY.json  k_scan_topk=LABEL:rows:dim:batch  k_mma_topk=LABEL:rows:dim:batch
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def to_bytes(pair):
    value, unit = pair
    return float(value) * UNITS[unit]


def main():
    summary_path = sys.argv[1]
    summary = json.load(open(summary_path))
    out = {}
    for spec in sys.argv[2:]:
        kernel, rest = spec.split("=", 1)
        label, rows, dim, batch = rest.split(":")
        launch = summary[label][0]
        if kernel not in launch["Kernel Name"][0]:
            raise SystemExit(f"{label} is a capture of {launch['Kernel Name'][0]}, not of {kernel}")
        out[kernel] = {
            "bytes_per_launch": to_bytes(launch["dram__bytes_read.sum"]) + to_bytes(launch["dram__bytes_write.sum"]),
            "rows": int(rows), "dim": int(dim), "batch": int(batch),
            "kernel_name": launch["Kernel Name"][0],
            "kernel_source_hash": kernel_source_hash(kernel),
            "source": f"profiles/{os.path.basename(summary_path)} [{label}] (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)",
        }
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
