#!/usr/bin/env python
"""Turn `ncu --set full` reports (gpurun_out/*.ncu-rep) into the JSON summaries committed under profiles/.
    python profiles/summarise_ncu.py OUT.json LABEL=report.ncu-rep [LABEL=report.ncu-rep ...]
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
]


def summarise(report):
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        out.append({m: [r[header.index(m)], units[header.index(m)]] for m in METRICS if m in header})
    return out


def main():
    dest, result = sys.argv[1], {}
    for spec in sys.argv[2:]:
        label, report = spec.split("=", 1)
        result[label] = summarise(report)
    with open(dest, "w") as f:
        json.dump(result, f, indent=1)
    for label, launches in result.items():
        for l in launches:
            print(label, "|", l["gpu__time_duration.sum"], l["dram__bytes_read.sum"], l.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                  l["launch__registers_per_thread"], l["smsp__inst_executed.sum"])


if __name__ == "__main__":
    main()
