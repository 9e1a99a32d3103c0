"""bench.py's JSON contract, checked on the CPU: the reference arm (`--impl reference`, the oracle's C port on the
host cores) must print ONE line with the driver's keys, only rank 0 may print, and our arm must refuse to run
without a B200 (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                          timeout=300)


def test_reference_arm_prints_one_contract_line():
    out = run_bench("--impl", "reference", "--rows", "200000", "--cpu-sample-rows", "50000", "--steps", "2", "--warmup", "1")
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "queries/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "bf16"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    out = run_bench("--impl", "reference", "--gpus", "2", "--rows", "100000", "--cpu-sample-rows", "20000", "--steps", "1",
                    "--warmup", "1", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_our_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        return  # on the GPU box the arm runs; the refusal is what is checked here
    out = run_bench("--rows", "1000", "--steps", "1", "--warmup", "1")
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_committed_traffic_stamps_match_the_kernel_sources():
    """profiles/traffic.json is only reported while its stamps equal the hashes of the kernel sources in the tree
    (bench.traffic_for).  A kernel edit without a fresh `ncu --set full` capture must be noticed here, not by a
    `"traffic": null` in the driver's bench line: re-capture (tests/gpu_round2_final5.sh) and re-stamp
    (profiles/make_traffic.py)."""
    import json
    import os

    import bench

    doc = json.load(open(os.path.join(bench.ROOT, "profiles", "traffic.json")))
    for kernel in ("k_scan_topk", "k_mma_topk"):
        assert doc[kernel]["kernel_source_hash"] == bench.kernel_source_hash(kernel), \
            f"{kernel}: sources changed since the ncu capture in {doc[kernel]['source']}"
    got, note = bench.traffic_for("k_scan_topk", 100_000_000, 384, 1)
    assert got is not None and abs(got / (100_000_000 * 772) - 1) < 0.01, note          # traffic == algorithmic bytes
    got, note = bench.traffic_for("k_mma_topk", 100_000_000, 384, 1024)
    assert got is not None and abs(got / (100_000_000 * 772) - 1) < 0.01, note
    assert bench.traffic_for("k_mma_topk", 100_000_000, 384, 64)[0] is None             # no capture at that batch: never a guess
