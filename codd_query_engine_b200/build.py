"""Build libcoddvec.so (hand-written sm_100a CUDA + C-ABI) in-tree with nvcc.

The translation units compile in parallel; objects are cached under build/ by source mtime.
`python -m codd_query_engine_b200.build [--force]`
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OUT = PKG / "libcoddvec.so"
OBJ_DIR = ROOT / "build" / "obj"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-diag-suppress", "177",
]
UNITS = ["coddvec.cu", "k2_inst_qb1.cu", "k2_inst_qb2.cu", "k2_inst_qb4.cu", "k3_mma.cu", "k3_inst_cg2.cu", "k3_inst_pilot.cu"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps_mtime() -> float:
    files = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.hpp")) + [ROOT / "include" / "coddvec.h", Path(__file__)]
    return max(f.stat().st_mtime for f in files)


VARIANTS = {  # name -> (object directory, extra nvcc flags, library name)
    "": ("obj", [], "libcoddvec.so"),
    "checked": ("obj_checked", ["-DCV_CHECKED"], "libcoddvec_checked.so"),
    "timeline": ("obj_timeline", ["-DCV_K3_TIMELINE"], "libcoddvec_timeline.so"),
}


def _compile(unit: str, force: bool, variant: str = "") -> Path:
    src = CSRC / unit
    obj_dir = OBJ_DIR.parent / VARIANTS[variant][0]
    obj_dir.mkdir(parents=True, exist_ok=True)
    obj = obj_dir / (src.stem + ".o")
    log = obj_dir / (src.stem + ".ptxas.log")
    newest = max(src.stat().st_mtime, _deps_mtime())
    if not force and obj.exists() and obj.stat().st_mtime >= newest:
        return obj
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    extra = os.environ.get("CODDVEC_NVCC_EXTRA", "").split()  # (-D switches of A/B experiment builds)
    cmd = [_nvcc(), *NVCC_FLAGS, *VARIANTS[variant][1], *extra, "-ccbin", "/usr/bin/g++", "-c", str(src), "-o", str(obj)]
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    log.write_text(proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed for {unit}:\n{proc.stderr[-4000:]}")
    return obj


def build(force: bool = False, verbose: bool = False, checked: bool = False, variant: str = "") -> Path:
    """variant "checked": libcoddvec_checked.so, the same sources with -DCV_CHECKED (device-side bounds assertions that
    trap) -- the assertion pass that stands in for compute-sanitizer.  variant "timeline": libcoddvec_timeline.so with
    -DCV_K3_TIMELINE (per-CTA phase stamps and a per-tile trace of K3, printed with CODDVEC_K3_TIMELINE=1 /
    CODDVEC_K3_TRACE=1).  Select either with CODDVEC_LIB; the product library carries neither."""
    variant = "checked" if checked else variant
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    OUT = PKG / VARIANTS[variant][2]
    units = [u for u in UNITS if (CSRC / u).exists()]
    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(lambda u: _compile(u, force, variant), units))
    if force or not OUT.exists() or any(o.stat().st_mtime > OUT.stat().st_mtime for o in objs):
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        cmd = [_nvcc(), "-shared", "-cudart", "static", "-ccbin", "/usr/bin/g++",
               "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(OUT), *map(str, objs)]
        proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if proc.returncode != 0:
            raise RuntimeError(f"link failed:\n{proc.stderr[-4000:]}")
    if verbose:
        print(f"built {OUT}")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True,
          variant="checked" if "--checked" in sys.argv else ("timeline" if "--timeline" in sys.argv else ""))
