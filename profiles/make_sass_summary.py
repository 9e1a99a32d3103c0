#!/usr/bin/env python
"""Opcode histogram per kernel of libcoddvec.so, from `cuobjdump -sass` -- the committed evidence that the hot
kernels are Blackwell-native (UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor loads, LDTM = tcgen05.ld, UTCBAR =
tcgen05.commit, SYNCS = mbarrier) rather than recompiled mma.sync code.

    python profiles/make_sass_summary.py > profiles/sass_summary.txt
"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "codd_query_engine_b200", "libcoddvec.so")
KEY = ("UTCHMMA", "UTMALDG", "UTMAPF", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR", "HMMA", "LDG", "LDGSTS", "STG", "ATOMG",
       "RED", "FFMA", "FMUL", "FMNMX", "FMNMX3", "DFMA", "SHFL", "LDS", "STS", "BAR", "MEMBAR", "ATOMS")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
            full = m.group(1) + m.group(2)
            if m.group(1) in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS"):
                kernels[name]["  " + full] += 1
    sha = hashlib.sha256(open(LIB, "rb").read()).hexdigest()[:16]
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"# cuobjdump -sass codd_query_engine_b200/libcoddvec.so  (sha256 {sha}, built from the tree at/after commit {head})")
    print("# per kernel: total SASS instructions, then the opcodes that identify the hardware path\n")
    for kname, ops in kernels.items():
        short = re.sub(r"\(.*", "", kname)
        total = sum(v for k, v in ops.items() if not k.startswith("  "))
        print(f"{short}: {total} instructions")
        for op in KEY:
            if ops.get(op):
                variants = ", ".join(f"{k.strip()} x{v}" for k, v in sorted(ops.items()) if k.startswith("  " + op))
                print(f"    {op:10s} {ops[op]:6d}" + (f"   [{variants}]" if variants else ""))
        print()


if __name__ == "__main__":
    sys.exit(main())
