#!/usr/bin/env python
"""SASS evidence for the generated plans (run on the CPU box after __graft_entry__.build()): per kernel of every
cubin under build/cuda_check the instruction count and the mnemonics that prove how it moves data -- UTMALDG (TMA
tensor loads), SYNCS (mbarrier), LDS / STS (shared memory), LDG.E.128 / STG.E.128 (vectorised global accesses),
ATOM / RED (atomics), SHFL / MATCH / VOTE (warp primitives), MUFU (transcendentals) -- plus registers and shared
memory from cuobjdump -res-usage.   python scripts/sass_summary.py > profiles/<tag>_sass_summary.txt"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTMALDG", "SYNCS", "LDS", "STS", "LDG.E.128", "STG.E.128", "LDG", "STG", "ATOM", "RED", "SHFL", "MATCH", "VOTE", "MUFU", "FFMA", "DFMA"]


def main():
    print("# cuobjdump -sass / -res-usage of the nvcc -gencode arch=compute_100a,code=sm_100a cubins of build/cuda_check")
    for cubin in sorted(glob.glob(os.path.join(ROOT, "build", "cuda_check", "*.cubin"))):
        sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", cubin], capture_output=True, text=True).stdout
        usage = {}
        cur = None
        for ln in res.splitlines():
            m = re.search(r"Function (\w+):", ln)
            if m:
                cur = m.group(1)
            m = re.search(r"REG:(\d+).*SHARED:(\d+)", ln)
            if m and cur:
                usage[cur] = (int(m.group(1)), int(m.group(2)))
        print("\n== %s" % os.path.basename(cubin))
        kernels = collections.OrderedDict()
        cur = None
        for ln in sass.splitlines():
            m = re.search(r"Function : (\w+)", ln)
            if m:
                cur = m.group(1)
                kernels[cur] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if m and cur:
                op = m.group(1)
                kernels[cur]["_n"] += 1
                for k in KEYS:
                    if op == k or op.startswith(k + "."):
                        kernels[cur][k] += 1
                if op.startswith("LDG.E.128"):
                    kernels[cur]["LDG.E.128"] += 0
        for k, c in kernels.items():
            if c["_n"] < 40 and not k.startswith("th_pcg"):
                continue
            r = usage.get(k, ("?", "?"))
            hits = " ".join("%s=%d" % (key, c[key]) for key in KEYS if c[key])
            print("  %-24s insts %5d regs %4s smem %6s  %s" % (k, c["_n"], r[0], r[1], hits))


if __name__ == "__main__":
    main()
