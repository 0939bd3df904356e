#!/usr/bin/env python
"""SASS listings of the hot kernels at the configured sizes -> profiles/sass/ (nvcc cross-compiles sm_100a without a GPU;
the plans JIT-compile the same source with NVRTC).  Encoding columns are stripped to keep the listings readable."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import energies                                    # noqa: E402
from thallo_b200.frontend import codegen           # noqa: E402

CASES = [("image_warping", [2048, 2048], "levenberg_marquardt", {}, ["th_pcg_a", "th_pcg_b"]),
         ("arap_mesh_deformation", [4000000, 23984002], "gauss_newton", dict(schedule="gather"), ["th_gather_s0", "th_step3"]),
         ("shape_from_shading", [8192, 8192], "gauss_newton", {}, ["th_pcg_a"]),
         ("volumetric_mesh_deformation", [160, 160, 160], "gauss_newton", {}, ["th_pcg_a"]),
         ("bundle_adjustment", [10000, 5000000, 25000000], "levenberg_marquardt", {}, ["th_matj_g0", "th_gather_s0", "th_gather_s1"])]


def main(tag="r02"):
    out = os.path.join(ROOT, "profiles", "sass")
    os.makedirs(out, exist_ok=True)
    skel = os.path.join(ROOT, "thallo_b200", "csrc", "skeleton")
    for name, dims, kind, kw, funs in CASES:
        low = codegen.lower(energies.load(name), dims, kind, name, **kw)
        cu, cubin = "/tmp/sass_%s.cu" % name, "/tmp/sass_%s.cubin" % name
        open(cu, "w").write(low.source)
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-lineinfo", "-O3", "-I", skel, "-cubin", "-o", cubin, cu])
        for f in funs:
            txt = subprocess.run(["cuobjdump", "-sass", "-fun", f, cubin], capture_output=True, text=True).stdout
            lines = [l.split("/* 0x")[0].rstrip() for l in txt.splitlines() if not l.strip().startswith("/* 0x")]
            path = os.path.join(out, "%s_%s_%s.sass" % (tag, name, f))
            with open(path, "w") as fh:
                fh.write("// %s of energy %s %s %s (nvcc 12.9 -O3 -lineinfo, sm_100a)\n" % (f, name, dims, kind) + "\n".join(lines) + "\n")
            print(path, len(lines))


if __name__ == "__main__":
    main(*sys.argv[1:])
