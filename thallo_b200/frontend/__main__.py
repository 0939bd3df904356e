"""Command-line front end:  python -m thallo_b200.frontend --energy FILE --kind K --dims a,b --out DIR

Invoked by Thallo_ProblemPlan (csrc/th_api.cpp run_frontend) when a problem was defined
from a file name, the way reference programs do (tests/*/main.cpp,
examples/shared/ThalloSolver.h:43-60).  Writes DIR/plan.desc and DIR/energy.cu.
`--query-ndims` writes DIR/ndims.txt (number of Dims() the energy declares) instead.
`--dims-on-stdin` does both in ONE process: it prints "ndims N" on stdout, then reads the N sizes (one line, comma or
space separated) from stdin and lowers -- the library cannot know how many entries of the caller's `dimensions` array to
read before the energy has been parsed.
"""
import argparse
import os
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--energy", required=True)
    ap.add_argument("--kind", default="gauss_newton")
    ap.add_argument("--dims", default="")
    ap.add_argument("--double", type=int, default=0)
    ap.add_argument("--out", required=True)
    ap.add_argument("--query-ndims", action="store_true")
    ap.add_argument("--lm-as-committed", action="store_true")
    ap.add_argument("--schedule", default="auto")
    ap.add_argument("--dims-on-stdin", action="store_true")
    a = ap.parse_args()
    import energies
    from thallo_b200.frontend import codegen, dsl
    define, mod = energies.define_for(a.energy)
    if define is None:
        sys.stderr.write("energy file '%s' does not exist and no transcription is registered under that name "
                         "(energies/__init__.py REGISTRY)\n" % a.energy)
        return 2
    if a.query_ndims or a.dims_on_stdin:
        L = dsl.SymbolicL([1] * 8)
        try:
            define(L)
        except Exception:
            pass
        if a.query_ndims:
            with open(os.path.join(a.out, "ndims.txt"), "w") as f:
                f.write(str(len(L.dims)))
            return 0
        sys.stdout.write("ndims %d\n" % len(L.dims))
        sys.stdout.flush()
        a.dims = sys.stdin.readline().replace(" ", ",")
    dims = [int(x) for x in a.dims.strip().split(",") if x]
    low = codegen.lower(define, dims, a.kind, mod, bool(a.double), a.schedule, a.lm_as_committed)
    with open(os.path.join(a.out, "plan.desc"), "w") as f:
        f.write(codegen.descriptor_text(low.desc))
    with open(os.path.join(a.out, "energy.cu"), "w") as f:
        f.write(low.source)
    return 0


if __name__ == "__main__":
    sys.exit(main())
