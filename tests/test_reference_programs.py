"""The reference's own C++ test programs (every tests/*/main.cpp, 14 of them), UNMODIFIED, compiled
where they lie against this repository's include/Thallo.h and linked with libThallo.so (`make -C oracle ref`): the
drop-in check of the C ABI at the source and the link level.  Running them needs a GPU; that part is below, marked
`gpu`.  Nothing from the reference is copied: the energy files the programs ask for by name (`laplacian.t`) are
written into the working directory from the texts below, which restate the two energies."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref")

IMAGE_LAPLACIAN_T = """
-- fit to A plus forward differences in x and y (the revision of the energy that produced tests/minimal/gold.png)
local W,H = Dims("W","H")
Inputs { X = Unknown(float,{W,H},0), A = Array(float,{W,H},1) }
local x,y = W(),H()
local w_fit = .2
r = Residuals {
    fit = w_fit*(X(x,y) - A(x,y)),
    reg = { Select(InBounds(x+1,y), X(x,y) - X(x+1,y), 0), Select(InBounds(x,y+1), X(x,y) - X(x,y+1), 0) }
}
"""

GRAPH_LAPLACIAN_T = """
-- fit to A plus differences along the edges (v0, v1) of a graph
local N,E = Dims("N","E")
Inputs { X = Unknown(float,{N},0), A = Array(float,{N},1), v0 = Sparse({E},{N},2), v1 = Sparse({E},{N},3) }
local n,e = N(),E()
r = Residuals { fit = .5*(X(n) - A(n)), reg = X(v0(e)) - X(v1(e)) }
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_callers_compile_and_link_unmodified():
    from thallo_b200 import api
    api.build_library()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    programs = sorted(os.path.basename(os.path.dirname(p)) for p in __import__("glob").glob(os.path.join(REF, "tests", "*", "main.cpp")))
    assert len(programs) >= 14 and "minimal" in programs and "minimal_graph" in programs
    for name in ["ref_" + t for t in programs]:
        path = os.path.join(BIN, name)
        assert os.path.isfile(path) and os.access(path, os.X_OK)
        undefined = subprocess.run(["nm", "-D", "--undefined-only", path], capture_output=True, text=True).stdout
        used = set(re.findall(r"\bThallo_\w+", undefined))
        assert {"Thallo_NewState", "Thallo_ProblemDefine", "Thallo_ProblemPlan"} <= used and \
            ({"Thallo_ProblemSolve"} <= used or {"Thallo_ProblemInit", "Thallo_ProblemStep"} <= used), (name, used)
        ldd = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
        assert "libThallo.so" in ldd and "not found" not in ldd.split("libThallo.so")[1].splitlines()[0]


def _libc_uniform(n, seed=1):
    """What the programs fill their input with: glibc rand() / RAND_MAX from the default seed (or the one they set)."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(ctypes.c_uint(seed))
    return np.array([libc.rand() / 2147483647.0 for _ in range(n)], np.float64).astype(np.float32)


def _run(binary, tmp_path, energy_text):
    (tmp_path / "laplacian.t").write_text(energy_text)
    r = subprocess.run([os.path.join(BIN, binary)], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    m = re.search(r"minimal\w* ([-+0-9.eE]+|nan|inf)", r.stdout)
    assert m, r.stdout[-2000:]
    assert (tmp_path / "result.png").exists()
    return float(m.group(1))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isfile(os.path.join(BIN, "ref_minimal")), reason="oracle/_ref not built")
def test_reference_minimal_program_runs_against_this_library(tmp_path):
    import torch
    from thallo_b200.api import ThalloSolver
    cost = _run("ref_minimal", tmp_path, IMAGE_LAPLACIAN_T)
    A = _libc_uniform(512 * 512)
    dX, dA = torch.from_numpy(A.copy()).cuda(), torch.from_numpy(A.copy()).cuda()
    s = ThalloSolver([512, 512], "laplacian", "gauss_newton")
    want = s.solve([dX, dA])                      # the library's defaults, like the program: GN 10 x 10
    assert abs(cost - want) <= 1e-5 * abs(want)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isfile(os.path.join(BIN, "ref_minimal_graph")), reason="oracle/_ref not built")
def test_reference_minimal_graph_program_runs_against_this_library(tmp_path):
    import torch
    from thallo_b200.api import ThalloSolver
    cost = _run("ref_minimal_graph", tmp_path, GRAPH_LAPLACIAN_T)
    A = _libc_uniform(512)
    v0 = np.arange(511, dtype=np.int32)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    s = ThalloSolver([512, 511], "graph_laplacian", "gauss_newton")
    want = s.solve([dev(A.copy()), dev(A.copy()), dev(v0), dev(v0 + 1)])
    assert abs(cost - want) <= 1e-5 * abs(want)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_example_harness_compiles_against_this_header(tmp_path):
    """examples/shared/ThalloSolver.h -- the wrapper every reference example drives the C ABI through (NewState ->
    ProblemDefine -> ProblemPlan, SetSolverParameter*, Solve or Init / Step / CurrentCost, GetPerformanceSummary) --
    compiled as the examples compile it (nvcc) with this repository's Thallo.h in place of the reference's."""
    tu = tmp_path / "harness.cu"
    tu.write_text('extern "C" {\n#include "Thallo.h"\n}\n#include "ThalloSolver.h"\nint main() { return 0; }\n')
    r = subprocess.run(["nvcc", "-std=c++14", "-w", "-c", "-o", str(tmp_path / "harness.o"), "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(REF, "examples", "shared"), str(tu)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_minimal_links_statically_against_libthallo_a(tmp_path):
    """The reference ships a static libThallo.a (API/Makefile:8-12); so does this repository."""
    from thallo_b200 import api
    api.build_library()
    lib_a = os.path.join(os.path.dirname(api.LIB_PATH), "libThallo.a")
    out = tmp_path / "ref_minimal_static"
    r = subprocess.run(["g++", "-std=c++14", "-w", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                        os.path.join(REF, "tests", "minimal", "main.cpp"), "-o", str(out), lib_a, "-L/usr/local/cuda/lib64",
                        "-lnvrtc", "-lcudart", "-ldl", "-lrt", "-lpthread"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "libThallo" not in subprocess.run(["ldd", str(out)], capture_output=True, text=True).stdout


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_struct_layouts_equal_the_reference_header(tmp_path):
    """Thallo_NewState takes its parameters BY VALUE and Thallo_GetPerformanceSummary fills a caller-owned struct: sizes
    and field offsets of include/Thallo.h must equal those of the reference's API/release/include/Thallo.h."""
    probe = tmp_path / "probe.c"
    probe.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include HEADER
int main(void) {
    printf("init %zu %zu %zu %zu %zu %zu %zu\n", sizeof(Thallo_InitializationParameters),
           offsetof(Thallo_InitializationParameters, doublePrecision), offsetof(Thallo_InitializationParameters, verbosityLevel),
           offsetof(Thallo_InitializationParameters, timingLevel), offsetof(Thallo_InitializationParameters, threadsPerBlock),
           offsetof(Thallo_InitializationParameters, useAutoscheduler), offsetof(Thallo_InitializationParameters, cpuOnly));
    printf("entry %zu %zu %zu %zu %zu %zu\n", sizeof(Thallo_PerformanceEntry), offsetof(Thallo_PerformanceEntry, count),
           offsetof(Thallo_PerformanceEntry, minMS), offsetof(Thallo_PerformanceEntry, maxMS),
           offsetof(Thallo_PerformanceEntry, meanMS), offsetof(Thallo_PerformanceEntry, stddevMS));
    printf("summary %zu %zu %zu %zu %zu %zu\n", sizeof(Thallo_PerformanceSummary), offsetof(Thallo_PerformanceSummary, total),
           offsetof(Thallo_PerformanceSummary, nonlinearIteration), offsetof(Thallo_PerformanceSummary, nonlinearSetup),
           offsetof(Thallo_PerformanceSummary, linearSolve), offsetof(Thallo_PerformanceSummary, nonlinearResolve));
    return 0;
}
''')
    outs = []
    for tag, header in (("ours", os.path.join(ROOT, "include", "Thallo.h")), ("ref", os.path.join(REF, "API", "release", "include", "Thallo.h"))):
        exe = tmp_path / ("probe_" + tag)
        subprocess.check_call(["gcc", "-w", "-DHEADER=\"%s\"" % header, str(probe), "-o", str(exe)])
        outs.append(subprocess.run([str(exe)], capture_output=True, text=True).stdout)
    assert outs[0] == outs[1] and outs[0].startswith("init 24 0 4 8 12 16 20")
    # and the ctypes mirror used by the Python host side agrees with both
    import ctypes as C
    from thallo_b200 import api
    assert C.sizeof(api.InitializationParameters) == 24
    assert ("summary %d" % C.sizeof(api.PerformanceSummary)) in outs[0] and ("entry %d" % C.sizeof(api.PerformanceEntry)) in outs[0]


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isfile(os.path.join(BIN, "ref_create_delete_cycle")), reason="oracle/_ref not built")
def test_reference_create_delete_cycle_program_runs_against_this_library(tmp_path):
    """tests/create_delete_cycle/main.cpp:22-26: Thallo_ProblemPlan / Thallo_PlanFree ten times on the same problem, then a
    solve.  The front end is started once (posix_spawn, no shell) and the other ten plans come out of the in-process
    lowering cache; the run must finish well within the time ten interpreter start-ups would take."""
    import time
    (tmp_path / "laplacian.t").write_text(IMAGE_LAPLACIAN_T)
    t0 = time.time()
    r = subprocess.run([os.path.join(BIN, "ref_create_delete_cycle")], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    dt = time.time() - t0
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "Iteration: 9" in r.stdout
    assert (tmp_path / "result.png").exists()
    assert dt < 60, dt


MINIMAL_EXCLUDE_T = """
-- fit to A plus unguarded forward differences (out-of-bounds reads are zero, thallo.t:876-882)
W,H = Dims("W","H")
Inputs { X = Unknown(float,{W,H},0), A = Array(float,{W,H},1) }
w_fit = .2
x,y = W(), H()
r = Residuals { fit = w_fit*(X(x,y) - A(x,y)), reg = { (X(x,y) - X(x+1,y)), (X(x,y) - X(x,y+1)) } }
"""

MINIMAL_MATERIALIZE_T = """
-- the x-difference goes through a ComputedArray (v:get), so its values and gradient image are stored by `precompute`
W,H = Dims("W","H")
Inputs { X = Unknown(float,{W,H},0), A = Array(float,{W,H},1) }
x,y = W(),H()
v = X(x,y) - X(x+1,y)
v = Select(InBounds(x+1,y), v, 0.0)
local reg = v:get(x,y)
r = Residuals { fit = (X(x,y) - A(x,y)), reg = reg }
"""


@pytest.mark.gpu
@pytest.mark.parametrize("binary,fname,text,seed", [("ref_minimal_exclude", "minimal_exclude.t", MINIMAL_EXCLUDE_T, 1),
                                                   ("ref_minimal_materialize", "minimal_materialize.t", MINIMAL_MATERIALIZE_T, 0xF8127324)],
                         ids=["minimal_exclude", "minimal_materialize"])
def test_more_reference_programs_run_against_this_library(tmp_path, binary, fname, text, seed):
    """tests/minimal_exclude and tests/minimal_materialize (a ComputedArray), unmodified, 512 x 512, GN 10 x 10: the cost the
    program prints equals the cost of the same solve driven through the Python mirror of the C ABI from the same file."""
    if not os.path.isfile(os.path.join(BIN, binary)):
        pytest.skip("oracle/_ref not built")
    import torch
    from thallo_b200.api import ThalloSolver
    (tmp_path / fname).write_text(text)
    r = subprocess.run([os.path.join(BIN, binary)], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    m = re.search(r"%s ([-+0-9.eE]+|nan|inf)" % binary[len("ref_"):], r.stdout)
    assert m, r.stdout[-2000:]
    cost = float(m.group(1))
    A = _libc_uniform(512 * 512, seed)
    dX, dA = torch.from_numpy(A.copy()).cuda(), torch.from_numpy(A.copy()).cuda()
    s = ThalloSolver([512, 512], str(tmp_path / fname), "gauss_newton", via_file=True)
    want = s.solve([dX, dA])
    assert np.isfinite(cost) and abs(cost - want) <= 1e-5 * abs(want), (cost, want)
