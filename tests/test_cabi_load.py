"""The C-ABI library loads on a CPU-only box and exports every symbol the headers declare;
every configured energy lowers to CUDA C++ that NVRTC compiles for sm_100a (no GPU needed)."""
import os
import re

import pytest

from thallo_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(Thallo(?:B200)?_[A-Za-z]+)\s*\(", txt)))


def test_library_builds_and_exports_all_declared_symbols():
    api.build_library()
    L = api.lib()
    names = _declared("Thallo.h") + _declared("thallo_b200.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libThallo.so does not export " + n
    assert b"sm_100a" in L.ThalloB200_Version()


def test_reference_entry_points_present():
    # exactly the twelve functions of reference API/release/include/Thallo.h:41-106
    want = {"Thallo_NewState", "Thallo_ProblemDefine", "Thallo_ProblemDelete", "Thallo_ProblemPlan", "Thallo_PlanFree",
            "Thallo_SetSolverParameter", "Thallo_GetSolverParameter", "Thallo_ProblemSolve", "Thallo_ProblemInit",
            "Thallo_ProblemStep", "Thallo_ProblemCurrentCost", "Thallo_GetPerformanceSummary"}
    assert want == set(_declared("Thallo.h"))


ENERGIES = [
    ("laplacian", [64, 64], "gauss_newton", {}),
    ("laplacian", [64, 64], "levenberg_marquardt", dict(schedule="residualwise")),
    ("graph_laplacian", [512, 511], "gauss_newton", {}),
    ("image_warping", [128, 96], "levenberg_marquardt", {}),
    ("image_warping", [128, 96], "gauss_newton", dict(schedule="residualwise")),
    ("optical_flow", [64, 64], "gauss_newton", {}),
    ("arap_mesh_deformation", [100, 600], "gauss_newton", {}),
    ("volumetric_mesh_deformation", [8, 8, 8], "gauss_newton", {}),
    ("bundle_adjustment", [10, 100, 500], "levenberg_marquardt", {}),
]


@pytest.mark.parametrize("name,dims,kind,kw", ENERGIES)
def test_energy_lowers_and_compiles_for_sm100a(name, dims, kind, kw):
    import energies
    from thallo_b200.frontend import codegen
    api.build_library()
    low = codegen.lower(energies.load(name), dims, kind, name, False, kw.get("schedule", "auto"))
    ok, log, size = api.compile_only(low.source)
    assert ok, log[-4000:]
    assert size > 1000


def test_double_precision_compiles():
    import energies
    from thallo_b200.frontend import codegen
    low = codegen.lower(energies.load("image_warping"), [64, 64], "levenberg_marquardt", "image_warping", True)
    ok, log, size = api.compile_only(low.source)
    assert ok, log[-4000:]


def test_warp_primitive_known_answer_kernels_compile():
    # the translation unit ThalloB200_WarpSelfTest runs on a GPU box (tests/test_gpu_warp.py)
    ok, log, size = api.compile_only('#define TH_WARP_KAT 1\n#include "thallo_warp.cuh"\n')
    assert ok, log[-2000:]
    assert size > 1000


# ---- error behaviour of the reference-facing entry points, as far as it can be exercised without a GPU
def test_problem_define_rejects_unknown_solver_kinds_and_plan_fails_loudly_without_a_gpu(tmp_path):
    import ctypes as C
    import torch
    from thallo_b200 import api
    L = api.lib()
    st = L.Thallo_NewState(api.InitializationParameters(0, 0, 0, 0, 1, 0))
    assert st
    # thallo.t:74 asserts the kind is exactly one of the two names ("LMGPU" is Opt's, not Thallo's)
    assert not L.Thallo_ProblemDefine(st, b"image_warping.t", b"LMGPU")
    assert b"unknown solver kind" in L.ThalloB200_LastError()
    p = L.Thallo_ProblemDefine(st, b"image_warping.t", b"levenberg_marquardt")
    assert p                                           # only records (file name, kind), like the reference (thallo.t:5954-5958)
    dims = (C.c_uint * 2)(64, 64)
    missing = L.Thallo_ProblemDefine(st, b"no_such_energy.t", b"gauss_newton")
    assert missing and not L.Thallo_ProblemPlan(st, missing, dims)         # the file is read at plan time
    assert b"does not exist" in L.ThalloB200_LastError()
    broken = tmp_path / "broken.t"
    broken.write_text("local W,H = Dims('W','H')\nInputs { X = Unknown(float,{W,H},0) }\nr = Residuals { fit = X(W(),H()) + }\n")
    bp = L.Thallo_ProblemDefine(st, str(broken).encode(), b"gauss_newton")
    assert bp and not L.Thallo_ProblemPlan(st, bp, dims)
    assert b"broken.t:3" in L.ThalloB200_LastError()                        # syntax error reported with file and line
    # the front end is started with an argument vector, not through a shell: quotes and spaces in a path are just characters
    odd = tmp_path / "it's an \"energy\" $(file).t"
    odd.write_text("local W,H = Dims('W','H')\nInputs { X = Unknown(float,{W,H},0), A = Array(float,{W,H},1) }\n"
                   "r = Residuals { fit = X(W(),H()) - A(W(),H()) }\n")
    op = L.Thallo_ProblemDefine(st, str(odd).encode(), b"gauss_newton")
    plan = L.Thallo_ProblemPlan(st, op, dims)
    if torch.cuda.is_available():
        assert plan
        L.Thallo_PlanFree(st, plan)
    else:
        assert not plan and b"no CUDA device" in L.ThalloB200_LastError()     # i.e. the lowering itself succeeded
    if not torch.cuda.is_available():
        assert not L.Thallo_ProblemPlan(st, p, dims)                          # no CPU fallback: NULL + message, never a silent CPU path
        assert b"no CUDA device" in L.ThalloB200_LastError() and b"no CPU fallback" in L.ThalloB200_LastError()
    L.Thallo_ProblemDelete(st, p)
    assert L.ThalloB200_Version().startswith(b"thallo_b200")
