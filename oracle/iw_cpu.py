"""ORACLE (test infrastructure / CPU baseline, not product code): ctypes wrapper of
oracle/iw_cpu.c, the plain-C restatement of the reference's cpuOnly path for the
image_warping energy (see the header of iw_cpu.c for the reference file:line map).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
fp = C.POINTER(C.c_float)


class IWProblem(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("Offset", fp), ("Angle", fp), ("UrShape", fp),
                ("Constraints", fp), ("Mask", fp), ("w_fitSqrt", C.c_float), ("w_regSqrt", C.c_float)]


class IWSolverParams(C.Structure):
    _fields_ = [("lm", C.c_int), ("nIterations", C.c_int), ("lIterations", C.c_int), ("residual_reset_period", C.c_int),
                ("min_relative_decrease", C.c_float), ("min_trust_region_radius", C.c_float),
                ("max_trust_region_radius", C.c_float), ("q_tolerance", C.c_float), ("function_tolerance", C.c_float),
                ("trust_region_radius", C.c_float), ("radius_decrease_factor", C.c_float),
                ("min_lm_diagonal", C.c_float), ("max_lm_diagonal", C.c_float), ("max_pcg_iterations", C.c_longlong)]


class IWResult(C.Structure):
    _fields_ = [("n_nonlinear", C.c_int), ("n_pcg", C.c_longlong), ("seconds_total", C.c_double),
                ("seconds_pcg", C.c_double), ("cost", C.c_double * 260), ("n_lin", C.c_int * 256), ("n_cost", C.c_int)]


_libs = {}


def build():
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib(acc64=False):
    name = "libiw_cpu_acc64.so" if acc64 else "libiw_cpu.so"
    if name not in _libs:
        path = os.path.join(_HERE, "lib", name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.iw_default_params.argtypes = [C.POINTER(IWSolverParams)]
        L.iw_solve.restype, L.iw_solve.argtypes = C.c_int, [C.POINTER(IWProblem), C.POINTER(IWSolverParams), C.POINTER(IWResult)]
        L.iw_cost.restype, L.iw_cost.argtypes = C.c_double, [C.POINTER(IWProblem)]
        L.iw_num_threads.restype = C.c_int
        L.iw_set_threads.argtypes = [C.c_int]
        _libs[name] = L
    return _libs[name]


def solve(W, H, d, kind="levenberg_marquardt", acc64=False, max_pcg=0, **params):
    """d: dict from workloads.image_warping_inputs (Offset and Angle are updated in place).
    Returns dict(costs, n_lin, n_pcg, seconds_total, seconds_pcg, threads)."""
    L = lib(acc64)
    arrs = {k: np.ascontiguousarray(d[k], np.float32) for k in ("Offset", "Angle", "UrShape", "Constraints", "Mask")}
    P = IWProblem(W, H, *[arrs[k].ctypes.data_as(fp) for k in ("Offset", "Angle", "UrShape", "Constraints", "Mask")],
                  float(d["w_fitSqrt"]), float(d["w_regSqrt"]))
    sp = IWSolverParams()
    L.iw_default_params(C.byref(sp))
    sp.lm = int(kind == "levenberg_marquardt")
    sp.max_pcg_iterations = int(max_pcg)
    for k, v in params.items():
        setattr(sp, k, v)
    R = IWResult()
    rc = L.iw_solve(C.byref(P), C.byref(sp), C.byref(R))
    assert rc == 0
    for k in ("Offset", "Angle"):
        if arrs[k] is not d[k]:
            np.copyto(np.asarray(d[k]).reshape(arrs[k].shape), arrs[k])
    return dict(costs=[R.cost[i] for i in range(R.n_cost)], n_lin=[R.n_lin[i] for i in range(min(R.n_nonlinear, 256))],
                n_pcg=int(R.n_pcg), n_nonlinear=int(R.n_nonlinear), seconds_total=R.seconds_total,
                seconds_pcg=R.seconds_pcg, threads=int(L.iw_num_threads()))
