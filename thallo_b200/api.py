"""ctypes binding of libThallo.so and a host-side mirror of the reference's C++ solver
wrapper (examples/shared/ThalloSolver.h:40-112: ctor = NewState -> ProblemDefine ->
ProblemPlan; solve = SetSolverParameter* -> Solve or Init/Step loop -> summary).

PyTorch is used only as the owner of device memory / streams; every compute call goes
through the C ABI declared in include/Thallo.h and include/thallo_b200.h.  The library
is required: there is no Python or CPU fallback.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libThallo.so")


class InitializationParameters(C.Structure):        # Thallo.h:10-36
    _fields_ = [("doublePrecision", C.c_int), ("verbosityLevel", C.c_int), ("timingLevel", C.c_int),
                ("threadsPerBlock", C.c_int), ("useAutoscheduler", C.c_int), ("cpuOnly", C.c_int)]


class PerformanceEntry(C.Structure):                # Thallo.h:85-92
    _fields_ = [("count", C.c_uint), ("minMS", C.c_double), ("maxMS", C.c_double), ("meanMS", C.c_double),
                ("stddevMS", C.c_double)]


class PerformanceSummary(C.Structure):              # Thallo.h:94-104
    _fields_ = [("total", PerformanceEntry), ("nonlinearIteration", PerformanceEntry),
                ("nonlinearSetup", PerformanceEntry), ("linearSolve", PerformanceEntry),
                ("nonlinearResolve", PerformanceEntry)]


FLOAT_PARAMS = ("min_relative_decrease", "min_trust_region_radius", "max_trust_region_radius", "q_tolerance",
                "function_tolerance", "trust_region_radius", "radius_decrease_factor", "min_lm_diagonal",
                "max_lm_diagonal", "max_solver_time_in_seconds")
INT_PARAMS = ("residual_reset_period", "nIter", "nIterations", "lIterations")

_lib = None


def build_library(force=False):
    """make -C thallo_b200/csrc (host C++ only; kernels are NVRTC-compiled per plan)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libThallo.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C thallo_b200/csrc`); thallo_b200 has no fallback path")
    L = C.CDLL(LIB_PATH)
    vp, cp = C.c_void_p, C.c_char_p
    L.Thallo_NewState.restype, L.Thallo_NewState.argtypes = vp, [InitializationParameters]
    L.Thallo_ProblemDefine.restype, L.Thallo_ProblemDefine.argtypes = vp, [vp, cp, cp]
    L.Thallo_ProblemDelete.restype, L.Thallo_ProblemDelete.argtypes = None, [vp, vp]
    L.Thallo_ProblemPlan.restype, L.Thallo_ProblemPlan.argtypes = vp, [vp, vp, C.POINTER(C.c_uint)]
    L.Thallo_PlanFree.restype, L.Thallo_PlanFree.argtypes = None, [vp, vp]
    L.Thallo_SetSolverParameter.restype, L.Thallo_SetSolverParameter.argtypes = None, [vp, vp, cp, vp]
    L.Thallo_GetSolverParameter.restype, L.Thallo_GetSolverParameter.argtypes = None, [vp, vp, cp, vp]
    L.Thallo_ProblemSolve.restype, L.Thallo_ProblemSolve.argtypes = None, [vp, vp, C.POINTER(vp)]
    L.Thallo_ProblemInit.restype, L.Thallo_ProblemInit.argtypes = None, [vp, vp, C.POINTER(vp)]
    L.Thallo_ProblemStep.restype, L.Thallo_ProblemStep.argtypes = C.c_int, [vp, vp, C.POINTER(vp)]
    L.Thallo_ProblemCurrentCost.restype, L.Thallo_ProblemCurrentCost.argtypes = C.c_double, [vp, vp]
    L.Thallo_GetPerformanceSummary.restype, L.Thallo_GetPerformanceSummary.argtypes = None, [vp, vp, C.POINTER(PerformanceSummary)]
    L.ThalloB200_ProblemDefineFromSource.restype, L.ThalloB200_ProblemDefineFromSource.argtypes = vp, [vp, cp, cp, cp]
    L.ThalloB200_CompileOnly.restype = C.c_int
    L.ThalloB200_CompileOnly.argtypes = [cp, C.c_char_p, C.c_ulong, C.POINTER(C.c_ulong)]
    L.ThalloB200_SetStream.restype, L.ThalloB200_SetStream.argtypes = None, [vp, vp]
    L.ThalloB200_PlanLaunchCount.restype, L.ThalloB200_PlanLaunchCount.argtypes = C.c_ulonglong, [vp, vp]
    L.ThalloB200_PlanLastLinearIterations.restype, L.ThalloB200_PlanLastLinearIterations.argtypes = C.c_int, [vp, vp]
    L.ThalloB200_PlanTotalLinearIterations.restype, L.ThalloB200_PlanTotalLinearIterations.argtypes = C.c_ulonglong, [vp, vp]
    L.ThalloB200_PlanReadVector.restype = C.c_longlong
    L.ThalloB200_PlanReadVector.argtypes = [vp, vp, cp, vp, C.c_longlong]
    L.ThalloB200_PlanVectorPointer.restype, L.ThalloB200_PlanVectorPointer.argtypes = vp, [vp, vp, cp]
    L.ThalloB200_PlanExportJacobian.restype = C.c_longlong
    L.ThalloB200_PlanExportJacobian.argtypes = [vp, vp, C.c_int, vp, vp, C.c_longlong]
    L.ThalloB200_PlanKernelTimes.restype = C.c_longlong
    L.ThalloB200_PlanKernelTimes.argtypes = [vp, vp, C.c_char_p, C.c_longlong]
    L.ThalloB200_NcclUniqueId.restype, L.ThalloB200_NcclUniqueId.argtypes = C.c_int, [vp, C.c_int]
    L.ThalloB200_PlanInitComm.restype, L.ThalloB200_PlanInitComm.argtypes = C.c_int, [vp, vp, vp, C.c_int, C.c_int]
    L.ThalloB200_PlanIpcHandle.restype = C.c_int
    L.ThalloB200_PlanIpcHandle.argtypes = [vp, vp, vp, C.POINTER(C.c_longlong)]
    L.ThalloB200_PlanConnect.restype = C.c_int
    L.ThalloB200_PlanConnect.argtypes = [vp, vp, vp, C.c_longlong, vp, C.c_longlong]
    L.ThalloB200_PlanConnectGraph.restype = C.c_int
    L.ThalloB200_PlanConnectGraph.argtypes = [vp, vp, vp, C.c_longlong, C.c_longlong, vp, C.c_longlong, C.c_longlong]
    L.ThalloB200_PlanPeerInfo.restype, L.ThalloB200_PlanPeerInfo.argtypes = C.c_int, [vp, vp, C.POINTER(C.c_longlong)]
    L.ThalloB200_PlanConnectAll.restype = C.c_int
    L.ThalloB200_PlanConnectAll.argtypes = [vp, vp, C.c_int, vp, C.POINTER(C.c_longlong)]
    L.ThalloB200_WarpSelfTest.restype = C.c_int
    L.ThalloB200_WarpSelfTest.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int]
    L.ThalloB200_LastError.restype, L.ThalloB200_LastError.argtypes = cp, []
    L.ThalloB200_Version.restype, L.ThalloB200_Version.argtypes = cp, []
    _lib = L
    return L


def warp_self_test(which, nkeys=4):
    """Known-answer tests of the warp primitives (include/thallo_b200.h ThalloB200_WarpSelfTest)."""
    out = (C.c_double * 64)()
    n = lib().ThalloB200_WarpSelfTest(int(which), int(nkeys), out, 64)
    if n < 0:
        raise RuntimeError("ThalloB200_WarpSelfTest failed: " + lib().ThalloB200_LastError().decode())
    return [out[i] for i in range(n)]


def compile_only(source):
    """NVRTC-compile a generated translation unit for sm_100a (works without a GPU).
    Returns (ok, log, cubin_bytes)."""
    L = lib()
    buf = C.create_string_buffer(1 << 16)
    sz = C.c_ulong(0)
    rc = L.ThalloB200_CompileOnly(source.encode(), buf, len(buf), C.byref(sz))
    return rc == 0, buf.value.decode(errors="replace"), sz.value


def assemble_jacobian(desc, fetch):
    """CSR J from per-group (value, column) entry lists in the export layout of ThalloB200_PlanExportJacobian:
    `fetch(group, n)` returns the n = count * nnz_per_elem entries of a group, element-major, within an element row
    by row with row_nnz[k] entries in row k; column -1 marks an access outside the domain.  Rows are numbered like the
    reference's (gauss_newton.t:401-402): group base + element * terms + term."""
    import numpy as np
    import scipy.sparse as sp
    rows, cols, vals, base = [], [], [], 0
    for gi, g in enumerate(desc["groups"]):
        n = g["count"] * g["nnz_per_elem"]
        v, c = fetch(gi, n)
        term_of = np.repeat(np.arange(g["nterms"]), g["row_nnz"])                  # entry within an element -> row
        r = base + (np.arange(g["count"])[:, None] * g["nterms"] + term_of[None, :]).reshape(-1)
        ok = np.asarray(c) >= 0
        rows.append(r[ok]); cols.append(np.asarray(c)[ok]); vals.append(np.asarray(v)[ok])
        base += g["count"] * g["nterms"]
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(base, desc["nunk"]))


class ThalloSolver:
    """Mirror of examples/shared/ThalloSolver.h.  `energy` is either a reference-style file
    name ("image_warping.t": the C library runs the front end itself, as a reference program
    would) or an energy module name lowered in-process and handed over through
    ThalloB200_ProblemDefineFromSource."""

    def __init__(self, dims, energy, kind="gauss_newton", double=False, verbosity=0, timing=1,
                 via_file=False, schedule="auto", define_kwargs=None, stream=None, partition=None):
        L = lib()
        self.L = L
        self.dims = [int(d) for d in dims]
        self.double = bool(double)
        ip = InitializationParameters(int(double), int(verbosity), int(timing), 0, 1, 0)
        self.state = L.Thallo_NewState(ip)
        if stream is not None:
            L.ThalloB200_SetStream(self.state, C.c_void_p(int(stream)))
        self.lowered = None
        if via_file:
            self.problem = L.Thallo_ProblemDefine(self.state, energy.encode(), kind.encode())
        else:
            import energies
            from .frontend import codegen
            define, mod = energies.define_for(energy)
            if define is None:
                raise RuntimeError("no energy file or registered energy named '%s'" % energy)
            low = codegen.lower(define, self.dims, kind, mod, double, schedule, partition=partition,
                                **(define_kwargs or {}))
            self.lowered = low
            self.problem = L.ThalloB200_ProblemDefineFromSource(
                self.state, codegen.descriptor_text(low.desc).encode(), low.source.encode(), kind.encode())
        if not self.problem:
            raise RuntimeError("Thallo_ProblemDefine failed: " + L.ThalloB200_LastError().decode())
        arr = (C.c_uint * len(self.dims))(*self.dims)
        self.plan = L.Thallo_ProblemPlan(self.state, self.problem, arr)
        if not self.plan:
            raise RuntimeError("Thallo_ProblemPlan failed: " + L.ThalloB200_LastError().decode())
        self._keep = []

    # ---- parameter marshalling (examples/shared/NamedParameters.h): device tensors -> pointers,
    # python / numpy scalars -> host scalars of the declared C type
    def _params(self, params):
        import numpy as np
        n = len(params)
        arr = (C.c_void_p * n)()
        keep = []
        for i, p in enumerate(params):
            if hasattr(p, "data_ptr"):
                arr[i] = p.data_ptr()
                keep.append(p)
            elif isinstance(p, (np.ndarray, np.generic)):
                a = np.ascontiguousarray(p)
                arr[i] = a.ctypes.data
                keep.append(a)
            elif isinstance(p, float):
                a = np.array([p], np.float32)
                arr[i] = a.ctypes.data
                keep.append(a)
            elif isinstance(p, int):
                arr[i] = p
            else:
                raise TypeError("parameter %d: unsupported type %r" % (i, type(p)))
        self._keep = keep
        return arr

    def set_parameters(self, **kw):
        for k, v in kw.items():
            if k in INT_PARAMS:
                val = C.c_int(int(v))
            else:
                val = C.c_float(float(v))
            self.L.Thallo_SetSolverParameter(self.state, self.plan, k.encode(), C.byref(val))

    def get_parameter(self, k):
        val = C.c_int(0) if k in INT_PARAMS else C.c_float(0)
        self.L.Thallo_GetSolverParameter(self.state, self.plan, k.encode(), C.byref(val))
        return val.value

    def solve(self, params, **solver_params):
        self.set_parameters(**solver_params)
        self.L.Thallo_ProblemSolve(self.state, self.plan, self._params(params))
        return self.current_cost()

    def init(self, params):
        self._arr = self._params(params)
        self.L.Thallo_ProblemInit(self.state, self.plan, self._arr)

    def step(self, params=None):
        if params is not None:
            self._arr = self._params(params)
        return self.L.Thallo_ProblemStep(self.state, self.plan, self._arr)

    def current_cost(self):
        return self.L.Thallo_ProblemCurrentCost(self.state, self.plan)

    def summary(self):
        s = PerformanceSummary()
        self.L.Thallo_GetPerformanceSummary(self.state, self.plan, C.byref(s))
        return s

    def launches(self):
        return int(self.L.ThalloB200_PlanLaunchCount(self.state, self.plan))

    def last_linear_iterations(self):
        return int(self.L.ThalloB200_PlanLastLinearIterations(self.state, self.plan))

    def total_linear_iterations(self):
        return int(self.L.ThalloB200_PlanTotalLinearIterations(self.state, self.plan))

    def kernel_times(self):
        """{kernel name: (launches, total device ms)}; needs timing >= 2."""
        buf = C.create_string_buffer(1 << 16)
        self.L.ThalloB200_PlanKernelTimes(self.state, self.plan, buf, len(buf))
        out = {}
        for ln in buf.value.decode().splitlines():
            name, cnt, ms = ln.split()
            out[name] = (int(cnt), float(ms))
        return out

    def export_jacobian(self):
        """J at the current unknowns (after init) as a SciPy CSR matrix in the reference's row order
        (group base + element * terms + term), assembled from ThalloB200_PlanExportJacobian of every group."""
        import numpy as np
        rt = np.float64 if self.double else np.float32

        def fetch(gi, n):
            v = np.zeros(max(n, 1), rt)
            c = np.zeros(max(n, 1), np.int64)
            got = self.L.ThalloB200_PlanExportJacobian(self.state, self.plan, gi, v.ctypes.data, c.ctypes.data, n)
            if got != n:
                raise RuntimeError("ThalloB200_PlanExportJacobian failed for group %d" % gi)
            return v[:n], c[:n]
        return assemble_jacobian(self.lowered.desc, fetch)

    def read_vector(self, name, count):
        import numpy as np
        out = np.zeros(count, np.float64 if self.double else np.float32)
        self.L.ThalloB200_PlanReadVector(self.state, self.plan, name.encode(), out.ctypes.data, count)
        return out

    def vector(self, name):
        """Zero-copy torch view of solver vector `name` on the device (nunk reals)."""
        import torch
        ptr = self.L.ThalloB200_PlanVectorPointer(self.state, self.plan, name.encode())
        if not ptr:
            raise KeyError(name)
        n = int(self.lowered.desc["nunk"]) if self.lowered is not None else None
        assert n is not None, "vector(): needs a plan lowered in-process"

        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = dict(shape=(n,), typestr="<f8" if self.double else "<f4", data=(int(ptr), False), version=3,
                                          strides=None)
        return torch.as_tensor(v, device="cuda")

    def close(self):
        if getattr(self, "plan", None):
            self.L.Thallo_PlanFree(self.state, self.plan)
            self.plan = None
        if getattr(self, "problem", None):
            self.L.Thallo_ProblemDelete(self.state, self.problem)
            self.problem = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
