"""End to end from an energy FILE, the way reference programs use the library: a `.t` file on disk is
handed to Thallo_ProblemDefine by name (examples/shared/ThalloSolver.h:43-60), the library reads and
lowers it (thallo_b200/frontend/tlang.py), JIT-compiles it and solves.  The oracle evaluates the same
file through its own DSL namespace.  The energy texts are written for these tests."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle.solver import OracleSolver
from thallo_b200.frontend import tlang
from test_tlang import HEAT_T

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _trajectory_oracle(define, dims, kind, dtype, mode, params, nit, lit, **kw):
    o = OracleSolver(define, dims, kind, dtype, mode, **kw)
    o.set("nIterations", nit); o.set("lIterations", lit)
    o.init(params)
    c = [o.current_cost()]
    while o.step(params):
        c.append(o.current_cost())
    c.append(o.current_cost())
    return o, c


def _trajectory_gpu(path, dims, kind, dtype, params, nit, lit):
    from thallo_b200.api import ThalloSolver
    s = ThalloSolver(dims, path, kind, double=(dtype == np.float64), via_file=True)
    s.set_parameters(nIterations=nit, lIterations=lit)
    s.init(params)
    c, lin = [s.current_cost()], []
    while s.step():
        c.append(s.current_cost())
        lin.append(s.last_linear_iterations())
    c.append(s.current_cost())
    s.close()
    return c, lin


def _heat_inputs(W, H, dtype):
    rs = np.random.RandomState(11)
    yy, xx = np.mgrid[0:H, 0:W]
    T = np.stack([np.sin(xx / 5.0) + 0.1 * rs.randn(H, W), np.cos(yy / 7.0) + 0.1 * rs.randn(H, W)], -1).astype(dtype)
    U = (T + 0.3 * rs.randn(H, W, 2)).astype(dtype)
    M = (rs.rand(H, W) < 0.15).astype(dtype)
    return U, T, M


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 1e-5)])
@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
def test_image_energy_from_t_file(tmp_path, dtype, tol, kind):
    W, H = 80, 56
    p = tmp_path / "heat.t"
    p.write_text(HEAT_T)
    U, T, M = _heat_inputs(W, H, dtype)
    wd = np.float32(0.8)
    Uo = U.copy()
    o, cref = _trajectory_oracle(tlang.load(str(p)), [W, H], kind, dtype, "at_output", [Uo, T, M, wd], 3, 12)
    dU, dT, dM = dev(U), dev(T), dev(M)
    c, lin = _trajectory_gpu(str(p), [W, H], kind, dtype, [dU, dT, dM, wd], 3, 12)
    assert len(c) == len(cref), (c, cref)
    for a, b in zip(c, cref):
        assert abs(a - b) <= tol * max(abs(b), 1e-6), (c, cref)
    if kind == "levenberg_marquardt":
        assert lin == [it["n_lin"] for it in o.trace if "n_lin" in it][:len(lin)]
    got = dU.cpu().numpy()
    assert np.abs(got - Uo).max() <= (1e-9 if dtype == np.float64 else 2e-4)
    masked = M != 0
    assert np.array_equal(got[masked], U[masked])          # excluded unknowns are never written


GET_T = """
-- smoothness of a nonlinear per-vertex quantity over graph edges; the quantity is fetched through
-- :get(), i.e. stored once per vertex (value + derivative) and read through the edge's index arrays
local N,E = Dims("N","E")
Inputs {
    X  = Unknown(thallo_float,{N},0),
    A  = Array(thallo_float,{N},1),
    v0 = Sparse({E},{N},2),
    v1 = Sparse({E},{N},3)
}
local n,e = N(),E()
local q = sin(X(n)) * (1.0 + A(n))
r = Residuals {
    fit = X(n) - A(n),
    reg = 0.5*(q:get(v0(e)) - q:get(v1(e)))
}
"""


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 1e-5)])
def test_graph_energy_with_computed_array_from_t_file(tmp_path, dtype, tol):
    N = 600
    rs = np.random.RandomState(5)
    v0 = np.concatenate([np.arange(N - 1), rs.randint(0, N, 400)]).astype(np.int32)
    v1 = np.concatenate([np.arange(1, N), rs.randint(0, N, 400)]).astype(np.int32)
    E = len(v0)
    A = rs.rand(N).astype(dtype)
    X = (A + 0.2 * rs.randn(N)).astype(dtype)
    p = tmp_path / "graph_get.t"
    p.write_text(GET_T)
    Xo = X.copy()
    o, cref = _trajectory_oracle(tlang.load(str(p)), [N, E], "gauss_newton", dtype, "residualwise", [Xo, A, v0, v1], 4, 10)
    dX = dev(X)
    c, _ = _trajectory_gpu(str(p), [N, E], "gauss_newton", dtype, [dX, dev(A), dev(v0), dev(v1)], 4, 10)
    assert len(c) == len(cref), (c, cref)
    for a, b in zip(c, cref):
        assert abs(a - b) <= tol * max(abs(b), 1e-6), (c, cref)
    assert np.abs(dX.cpu().numpy() - Xo).max() <= (1e-9 if dtype == np.float64 else 2e-4)


POSES_T = """
-- frame-to-frame alignment of point correspondences: one se(3) pose per frame, stored as a 4x4 matrix per frame
-- (computed array with 12 non-constant entries) and fetched per correspondence through the index arrays
local T,C = Dims("T","C")
Inputs {
    Rot   = Unknown(thallo_float3,{T},0),
    Trans = Unknown(thallo_float3,{T},1),
    Pa    = Array(thallo_float3,{C},2),
    Pb    = Array(thallo_float3,{C},3),
    fa    = Sparse({C},{T},4),
    fb    = Sparse({C},{T},5),
    w     = Param(float,6)
}
UsePreconditioner(true)
local t,c = T(),C()
local pose = PoseToMatrix(Rot(t), Trans(t))
local Ma, Mb = pose:get(fa(c)), pose:get(fb(c))
r = Residuals {
    align = Sqrt(w) * (rigid_trans(Ma, Pa(c)) - rigid_trans(Mb, Pb(c))),
    prior = { 0.3*Rot(t), 0.3*Trans(t) }
}
"""


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 1e-5)])
def test_pose_energy_with_stored_matrices_from_t_file(tmp_path, dtype, tol):
    rs = np.random.RandomState(9)
    T, C = 12, 300
    fa = rs.randint(0, T, C).astype(np.int32)
    fb = ((fa + 1 + rs.randint(0, T - 1, C)) % T).astype(np.int32)          # never the same frame twice
    Pa = rs.randn(C, 3).astype(dtype)
    Pb = (Pa + 0.05 * rs.randn(C, 3)).astype(dtype)
    Rot = (0.2 * rs.randn(T, 3)).astype(dtype)
    Trans = (0.3 * rs.randn(T, 3)).astype(dtype)
    w = np.float32(2.0)
    p = tmp_path / "poses.t"
    p.write_text(POSES_T)
    Ro, To = Rot.copy(), Trans.copy()
    o, cref = _trajectory_oracle(tlang.load(str(p)), [T, C], "gauss_newton", dtype, "residualwise", [Ro, To, Pa, Pb, fa, fb, w], 4, 12)
    dR, dT = dev(Rot), dev(Trans)
    c, _ = _trajectory_gpu(str(p), [T, C], "gauss_newton", dtype, [dR, dT, dev(Pa), dev(Pb), dev(fa), dev(fb), w], 4, 12)
    assert len(c) == len(cref), (c, cref)
    for a, b in zip(c, cref):
        assert abs(a - b) <= tol * max(abs(b), 1e-6), (c, cref)
    assert np.abs(dR.cpu().numpy() - Ro).max() <= (1e-8 if dtype == np.float64 else 2e-4)
