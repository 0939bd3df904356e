"""Parity tests proper: the CUDA path, called through the C ABI (Thallo.h), against the
oracle on the same seeded inputs, and against the reference's golden images.

Tolerances (BASELINE.json north_star): per-iteration and final cost within 1e-5 relative
in float32 (1e-10 in float64), identical iteration counts where a convergence test fires.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import energies
from oracle.solver import OracleSolver
from thallo_b200 import workloads as wl

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_solver(*a, **k):
    from thallo_b200.api import ThalloSolver
    return ThalloSolver(*a, **k)


# ---------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("schedule", ["at_output", "residualwise"])
def test_kat1_minimal_golden(schedule):
    gold = np.load(os.path.join(GOLD, "kat_minimal.npz"))["gold"]
    X, A = wl.minimal_inputs(512, 512)
    dX, dA = dev(X), dev(A)
    s = make_solver([512, 512], "laplacian", "gauss_newton", schedule=schedule)
    s.solve([dX, dA])
    out = (dX.cpu().numpy().reshape(512, 512) * 255).astype(np.uint8)
    assert np.array_equal(out, gold)


def test_kat2_minimal_graph_golden():
    gold = np.load(os.path.join(GOLD, "kat_minimal_graph.npz"))["gold"].reshape(-1)
    X, A, v0, v1 = wl.minimal_graph_inputs(512)
    dX, dA, d0, d1 = dev(X), dev(A), dev(v0), dev(v1)
    s = make_solver([512, 511], "graph_laplacian", "gauss_newton")
    s.solve([dX, dA, d0, d1])
    out = (dX.cpu().numpy() * 255).astype(np.uint8)
    assert np.array_equal(out, gold)


def test_kat1_via_reference_style_file_name():
    """Thallo_ProblemDefine(state, "…/tests/minimal/laplacian.t", …): the library runs the front end itself."""
    gold = np.load(os.path.join(GOLD, "kat_minimal.npz"))["gold"]
    X, A = wl.minimal_inputs(512, 512)
    dX, dA = dev(X), dev(A)
    s = make_solver([512, 512], "tests/minimal/laplacian.t", "gauss_newton", via_file=True)
    s.solve([dX, dA])
    out = (dX.cpu().numpy().reshape(512, 512) * 255).astype(np.uint8)
    assert np.array_equal(out, gold)


# ---------------------------------------------------------------- oracle parity, config 1 (256x256)
@pytest.mark.parametrize("schedule", ["at_output", "residualwise"])
def test_config1_minimal_256_matches_oracle(schedule):
    X, A = wl.minimal_inputs(256, 256)
    Xo = X.copy()
    o = OracleSolver(energies.load("laplacian"), [256, 256], "gauss_newton", np.float32, schedule)
    c_ref = o.solve([Xo, A])
    dX, dA = dev(X), dev(A)
    s = make_solver([256, 256], "laplacian", "gauss_newton", schedule=schedule)
    c = s.solve([dX, dA])
    assert abs(c - c_ref) <= 1e-5 * abs(c_ref)
    assert np.abs(dX.cpu().numpy() - Xo).max() < 1e-4


# ---------------------------------------------------------------- oracle parity, image_warping
def _iw(W, H, kind, schedule, dtype=np.float32, nit=6, lit=40):
    d = wl.image_warping_inputs(W, H)
    po = wl.image_warping_params(d)
    po = [np.array(p, dtype=dtype) if i < 5 else p for i, p in enumerate(po)]
    o = OracleSolver(energies.load("image_warping"), [W, H], kind, dtype, schedule)
    o.set("nIterations", nit); o.set("lIterations", lit)
    o.init(po)
    costs_ref = [o.current_cost()]
    while o.step(po):
        costs_ref.append(o.current_cost())
    costs_ref.append(o.current_cost())

    pg = wl.image_warping_params(wl.image_warping_inputs(W, H))
    pg = [np.array(p, dtype=dtype) if i < 5 else p for i, p in enumerate(pg)]
    dp = [dev(p) if i < 5 else p for i, p in enumerate(pg)]
    s = make_solver([W, H], "image_warping", kind, double=(dtype == np.float64), schedule=schedule)
    s.set_parameters(nIterations=nit, lIterations=lit)
    s.init(dp)
    costs = [s.current_cost()]
    lin = []
    while s.step():
        costs.append(s.current_cost())
        lin.append(s.last_linear_iterations())
    costs.append(s.current_cost())
    return o, costs_ref, costs, lin, po, dp


@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
@pytest.mark.parametrize("schedule", ["at_output", "residualwise"])
def test_image_warping_cost_trajectory_f32(kind, schedule):
    o, cref, c, lin, po, dp = _iw(96, 80, kind, schedule)
    assert len(c) == len(cref), (c, cref)
    for a, b in zip(c, cref):
        assert abs(a - b) <= 1e-5 * max(abs(b), 1e-3), (c, cref)
    if kind == "levenberg_marquardt":
        ref_lin = [it["n_lin"] for it in o.trace if "n_lin" in it]
        assert lin == ref_lin[:len(lin)], (lin, ref_lin)
    off = dp[0].cpu().numpy()
    assert np.abs(off - po[0]).max() < 2e-3
    mask = po[4].reshape(-1) != 0         # excluded unknowns are never written
    init = wl.image_warping_inputs(96, 80)
    assert np.array_equal(off.reshape(-1, 2)[mask], init["Offset"].reshape(-1, 2)[mask])


def test_image_warping_cost_trajectory_f64():
    o, cref, c, lin, po, dp = _iw(64, 48, "levenberg_marquardt", "at_output", np.float64)
    assert len(c) == len(cref)
    for a, b in zip(c, cref):
        assert abs(a - b) <= 1e-10 * max(abs(b), 1e-6), (c, cref)
    ref_lin = [it["n_lin"] for it in o.trace if "n_lin" in it]
    assert lin == ref_lin[:len(lin)]


def test_solver_parameter_roundtrip_and_summary():
    X, A = wl.minimal_inputs(64, 64)
    s = make_solver([64, 64], "laplacian", "gauss_newton")
    assert s.get_parameter("nIterations") == 10 and s.get_parameter("lIterations") == 10   # gauss_newton.t:53-54
    assert abs(s.get_parameter("trust_region_radius") - 1e4) < 1
    s.set_parameters(nIterations=3, lIterations=5, q_tolerance=0.5)
    assert s.get_parameter("nIterations") == 3 and abs(s.get_parameter("q_tolerance") - 0.5) < 1e-7
    dX, dA = dev(X), dev(A)
    s.solve([dX, dA])
    sm = s.summary()
    assert sm.total.count == 1 and sm.nonlinearIteration.count == 3 and sm.linearSolve.count == 3
    assert sm.total.meanMS > 0
    assert s.launches() >= 3 * (1 + 5 * 2)      # per nonlinear iteration: init + lIterations x (th_pcg_a, th_pcg_b)


def test_lm_as_committed_switch_runs_gauss_newton(monkeypatch):
    """THALLO_LM_AS_COMMITTED=1: a "levenberg_marquardt" problem defined by file name behaves like the reference
    snapshot, where that kind silently runs Gauss-Newton (thallo.t:463)."""
    W, H = 64, 40

    def run(kind):
        p = wl.image_warping_params(wl.image_warping_inputs(W, H))
        dp = [dev(x) if i < 5 else x for i, x in enumerate(p)]
        s = make_solver([W, H], "image_warping.t", kind, via_file=True)
        s.set_parameters(nIterations=3, lIterations=15)
        s.init(dp)
        c = [s.current_cost()]
        while s.step():
            c.append(s.current_cost())
        c.append(s.current_cost())
        s.close()
        return c
    gn = run("gauss_newton")
    lm = run("levenberg_marquardt")
    monkeypatch.setenv("THALLO_LM_AS_COMMITTED", "1")
    compat = run("levenberg_marquardt")
    assert compat == gn and lm != gn
