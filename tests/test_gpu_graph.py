"""Parity of the graph-domain schedules against the oracle through the C ABI: the gather schedule
(adjacency lists, no atomics; the default) and the residualwise scatter schedule (float atomics, the
reference's form) on config 4b's ARAP mesh energy (arap_mesh_deformation.t, small triangulated
grid); the materialised-Jacobian path on config 5's bundle adjustment energy and on config 1b
(tests/minimal/laplacian.t with its materialize directives).
Tolerance 1e-5 relative on every cost of the trajectory (float32), identical LM iteration counts."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

import energies
from oracle.solver import OracleSolver
from thallo_b200 import workloads as wl

pytestmark = pytest.mark.gpu


from _parity import dev, assert_costs_close, oracle_trajectory


def _oracle_costs(name, dims, kind, params, nit, lit, dtype=np.float32, **kw):
    o = OracleSolver(energies.load(name), dims, kind, dtype, "residualwise", **kw)
    o.set("nIterations", nit); o.set("lIterations", lit)
    o.init(params)
    c = [o.current_cost()]
    while o.step(params):
        c.append(o.current_cost())
    c.append(o.current_cost())
    return o, c


@pytest.mark.parametrize("schedule", ["gather", "residualwise"])
@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
def test_arap_mesh_matches_oracle(kind, schedule):
    from thallo_b200.api import ThalloSolver
    from _parity import ulp_perturbed
    nx, ny = 24, 18
    nit, lit = 4, 30
    d = wl.arap_mesh_inputs(nx, ny)
    dims = [nx * ny, len(d["V0"])]
    po = wl.arap_mesh_params(d)
    o, cref = _oracle_costs("arap_mesh_deformation", dims, kind, po, nit, lit)
    p64 = [np.array(x, np.float64) if (i >= 2 and x.dtype == np.float32) else x for i, x in enumerate(wl.arap_mesh_params(wl.arap_mesh_inputs(nx, ny)))]
    _, cref64 = _oracle_costs("arap_mesh_deformation", dims, kind, p64, nit, lit, np.float64)
    pert = []
    for seed in (1, 2):      # float32 rounding sensitivity of this (far from converged, truncated-PCG) trajectory
        dq = wl.arap_mesh_inputs(nx, ny)
        dq["Position"] = ulp_perturbed(dq["Position"], seed)
        pert.append(_oracle_costs("arap_mesh_deformation", dims, kind, wl.arap_mesh_params(dq), nit, lit)[1])

    pg = wl.arap_mesh_params(wl.arap_mesh_inputs(nx, ny))
    dp = [dev(p) if i >= 2 else p for i, p in enumerate(pg)]
    s = ThalloSolver(dims, "arap_mesh_deformation", kind, schedule=schedule)
    assert s.lowered.desc["schedule"] == schedule
    s.set_parameters(nIterations=nit, lIterations=lit)
    s.init(dp)
    c, lin = [s.current_cost()], []
    while s.step():
        c.append(s.current_cost())
        lin.append(s.last_linear_iterations())
    c.append(s.current_cost())
    assert_costs_close(c, cref, 1e-5, 1e-3, cref64, pert)
    assert abs(c[1] - cref[1]) <= 1e-5 * abs(cref[1])          # the first nonlinear step is well conditioned: plain 1e-5 rule
    if kind == "levenberg_marquardt":
        assert lin == [it["n_lin"] for it in o.trace][:len(lin)]
    assert np.abs(dp[2].cpu().numpy() - po[2]).max() < 5e-2


def _first_pcg_iteration_reference(name, dims, kind, params, dtype=np.float64, **kw):
    """r0 = -J^T F, Jacobi preconditioner, p0, A p0, alpha, delta1, r1 of the first PCG iteration of
    the first nonlinear iteration, from the oracle's J (float64)."""
    from oracle.npdsl import evaluate
    L, F, J = evaluate(energies.load(name), dims, params, dtype, **kw)
    J = J.tocsr()
    g = J.T @ F
    d = np.asarray(J.multiply(J).sum(axis=0)).reshape(-1)
    r0 = -g
    if kind == "levenberg_marquardt":
        radius = 1e4
        pre_gn = 1.0 / (1.0 + np.sqrt(d)) ** 2 if L.usepreconditioner else np.ones_like(d)
        ctc_raw = d / radius
        mult = (1.0 / pre_gn) / radius
        ctc = np.minimum(np.maximum(ctc_raw, 1e-6 * mult), 1e32 * mult)
        pre = 1.0 / (ctc + radius * ctc_raw)
    else:
        pre = 1.0 / (1.0 + np.sqrt(d)) ** 2 if L.usepreconditioner else np.ones_like(d)
        ctc = np.zeros_like(d)
    p0 = pre * r0
    Ap = J.T @ (J @ p0) + ctc * p0
    alpha = (r0 @ p0) / (p0 @ Ap)
    return dict(preconditioner=pre, Ap_X=Ap, delta=alpha * p0, r=r0 - alpha * Ap)


@pytest.mark.parametrize("case", ["arap_gather", "arap_residualwise", "arap_jtjp", "ba_materialised", "ba_matrix_free", "laplacian_materialised"])
@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
def test_first_pcg_iteration_vectors_match_oracle(case, kind):
    """Operator-level parity, independent of how truncated PCG amplifies rounding: after exactly one
    PCG iteration the solver vectors (Jacobi preconditioner, A p0, delta = alpha p0, r1 = r0 - alpha A p0)
    equal the float64 oracle's to float32 accuracy."""
    from thallo_b200.api import ThalloSolver
    kw, okw = {}, {}
    if case.startswith("arap"):
        nx, ny = 20, 14
        d = wl.arap_mesh_inputs(nx, ny)
        rng = np.random.RandomState(7)
        d["Position"] = d["Position"] + 0.2 * rng.randn(*d["Position"].shape).astype(np.float32)
        d["Angle"] = d["Angle"] + 0.3 * rng.randn(*d["Angle"].shape).astype(np.float32)
        dims, name, params = [nx * ny, len(d["V0"])], "arap_mesh_deformation", wl.arap_mesh_params(d)
        if case == "arap_jtjp":         # Jt[Jp]: J p stored per edge, transposed partials gathered per vertex
            kw["define_kwargs"] = okw = dict(jp=True)
        else:
            kw["schedule"] = case.split("_", 1)[1]
        devslots = range(2, 8)
    elif case.startswith("ba"):
        d = wl.bundle_adjustment_inputs(10, 200, 4)
        dims, name, params = [10, 200, len(d["oToC"])], "bundle_adjustment", wl.bundle_adjustment_params(d)
        mat = case == "ba_materialised"
        kw["define_kwargs"] = okw = dict(materialize=mat)
        devslots = range(0, 5)
    else:
        X, A = wl.minimal_inputs(40, 30)
        X = X + 0.1 * np.random.RandomState(0).randn(X.size).astype(np.float32)
        dims, name, params = [40, 30], "laplacian", [X, A]
        kw["define_kwargs"] = okw = dict(materialize=True)
        devslots = range(0, 2)
    p64 = [np.array(x, np.float64) if (hasattr(x, "dtype") and x.dtype == np.float32) else x for x in params]
    ref = _first_pcg_iteration_reference(name, dims, kind, p64, **okw)
    dp = [dev(p) if i in devslots else p for i, p in enumerate(params)]
    s = ThalloSolver(dims, name, kind, **kw)
    s.set_parameters(nIterations=1, lIterations=1)
    s.init(dp)
    s.step()
    n = ref["r"].size
    for vec in ("preconditioner", "Ap_X", "delta", "r"):
        got = s.read_vector(vec, n).astype(np.float64)
        scale = max(np.abs(ref[vec]).max(), 1e-12)
        assert np.abs(got - ref[vec]).max() <= 2e-5 * scale, (case, kind, vec, np.abs(got - ref[vec]).max() / scale)


def _trajectory(s, dp):
    s.init(dp)
    c, lin = [s.current_cost()], []
    while s.step():
        c.append(s.current_cost())
        lin.append(s.last_linear_iterations())
    c.append(s.current_cost())
    return c, lin


def test_arap_mesh_unsorted_edges_use_a_permutation():
    """Edge lists in arbitrary order: the adjacency of both endpoints needs a permutation; same costs as sorted."""
    from thallo_b200.api import ThalloSolver
    nx, ny = 20, 16
    d = wl.arap_mesh_inputs(nx, ny)
    dims = [nx * ny, len(d["V0"])]
    costs = []
    for shuffle in (False, True):
        d = wl.arap_mesh_inputs(nx, ny)
        if shuffle:
            perm = np.random.RandomState(0).permutation(len(d["V0"]))
            d["V0"], d["V1"] = d["V0"][perm].copy(), d["V1"][perm].copy()
        pg = wl.arap_mesh_params(d)
        dp = [dev(p) if i >= 2 else p for i, p in enumerate(pg)]
        s = ThalloSolver(dims, "arap_mesh_deformation", "gauss_newton")
        s.set_parameters(nIterations=3, lIterations=20)
        c, _ = _trajectory(s, dp)
        costs.append(c)
    for a, b in zip(*costs):
        assert abs(a - b) <= 1e-4 * max(abs(b), 1e-3), costs


@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
def test_bundle_adjustment_materialised_matches_oracle(kind):
    """Config 5 (small): sparse-materialised J (stored partials, J p per observation, transposed gathers per
    camera and per point); LM residual reset reproduces the reference's materialised-schedule quirk
    (A*delta = CtC*delta)."""
    from thallo_b200.api import ThalloSolver
    from _parity import assert_lm_parity, ulp_perturbed
    C_, P_ = 12, 300
    d = wl.bundle_adjustment_inputs(C_, P_, 5)
    dims = [C_, P_, len(d["oToC"])]
    nit, lit = 4, 25
    make_params = lambda: wl.bundle_adjustment_params(wl.bundle_adjustment_inputs(C_, P_, 5))
    make_oracle = lambda: OracleSolver(energies.load("bundle_adjustment"), dims, kind, np.float32, "residualwise", materialized=True)
    dp = [dev(p) for p in make_params()]
    s = ThalloSolver(dims, "bundle_adjustment", kind)
    assert s.lowered.desc["schedule"] == "gather" and s.lowered.desc["groups"][0]["materialize"] == 1
    s.set_parameters(nIterations=nit, lIterations=lit)
    c, lin = _trajectory(s, dp)
    if kind == "levenberg_marquardt":
        assert_lm_parity(c, lin, make_oracle, make_params, nit, lit, 1e-5, 1e-3)
    else:
        _, cref = _oracle_costs("bundle_adjustment", dims, kind, make_params(), nit, lit, materialized=True)
        pert = []
        for seed in (1, 2):
            dq = wl.bundle_adjustment_inputs(C_, P_, 5)
            dq["points"] = ulp_perturbed(dq["points"], seed)
            pert.append(_oracle_costs("bundle_adjustment", dims, kind, wl.bundle_adjustment_params(dq), nit, lit, materialized=True)[1])
        assert_costs_close(c, cref, 1e-5, 1e-3, None, pert)


def test_bundle_adjustment_wide_camera_gather():
    """Many observations per camera (>= 64): the camera space is gathered by one warp per camera; same
    trajectory as the matrix-free form of the energy (materialize=False)."""
    from thallo_b200.api import ThalloSolver
    C_, P_ = 6, 400
    d = wl.bundle_adjustment_inputs(C_, P_, 3)
    dims = [C_, P_, len(d["oToC"])]
    out = []
    for mat in (True, False):
        dp = [dev(p) for p in wl.bundle_adjustment_params(wl.bundle_adjustment_inputs(C_, P_, 3))]
        s = ThalloSolver(dims, "bundle_adjustment", "gauss_newton", define_kwargs=dict(materialize=mat))
        assert s.lowered.desc["gather"]["spaces"][0]["lanes"] == 32
        s.set_parameters(nIterations=3, lIterations=20)
        c, _ = _trajectory(s, dp)
        out.append(c)
    for a, b in zip(*out):
        assert abs(a - b) <= 1e-4 * max(abs(b), 1e-3), out


def test_config1b_materialised_laplacian_matches_oracle_and_golden():
    """tests/minimal/laplacian.t as written, with its materialize directives (:16-20): dense-offset endpoints."""
    import os
    from thallo_b200.api import ThalloSolver
    X, A = wl.minimal_inputs(256, 256)
    Xo = X.copy()
    o = OracleSolver(energies.load("laplacian"), [256, 256], "gauss_newton", np.float32, "residualwise", materialized=True)
    c_ref = o.solve([Xo, A])
    dX, dA = dev(X), dev(A)
    s = ThalloSolver([256, 256], "laplacian", "gauss_newton", define_kwargs=dict(materialize=True))
    assert s.lowered.desc["schedule"] == "gather"
    c = s.solve([dX, dA])
    assert abs(c - c_ref) <= 1e-5 * abs(c_ref)
    assert np.abs(dX.cpu().numpy() - Xo).max() < 1e-4
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "kat_minimal.npz"))["gold"]
    X, A = wl.minimal_inputs(512, 512)
    dX, dA = dev(X), dev(A)
    s = ThalloSolver([512, 512], "laplacian", "gauss_newton", define_kwargs=dict(materialize=True))
    s.solve([dX, dA])
    assert np.array_equal((dX.cpu().numpy().reshape(512, 512) * 255).astype(np.uint8), gold)


@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
def test_arap_mesh_jtjp_schedule_matches_oracle_and_inline(kind):
    """`r.reg.Jp:set_materialize(true)` (schedule Jt[Jp], thallo.t:4121): same operator as the
    inline schedule, so GN follows the inline trajectory to float32 rounding; in LM the residual reset
    (every 10th PCG iteration) leaves the group out of A delta as the reference does (no applyJTJ exists
    for it, gauss_newton.t:1058-1065), which the oracle restates per group."""
    from thallo_b200.api import ThalloSolver
    nx, ny = 24, 18
    nit, lit = 3, 25
    dims = [nx * ny, len(wl.arap_mesh_inputs(nx, ny)["V0"])]
    o, cref = _oracle_costs("arap_mesh_deformation", dims, kind, wl.arap_mesh_params(wl.arap_mesh_inputs(nx, ny)), nit, lit,
                            define_kwargs=dict(jp=True))
    p64 = [np.array(x, np.float64) if (i >= 2 and x.dtype == np.float32) else x
           for i, x in enumerate(wl.arap_mesh_params(wl.arap_mesh_inputs(nx, ny)))]
    _, cref64 = _oracle_costs("arap_mesh_deformation", dims, kind, p64, nit, lit, np.float64, define_kwargs=dict(jp=True))
    runs = {}
    for jp in (True, False):
        pg = wl.arap_mesh_params(wl.arap_mesh_inputs(nx, ny))
        dp = [dev(p) if i >= 2 else p for i, p in enumerate(pg)]
        s = ThalloSolver(dims, "arap_mesh_deformation", kind, define_kwargs=dict(jp=jp))
        assert [g["materialize"] for g in s.lowered.desc["groups"]] == ([0, 2] if jp else [0, 0])
        s.set_parameters(nIterations=nit, lIterations=lit)
        runs[jp] = _trajectory(s, dp)
        s.close()
    c, lin = runs[True]
    assert_costs_close(c, cref, 1e-5, 1e-3, cref64)
    assert abs(c[1] - cref[1]) <= 1e-5 * abs(cref[1])
    if kind == "levenberg_marquardt":
        assert lin == [it["n_lin"] for it in o.trace][:len(lin)]
    else:
        assert_costs_close(c, runs[False][0], 1e-5, 1e-3, cref64)
