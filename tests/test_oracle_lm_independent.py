"""The LM blocks of the oracle have no golden in the reference (LM is unreachable as committed), so their
trajectory is pinned only by our two restatements agreeing.  This adds an independent anchor for WHERE they converge:
SciPy's trust-region least-squares solver on the same residuals and Jacobian (oracle/npdsl.py) must reach the same
minimum as the oracle's GN and LM loops from the same start."""
import numpy as np
import pytest
from scipy.optimize import least_squares

import energies
from oracle.npdsl import evaluate
from oracle.solver import OracleSolver
from thallo_b200 import workloads as wl


def _problem(W, H):
    d = wl.image_warping_inputs(W, H)
    rs = np.random.RandomState(3)
    d["Offset"] = d["Offset"] + 0.4 * rs.randn(*d["Offset"].shape).astype(np.float32)
    d["Mask"] = np.zeros_like(d["Mask"])
    d["Constraints"] = d["UrShape"] + 0.5 * rs.randn(*d["UrShape"].shape).astype(np.float32)      # every pixel constrained: well posed
    return [np.asarray(p, np.float64) for p in wl.image_warping_params(d)]


@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
def test_oracle_converges_to_the_minimum_an_independent_solver_finds(kind):
    W, H = 14, 11
    define = energies.load("image_warping")
    p0 = _problem(W, H)
    n_off = p0[0].size

    def unpack(x):
        p = [np.array(a, copy=True) for a in p0]
        p[0] = x[:n_off].reshape(p0[0].shape)
        p[1] = x[n_off:].reshape(p0[1].shape)
        return p

    def fun(x):
        return evaluate(define, [W, H], unpack(x), np.float64)[1]

    def jac(x):
        return evaluate(define, [W, H], unpack(x), np.float64)[2].toarray()
    x0 = np.concatenate([p0[0].reshape(-1), p0[1].reshape(-1)])
    ref = least_squares(fun, x0, jac=jac, method="trf", xtol=1e-14, ftol=1e-14, gtol=1e-12)
    ref_cost = 0.5 * float(ref.fun @ ref.fun)
    o = OracleSolver(define, [W, H], kind, np.float64, "at_output")
    o.set("nIterations", 40); o.set("lIterations", 200)
    o.set("function_tolerance", 0.0); o.set("q_tolerance", 1e-9)
    params = [np.array(a, copy=True) for a in p0]
    o.init(params)
    while o.step(params):
        pass
    cost = o.current_cost()
    assert ref.cost == pytest.approx(ref_cost)
    assert abs(cost - ref_cost) <= 1e-6 * ref_cost, (kind, cost, ref_cost)
    assert np.abs(np.concatenate([params[0].reshape(-1), params[1].reshape(-1)]) - ref.x).max() <= 1e-4
