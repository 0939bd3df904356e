"""Parity of the residualwise (graph-domain) schedule against the oracle through the C ABI:
config 4b's ARAP mesh energy (arap_mesh_deformation.t) on a small triangulated grid.
Tolerance 1e-5 relative on every cost of the trajectory (float32), identical LM iteration counts."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

import energies
from oracle.solver import OracleSolver
from thallo_b200 import workloads as wl

pytestmark = pytest.mark.gpu


from _parity import dev, assert_costs_close, oracle_trajectory


@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
def test_arap_mesh_matches_oracle(kind):
    from thallo_b200.api import ThalloSolver
    nx, ny = 24, 18
    d = wl.arap_mesh_inputs(nx, ny)
    dims = [nx * ny, len(d["V0"])]
    po = wl.arap_mesh_params(d)
    o = OracleSolver(energies.load("arap_mesh_deformation"), dims, kind, np.float32, "residualwise")
    o.set("nIterations", 4); o.set("lIterations", 30)
    o.init(po)
    cref = [o.current_cost()]
    while o.step(po):
        cref.append(o.current_cost())
    cref.append(o.current_cost())

    pg = wl.arap_mesh_params(wl.arap_mesh_inputs(nx, ny))
    dp = [dev(p) if i >= 2 else p for i, p in enumerate(pg)]
    s = ThalloSolver(dims, "arap_mesh_deformation", kind)
    s.set_parameters(nIterations=4, lIterations=30)
    s.init(dp)
    c, lin = [s.current_cost()], []
    while s.step():
        c.append(s.current_cost())
        lin.append(s.last_linear_iterations())
    c.append(s.current_cost())
    p64 = [np.array(x, np.float64) if (i >= 2 and x.dtype == np.float32) else x for i, x in enumerate(wl.arap_mesh_params(wl.arap_mesh_inputs(nx, ny)))]
    _, cref64 = oracle_trajectory("arap_mesh_deformation", dims, kind, p64, np.float64, 4, 30, "residualwise")
    # float atomics in the scatter (like the reference's): float32 noise floor, see tests/_parity.py
    assert_costs_close(c, cref, 1e-5, 1e-3, cref64)
    if kind == "levenberg_marquardt":
        assert lin == [it["n_lin"] for it in o.trace][:len(lin)]
    assert np.abs(dp[2].cpu().numpy() - po[2]).max() < 2e-3
