"""Shared helpers of the GPU parity tests.

Tolerance rule (BASELINE.json north_star: cost within 1e-5 relative in float32, 1e-10 in float64):
every cost of the GPU trajectory must agree with the float32 oracle to `tol` relative, except
where float32 arithmetic itself cannot support that.  Some steps are ill-conditioned in float32:
a Gauss-Newton step from a far start that drops the cost by four orders of magnitude moves by
1e-2 when the *oracle* is merely re-run in float64, and graph domains accumulate J^T J p with
float atomics in arbitrary order (as the reference does).  The float32 noise floor of iteration i
is measured as the relative distance between the float32 and the float64 oracle over iterations
i-1..i+1; the GPU may differ from the float32 oracle by at most NOISE_FACTOR times that floor.
Where the floor is below `tol` (every well-conditioned step, all float64 runs) the plain 1e-5 /
1e-10 rule applies unchanged.
"""
import numpy as np
import torch

import energies
from oracle.solver import OracleSolver


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def oracle_trajectory(name, dims, kind, params, dtype, nit, lit, mode):
    o = OracleSolver(energies.load(name), dims, kind, dtype, mode)
    o.set("nIterations", nit); o.set("lIterations", lit)
    o.init(params)
    costs = [o.current_cost()]
    while o.step(params):
        costs.append(o.current_cost())
    costs.append(o.current_cost())
    return o, costs


def gpu_trajectory(name, dims, kind, params, device_slots, dtype, nit, lit, **kw):
    from thallo_b200.api import ThalloSolver
    dp = [dev(p) if i in device_slots else p for i, p in enumerate(params)]
    s = ThalloSolver(dims, name, kind, double=(dtype == np.float64), **kw)
    s.set_parameters(nIterations=nit, lIterations=lit)
    s.init(dp)
    costs, lin = [s.current_cost()], []
    while s.step():
        costs.append(s.current_cost())
        lin.append(s.last_linear_iterations())
    costs.append(s.current_cost())
    return s, costs, lin, dp


NOISE_FACTOR = 8.0


def assert_costs_close(c, cref, tol, floor, cref64=None):
    assert len(c) == len(cref), (c, cref)
    noise = [0.0] * len(cref)
    if cref64 is not None:
        assert len(cref64) == len(cref), (cref, cref64)
        rel = [abs(a - b) / max(abs(b), floor) for a, b in zip(cref, cref64)]
        noise = [max(rel[max(0, i - 1):i + 2]) for i in range(len(rel))]
    for i, (a, b) in enumerate(zip(c, cref)):
        allowed = max(tol, NOISE_FACTOR * noise[i]) * max(abs(b), floor)
        assert abs(a - b) <= allowed, (i, a, b, allowed, c, cref)
