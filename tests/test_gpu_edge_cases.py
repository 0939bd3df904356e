"""Edge cases of the hot path through the C ABI against the oracle: domains smaller than a tile or than the stencil
halo, single rows / columns, graphs with isolated vertices, self-loops, duplicate and unsorted edges, a single
edge, every unknown excluded.  Same tolerances as the other parity tests (1e-5 float32, 1e-10 float64)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

import energies
from oracle.solver import OracleSolver
from thallo_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _both(name, dims, kind, params, device_slots, dtype, nit, lit, mode, tol, **kw):
    from thallo_b200.api import ThalloSolver
    po = [np.array(p, copy=True) for p in params]
    o = OracleSolver(energies.load(name), dims, kind, dtype, mode)
    o.set("nIterations", nit); o.set("lIterations", lit)
    o.init(po)
    cref = [o.current_cost()]
    while o.step(po):
        cref.append(o.current_cost())
    cref.append(o.current_cost())
    dp = [dev(p) if i in device_slots else p for i, p in enumerate(params)]
    s = ThalloSolver(dims, name, kind, double=(dtype == np.float64), **kw)
    s.set_parameters(nIterations=nit, lIterations=lit)
    s.init(dp)
    c = [s.current_cost()]
    while s.step():
        c.append(s.current_cost())
    c.append(s.current_cost())
    s.close()
    assert len(c) == len(cref), (c, cref)
    for a, b in zip(c, cref):
        assert np.isfinite(a) and abs(a - b) <= tol * max(abs(b), 1e-6), (dims, c, cref)
    return dp, po


@pytest.mark.parametrize("dims", [(1, 1), (1, 7), (9, 1), (2, 2), (3, 5), (33, 2), (31, 9)])
@pytest.mark.parametrize("schedule", ["at_output", "residualwise"])
def test_image_domains_smaller_than_a_tile(dims, schedule):
    W, H = dims
    rs = np.random.RandomState(W * 31 + H)
    A = rs.rand(W * H)
    X = A + 0.3 * rs.randn(W * H)
    dp, po = _both("laplacian", [W, H], "gauss_newton", [X, A], [0, 1], np.float64, 3, 8, schedule, 1e-10, schedule=schedule)
    assert np.abs(dp[0].cpu().numpy() - po[0]).max() <= 1e-9


@pytest.mark.parametrize("dims", [(5, 4), (34, 3), (8, 9)])
@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
def test_image_warping_on_tiny_images(dims, kind):
    W, H = dims
    d = wl.image_warping_inputs(W, H)
    rs = np.random.RandomState(W + H)
    d["Offset"] = d["Offset"] + 0.5 * rs.randn(*d["Offset"].shape).astype(np.float32)      # the synthetic handles need a larger image
    d["Angle"] = d["Angle"] + 0.2 * rs.randn(*d["Angle"].shape).astype(np.float32)
    d["Mask"] = np.zeros_like(d["Mask"])
    p = [np.array(x, np.float64) if i < 5 else x for i, x in enumerate(wl.image_warping_params(d))]
    _both("image_warping", [W, H], kind, p, range(5), np.float64, 3, 12, "at_output", 1e-10)


@pytest.mark.parametrize("dims", [(3, 3, 2), (9, 2, 5), (2, 9, 1)])
def test_volume_smaller_than_a_tile(dims):
    W, H, D = dims
    p = wl.volumetric_params(wl.volumetric_inputs(W, H, D))
    p = [np.array(x, np.float64) if i < 4 else x for i, x in enumerate(p)]
    _both("volumetric_mesh_deformation", [W, H, D], "gauss_newton", p, range(4), np.float64, 2, 10, "at_output", 1e-10)


def _graph_case(case):
    rs = np.random.RandomState(4)
    N = 40
    if case == "isolated_vertices":          # vertices 30..39 have no edge at all
        v0 = rs.randint(0, 30, 90); v1 = rs.randint(0, 30, 90)
    elif case == "self_loops_and_duplicates":
        v0 = rs.randint(0, N, 60); v1 = rs.randint(0, N, 60)
        v1[:10] = v0[:10]                                        # self loops: residual X(v) - X(v) = 0
        v0 = np.concatenate([v0, v0[10:30]]); v1 = np.concatenate([v1, v1[10:30]])       # duplicates
    elif case == "single_edge":
        v0 = np.array([3]); v1 = np.array([17])
    else:                                                        # star: one vertex of degree N - 1
        v0 = np.zeros(N - 1, np.int64); v1 = np.arange(1, N)
    A = rs.rand(N)
    X = A + 0.5 * rs.randn(N)
    return N, X, A, v0.astype(np.int32), v1.astype(np.int32)


@pytest.mark.parametrize("case", ["isolated_vertices", "self_loops_and_duplicates", "single_edge", "star"])
@pytest.mark.parametrize("schedule", ["gather", "residualwise"])
def test_irregular_graphs(case, schedule):
    N, X, A, v0, v1 = _graph_case(case)
    dp, po = _both("graph_laplacian", [N, len(v0)], "gauss_newton", [X, A, v0, v1], range(4), np.float64, 3, 15, "residualwise",
                   1e-10, schedule=schedule)
    assert np.abs(dp[0].cpu().numpy() - po[0]).max() <= 1e-9


@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
def test_every_unknown_excluded_leaves_the_unknowns_untouched(kind):
    """Mask != 0 everywhere: Exclude() removes every unknown (image_warping.t:14-15); nothing may be written and
    the cost must stay what it was."""
    from thallo_b200.api import ThalloSolver
    W, H = 40, 24
    d = wl.image_warping_inputs(W, H)
    d["Mask"] = np.ones_like(d["Mask"])
    p = wl.image_warping_params(d)
    dp = [dev(x) if i < 5 else x for i, x in enumerate(p)]
    before = [dp[0].clone(), dp[1].clone()]
    s = ThalloSolver([W, H], "image_warping", kind)
    s.set_parameters(nIterations=3, lIterations=10)
    s.init(dp)
    c0 = s.current_cost()
    n = 0
    while s.step() and n < 10:
        n += 1
    assert s.current_cost() == c0
    assert torch.equal(dp[0], before[0]) and torch.equal(dp[1], before[1])
    s.close()
