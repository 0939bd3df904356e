"""The plain-C restatement of the reference CPU path (oracle/iw_cpu.c, the CPU baseline that
bench.py times) against the NumPy dual-number oracle, which is itself pinned on the reference's
golden PNGs (tests/test_oracle_golden.py).  Tolerance: 1e-5 relative on every cost of the
trajectory with double accumulators; the float-accumulator build (the reference CPU path's own
arithmetic: one running float sum per dot product) is held to 1e-3."""
import numpy as np
import pytest

import energies
from oracle import iw_cpu
from oracle.solver import OracleSolver
from thallo_b200 import workloads as wl


def _oracle(W, H, kind, nit, lit):
    d = wl.image_warping_inputs(W, H)
    po = wl.image_warping_params(d)
    o = OracleSolver(energies.load("image_warping"), [W, H], kind, np.float32, "at_output")
    o.set("nIterations", nit); o.set("lIterations", lit)
    o.init(po)
    c0 = o.current_cost()
    while o.step(po):
        pass
    return o, c0, o.current_cost(), po


@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
def test_c_restatement_matches_numpy_oracle(kind):
    W, H, nit, lit = 72, 56, 5, 30
    o, c0, cfin, po = _oracle(W, H, kind, nit, lit)
    d = wl.image_warping_inputs(W, H)
    r = iw_cpu.solve(W, H, d, kind, acc64=True, nIterations=nit, lIterations=lit)
    assert abs(r["costs"][0] - c0) <= 1e-5 * abs(c0)
    assert abs(r["costs"][-1] - cfin) <= 1e-5 * abs(cfin), (r["costs"], cfin)
    if kind == "levenberg_marquardt":
        ref_new = [it["cost"] for it in o.trace if "cost" in it]
        got = r["costs"][1:-1]
        assert len(got) == len(ref_new)
        for a, b in zip(got, ref_new):
            assert abs(a - b) <= 1e-5 * abs(b), (got, ref_new)
        assert r["n_lin"] == [it["n_lin"] for it in o.trace]
    assert np.abs(d["Offset"].reshape(-1) - po[0].reshape(-1)).max() < 2e-3


def test_c_restatement_float_accumulators_close():
    W, H = 72, 56
    o, c0, cfin, po = _oracle(W, H, "gauss_newton", 4, 25)
    d = wl.image_warping_inputs(W, H)
    r = iw_cpu.solve(W, H, d, "gauss_newton", acc64=False, nIterations=4, lIterations=25)
    assert abs(r["costs"][-1] - cfin) <= 1e-3 * abs(cfin)


def test_bounded_sample_stops_after_requested_pcg_iterations():
    d = wl.image_warping_inputs(64, 64)
    r = iw_cpu.solve(64, 64, d, "levenberg_marquardt", max_pcg=7, nIterations=8, lIterations=100)
    assert r["n_pcg"] == 7
