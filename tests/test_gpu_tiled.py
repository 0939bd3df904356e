"""Parity of the tiled at-output schedule (th_pcg_a: TMA-staged shared-memory tiles, fused
p-update, hoisted invariants; th_pcg_b) against the oracle, through the C ABI.  Covers the TMA
and the cooperative-load variants, partial tiles, 3-D domains, sampled images, float64, and
bit-equality between the variants (they perform the same arithmetic on the same values).
Tolerances: 1e-5 relative on every cost of the trajectory in float32, 1e-10 in float64."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import energies
from oracle.solver import OracleSolver
from thallo_b200 import workloads as wl

pytestmark = pytest.mark.gpu


from _parity import dev, assert_costs_close as _close_nf, oracle_trajectory, gpu_trajectory as _trajectory_gpu


def _trajectory_oracle(name, dims, kind, params, dtype, nit, lit):
    return oracle_trajectory(name, dims, kind, params, dtype, nit, lit, "at_output")


def _close(c, cref, tol, floor, cref64=None):
    _close_nf(c, cref, tol, floor, cref64)


def _iw_params(W, H, dtype):
    p = wl.image_warping_params(wl.image_warping_inputs(W, H))
    return [np.array(x, dtype=dtype) if i < 5 else x for i, x in enumerate(p)]


@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
@pytest.mark.parametrize("size", [(100, 83), (75, 50)])       # partial tiles; 75 -> rows not 16-byte aligned -> cooperative loads
def test_image_warping_tiled_matches_oracle(kind, size):
    W, H = size
    o, cref = _trajectory_oracle("image_warping", [W, H], kind, _iw_params(W, H, np.float32), np.float32, 5, 35)
    _, cref64 = _trajectory_oracle("image_warping", [W, H], kind, _iw_params(W, H, np.float64), np.float64, 5, 35)
    s, c, lin, dp = _trajectory_gpu("image_warping", [W, H], kind, _iw_params(W, H, np.float32), range(5), np.float32, 5, 35)
    assert s.lowered.desc["tiled"] == 1 and s.lowered.desc["ncoef"] == 2
    _close(c, cref, 1e-5, 1e-3, cref64)       # float32 noise floor, see tests/_parity.py
    if kind == "levenberg_marquardt":
        assert lin == [it["n_lin"] for it in o.trace][:len(lin)]


def test_tma_and_cooperative_load_variants_are_bit_identical():
    W, H = 96, 72
    outs = []
    for no_tma in (False, True):
        if no_tma:
            os.environ["THALLO_B200_NO_TMA"] = "1"
        try:
            s, c, lin, dp = _trajectory_gpu("image_warping", [W, H], "levenberg_marquardt", _iw_params(W, H, np.float32),
                                            range(5), np.float32, 4, 30)
        finally:
            os.environ.pop("THALLO_B200_NO_TMA", None)
        outs.append((c, lin, dp[0].cpu().numpy().copy(), dp[1].cpu().numpy().copy()))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    assert np.array_equal(outs[0][2], outs[1][2]) and np.array_equal(outs[0][3], outs[1][3])


def test_hoisted_invariants_do_not_change_results():
    """cos/sin of the angle read from the precomputed coefficient image are the same bits as
    recomputing them at every tap."""
    W, H = 64, 48
    outs = []
    for hoist in (True, False):
        s, c, lin, dp = _trajectory_gpu("image_warping", [W, H], "gauss_newton", _iw_params(W, H, np.float32), range(5),
                                        np.float32, 3, 20, define_kwargs=dict(hoist=hoist))
        assert s.lowered.desc["ncoef"] == (2 if hoist else 0)
        outs.append((c, dp[0].cpu().numpy().copy()))
    _close(outs[0][0], outs[1][0], 1e-6, 1e-3)
    assert np.abs(outs[0][1] - outs[1][1]).max() < 1e-4


def test_image_warping_tiled_f64():
    W, H = 68, 44
    o, cref = _trajectory_oracle("image_warping", [W, H], "levenberg_marquardt", _iw_params(W, H, np.float64), np.float64, 5, 40)
    s, c, lin, dp = _trajectory_gpu("image_warping", [W, H], "levenberg_marquardt", _iw_params(W, H, np.float64), range(5),
                                    np.float64, 5, 40)
    _close(c, cref, 1e-10, 1e-6)
    assert lin == [it["n_lin"] for it in o.trace][:len(lin)]


@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
def test_volumetric_3d_tiled_matches_oracle(kind):
    W, H, D = 12, 10, 9
    po = wl.volumetric_params(wl.volumetric_inputs(W, H, D))
    o, cref = _trajectory_oracle("volumetric_mesh_deformation", [W, H, D], kind, po, np.float32, 4, 25)
    pg = wl.volumetric_params(wl.volumetric_inputs(W, H, D))
    s, c, lin, dp = _trajectory_gpu("volumetric_mesh_deformation", [W, H, D], kind, pg, range(4), np.float32, 4, 25)
    assert s.lowered.desc["tiled"] == 1 and s.lowered.desc["ncoef"] == 6
    _close(c, cref, 1e-5, 1e-3)
    assert np.abs(dp[0].cpu().numpy() - po[0]).max() < 2e-3


def test_optical_flow_tiled_matches_oracle():
    W, H = 64, 40
    po = wl.optical_flow_params(wl.optical_flow_inputs(W, H))
    o, cref = _trajectory_oracle("optical_flow", [W, H], "gauss_newton", po, np.float32, 2, 30)
    pg = wl.optical_flow_params(wl.optical_flow_inputs(W, H))
    s, c, lin, dp = _trajectory_gpu("optical_flow", [W, H], "gauss_newton", pg, range(2, 7), np.float32, 2, 30)
    assert s.lowered.desc["tiled"] == 1
    _close(c, cref, 1e-5, 1e-3)
    assert np.abs(dp[2].cpu().numpy() - po[2]).max() < 2e-3


# ---- computed arrays (shape_from_shading): precompute kernels, gradient images staged through the tiles (halo 2)
def _sfs_params(d, dtype):
    # images take the solver's precision; scalar Params are declared `float` in the energy
    # (shape_from_shading.t:4-19) and stay float32 host scalars in both precisions
    return [np.array(p, dtype=dtype) if (np.asarray(p).dtype == np.float32 and np.asarray(p).size > 1) else p
            for p in wl.sfs_params(d)]


@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
def test_shape_from_shading_reference_crop_matches_oracle(kind):
    """A crop of the reference's own example input (tests/golden/sfs_crop.npz), 6 nonlinear x 10 PCG
    iterations as examples/shape_from_shading/src/main.cpp:44-45 (60 x 10 there)."""
    d, W, H = wl.sfs_fixture_inputs(os.path.join(os.path.dirname(__file__), "golden", "sfs_crop.npz"))
    o, cref = _trajectory_oracle("shape_from_shading", [W, H], kind, _sfs_params(d, np.float32), np.float32, 6, 10)
    _, cref64 = _trajectory_oracle("shape_from_shading", [W, H], kind, _sfs_params(d, np.float64), np.float64, 6, 10)
    s, c, lin, dp = _trajectory_gpu("shape_from_shading", [W, H], kind, _sfs_params(d, np.float32), range(16, 21), np.float32, 6, 10)
    assert s.lowered.desc["tiled"] == 1 and len(s.lowered.desc["computed"]) == 2
    _close(c, cref, 1e-5, 1e-3, cref64)
    if kind == "levenberg_marquardt":
        assert lin == [it["n_lin"] for it in o.trace][:len(lin)]


def test_shape_from_shading_synthetic_f64():
    W, H = 72, 52
    d = wl.sfs_inputs(W, H)
    o, cref = _trajectory_oracle("shape_from_shading", [W, H], "gauss_newton", _sfs_params(d, np.float64), np.float64, 4, 10)
    s, c, lin, dp = _trajectory_gpu("shape_from_shading", [W, H], "gauss_newton", _sfs_params(d, np.float64), range(16, 21), np.float64, 4, 10)
    _close(c, cref, 1e-10, 1e-6)
    assert cref[-1] < cref[0]


# ---- the two variants of the tile operator (shifted-instance form / two-phase form with J p in shared memory) and the
# two specialisations of each (edge tiles with bounds predicates / interior tiles without): same operator, so the same
# float64 trajectory to 1e-10 and the oracle's float32 trajectory to the usual tolerance, on domains with partial tiles
# (edge variant only), and on domains large enough to contain interior tiles
@pytest.mark.parametrize("two_phase", ["0", "1"])
@pytest.mark.parametrize("case", ["sfs", "volume", "image_warping", "optical_flow"])
def test_tile_operator_variants_match_oracle_f64(case, two_phase):
    if case == "sfs":
        W, H = 150, 61
        name, dims, kind, slots, nit, lit = "shape_from_shading", [W, H], "gauss_newton", range(16, 21), 3, 10
        mk = lambda: _sfs_params(wl.sfs_inputs(W, H), np.float64)
    elif case == "volume":
        name, dims, kind, slots, nit, lit = "volumetric_mesh_deformation", [28, 27, 14], "gauss_newton", range(4), 2, 15
        mk = lambda: [np.array(p, np.float64) if np.asarray(p).size > 1 else p for p in wl.volumetric_params(wl.volumetric_inputs(28, 27, 14))]
    elif case == "image_warping":
        name, dims, kind, slots, nit, lit = "image_warping", [136, 60], "levenberg_marquardt", range(5), 3, 20
        mk = lambda: _iw_params(136, 60, np.float64)
    else:
        name, dims, kind, slots, nit, lit = "optical_flow", [132, 52], "gauss_newton", range(2, 7), 2, 20
        mk = lambda: [np.array(p, np.float64) if np.asarray(p).size > 1 else p for p in wl.optical_flow_params(wl.optical_flow_inputs(132, 52))]
    o, cref = _trajectory_oracle(name, dims, kind, mk(), np.float64, nit, lit)
    os.environ["THALLO_B200_TWO_PHASE"] = two_phase
    try:
        s, c, lin, dp = _trajectory_gpu(name, dims, kind, mk(), slots, np.float64, nit, lit)
    finally:
        os.environ.pop("THALLO_B200_TWO_PHASE", None)
    assert bool(s.lowered.generator.two_phase) == (two_phase == "1")
    _close(c, cref, 1e-10, 1e-6)
    if kind == "levenberg_marquardt":
        assert lin == [it["n_lin"] for it in o.trace][:len(lin)]


@pytest.mark.parametrize("case", ["sfs", "volume"])
def test_interior_tile_specialisation_matches(case):
    """THALLO_B200_EDGE_SPECIALIZE=1: interior tiles drop the bounds predicates (all true there).  The selects they guard
    fold away, which changes how the compiler contracts the sums into fused multiply-adds: same trajectory to float32
    rounding, not bit for bit (measured, profiles/r02g_*)."""
    outs = []
    for specialise in (False, True):
        if specialise:
            os.environ["THALLO_B200_EDGE_SPECIALIZE"] = "1"
        try:
            if case == "sfs":
                W, H = 150, 61
                s, c, lin, dp = _trajectory_gpu("shape_from_shading", [W, H], "gauss_newton", _sfs_params(wl.sfs_inputs(W, H), np.float32),
                                                range(16, 21), np.float32, 3, 10)
                outs.append((c, dp[16].cpu().numpy()))
            else:
                p = wl.volumetric_params(wl.volumetric_inputs(28, 27, 14))
                s, c, lin, dp = _trajectory_gpu("volumetric_mesh_deformation", [28, 27, 14], "gauss_newton", p, range(4), np.float32, 2, 15)
                outs.append((c, dp[0].cpu().numpy()))
        finally:
            os.environ.pop("THALLO_B200_EDGE_SPECIALIZE", None)
    for a, b in zip(outs[0][0], outs[1][0]):
        assert abs(a - b) <= 1e-5 * max(abs(a), 1e-3), (outs[0][0], outs[1][0])
    assert np.abs(outs[0][1] - outs[1][1]).max() <= 1e-4 * max(1.0, np.abs(outs[0][1]).max())
