"""Host-side check of the lowering: the generated at-output operators (evaluated by the
NumPy DAG interpreter) equal the oracle's J^T F, diag(J^T J) and J^T J p (J from
dual-number AD) on seeded inputs, in float64 to 1e-10."""
import numpy as np
import pytest

import energies
from oracle.npdsl import evaluate
from thallo_b200 import workloads as wl
from thallo_b200.frontend import codegen, interp


def _check(name, dims, params, kw=None):
    low = codegen.lower(energies.load(name), dims, "gauss_newton", name, True, "at_output", **(kw or {}))
    gen = low.generator
    L, F, J = evaluate(energies.load(name), dims, params, np.float64, **(kw or {}))
    rng = np.random.RandomState(3)
    p = rng.randn(J.shape[1])
    g, d, out = interp.unknownwise(gen, params, p)
    JT = J.T.tocsr()
    g0 = JT @ F
    d0 = np.asarray(J.multiply(J).sum(axis=0)).reshape(-1)
    o0 = JT @ (J @ p)
    sc = max(1.0, np.abs(g0).max())
    assert np.abs(g - g0).max() <= 1e-10 * sc
    assert np.abs(d - d0).max() <= 1e-10 * max(1.0, d0.max())
    assert np.abs(out - o0).max() <= 1e-10 * max(1.0, np.abs(o0).max())
    if gen.tiled:
        # the two-phase form of the tile operator (J p per residual position first, then the transposed products)
        import os
        os.environ["THALLO_B200_TWO_PHASE"] = "1"
        try:
            gen2 = codegen.lower(energies.load(name), dims, "gauss_newton", name, True, "at_output", **(kw or {})).generator
        finally:
            os.environ.pop("THALLO_B200_TWO_PHASE")
        assert gen2.two_phase
        out2 = interp.unknownwise_two_phase(gen2, params, p)
        assert np.abs(out2 - o0).max() <= 1e-10 * max(1.0, np.abs(o0).max())


def test_laplacian_at_output_matches_oracle():
    X, A = wl.minimal_inputs(37, 29)
    X = X + 0.1 * np.random.RandomState(0).randn(X.size).astype(np.float32)
    _check("laplacian", [37, 29], [X.astype(np.float64), A.astype(np.float64)])


def test_laplacian_committed_variant():
    X, A = wl.minimal_inputs(16, 16)
    _check("laplacian", [16, 16], [X.astype(np.float64) * 1.3, A.astype(np.float64)], dict(variant="committed"))


def test_image_warping_at_output_matches_oracle():
    W, H = 48, 40
    d = wl.image_warping_inputs(W, H)
    rng = np.random.RandomState(1)
    d["Offset"] = d["Offset"] + rng.randn(*d["Offset"].shape).astype(np.float32)
    d["Angle"] = d["Angle"] + 0.3 * rng.randn(*d["Angle"].shape).astype(np.float32)
    # un-mask the border so that the "ghost residual" handling is exercised too
    d["Mask"] = d["Mask"].reshape(H, W).copy()
    d["Mask"][0, :] = 0
    d["Mask"][:, 0] = 0
    d["Mask"] = d["Mask"].reshape(-1)
    params = [np.asarray(p, np.float64) for p in wl.image_warping_params(d)]
    _check("image_warping", [W, H], params)


def test_volumetric_at_output_matches_oracle():
    W, H, D = 6, 5, 4
    rng = np.random.RandomState(2)
    n = W * H * D
    zz, yy, xx = np.mgrid[0:D, 0:H, 0:W]
    ur = np.stack([xx, yy, zz], -1).reshape(n, 3).astype(np.float64)
    off = ur + 0.2 * rng.randn(n, 3)
    ang = 0.2 * rng.randn(n, 3)
    cons = np.full((n, 3), -1e6)
    cons[::7] = ur[::7] + 0.5
    params = [off, ang, ur, cons, np.array([1.0]), np.array([np.sqrt(0.05)])]
    _check("volumetric_mesh_deformation", [W, H, D], params)


def test_optical_flow_at_output_matches_oracle():
    W, H = 24, 20
    rng = np.random.RandomState(4)
    n = W * H
    X = 0.7 * rng.randn(n, 2)
    I, Ih, Ix, Iy = (rng.rand(n) for _ in range(4))
    params = [np.array([np.sqrt(10.0)]), np.array([np.sqrt(0.1)]), X, I, Ih, Ix, Iy]
    _check("optical_flow", [W, H], params)


# ---- gather schedule (graph domains / materialised Jacobians): endpoint functions summed per unknown
def _check_gather(name, dims, params, kw=None):
    low = codegen.lower(energies.load(name), dims, "gauss_newton", name, True, "gather", **(kw or {}))
    gen = low.generator
    L, F, J = evaluate(energies.load(name), dims, params, np.float64, **(kw or {}))
    p = np.random.RandomState(5).randn(J.shape[1])
    o0 = J.T.tocsr() @ (J @ p)
    for mat in (False, True):       # matrix-free endpoint functions, then the stored-value (materialised J) functions
        out = interp.gather_apply(gen, params, p, materialised=mat)
        assert np.abs(out - o0).max() <= 1e-10 * max(1.0, np.abs(o0).max()), mat
    # gathered PCGInit1: -J^T F and diag(J^T J)
    r, dg = interp.gather_jtf(gen, params)
    g0 = -(J.T.tocsr() @ F)
    d0 = np.asarray(J.multiply(J).sum(axis=0)).reshape(-1)
    assert np.abs(r - g0).max() <= 1e-10 * max(1.0, np.abs(g0).max())
    assert np.abs(dg - d0).max() <= 1e-10 * max(1.0, d0.max())
    return low


def test_arap_mesh_gather_matches_oracle():
    nx, ny = 9, 7
    d = wl.arap_mesh_inputs(nx, ny)
    rng = np.random.RandomState(6)
    d["Position"] = d["Position"] + 0.3 * rng.randn(*d["Position"].shape).astype(np.float32)
    d["Angle"] = d["Angle"] + 0.4 * rng.randn(*d["Angle"].shape).astype(np.float32)
    params = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.arap_mesh_params(d)]
    low = _check_gather("arap_mesh_deformation", [nx * ny, len(d["V0"])], params)
    assert low.desc["schedule"] == "gather" and len(low.desc["gather"]["sparse_endpoints"]) == 2


def test_graph_laplacian_gather_matches_oracle():
    X, A, v0, v1 = wl.minimal_graph_inputs(64)
    _check_gather("graph_laplacian", [64, 63], [X.astype(np.float64) * 1.1, A.astype(np.float64), v0, v1])


def test_bundle_adjustment_gather_matches_oracle():
    d = wl.bundle_adjustment_inputs(6, 40, 4)
    params = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.bundle_adjustment_params(d)]
    low = _check_gather("bundle_adjustment", [6, 40, len(d["oToC"])], params)
    g = low.desc["gather"]
    assert g["groups"][0]["nnzp"] == 24 and low.desc["groups"][0]["materialize"] == 1


def test_materialised_laplacian_gather_matches_oracle():
    X, A = wl.minimal_inputs(13, 11)
    low = _check_gather("laplacian", [13, 11], [X.astype(np.float64) * 1.3, A.astype(np.float64)],
                        dict(variant="committed", materialize=True))
    assert all(g["materialize"] for g in low.desc["groups"])


# ---- computed arrays (`exp:get()`): value + gradient images, chain rule through the gradient image
def _sfs_params64(d):
    return [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.sfs_params(d)]


def test_shape_from_shading_synthetic_matches_oracle():
    W, H = 40, 32
    d = wl.sfs_inputs(W, H)
    low_check = _check("shape_from_shading", [W, H], _sfs_params64(d))


def test_shape_from_shading_reference_crop_matches_oracle():
    import os
    d, W, H = wl.sfs_fixture_inputs(os.path.join(os.path.dirname(__file__), "golden", "sfs_crop.npz"))
    _check("shape_from_shading", [W, H], _sfs_params64(d))


def test_shape_from_shading_lowering_has_two_computed_arrays():
    low = codegen.lower(energies.load("shape_from_shading"), [64, 48], "gauss_newton", "shape_from_shading")
    assert low.desc["computed"] == [dict(elements=64 * 48, ngrad=3), dict(elements=64 * 48, ngrad=0)]
    assert low.desc["tiled"] == 1 and low.desc["tile"]["halo"][:2] == [2, 2]
    assert "th_precompute_c0" in low.source or "TH_COMPUTED_LIST(X) X(0) X(1)" in low.source


# ---- Jt[Jp] schedule (APPLY_SEPARATELY, thallo.t:4121; a10 of SURVEY 8): J p stored per residual row, transposed
# partials gathered per unknown
def test_arap_mesh_jtjp_schedule_matches_oracle():
    nx, ny = 8, 7
    d = wl.arap_mesh_inputs(nx, ny)
    rng = np.random.RandomState(9)
    d["Position"] = d["Position"] + 0.3 * rng.randn(*d["Position"].shape).astype(np.float32)
    d["Angle"] = d["Angle"] + 0.4 * rng.randn(*d["Angle"].shape).astype(np.float32)
    params = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.arap_mesh_params(d)]
    low = _check_gather("arap_mesh_deformation", [nx * ny, len(d["V0"])], params, dict(jp=True))
    assert [g["materialize"] for g in low.desc["groups"]] == [0, 2]          # fit inline, reg Jt[Jp]
    assert "th_applyj_g" in low.source or "TH_JP_LIST(X) X(1)" in low.source
    assert "jtp_ep" in low.source


def test_graph_laplacian_jtjp_schedule_matches_oracle():
    X, A, v0, v1 = wl.minimal_graph_inputs(48)
    low = _check_gather("graph_laplacian", [48, 47], [X.astype(np.float64) * 0.9, A.astype(np.float64), v0, v1], dict(jp=True))
    assert [g["materialize"] for g in low.desc["groups"]] == [0, 2]


# ---- Jacobian export (ThalloB200_PlanExportJacobian): layout + assembly against the oracle's J
@pytest.mark.parametrize("case", ["image_warping", "arap_mesh", "bundle_adjustment"])
def test_exported_jacobian_layout_assembles_to_the_oracle_jacobian(case):
    from thallo_b200 import api
    if case == "image_warping":
        W, H = 20, 14
        d = wl.image_warping_inputs(W, H)
        rng = np.random.RandomState(1)
        d["Offset"] = d["Offset"] + rng.randn(*d["Offset"].shape).astype(np.float32)
        d["Angle"] = d["Angle"] + 0.3 * rng.randn(*d["Angle"].shape).astype(np.float32)
        name, dims, params = "image_warping", [W, H], [np.asarray(p, np.float64) for p in wl.image_warping_params(d)]
    elif case == "arap_mesh":
        d = wl.arap_mesh_inputs(7, 6)
        name, dims = "arap_mesh_deformation", [42, len(d["V0"])]
        params = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.arap_mesh_params(d)]
    else:
        d = wl.bundle_adjustment_inputs(5, 30, 3)
        name, dims = "bundle_adjustment", [5, 30, len(d["oToC"])]
        params = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.bundle_adjustment_params(d)]
    low = codegen.lower(energies.load(name), dims, "gauss_newton", name, True)
    _, F, J = evaluate(energies.load(name), dims, params, np.float64)
    Jx = api.assemble_jacobian(low.desc, lambda gi, n: interp.jacobian_entries(low.generator, params, gi))
    assert Jx.shape == J.shape
    assert abs(Jx - J).max() <= 1e-12 * max(1.0, abs(J).max())


# ---- two-pass operator on image domains: gather schedule with every group in the Jt[Jp] form (lower(..., jp_all=True))
@pytest.mark.parametrize("name", ["shape_from_shading", "volumetric_mesh_deformation", "image_warping"])
def test_two_pass_operator_on_image_domains_matches_oracle(name):
    if name == "shape_from_shading":
        dims = [40, 32]
        params = _sfs_params64(wl.sfs_inputs(*dims))
    elif name == "volumetric_mesh_deformation":
        dims = [6, 5, 4]
        params = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.volumetric_params(wl.volumetric_inputs(*dims))]
    else:
        dims = [24, 20]
        d = wl.image_warping_inputs(*dims)
        d["Angle"] = d["Angle"] + 0.3 * np.random.RandomState(1).randn(*d["Angle"].shape).astype(np.float32)
        params = [np.asarray(p, np.float64) for p in wl.image_warping_params(d)]
    low = codegen.lower(energies.load(name), dims, "gauss_newton", name, True, "gather", jp_all=True)
    assert all(g["materialize"] == 2 for g in low.desc["groups"])
    _, F, J = evaluate(energies.load(name), dims, params, np.float64)
    p = np.random.RandomState(5).randn(J.shape[1])
    want = J.T.tocsr() @ (J @ p)
    got = interp.gather_apply(low.generator, params, p)
    assert np.abs(got - want).max() <= 1e-10 * max(1.0, np.abs(want).max())
