"""Whole front-end pipeline on the reference's OTHER energy files (beyond the five configured workloads), read as
written: `.t` text -> DSL -> symbolic AD -> lowered operators, evaluated on the CPU and compared with the oracle's
dual-number J on seeded random inputs (J^T F, per-access diagonal, J^T J p to 1e-10 in float64), and NVRTC-compiled for
sm_100a.  Needs the reference checkout (skipped on the GPU box); nothing is copied from it."""
import os

import numpy as np
import pytest

from oracle import npdsl
from thallo_b200.frontend import codegen, dsl, interp, tlang

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")


def random_params(define, dims, seed=0):
    """Seeded inputs for whatever Inputs{} the energy declares."""
    L = dsl.build_spec(define, dims)
    rs = np.random.RandomState(seed)
    n = 1 + max([im.pidx for im in L.images if im.pidx >= 0] + [s.pidx for s in L.sparses] + [p.pidx for p in L.params])
    params = [None] * n
    for im in L.images:
        if im.pidx < 0:
            continue
        shape = (im.elements, im.channels) if im.channels > 1 else (im.elements,)
        if im.ctype == "uchar":
            params[im.pidx] = (rs.rand(*shape) < 0.8).astype(np.uint8)
        elif im.ctype == "int":
            params[im.pidx] = rs.randint(0, 3, shape).astype(np.int32)
        else:
            params[im.pidx] = 0.5 + rs.rand(*shape)
    for s in L.sparses:
        E = int(np.prod([d.size for d in s.frm]))
        params[s.pidx] = rs.randint(0, s.to[0].size, E).astype(np.int32)
    for p in L.params:
        params[p.pidx] = np.array([0.3 + rs.rand()], np.float32 if p.ctype == "float" else np.int32)
    return params


CASES = [
    ("examples/cotangent_mesh_smoothing/cotangent_mesh_smoothing.t", [30, 70], "gather"),
    ("examples/robust_nonrigid_alignment/robust_nonrigid_alignment.t", [30, 70], "gather"),
    ("examples/embedded_mesh_deformation/embedded_mesh_deformation.t", [30, 70], "gather"),
    ("examples/poisson_image_editing/poisson_image_editing.t", [14, 11], "at_output"),
    ("tests/minimal_exclude/minimal_exclude.t", [14, 11], "at_output"),
    ("tests/minimal_materialize/minimal_materialize.t", [14, 11], "at_output"),
    ("tests/create_delete_cycle/laplacian.t", [9, 13], "at_output"),
    ("tests/energy_unit_tests/laplacian.t", [9, 13], "at_output"),
    ("tests/dense/curveFitting.t", [20, 3, 40], "gather"),
    ("examples/sparse_bundle_fusion/bundle_fusion_solve.t", [8, 8, 12, 60], "gather"),     # SE(3) poses stored per frame, fetched per correspondence
]


@pytest.mark.parametrize("path,dims,schedule", CASES, ids=[c[0].split("/")[-2] + "/" + c[0].split("/")[-1] for c in CASES])
@pytest.mark.parametrize("kind", ["gauss_newton", "levenberg_marquardt"])
def test_reference_energy_file_lowers_correctly(path, dims, schedule, kind):
    define = tlang.load(os.path.join(REF, path))
    params = random_params(define, dims)
    low = codegen.lower(define, dims, kind, "energy", True)
    assert low.desc["schedule"] == schedule
    gen = low.generator
    Ln, F, J = npdsl.evaluate(define, dims, params, np.float64)
    p = np.random.RandomState(5).randn(J.shape[1])
    JT = J.T.tocsr()
    g0, d0, o0 = JT @ F, Ln.diag_sq, JT @ (J @ p)
    if schedule == "at_output":
        g, d, o = interp.unknownwise(gen, params, p)
    else:
        o = interp.gather_apply(gen, params, p)
        r, d = interp.gather_jtf(gen, params)
        g = -r
    for got, want in ((g, g0), (d, d0), (o, o0)):
        assert np.abs(got - want).max() <= 1e-10 * max(1.0, np.abs(want).max())
    if kind == "gauss_newton":
        from thallo_b200 import api
        ok, log, size = api.compile_only(codegen.lower(define, dims, kind, "energy").source)
        assert ok and size > 0, log[-2000:]
