"""The Jacobian the device materialises (ThalloB200_PlanExportJacobian: the reference's precomputeJ layout,
gauss_newton.t:325-487) against the oracle's dual-number J, entry by entry: partial derivatives, column indices of
dense and sparse accesses, out-of-domain accesses.  1e-5 relative to the largest entry in float32, 1e-12 in float64."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

import energies
from oracle.npdsl import evaluate
from thallo_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _case(case, dtype):
    if case == "image_warping":
        W, H = 45, 26
        d = wl.image_warping_inputs(W, H)
        rng = np.random.RandomState(1)
        d["Offset"] = d["Offset"] + rng.randn(*d["Offset"].shape).astype(np.float32)
        d["Angle"] = d["Angle"] + 0.3 * rng.randn(*d["Angle"].shape).astype(np.float32)
        name, dims, params, kw = "image_warping", [W, H], wl.image_warping_params(d), {}
    elif case == "arap_mesh":
        d = wl.arap_mesh_inputs(13, 9)
        rng = np.random.RandomState(2)
        d["Angle"] = d["Angle"] + 0.3 * rng.randn(*d["Angle"].shape).astype(np.float32)
        name, dims, params, kw = "arap_mesh_deformation", [117, len(d["V0"])], wl.arap_mesh_params(d), {}
    else:
        d = wl.bundle_adjustment_inputs(6, 50, 4)
        name, dims, params, kw = "bundle_adjustment", [6, 50, len(d["oToC"])], wl.bundle_adjustment_params(d), {}
    params = [np.asarray(p, dtype) if np.asarray(p).dtype == np.float32 and np.asarray(p).size > 1 else p for p in params]
    return name, dims, params, kw


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
@pytest.mark.parametrize("case", ["image_warping", "arap_mesh", "bundle_adjustment"])
def test_device_jacobian_equals_the_oracle_jacobian(case, dtype, tol):
    from thallo_b200.api import ThalloSolver
    name, dims, params, kw = _case(case, dtype)
    _, F, J = evaluate(energies.load(name), dims, [np.array(p, copy=True) for p in params], dtype)
    dp = [dev(p) if (hasattr(p, "size") and np.asarray(p).size > 1) else p for p in params]
    s = ThalloSolver(dims, name, "gauss_newton", double=(dtype == np.float64), **kw)
    s.init(dp)
    Jd = s.export_jacobian()
    s.close()
    assert Jd.shape == J.shape
    scale = max(1.0, abs(J).max())
    assert abs(Jd - J.astype(Jd.dtype)).max() <= tol * scale
