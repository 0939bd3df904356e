"""The locality claim behind the full-size parity record (oracle/fullsize.py): the float64 oracle evaluated on a crop
of a problem reproduces r0 = -J^T F, the preconditioner and A p of the full problem on the crop's interior -- for every
configured energy, including index VALUES (shape_from_shading's camera model), sampled images (optical_flow), computed
arrays, a graph domain and bundle adjustment.  CPU only: oracle against oracle."""
import numpy as np
import pytest

import energies
from oracle import fullsize
from oracle.solver import OracleSolver
from thallo_b200 import configs, workloads as wl
from thallo_b200.frontend import codegen


def _case_params(key):
    if key == "2":
        c = configs.case("2", dims=(40, 64))
        d = wl.image_warping_inputs(40, 64)
        return c, [d[k] for k in ("Offset", "Angle", "UrShape", "Constraints", "Mask")] + [np.array([d["w_fitSqrt"]], np.float32), np.array([d["w_regSqrt"]], np.float32)]
    if key == "3a":
        c = configs.case("3a", dims=(40, 72))
        return c, wl.optical_flow_params(wl.optical_flow_inputs(40, 72))
    if key == "3b":
        c = configs.case("3b", dims=(48, 80))
        return c, wl.sfs_params(wl.sfs_inputs(48, 80))
    if key == "4a":
        c = configs.case("4a", dims=(10, 9, 40))
        return c, wl.volumetric_params(wl.volumetric_inputs(10, 9, 40))
    if key == "4b":
        c = configs.case("4b", n=40)
        d = wl.arap_mesh_inputs(40, 40)
        rng = np.random.RandomState(3)
        d["Angle"] = d["Angle"] + 0.2 * rng.randn(*d["Angle"].shape).astype(np.float32)
        return c, wl.arap_mesh_params(d)
    c = configs.case("5", cameras=12, points=1500)
    return c, wl.bundle_adjustment_params(wl.bundle_adjustment_inputs(12, 1500, 5))


@pytest.mark.parametrize("key", ["2", "3a", "3b", "4a", "4b", "5"])
def test_crop_reproduces_the_full_problem_on_its_interior(key):
    c, params = _case_params(key)
    dims = [int(x) for x in c.dims]
    p64 = [np.array(x, np.float64) if (hasattr(x, "dtype") and x.dtype == np.float32 and np.size(x) > 1) else x for x in params]
    low = codegen.lower(energies.load(c.energy), dims, c.kind, c.energy, schedule=c.schedule, **(c.define_kwargs or {}))
    desc = low.desc
    full = OracleSolver(energies.load(c.energy), dims, c.kind, np.float64, c.oracle_mode, materialized=c.materialized,
                        define_kwargs=c.define_kwargs)
    for k, v in c.solver_params.items():
        full.set(k, v)
    sv = full.setup_vectors(p64)
    rng = np.random.RandomState(1)
    p = rng.randn(sv["r"].size)
    Ap = sv["applyA"](p)
    checked = 0
    for label, crop in fullsize.crops_for(c, dims, desc, band=3)(p64):
        o = OracleSolver(energies.load(c.energy), crop.dims, c.kind, np.float64, c.oracle_mode, materialized=c.materialized,
                         define_kwargs=c.define_kwargs, origin=crop.origin)
        for k, v in c.solver_params.items():
            o.set(k, v)
        loc = o.setup_vectors(crop.params)
        mask = fullsize._masks(desc, crop)

        def gather(vec):
            parts = []
            for u in desc["unknowns"]:
                e0, e1, _, _ = crop.ranges[u["pidx"]]
                parts.append(vec[u["offset"] + e0 * u["channels"]:u["offset"] + e1 * u["channels"]])
            return np.concatenate(parts)
        assert mask.any(), (key, label)
        for name, a, b in (("r0", gather(sv["r"]), loc["r"]), ("M", gather(sv["M"]), loc["M"]), ("Ap", gather(Ap), loc["applyA"](gather(p)))):
            scale = max(np.abs(a[mask]).max(), 1e-300)
            assert np.abs(a[mask] - b[mask]).max() <= 1e-12 * scale, (key, label, name, np.abs(a[mask] - b[mask]).max() / scale)
        checked += int(mask.sum())
    assert checked > 0
