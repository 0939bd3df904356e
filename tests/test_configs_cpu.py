"""CPU checks of the configured-workload builders and of bench.py's record helpers: sizes as BASELINE.json names them,
the SURVEY 8d byte formulas, row-restricted device generators equal to slices of the full fields (what a rank of a
partitioned run generates must be what a single-GPU run sees there), trajectory comparison records."""
import importlib.util
import os

import numpy as np
import pytest

from thallo_b200 import configs, workloads as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_configured_sizes_match_baseline_json():
    c = {k: configs.case(k) for k in configs.CASES}
    assert tuple(c["1"].dims) == (256, 256) and c["1"].kind == "gauss_newton" and (c["1"].nit, c["1"].lit) == (10, 10)
    assert tuple(c["2"].dims) == (2048, 2048) and c["2"].kind == "levenberg_marquardt" and (c["2"].nit, c["2"].lit) == (8, 100)
    assert tuple(c["3a"].dims) == (8192, 8192) and tuple(c["3b"].dims) == (8192, 8192)
    assert int(np.prod(c["4a"].dims)) == 160 ** 3
    assert c["4b"].dims == (4000000, 23984002)                   # 2000 x 2000 grid, 6-neighbourhood, directed edges
    assert c["5"].dims == (10000, 5000000, 25000000) and c["5"].solver_params["q_tolerance"] == pytest.approx(0.1)
    # SURVEY 8d byte counts per PCG iteration
    assert c["2"].survey_iteration_bytes() == 204 * 2048 * 2048
    assert c["3a"].survey_iteration_bytes() == 112 * 8192 * 8192
    assert c["4a"].survey_iteration_bytes() == 348 * 160 ** 3
    assert abs(c["4b"].survey_iteration_bytes() - 1.87e9) < 0.02e9
    assert c["4b"]._edges() == 2 * (2 * 2000 * 1999 + 1999 * 1999)


def test_mesh_edge_count_formula_matches_the_generator():
    for n in (5, 17, 40):
        c = configs.case("4b", n=n)
        assert c.dims == (n * n, len(wl.arap_mesh_inputs(n, n)["V0"]))


@pytest.mark.parametrize("gen,keys", [(wl.optical_flow_inputs_torch, ("I", "I_hat_im", "I_hat_dx", "I_hat_dy")),
                                      (wl.sfs_inputs_torch, ("X", "D_i", "Im"))])
def test_row_restricted_generators_equal_slices_of_the_full_fields(gen, keys):
    W, H = 48, 37
    full = gen(W, H, "cpu")
    for rows in ((0, 9), (8, 30), (29, 37)):
        part = gen(W, H, "cpu", rows=rows)
        for k in keys:
            a = full[k].numpy().reshape(H, W)[rows[0]:rows[1]].reshape(-1)
            assert np.array_equal(a, part[k].numpy()), (k, rows)


def test_torch_generators_reproduce_the_numpy_fields():
    a, b = wl.optical_flow_inputs(40, 28), wl.optical_flow_inputs_torch(40, 28, "cpu")
    for k in ("I", "I_hat_im", "I_hat_dx", "I_hat_dy"):
        assert np.array_equal(a[k], b[k].numpy())
    a, b = wl.sfs_inputs(40, 28), wl.sfs_inputs_torch(40, 28, "cpu")
    for k in ("D_i", "Im"):
        assert np.array_equal(a[k], b[k].numpy())
    d = wl.bundle_adjustment_inputs_torch(12, 300, "cpu")
    o2c, o2p = d["oToC"].numpy(), d["oToP"].numpy()
    assert len(o2c) == 1500 and (np.diff(o2p) >= 0).all() and o2c.min() >= 0 and o2c.max() < 12
    assert all(len(set(o2c[5 * p:5 * p + 5])) == 5 for p in range(300))          # five distinct cameras per point
    assert np.isfinite(d["observations"].numpy()).all()


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_trajectory_comparison_record():
    b = _bench()
    r = b.compare_trajectories([10.0, 5.0, 2.5], [7, 9], [10.0, 5.0 * (1 + 2e-6), 2.5], [7, 9], "x")
    assert r["pcg_counts_equal"] and r["same_length"] and r["costs_compared"] == 3 and 1.9e-6 < r["max_rel"] < 2.1e-6
    assert b.with_noise(dict(r), 0.0)["within_1e-5_or_8x_noise_floor"] is True
    r2 = b.compare_trajectories([10.0, 5.0], [7], [10.0, 5.1], [8], "x")
    assert not r2["pcg_counts_equal"] and r2["max_rel"] > 1e-2
    assert b.with_noise(dict(r2), 1e-6)["within_1e-5_or_8x_noise_floor"] is False
    assert b.with_noise(dict(r2), 1e-2)["within_1e-5_or_8x_noise_floor"] is True
