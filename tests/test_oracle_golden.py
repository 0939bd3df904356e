"""Pins the oracle against the reference's only known-answer vectors
(reference tests/minimal/gold.png, tests/minimal_graph/gold.png)."""
import os

import numpy as np
import pytest

from energies import load
from oracle.solver import OracleSolver
from thallo_b200 import workloads

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_msvc_rand_matches_scalar_lcg():
    a = workloads.msvc_rand(1000)
    b = workloads.msvc_rand_fast(1000)
    assert np.array_equal(a, b)
    assert a[:3].tolist() == [41, 18467, 6334]   # the well-known first MSVC rand() outputs


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", ["at_output", "residualwise"])
def test_kat2_minimal_graph(dtype, mode):
    gold = np.load(os.path.join(GOLD, "kat_minimal_graph.npz"))["gold"].reshape(-1)
    X, A, v0, v1 = workloads.minimal_graph_inputs(512)
    X = X.astype(dtype)
    s = OracleSolver(load("graph_laplacian"), [512, 511], "gauss_newton", dtype, mode)
    s.solve([X, A, v0, v1])
    out = (X.astype(np.float32) * 255).astype(np.uint8)
    assert np.array_equal(out, gold)


@pytest.mark.parametrize("dtype", [np.float32])
def test_kat1_minimal(dtype):
    gold = np.load(os.path.join(GOLD, "kat_minimal.npz"))["gold"]
    X, A = workloads.minimal_inputs(512, 512)
    s = OracleSolver(load("laplacian"), [512, 512], "gauss_newton", dtype, "residualwise")
    s.solve([X, A])
    out = (X.reshape(512, 512) * 255).astype(np.uint8)
    assert np.array_equal(out, gold)


def test_kat1_committed_variant_differs_only_in_last_rows():
    gold = np.load(os.path.join(GOLD, "kat_minimal.npz"))["gold"]
    X, A = workloads.minimal_inputs(512, 512)
    s = OracleSolver(load("laplacian"), [512, 512], "gauss_newton", np.float32, "residualwise",
                     define_kwargs=dict(variant="committed"))
    s.solve([X, A])
    out = (X.reshape(512, 512) * 255).astype(np.uint8)
    diff = np.argwhere(out != gold)
    assert len(diff) > 0 and diff[:, 0].min() >= 400   # deviation confined to the bottom rows
