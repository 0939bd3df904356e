"""Pins the oracle's operator math on the headline energy against code the reference's authors wrote by hand: the
CUDA device functions of examples/image_warping/src/WarpingSolverEquations.h (cost, -J^T F, J^T J p of the 2-D ARAP
energy, hand-derived), compiled for the host from the reference tree where it lies (`make -C oracle hand` ->
oracle/_ref/libiw_hand.so; SURVEY 8c "secondary oracle").  The oracle differentiates the energy as Thallo defines it
with dual numbers; the hand solver minimises sum w e^2 where Thallo minimises 1/2 sum (sqrt(w) e)^2, hence the factor 2."""
import ctypes as C
import os

import numpy as np
import pytest

import energies
from oracle.npdsl import evaluate
from thallo_b200 import workloads as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libiw_hand.so")


def _lib():
    if os.path.isdir("/root/reference"):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "hand"], stdout=subprocess.DEVNULL)
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libiw_hand.so not built (needs the reference checkout)")
    lib = C.CDLL(LIB)
    fp = C.POINTER(C.c_float)
    common = [C.c_int, C.c_int, fp, fp, fp, fp, fp, C.c_float, C.c_float]
    lib.iw_hand_cost.restype, lib.iw_hand_cost.argtypes = C.c_double, common
    lib.iw_hand_minus_jtf.argtypes = common + [fp, fp]
    lib.iw_hand_apply_jtj.argtypes = common + [fp, fp, fp, fp]
    return lib, fp


@pytest.mark.parametrize("W,H,seed", [(23, 17, 1), (8, 31, 2), (40, 5, 3)])
def test_oracle_cost_gradient_and_jtj_match_the_reference_hand_derived_equations(W, H, seed):
    lib, fp = _lib()
    rs = np.random.RandomState(seed)
    d = wl.image_warping_inputs(W, H)
    d["Offset"] = (d["Offset"] + rs.randn(*d["Offset"].shape)).astype(np.float32)
    d["Angle"] = (0.4 * rs.randn(*d["Angle"].shape)).astype(np.float32)
    d["Mask"] = np.zeros_like(d["Mask"])                       # every pixel valid: the two formulations then agree term by term
    cons = -np.ones_like(d["Constraints"])
    sel = rs.rand(W * H) < 0.3
    cons[sel] = (d["UrShape"][sel] + 1.0 + rs.rand(int(sel.sum()), 2)).astype(np.float32)
    d["Constraints"] = cons
    wfit, wreg = 3.0, 0.7
    d["w_fitSqrt"], d["w_regSqrt"] = np.float32(np.sqrt(wfit)), np.float32(np.sqrt(wreg))
    _, F, J = evaluate(energies.load("image_warping"), [W, H], [np.asarray(p, np.float64) for p in wl.image_warping_params(d)], np.float64)
    n = W * H
    ptr = lambda a: a.ctypes.data_as(fp)
    x, A, ur, cn, mk = (np.ascontiguousarray(d[k], np.float32) for k in ("Offset", "Angle", "UrShape", "Constraints", "Mask"))
    args = (W, H, ptr(x), ptr(A), ptr(ur), ptr(cn), ptr(mk), wfit, wreg)
    # cost
    cost_thallo = 0.5 * float(F @ F)
    assert abs(lib.iw_hand_cost(*args) - 2 * cost_thallo) <= 1e-6 * 2 * cost_thallo
    # -J^T F
    b, bA = np.zeros((n, 2), np.float32), np.zeros(n, np.float32)
    lib.iw_hand_minus_jtf(*args, ptr(b), ptr(bA))
    g = -(J.T @ F)
    assert np.abs(b.reshape(-1) - 2 * g[:2 * n]).max() <= 2e-6 * np.abs(g).max()
    assert np.abs(bA - 2 * g[2 * n:]).max() <= 2e-6 * np.abs(g).max()
    # J^T J p
    pv = rs.randn(3 * n).astype(np.float32)
    pp, pa = np.ascontiguousarray(pv[:2 * n]), np.ascontiguousarray(pv[2 * n:])
    out, outA = np.zeros((n, 2), np.float32), np.zeros(n, np.float32)
    lib.iw_hand_apply_jtj(*args, ptr(pp), ptr(pa), ptr(out), ptr(outA))
    o = J.T @ (J @ pv.astype(np.float64))
    assert np.abs(out.reshape(-1) - 2 * o[:2 * n]).max() <= 2e-6 * np.abs(o).max()
    assert np.abs(outA - 2 * o[2 * n:]).max() <= 2e-6 * np.abs(o).max()


def _arap_lib():
    lib_path = os.path.join(ROOT, "oracle", "_ref", "libarap_hand.so")
    if os.path.isdir("/root/reference"):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "hand"], stdout=subprocess.DEVNULL)
    if not os.path.exists(lib_path):
        pytest.skip("oracle/_ref/libarap_hand.so not built (needs the reference checkout)")
    lib = C.CDLL(lib_path)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    common = [C.c_int, fp, fp, fp, fp, ip, ip, ip, C.c_float, C.c_float]
    lib.arap_hand_cost.restype, lib.arap_hand_cost.argtypes = C.c_double, common
    lib.arap_hand_minus_jtf.argtypes = common + [fp, fp]
    lib.arap_hand_apply_jtj.argtypes = common + [fp, fp, fp, fp]
    return lib, fp, ip


@pytest.mark.parametrize("nx,ny,seed", [(9, 7, 2), (5, 12, 3)])
def test_oracle_matches_the_reference_hand_derived_arap_mesh_equations(nx, ny, seed):
    """examples/arap_mesh_deformation/src/WarpingSolverEquations.h (+ RotationHelper.h): graph domain, 3-D rotations by
    three Euler angles per vertex; the hand solver walks per-vertex neighbour lists, Thallo's energy directed edges."""
    lib, fp, ip = _arap_lib()
    rs = np.random.RandomState(seed)
    d = wl.arap_mesh_inputs(nx, ny)
    N = nx * ny
    d["Position"] = (d["Position"] + 0.3 * rs.randn(N, 3)).astype(np.float32)
    d["Angle"] = (0.4 * rs.randn(N, 3)).astype(np.float32)
    wfit, wreg = 4.0, 1.3
    d["w_fitSqrt"], d["w_regSqrt"] = np.float32(np.sqrt(wfit)), np.float32(np.sqrt(wreg))
    p64 = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.arap_mesh_params(d)]
    _, F, J = evaluate(energies.load("arap_mesh_deformation"), [N, len(d["V0"])], p64, np.float64)
    V0, V1 = d["V0"].astype(np.int64), d["V1"].astype(np.int64)
    numN = np.bincount(V0, minlength=N).astype(np.int32)
    nOff = np.concatenate([[0], np.cumsum(numN)[:-1]]).astype(np.int32)
    nIdx = V1[np.argsort(V0, kind="stable")].astype(np.int32)
    target = d["Constraints"].astype(np.float32).copy()
    target[target[:, 0] < -999999.9] = -np.inf              # the hand solver marks "no target" with MINF
    ptr, iptr = (lambda a: a.ctypes.data_as(fp)), (lambda a: a.ctypes.data_as(ip))
    x, a, ur = (np.ascontiguousarray(d[k], np.float32) for k in ("Position", "Angle", "Original"))
    args = (N, ptr(x), ptr(a), ptr(target), ptr(ur), iptr(numN), iptr(nIdx), iptr(nOff), wfit, wreg)
    assert abs(lib.arap_hand_cost(*args) - float(F @ F)) <= 2e-6 * float(F @ F)
    b, bA = np.zeros((N, 3), np.float32), np.zeros((N, 3), np.float32)
    lib.arap_hand_minus_jtf(*args, ptr(b), ptr(bA))
    g = -(J.T @ F)
    assert np.abs(b.reshape(-1) - 2 * g[:3 * N]).max() <= 3e-6 * np.abs(g).max()
    assert np.abs(bA.reshape(-1) - 2 * g[3 * N:]).max() <= 3e-6 * np.abs(g).max()
    pv = rs.randn(6 * N).astype(np.float32)
    pp, pa = np.ascontiguousarray(pv[:3 * N]), np.ascontiguousarray(pv[3 * N:])
    out, outA = np.zeros((N, 3), np.float32), np.zeros((N, 3), np.float32)
    lib.arap_hand_apply_jtj(*args, ptr(pp), ptr(pa), ptr(out), ptr(outA))
    o = J.T @ (J @ pv.astype(np.float64))
    assert np.abs(out.reshape(-1) - 2 * o[:3 * N]).max() <= 3e-6 * np.abs(o).max()
    assert np.abs(outA.reshape(-1) - 2 * o[3 * N:]).max() <= 3e-6 * np.abs(o).max()


@pytest.mark.parametrize("dims,seed", [((5, 4, 6), 2), ((3, 7, 4), 5)])
def test_oracle_matches_the_reference_hand_derived_volumetric_equations(dims, seed):
    """examples/volumetric_mesh_deformation/src/WarpingSolverEquations.h: 3-D lattice, six neighbours, 3-D rotations.
    The hand solver's x is its slowest axis; the stencil is symmetric, so its (x, y, z) = this energy's (D, H, W)."""
    lib_path = os.path.join(ROOT, "oracle", "_ref", "libvol_hand.so")
    if os.path.isdir("/root/reference"):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "hand"], stdout=subprocess.DEVNULL)
    if not os.path.exists(lib_path):
        pytest.skip("oracle/_ref/libvol_hand.so not built (needs the reference checkout)")
    lib = C.CDLL(lib_path)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    common = [C.c_int, fp, fp, fp, fp, ip, C.c_float, C.c_float]
    lib.vol_hand_cost.restype, lib.vol_hand_cost.argtypes = C.c_double, common
    lib.vol_hand_minus_jtf.argtypes = common + [fp, fp]
    lib.vol_hand_apply_jtj.argtypes = common + [fp, fp, fp, fp]
    W, H, D = dims
    N = W * H * D
    rs = np.random.RandomState(seed)
    d = wl.volumetric_inputs(W, H, D)
    d["Offset"] = (d["Offset"] + 0.3 * rs.randn(N, 3)).astype(np.float32)
    d["Angle"] = (0.4 * rs.randn(N, 3)).astype(np.float32)
    wfit, wreg = 2.0, 0.6
    d["w_fitSqrt"], d["w_regSqrt"] = np.float32(np.sqrt(wfit)), np.float32(np.sqrt(wreg))
    p64 = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.volumetric_params(d)]
    _, F, J = evaluate(energies.load("volumetric_mesh_deformation"), [W, H, D], p64, np.float64)
    target = d["Constraints"].astype(np.float32).copy()
    target[target[:, 0] < -999999.9] = -np.inf
    nodes = np.array([D, H, W], np.int32)
    ptr = lambda a: a.ctypes.data_as(fp)
    x, a, ur = (np.ascontiguousarray(d[k], np.float32) for k in ("Offset", "Angle", "UrShape"))
    args = (N, ptr(x), ptr(a), ptr(target), ptr(ur), nodes.ctypes.data_as(ip), wfit, wreg)
    assert abs(lib.vol_hand_cost(*args) - float(F @ F)) <= 2e-6 * float(F @ F)
    b, bA = np.zeros((N, 3), np.float32), np.zeros((N, 3), np.float32)
    lib.vol_hand_minus_jtf(*args, ptr(b), ptr(bA))
    g = -(J.T @ F)
    assert np.abs(b.reshape(-1) - 2 * g[:3 * N]).max() <= 3e-6 * np.abs(g).max()
    assert np.abs(bA.reshape(-1) - 2 * g[3 * N:]).max() <= 3e-6 * np.abs(g).max()
    pv = rs.randn(6 * N).astype(np.float32)
    pp, pa = np.ascontiguousarray(pv[:3 * N]), np.ascontiguousarray(pv[3 * N:])
    out, outA = np.zeros((N, 3), np.float32), np.zeros((N, 3), np.float32)
    lib.vol_hand_apply_jtj(*args, ptr(pp), ptr(pa), ptr(out), ptr(outA))
    o = J.T @ (J @ pv.astype(np.float64))
    assert np.abs(out.reshape(-1) - 2 * o[:3 * N]).max() <= 3e-6 * np.abs(o).max()
    assert np.abs(outA.reshape(-1) - 2 * o[3 * N:]).max() <= 3e-6 * np.abs(o).max()


@pytest.mark.parametrize("W,H,seed", [(36, 28, 4), (17, 40, 6)])
def test_oracle_shading_term_and_its_gradient_match_the_reference_hand_derived_helper(W, H, seed):
    """shape_from_shading: the per-pixel shading error B - I (normal from three depth samples, nine spherical-harmonics
    lighting coefficients) and its derivatives with respect to those samples, hand-derived in
    examples/shape_from_shading/src/SFSSolverUtil.h:59-195, against the value and gradient image of the ComputedArray
    `B_I_comp` the oracle derives from the energy by dual-number AD (captured at its first :get)."""
    from oracle import npdsl
    lib_path = os.path.join(ROOT, "oracle", "_ref", "libsfs_hand.so")
    if os.path.isdir("/root/reference"):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "hand"], stdout=subprocess.DEVNULL)
    if not os.path.exists(lib_path):
        pytest.skip("oracle/_ref/libsfs_hand.so not built (needs the reference checkout)")
    lib = C.CDLL(lib_path)
    fp = C.POINTER(C.c_float)
    lib.sfs_hand_shading.argtypes = [C.c_int, C.c_int, fp, fp, fp, C.c_float, C.c_float, C.c_float, C.c_float, fp]
    d = wl.sfs_inputs(W, H)
    rs = np.random.RandomState(seed)
    depth = (0.5 + 0.05 * rs.rand(H, W)).astype(np.float32)       # every depth valid and positive: the hand code tests the refined
    d["X"] = depth.reshape(-1).copy()                              # depth for validity, the energy the input depth
    d["D_i"] = depth.reshape(-1).copy()
    p64 = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.sfs_params(d)]
    captured = []
    orig = npdsl.Dual.get

    def spy(self, *idx):
        captured.append(self)
        return orig(self, *idx)
    npdsl.Dual.get = spy
    try:
        npdsl.evaluate(energies.load("shape_from_shading"), [W, H], p64, np.float64)
    finally:
        npdsl.Dual.get = orig
    bi = captured[0]                                               # B_I_comp, shape_from_shading.t:79
    assert sorted(k[1][1] for k in bi.d) == [(-1, 0), (0, -1), (0, 0)]
    out = np.zeros((H, W, 4), np.float32)
    X, Im = np.ascontiguousarray(d["X"], np.float32), np.ascontiguousarray(d["Im"], np.float32)
    light = np.ascontiguousarray(np.array(d["light"], np.float32))
    ptr = lambda a: a.ctypes.data_as(fp)
    lib.sfs_hand_shading(W, H, ptr(X), ptr(Im), ptr(light), d["f_x"], d["f_y"], d["u_x"], d["u_y"], ptr(out))
    inner = (slice(1, H), slice(1, W))
    val = np.asarray(bi.val).reshape(H, W)
    assert np.abs(out[..., 3][inner] - val[inner]).max() <= 3e-6 * np.abs(val[inner]).max()
    channel = {(-1, 0): 0, (0, 0): 1, (0, -1): 2}                  # the helper's d0 = X(x-1, y), d1 = X(x, y), d2 = X(x, y-1)
    for key, dv in bi.d.items():
        dv = np.broadcast_to(np.asarray(dv), (H, W))
        got = out[..., channel[tuple(key[1][1])]]
        assert np.abs(got[inner] - dv[inner]).max() <= 3e-6 * np.abs(dv[inner]).max()
