"""The two configured energies for which the reference holds no hand-derived equations (optical_flow,
bundle_adjustment): the oracle's Jacobian (dual-number AD) against central finite differences of its own residuals in
float64, and bundle adjustment's residuals against an independent statement of the published Snavely camera model
(angle-axis rotation by Rodrigues' formula, perspective division with Bundler's sign, two-term radial distortion)."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import energies
from oracle.npdsl import evaluate
from thallo_b200 import workloads as wl


def _fd_check(name, dims, params, unknown_slots, h=1e-6, tol=2e-6, samples=60, seed=0):
    define = energies.load(name)
    _, F, J = evaluate(define, dims, params, np.float64)
    J = J.tocsc()
    rs = np.random.RandomState(seed)
    sizes = [np.asarray(params[s]).size for s in unknown_slots]
    total = sum(sizes)
    scale = max(1.0, abs(J).max())
    for col in rs.choice(total, size=min(samples, total), replace=False):
        k, off = 0, int(col)
        while off >= sizes[k]:
            off -= sizes[k]
            k += 1
        fs = []
        for sgn in (+1, -1):
            p = [np.array(a, copy=True) if hasattr(a, "shape") else a for a in params]
            flat = p[unknown_slots[k]].reshape(-1)
            flat[off] += sgn * h
            fs.append(evaluate(define, dims, p, np.float64)[1])
        fd = (fs[0] - fs[1]) / (2 * h)
        got = np.asarray(J[:, int(col)].todense()).reshape(-1)
        assert np.abs(got - fd).max() <= tol * scale, (name, int(col), np.abs(got - fd).max())


def _bilinear(img, x, y):
    """floor / ceil lerp with zero outside the image (thallo.t:899-907), written independently of the oracle's."""
    H, W = img.shape
    x0, y0 = np.floor(x).astype(int), np.floor(y).astype(int)
    x1, y1 = np.ceil(x).astype(int), np.ceil(y).astype(int)

    def at(ix, iy):
        ok = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
        return np.where(ok, img[np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)], 0.0)
    tx, ty = x - x0, y - y0
    top = (1 - tx) * at(x0, y0) + tx * at(x1, y0)
    bot = (1 - tx) * at(x0, y1) + tx * at(x1, y1)
    return (1 - ty) * top + ty * bot


def test_optical_flow_residuals_and_jacobian_follow_the_sampled_image_rules():
    """A SampledImage is differentiated through the derivative images the CALLER supplies, not through the
    interpolant (thallo.t:5803-5817): d fit / d X = -w * sample(I_hat_dx | I_hat_dy) at the warped position.  The
    regularisation rows are linear and checked by finite differences of the oracle's own residuals."""
    W, H = 14, 11
    d = wl.optical_flow_inputs(W, H)
    rs = np.random.RandomState(1)
    d["X"] = (0.8 * rs.randn(W * H, 2)).astype(np.float32)
    params = [np.asarray(p, np.float64) for p in wl.optical_flow_params(d)]
    wf, wr, X, I, Ih, Ix, Iy = params
    _, F, J = evaluate(energies.load("optical_flow"), [W, H], params, np.float64)
    J = J.toarray()
    n = W * H
    ys, xs = np.mgrid[0:H, 0:W]
    px, py = xs.reshape(-1) + X[:, 0], ys.reshape(-1) + X[:, 1]
    img = lambda a: a.reshape(H, W)
    fit = float(wf[0]) * (I - _bilinear(img(Ih), px, py))
    assert np.abs(F[:n] - fit).max() <= 1e-12                                    # groups sorted by name: fit first
    rows = np.arange(n)
    assert np.abs(J[rows, 2 * rows] + float(wf[0]) * _bilinear(img(Ix), px, py)).max() <= 1e-12
    assert np.abs(J[rows, 2 * rows + 1] + float(wf[0]) * _bilinear(img(Iy), px, py)).max() <= 1e-12
    Jfit = J[:n].copy()
    Jfit[rows, 2 * rows] = 0
    Jfit[rows, 2 * rows + 1] = 0
    assert not Jfit.any()                                                        # the fit residual touches only its own pixel
    # regularisation rows: finite differences
    h = 1e-6
    for col in rs.choice(2 * n, size=40, replace=False):
        fs = []
        for sgn in (+1, -1):
            p = [np.array(a, copy=True) for a in params]
            p[2].reshape(-1)[col] += sgn * h
            fs.append(evaluate(energies.load("optical_flow"), [W, H], p, np.float64)[1][n:])
        assert np.abs(J[n:, col] - (fs[0] - fs[1]) / (2 * h)).max() <= 1e-8


def test_bundle_adjustment_jacobian_matches_finite_differences():
    d = wl.bundle_adjustment_inputs(5, 30, 3)
    params = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.bundle_adjustment_params(d)]
    _fd_check("bundle_adjustment", [5, 30, len(d["oToC"])], params, [0, 1], h=1e-6, tol=5e-6)


def test_bundle_adjustment_residuals_match_the_snavely_camera_model():
    d = wl.bundle_adjustment_inputs(6, 40, 4)
    params = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.bundle_adjustment_params(d)]
    cams, pts, obs, o2c, o2p = params
    _, F, _ = evaluate(energies.load("bundle_adjustment"), [6, 40, len(o2c)], params, np.float64)
    c, X = cams[o2c], pts[o2p]
    P = Rotation.from_rotvec(c[:, :3]).apply(X) + c[:, 3:6]                  # Rodrigues
    xp = -P[:, :2] / P[:, 2:3]                                               # Bundler: the camera looks down -z
    r2 = (xp ** 2).sum(axis=1)
    pred = xp * (c[:, 6] * (1.0 + r2 * (c[:, 7] + c[:, 8] * r2)))[:, None]
    want = (obs - pred).reshape(-1)                                          # residual = observed - predicted (bundle_adjustment.t:34)
    assert np.abs(F - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
