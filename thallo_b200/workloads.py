"""Synthetic inputs for the configured workloads.

Pure NumPy, deterministic.  Used by tests, bench.py and smoke(); nothing here is
solver code.  `msvc_rand` is the MSVC `rand()` LCG (seed 1, RAND_MAX 32767) that
reference tests/minimal/main.cpp:52-54 and tests/minimal_graph/main.cpp implicitly
used to produce the golden PNGs.
"""
import numpy as np


def msvc_rand(n, seed=1):
    out = np.empty(n, dtype=np.int64)
    s = seed
    for i in range(n):
        s = (s * 214013 + 2531011) & 0xFFFFFFFF
        out[i] = (s >> 16) & 0x7FFF
    return out


def msvc_rand_fast(n, seed=1):
    """Vectorised LCG: s_k = a^k s_0 + c (a^k - 1)/(a - 1) mod 2^32, via doubling."""
    a, c = 214013, 2531011
    A = np.empty(n + 1, dtype=np.uint64)
    C = np.empty(n + 1, dtype=np.uint64)
    A[0], C[0] = 1, 0
    filled = 1
    mask = np.uint64(0xFFFFFFFF)
    A[1], C[1] = a, c
    filled = 2
    while filled < n + 1:
        m = min(filled - 1, n + 1 - filled)
        # state after (filled-1 + j) steps = compose(step^(filled-1), step^j)
        Ak, Ck = A[filled - 1], C[filled - 1]
        A[filled:filled + m] = (A[1:1 + m] * Ak) & mask
        C[filled:filled + m] = (A[1:1 + m] * Ck + C[1:1 + m]) & mask
        filled += m
    s = (A[1:] * np.uint64(seed) + C[1:]) & mask
    return ((s >> np.uint64(16)) & np.uint64(0x7FFF)).astype(np.int64)


def minimal_inputs(W, H, seed=1):
    """A[i] = float(rand()/RAND_MAX), row-major; X0 = A (tests/minimal/main.cpp:50-60)."""
    r = msvc_rand_fast(W * H, seed)
    A = (r.astype(np.float64) / 32767.0).astype(np.float32)
    return A.copy(), A


def minimal_graph_inputs(N, seed=1):
    r = msvc_rand_fast(N, seed)
    A = (r.astype(np.float64) / 32767.0).astype(np.float32)
    v0 = np.arange(N - 1, dtype=np.int32)
    v1 = v0 + 1
    return A.copy(), A, v0, v1


def image_warping_inputs(W, H, seed=1, w_fit=100.0, w_reg=0.01):
    """Config 2 synthetic shape: UrShape=(x,y), Offset0=UrShape, Angle0=0,
    Mask=0 except a seeded ~10% of pixels in discs (=1, excluded) and the outermost
    image ring (=1, so the at-output and residualwise forms of the energy agree on the
    border, see DESIGN.md "ghost residuals"); Constraints=-1 except the second ring
    pinned in place and an 8x8 lattice of handles displaced by a smooth ramp."""
    rng = np.random.RandomState(seed)
    ys, xs = np.mgrid[0:H, 0:W]
    ur = np.stack([xs, ys], axis=-1).astype(np.float32)
    offset = ur.copy()
    angle = np.zeros((H, W), np.float32)
    mask = np.zeros((H, W), np.float32)
    ndisc = 12
    rad = max(2, int(np.sqrt(0.10 * W * H / (np.pi * ndisc))))
    for _ in range(ndisc):
        cx, cy = rng.randint(0, W), rng.randint(0, H)
        mask[(xs - cx) ** 2 + (ys - cy) ** 2 <= rad * rad] = 1.0
    mask[0, :] = mask[-1, :] = 1.0
    mask[:, 0] = mask[:, -1] = 1.0
    cons = -np.ones((H, W, 2), np.float32)
    ring = np.zeros((H, W), bool)
    ring[1, 1:-1] = ring[-2, 1:-1] = True
    ring[1:-1, 1] = ring[1:-1, -2] = True
    cons[ring] = ur[ring]
    gx = np.linspace(W * 0.15, W * 0.85, 8).astype(int)
    gy = np.linspace(H * 0.15, H * 0.85, 8).astype(int)
    for j, y in enumerate(gy):
        for i, x in enumerate(gx):
            s = np.sin(np.pi * (i + 1) / 9.0) * np.sin(np.pi * (j + 1) / 9.0)
            cons[y, x, 0] = x + 15.0 * s * (W / 2048.0 if W > 256 else 0.25)
            cons[y, x, 1] = y - 10.0 * s * (H / 2048.0 if H > 256 else 0.25)
    return dict(Offset=offset.reshape(-1, 2), Angle=angle.reshape(-1), UrShape=ur.reshape(-1, 2),
                Constraints=cons.reshape(-1, 2), Mask=mask.reshape(-1),
                w_fitSqrt=np.float32(np.sqrt(w_fit)), w_regSqrt=np.float32(np.sqrt(w_reg)))


def image_warping_params(d):
    return [d["Offset"], d["Angle"], d["UrShape"], d["Constraints"], d["Mask"],
            np.array([d["w_fitSqrt"]], np.float32), np.array([d["w_regSqrt"]], np.float32)]


def optical_flow_inputs(W, H, seed=1, w_fit=10.0, w_reg=0.1):
    """Config 3a synthetic shape: I = band-limited texture (16 seeded sinusoids),
    I_hat = I translated by (0.6, -0.4) px, I_hat_dx/dy = Prewitt/8 with a zero border
    (examples/optical_flow/src/CombinedSolver.h:149-175), X0 = 0."""
    rng = np.random.RandomState(seed)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)

    def tex(x, y):
        r = np.random.RandomState(seed + 7)
        out = np.zeros_like(x)
        for _ in range(16):
            fx, fy = r.uniform(-0.35, 0.35, 2)
            ph, am = r.uniform(0, 2 * np.pi), r.uniform(0.2, 1.0)
            out += am * np.sin(fx * x + fy * y + ph)
        return out / 8.0 + 0.5
    I = tex(xs, ys)
    Ih = tex(xs + 0.6, ys - 0.4)
    dx = np.zeros_like(Ih)
    dy = np.zeros_like(Ih)
    dx[1:-1, 1:-1] = (-Ih[:-2, :-2] - Ih[1:-1, :-2] - Ih[2:, :-2] + Ih[:-2, 2:] + Ih[1:-1, 2:] + Ih[2:, 2:]) / 8.0
    dy[1:-1, 1:-1] = (-Ih[:-2, :-2] - Ih[:-2, 1:-1] - Ih[:-2, 2:] + Ih[2:, :-2] + Ih[2:, 1:-1] + Ih[2:, 2:]) / 8.0
    del rng
    f = lambda a: np.ascontiguousarray(a.reshape(-1), np.float32)
    return dict(w_fitSqrt=np.float32(np.sqrt(w_fit)), w_regSqrt=np.float32(np.sqrt(w_reg)),
                X=np.zeros((W * H, 2), np.float32), I=f(I), I_hat_im=f(Ih), I_hat_dx=f(dx), I_hat_dy=f(dy))


def optical_flow_params(d):
    return [np.array([d["w_fitSqrt"]], np.float32), np.array([d["w_regSqrt"]], np.float32),
            d["X"], d["I"], d["I_hat_im"], d["I_hat_dx"], d["I_hat_dy"]]


def volumetric_inputs(W, H, D, seed=1, w_fit=1.0, w_reg=0.05):
    """Config 4a synthetic shape: UrShape = lattice, top face (z = D-1) rotated by
    30 degrees about the z axis through the lattice centre, bottom face fixed, every other node
    unconstrained (sentinel below -999999.9, volumetric_mesh_deformation.t:18)."""
    zz, yy, xx = np.mgrid[0:D, 0:H, 0:W]
    ur = np.stack([xx, yy, zz], -1).reshape(-1, 3).astype(np.float32)
    n = W * H * D
    cons = np.full((n, 3), -1e7, np.float32)
    bottom = (zz == 0).reshape(-1)
    top = (zz == D - 1).reshape(-1)
    cons[bottom] = ur[bottom]
    th = np.deg2rad(30.0)
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    x, y = ur[top, 0] - cx, ur[top, 1] - cy
    cons[top, 0] = np.cos(th) * x - np.sin(th) * y + cx
    cons[top, 1] = np.sin(th) * x + np.cos(th) * y + cy
    cons[top, 2] = ur[top, 2]
    rng = np.random.RandomState(seed)
    ang = (0.01 * rng.randn(n, 3)).astype(np.float32)
    return dict(Offset=ur.copy(), Angle=ang, UrShape=ur, Constraints=cons,
                w_fitSqrt=np.float32(np.sqrt(w_fit)), w_regSqrt=np.float32(np.sqrt(w_reg)))


def volumetric_params(d):
    return [d["Offset"], d["Angle"], d["UrShape"], d["Constraints"],
            np.array([d["w_fitSqrt"]], np.float32), np.array([d["w_regSqrt"]], np.float32)]


def arap_mesh_inputs(nx, ny, seed=1, w_fit=4.0, w_reg=1.0, handle_fraction=0.01):
    """Config 4b synthetic shape: a triangulated nx x ny grid (interior valence 6),
    directed edges listed per vertex as examples/shared/ThalloGraph.h:67-79
    (createGraphFromNeighborLists) does, Original = grid + seeded z-noise, ~1% handle vertices
    displaced, every other vertex unconstrained (sentinel below -999999.9,
    arap_mesh_deformation.t:20)."""
    rng = np.random.RandomState(seed)
    n = nx * ny
    ys, xs = np.mgrid[0:ny, 0:nx]
    vid = (xs + nx * ys)
    heads, tails = [], []
    # neighbour offsets of a grid triangulated along one diagonal: 6-neighbourhood
    for dx, dy in [(-1, -1), (0, -1), (-1, 0), (1, 0), (0, 1), (1, 1)]:
        x2, y2 = xs + dx, ys + dy
        ok = (x2 >= 0) & (x2 < nx) & (y2 >= 0) & (y2 < ny)
        heads.append(np.where(ok, vid, -1))
        tails.append(np.where(ok, x2 + nx * y2, -1))
    heads = np.stack(heads, -1).reshape(-1)       # per vertex, then per neighbour: the reference's order
    tails = np.stack(tails, -1).reshape(-1)
    keep = heads >= 0
    v0 = heads[keep].astype(np.int32)
    v1 = tails[keep].astype(np.int32)
    orig = np.stack([xs, ys, 0.05 * rng.randn(ny, nx)], -1).reshape(n, 3).astype(np.float32)
    cons = np.full((n, 3), -1e7, np.float32)
    nh = max(2, int(handle_fraction * n))
    hidx = rng.choice(n, nh, replace=False)
    cons[hidx] = orig[hidx] + np.stack([0.3 * np.sin(orig[hidx, 1] * 0.1), 0.2 * np.cos(orig[hidx, 0] * 0.1),
                                        0.5 + 0 * orig[hidx, 0]], -1).astype(np.float32)
    ang = np.zeros((n, 3), np.float32)
    return dict(w_fitSqrt=np.float32(np.sqrt(w_fit)), w_regSqrt=np.float32(np.sqrt(w_reg)),
                Position=orig.copy(), Angle=ang, Original=orig, Constraints=cons, V0=v0, V1=v1)


def arap_mesh_params(d):
    return [np.array([d["w_fitSqrt"]], np.float32), np.array([d["w_regSqrt"]], np.float32),
            d["Position"], d["Angle"], d["Original"], d["Constraints"], d["V0"], d["V1"]]


def _angle_axis_rotate(aa, pt):
    """Rodrigues rotation, rows of `aa` (n,3) applied to rows of `pt` (n,3) (lib.t:514-555)."""
    th2 = np.sum(aa * aa, -1, keepdims=True)
    th = np.sqrt(np.maximum(th2, 1e-30))
    w = aa / th
    c, s = np.cos(th), np.sin(th)
    large = pt * c + np.cross(w, pt) * s + w * (np.sum(w * pt, -1, keepdims=True) * (1.0 - c))
    small = pt + np.cross(aa, pt)
    return np.where(th2 > 1e-8, large, small)


def bundle_adjustment_inputs(C, P, obs_per_point=5, seed=1, noise_px=0.5, perturb=0.01):
    """Config 5 synthetic shape: C cameras on a ring looking at a unit cube of P
    points, every point seen by `obs_per_point` distinct seeded-random cameras; observations =
    Snavely projection (bundle_adjustment.t:15-32) + N(0, noise_px) pixels; cameras and points
    then perturbed by `perturb` (relative).  Observations are listed point-major (like BAL files),
    so oToP is sorted and oToC is not.  Camera = angle-axis 3, translation 3, focal, k1, k2."""
    rng = np.random.RandomState(seed)
    k = int(obs_per_point)
    assert C >= k
    theta = 2.0 * np.pi * np.arange(C) / C
    cams = np.zeros((C, 9), np.float64)
    cams[:, 1] = theta                       # rotation about the y axis: the ring
    cams[:, 0] = 0.05 * rng.randn(C)
    cams[:, 2] = 0.05 * rng.randn(C)
    cams[:, 3:5] = 0.1 * rng.randn(C, 2)
    cams[:, 5] = -5.0                        # points end up at z ~ -5 in the camera frame (BAL looks down -z)
    cams[:, 6] = 800.0 + 20.0 * rng.randn(C)
    cams[:, 7] = 1e-2 * rng.randn(C)
    cams[:, 8] = 1e-3 * rng.randn(C)
    pts = rng.uniform(-1.0, 1.0, (P, 3))
    # k distinct cameras per point: a random start and k random distinct strides would correlate; use
    # the first k entries of a per-point random offset walk (distinct by construction, vectorised)
    base = rng.randint(0, C, P)
    steps = 1 + rng.randint(0, max(1, (C - 1) // k), (P, k))
    steps[:, 0] = 0
    cam_of = np.sort((base[:, None] + np.cumsum(steps, 1)) % C, 1)
    o2c = cam_of.reshape(-1).astype(np.int32)
    o2p = np.repeat(np.arange(P, dtype=np.int32), k)
    cam = cams[o2c]
    p = _angle_axis_rotate(cam[:, 0:3], pts[o2p]) + cam[:, 3:6]
    cod = -p[:, 0:2] / p[:, 2:3]
    r2 = np.sum(cod * cod, -1, keepdims=True)
    obs = cod * cam[:, 6:7] * (1.0 + r2 * (cam[:, 7:8] + cam[:, 8:9] * r2))
    obs = obs + noise_px * rng.randn(*obs.shape)
    cams0 = cams * (1.0 + perturb * rng.randn(*cams.shape))
    pts0 = pts + perturb * rng.randn(*pts.shape)
    return dict(cameras=cams0.astype(np.float32), points=pts0.astype(np.float32), observations=obs.astype(np.float32),
                oToC=o2c, oToP=o2p)


def bundle_adjustment_params(d):
    return [d["cameras"], d["points"], d["observations"], d["oToC"], d["oToP"]]


def sfs_inputs(W, H, seed=1, w_p=100.0, w_s=100.0, w_g=1.0):
    """Config 3b synthetic shape: target depth = spherical cap over a plane at 0.5 with
    a seeded 0.1 % ripple, invalid (-10000, what the reference's loader turns -inf into,
    examples/shape_from_shading/src/SimpleBuffer.cpp:30-40) outside a centred ellipse; initial depth =
    target + seeded noise on valid pixels; target intensity = band-limited texture in (0, 1); edge
    masks all 1; intrinsics f = W, u = (W/2, H/2); lighting and weights as in the reference's
    data/shape_from_shading/default.SFSSolverParameters (w_p 100, w_s 100, w_g 1)."""
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    u, v = (xs - W / 2.0) / (W / 2.0), (ys - H / 2.0) / (H / 2.0)
    r2 = u * u + v * v
    depth = 0.5 - 0.08 * np.sqrt(np.maximum(0.0, 1.0 - np.minimum(r2, 1.0)))
    depth += 0.0005 * np.sin(0.37 * xs + 0.11 * ys) * np.cos(0.23 * ys - 0.05 * xs)
    valid = r2 < 0.81
    D = np.where(valid, depth, -10000.0)
    rng = np.random.RandomState(seed)
    X0 = np.where(valid, depth + 0.0008 * rng.standard_normal((H, W)), -10000.0)
    Im = 0.5 + 0.2 * np.sin(0.05 * xs + 0.02 * ys) + 0.15 * np.cos(0.031 * ys - 0.017 * xs) + 0.05 * np.sin(0.4 * xs) * np.sin(0.3 * ys)
    f = lambda a: np.ascontiguousarray(a.reshape(-1), np.float32)
    light = [0.6908317804336548, 0.044598858803510666, 0.01812959648668766, -0.1773163229227066, -0.04067882522940636,
             0.14467650651931763, 0.02393525093793869, -0.24658696353435516, 0.005797004792839289]
    return dict(w_p=w_p, w_s=w_s, w_g=w_g, f_x=float(W), f_y=float(W), u_x=W / 2.0, u_y=H / 2.0, light=light,
                X=f(X0), D_i=f(D), Im=f(Im), edgeMaskR=np.ones(W * H, np.uint8), edgeMaskC=np.ones(W * H, np.uint8))


def sfs_params(d):
    """problemparams of energies/shape_from_shading.py: slots 0-15 host scalars, 16-20 images."""
    sc = [d["w_p"], d["w_s"], d["w_g"], d["f_x"], d["f_y"], d["u_x"], d["u_y"]] + list(d["light"])
    return [np.array([x], np.float32) for x in sc] + [d["X"], d["D_i"], d["Im"], d["edgeMaskR"], d["edgeMaskC"]]


def sfs_fixture_inputs(path):
    """The crop of the reference's own shape_from_shading input committed under tests/golden/
    (made by tests/golden/make_sfs_fixture.py)."""
    z = np.load(path)
    return dict(w_p=float(z["w_p"]), w_s=float(z["w_s"]), w_g=float(z["w_g"]), f_x=float(z["f_x"]), f_y=float(z["f_y"]),
                u_x=float(z["u_x"]), u_y=float(z["u_y"]), light=[float(x) for x in z["light"]],
                X=z["X"].astype(np.float32), D_i=z["D_i"].astype(np.float32), Im=z["Im"].astype(np.float32),
                edgeMaskR=z["edgeMaskR"].astype(np.uint8), edgeMaskC=z["edgeMaskC"].astype(np.uint8)), int(z["W"]), int(z["H"])


# ---------------------------------------------------------------------------------------------------
# Device-side generators for the configured FULL sizes (8192^2 images, 25 M observations): the same synthetic shapes
# as above, written in torch so that a 67 M-pixel field takes milliseconds on the GPU instead of a minute of NumPy
# per rank.  Every field is a pure function of the absolute element coordinates (and a seed), so a rank of a
# partitioned solve generates just its rows [y0, y1) and gets the same numbers a single-GPU run sees there.
def _hash01(t, x, y, seed):
    """Deterministic uniform(0, 1) field of the integer coordinates (float64 sine hash)."""
    v = t.sin(x * 12.9898 + y * 78.233 + seed * 37.719) * 43758.5453
    return v - t.floor(v)


def optical_flow_inputs_torch(W, H, device, rows=None, seed=1, w_fit=10.0, w_reg=0.1):
    """optical_flow_inputs on `device`, restricted to rows [y0, y1) of the W x H image (default: all)."""
    import torch as t
    y0, y1 = rows if rows is not None else (0, H)
    ys = t.arange(y0 - 1, y1 + 1, device=device, dtype=t.float64)[:, None]
    xs = t.arange(-1, W + 1, device=device, dtype=t.float64)[None, :]
    r = np.random.RandomState(seed + 7)
    coef = [(tuple(r.uniform(-0.35, 0.35, 2)), r.uniform(0, 2 * np.pi), r.uniform(0.2, 1.0)) for _ in range(16)]

    def tex(x, y):
        out = t.zeros((ys.shape[0], xs.shape[1]), device=device, dtype=t.float64)
        for (fx, fy), ph, am in coef:
            out += am * t.sin(fx * x + fy * y + ph)
        return out / 8.0 + 0.5
    I = tex(xs, ys)[1:-1, 1:-1]
    Ihp = tex(xs + 0.6, ys - 0.4)                         # rows y0-1 .. y1, columns -1 .. W
    Ih = Ihp[1:-1, 1:-1]
    dx = (-Ihp[:-2, :-2] - Ihp[1:-1, :-2] - Ihp[2:, :-2] + Ihp[:-2, 2:] + Ihp[1:-1, 2:] + Ihp[2:, 2:]) / 8.0
    dy = (-Ihp[:-2, :-2] - Ihp[:-2, 1:-1] - Ihp[:-2, 2:] + Ihp[2:, :-2] + Ihp[2:, 1:-1] + Ihp[2:, 2:]) / 8.0
    yy = t.arange(y0, y1, device=device)[:, None]
    xx = t.arange(0, W, device=device)[None, :]
    border = (yy == 0) | (yy == H - 1) | (xx == 0) | (xx == W - 1)       # zero border of the derivative images
    dx = t.where(border, t.zeros_like(dx), dx)
    dy = t.where(border, t.zeros_like(dy), dy)
    f = lambda a: a.reshape(-1).to(t.float32).contiguous()
    n = (y1 - y0) * W
    return dict(w_fitSqrt=np.float32(np.sqrt(w_fit)), w_regSqrt=np.float32(np.sqrt(w_reg)),
                X=t.zeros((n, 2), device=device, dtype=t.float32), I=f(I), I_hat_im=f(Ih), I_hat_dx=f(dx), I_hat_dy=f(dy))


def sfs_inputs_torch(W, H, device, rows=None, seed=1, w_p=100.0, w_s=100.0, w_g=1.0):
    """sfs_inputs on `device`, rows [y0, y1) (the seeded depth noise is a coordinate hash instead of a NumPy stream)."""
    import torch as t
    y0, y1 = rows if rows is not None else (0, H)
    ys = t.arange(y0, y1, device=device, dtype=t.float64)[:, None]
    xs = t.arange(0, W, device=device, dtype=t.float64)[None, :]
    u, v = (xs - W / 2.0) / (W / 2.0), (ys - H / 2.0) / (H / 2.0)
    r2 = u * u + v * v
    depth = 0.5 - 0.08 * t.sqrt(t.clamp(1.0 - t.clamp(r2, max=1.0), min=0.0))
    depth = depth + 0.0005 * t.sin(0.37 * xs + 0.11 * ys) * t.cos(0.23 * ys - 0.05 * xs)
    valid = r2 < 0.81
    bad = t.full_like(depth, -10000.0)
    D = t.where(valid, depth, bad)
    noise = (_hash01(t, xs, ys, seed) + _hash01(t, xs, ys, seed + 1) + _hash01(t, xs, ys, seed + 2) - 1.5) * 2.0     # ~N(0, 1)
    X0 = t.where(valid, depth + 0.0008 * noise, bad)
    Im = 0.5 + 0.2 * t.sin(0.05 * xs + 0.02 * ys) + 0.15 * t.cos(0.031 * ys - 0.017 * xs) + 0.05 * t.sin(0.4 * xs) * t.sin(0.3 * ys)
    f = lambda a: a.reshape(-1).to(t.float32).contiguous()
    light = [0.6908317804336548, 0.044598858803510666, 0.01812959648668766, -0.1773163229227066, -0.04067882522940636,
             0.14467650651931763, 0.02393525093793869, -0.24658696353435516, 0.005797004792839289]
    n = (y1 - y0) * W
    return dict(w_p=w_p, w_s=w_s, w_g=w_g, f_x=float(W), f_y=float(W), u_x=W / 2.0, u_y=H / 2.0, light=light,
                X=f(X0), D_i=f(D), Im=f(Im), edgeMaskR=t.ones(n, device=device, dtype=t.uint8),
                edgeMaskC=t.ones(n, device=device, dtype=t.uint8))


def bundle_adjustment_inputs_torch(C, P, device, obs_per_point=5, seed=1, noise_px=0.5, perturb=0.01):
    """bundle_adjustment_inputs on `device` (torch's generator instead of NumPy's: same shapes and distributions)."""
    import torch as t
    g = t.Generator(device=device)
    g.manual_seed(seed)
    k = int(obs_per_point)
    assert C >= k
    f64 = dict(device=device, dtype=t.float64)
    rn = lambda *s: t.randn(*s, generator=g, **f64)
    theta = 2.0 * np.pi * t.arange(C, **f64) / C
    cams = t.zeros((C, 9), **f64)
    cams[:, 1] = theta
    cams[:, 0] = 0.05 * rn(C)
    cams[:, 2] = 0.05 * rn(C)
    cams[:, 3:5] = 0.1 * rn(C, 2)
    cams[:, 5] = -5.0
    cams[:, 6] = 800.0 + 20.0 * rn(C)
    cams[:, 7] = 1e-2 * rn(C)
    cams[:, 8] = 1e-3 * rn(C)
    pts = t.rand((P, 3), generator=g, **f64) * 2.0 - 1.0
    base = t.randint(0, C, (P,), generator=g, device=device)
    steps = 1 + t.randint(0, max(1, (C - 1) // k), (P, k), generator=g, device=device)
    steps[:, 0] = 0
    cam_of = t.sort((base[:, None] + t.cumsum(steps, 1)) % C, 1).values
    o2c = cam_of.reshape(-1).to(t.int32).contiguous()
    o2p = t.arange(P, device=device, dtype=t.int32).repeat_interleave(k).contiguous()
    cam = cams[o2c.long()]
    X = pts[o2p.long()]
    aa = cam[:, 0:3]
    th2 = (aa * aa).sum(-1, keepdim=True)
    th = t.sqrt(t.clamp(th2, min=1e-30))
    w = aa / th
    c, s = t.cos(th), t.sin(th)
    large = X * c + t.linalg.cross(w, X) * s + w * ((w * X).sum(-1, keepdim=True) * (1.0 - c))
    small = X + t.linalg.cross(aa, X)
    p = t.where(th2 > 1e-8, large, small) + cam[:, 3:6]
    cod = -p[:, 0:2] / p[:, 2:3]
    r2 = (cod * cod).sum(-1, keepdim=True)
    obs = cod * cam[:, 6:7] * (1.0 + r2 * (cam[:, 7:8] + cam[:, 8:9] * r2))
    obs = obs + noise_px * rn(*obs.shape)
    cams0 = cams * (1.0 + perturb * rn(*cams.shape))
    pts0 = pts + perturb * rn(*pts.shape)
    return dict(cameras=cams0.to(t.float32).contiguous(), points=pts0.to(t.float32).contiguous(),
                observations=obs.to(t.float32).contiguous(), oToC=o2c, oToP=o2p)
