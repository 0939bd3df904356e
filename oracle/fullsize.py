"""ORACLE-side checker for FULL-SIZE runs (test infrastructure, not product code).

The NumPy oracle (oracle/npdsl.py + oracle/solver.py) cannot evaluate an 8192 x 8192 image or a 25 M-observation
bundle adjustment in useful time, but everything the solver computes before and inside its first PCG iteration is LOCAL:
r0 = -J^T F, the preconditioner, the LM diagonal and A p = (J^T J [+ C]) p of an unknown depend only on the residuals
that touch it and on the unknowns those residuals read.  So a crop of the problem (a band of rows / layers, a range of
vertices with the edges inside it, a range of points with their observations), evaluated by the float64 oracle, must
reproduce the corresponding elements of the GPU's full-size vectors -- away from the artificial crop boundary by the
reach of the stencil.  The dot products, which are global, are checked by recomputing alpha = <r0, p0> / <p0, A p0> in
float64 from the GPU's own full-size vectors and comparing it with the alpha the solver used (delta = alpha p0).

Used by tests/test_gpu_fullsize.py and by bench.py's `parity` record (after the timed region; checker only).
"""
import numpy as np

import energies
from .solver import OracleSolver


def _np64(t):
    a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
    return a.astype(np.float64) if a.dtype == np.float32 else a


# ---------------------------------------------------------------------------------------------------- crops
class Crop:
    """dims, params (NumPy, float64 where the GPU holds float32), origin, and per unknown image the element range
    [e0, e1) of the full problem the crop covers plus the sub-range [c0, c1) (in crop-local elements) whose vectors
    are comparable with the full problem's."""

    def __init__(self, dims, params, origin, ranges):
        self.dims, self.params, self.origin, self.ranges = dims, params, origin, ranges


def slab_crop(dims, params, a, b, reach, unknown_pidx):
    """Layers [a, b) of the slowest axis of an image / volume domain (every array parameter lives on that domain)."""
    layer = int(np.prod(dims[:-1]))
    D = int(dims[-1])
    a, b = max(0, a), min(D, b)
    out = []
    for p in params:
        if hasattr(p, "shape") and int(np.prod(p.shape)) > 1:
            assert p.shape[0] == layer * D, "array parameter does not live on the image domain"
            out.append(_np64(p[a * layer:b * layer]).copy())
        else:
            out.append(p)
    c0 = 0 if a == 0 else 2 * reach            # A p0 reaches p0 within `reach`, p0 = M r0 is local: one reach would do for
    c1 = (b - a) if b == D else (b - a) - 2 * reach      # r0 / M, two are needed nowhere -- kept for margin
    ranges = {pidx: (a * layer, b * layer, c0 * layer, c1 * layer) for pidx in unknown_pidx}
    return Crop(list(dims[:-1]) + [b - a], out, {len(dims) - 1: a}, ranges)


def graph_crop(dims, params, va, vb, vertex_slots, index_slots, unknown_pidx):
    """Vertices [va, vb) of a graph energy over Dims(N, E) with the edges that lie inside the range."""
    N = int(dims[0])
    va, vb = max(0, va), min(N, vb)
    idx = [_np64(params[s]).astype(np.int64).reshape(-1) for s in index_slots]
    inside = np.ones(len(idx[0]), bool)
    reach = 0
    for k in idx:
        inside &= (k >= va) & (k < vb)
    for k in idx[1:]:
        reach = max(reach, int(np.abs(k - idx[0]).max()))
    out = list(params)
    for s in vertex_slots:
        out[s] = _np64(params[s][va:vb]).copy()
    for s, k in zip(index_slots, idx):
        out[s] = (k[inside] - va).astype(np.int32)
    c0 = 0 if va == 0 else 2 * reach
    c1 = (vb - va) if vb == N else (vb - va) - 2 * reach
    ranges = {pidx: (va, vb, c0, c1) for pidx in unknown_pidx}
    return Crop([vb - va, int(inside.sum())], out, None, ranges)


def ba_crop(dims, params, pa, pb):
    """Points [pa, pb) of bundle_adjustment (Dims C, P, O; params cameras, points, observations, oToC, oToP) with all of
    their observations (oToP is sorted) and every camera; only the point unknowns are comparable."""
    C_, P_, O_ = [int(x) for x in dims]
    o2p = params[4]
    o2p_np = o2p.detach().cpu().numpy() if hasattr(o2p, "detach") else np.asarray(o2p)
    oa, ob = int(np.searchsorted(o2p_np, pa, "left")), int(np.searchsorted(o2p_np, pb, "left"))
    out = [_np64(params[0]).copy(), _np64(params[1][pa:pb]).copy(), _np64(params[2][oa:ob]).copy(),
           np.ascontiguousarray(_np64(params[3][oa:ob]).astype(np.int32)), (o2p_np[oa:ob].astype(np.int64) - pa).astype(np.int32)]
    # unknown pidx 0 = cameras (whole image present, nothing comparable: their sums need every observation), 1 = points
    ranges = {0: (0, C_, 0, 0), 1: (pa, pb, 0, pb - pa)}
    return Crop([C_, pb - pa, ob - oa], out, None, ranges)


# ---------------------------------------------------------------------------------------------------- the check
def _gather_local(vec, desc, crop):
    """Full-size device vector -> the crop's unknown vector (float64 NumPy), unknown images back to back."""
    parts = []
    for u in desc["unknowns"]:
        e0, e1, _, _ = crop.ranges[u["pidx"]]
        ch = u["channels"]
        parts.append(vec[u["offset"] + e0 * ch:u["offset"] + e1 * ch].double().cpu().numpy())
    return np.concatenate(parts)


def _masks(desc, crop):
    m, off = [], 0
    for u in desc["unknowns"]:
        e0, e1, c0, c1 = crop.ranges[u["pidx"]]
        ch = u["channels"]
        k = np.zeros((e1 - e0) * ch, bool)
        k[max(0, c0) * ch:max(0, c1) * ch] = True
        m.append(k)
    return np.concatenate(m)


def first_iteration_parity(make_solver, fresh_params, energy, kind, crops, mode, define_kwargs=None, materialized=False,
                           solver_params=None):
    """`make_solver()` -> a ThalloSolver of the full-size problem; `fresh_params()` -> its parameter list in the initial
    state (device tensors); `crops(params)` -> list of (label, Crop).  Returns a dict of relative differences:
    r0, preconditioner, Ap per crop (max |gpu - oracle| / max |oracle| over the comparable elements), and alpha."""
    import torch
    sp = dict(solver_params or {})
    s = make_solver()
    desc = s.lowered.desc
    pname = "z" if desc.get("tiled") else "p"
    # run A: PCGInit only
    pa = fresh_params()
    s.set_parameters(**dict(sp, nIterations=1, lIterations=0, trust_region_radius=1e4))
    s.init(pa)
    s.step()
    torch.cuda.synchronize()
    r0, pre, p0 = s.vector("r").clone(), s.vector("preconditioner").clone(), s.vector(pname).clone()
    # run B: one PCG iteration from the same state
    pb = fresh_params()
    s.set_parameters(**dict(sp, nIterations=1, lIterations=1, trust_region_radius=1e4))
    s.init(pb)
    s.step()
    torch.cuda.synchronize()
    Ap, delta = s.vector("Ap_X"), s.vector("delta")
    p64 = p0.double()
    pp = float(torch.dot(p64, p64))
    alpha_gpu = float(torch.dot(delta.double(), p64)) / pp if pp > 0 else 0.0
    den = float(torch.dot(p64, Ap.double()))
    alpha_ref = float(torch.dot(r0.double(), p64)) / den if den != 0 else 0.0
    rec = {"alpha_rel": abs(alpha_gpu - alpha_ref) / max(abs(alpha_ref), 1e-300), "alpha": alpha_gpu, "crops": {}}
    define = energies.load(energy)
    params0 = fresh_params()
    worst = 0.0
    for label, crop in crops(params0):
        o = OracleSolver(define, crop.dims, kind, np.float64, mode, materialized=materialized, define_kwargs=define_kwargs,
                         origin=crop.origin)
        for k, v in sp.items():
            o.set(k, v)
        sv = o.setup_vectors(crop.params, radius=1e4)
        mask = _masks(desc, crop)
        got = {"r0": _gather_local(r0, desc, crop), "preconditioner": _gather_local(pre, desc, crop),
               "Ap": _gather_local(Ap, desc, crop)}
        ref = {"r0": sv["r"], "preconditioner": sv["M"], "Ap": sv["applyA"](_gather_local(p0, desc, crop))}
        out = {"elements_compared": int(mask.sum())}
        for k in got:
            scale = max(float(np.abs(ref[k][mask]).max()), 1e-300) if mask.any() else 1.0
            out[k] = float(np.abs(got[k][mask] - ref[k][mask]).max() / scale) if mask.any() else 0.0
            worst = max(worst, out[k])
        rec["crops"][label] = out
    rec["operator_max_rel"] = worst
    s.close()
    return rec


# ---------------------------------------------------------------------------------------------------- crops of the configured cases
def crops_for(case, dims, desc, band=12):
    """`crops(params)` callable for a thallo_b200.configs case at global dims `dims`: three crops per case -- at the
    start, in the middle and at the end of the partitioned axis (the end is where byte offsets are largest)."""
    unknown_pidx = [u["pidx"] for u in desc["unknowns"]]
    part = case.partition or "slab"

    def crops(params):
        out = []
        if part == "slab":
            D = int(dims[-1])
            reach = int(desc["tile"]["halo"][len(dims) - 1]) if desc.get("tiled") else 2
            h = band + 4 * reach
            for label, a in (("first", 0), ("middle", max(0, D // 2 - h // 2)), ("last", max(0, D - h))):
                out.append((label, slab_crop(dims, params, a, a + h, reach, unknown_pidx)))
        elif part == "graph":
            N = int(dims[0])
            nx = int(round(N ** 0.5))
            h = (band + 4) * nx + 4 * (nx + 1)
            vs = [i for i, u in enumerate(params) if hasattr(u, "shape") and len(u.shape) == 2 and u.shape[0] == N]
            ix = [i for i, u in enumerate(params) if hasattr(u, "shape") and len(u.shape) == 1 and u.shape[0] == int(dims[1])]
            for label, a in (("first", 0), ("middle", max(0, N // 2 - h // 2)), ("last", max(0, N - h))):
                out.append((label, graph_crop(dims, params, a, a + h, vs, ix, unknown_pidx)))
        else:
            P_ = int(dims[1])
            n = min(P_, 400)
            for label, a in (("first", 0), ("middle", max(0, P_ // 2 - n // 2)), ("last", max(0, P_ - n))):
                out.append((label, ba_crop(dims, params, a, a + n)))
        return out
    return crops
