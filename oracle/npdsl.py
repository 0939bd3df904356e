"""ORACLE (test infrastructure, not product code): numeric evaluation of the
energy DSL with NumPy forward-mode dual numbers.

Implements the DSL namespace `L` that `energies/*.py` are written against, but
instead of building a symbolic DAG (the product's front end) every expression
is evaluated immediately over the whole residual domain as arrays, carrying
exact partial derivatives per unknown access.  The result is the residual
vector F and the Jacobian J as a SciPy CSR matrix.  This is independent of
`thallo_b200.frontend` (different AD method, different evaluator), so it can
check the generated CUDA functions.

Reference semantics restated here:
  * image layout / offsets: API/src/thallo.t:609-738 (x fastest), :759-1017
  * out-of-bounds `get` returns 0: thallo.t:876-882; scatters dropped OOB: :3355-3390
  * InBounds / InBoundsExpanded: thallo.t:2091-2112
  * bilinear sample floor/ceil lerp: thallo.t:899-907; sampled-image partials: :5803-5817
  * derivative rules: API/src/ad.t:698-836 (select, abs, comparisons have zero
    derivative w.r.t. the condition)
  * unknown numbering for J columns: gauss_newton.t:157-160,448-451
Pinned (parity pinned): with oracle/solver.py on the reference's two golden images (tests/test_oracle_golden.py),
and -- cost, -J^T F and J^T J p of image_warping, arap_mesh_deformation and volumetric_mesh_deformation -- on the
reference authors' hand-derived CUDA equations compiled for the host from the reference tree
(oracle/*_hand_host.cpp, tests/test_oracle_hand_equations.py), as is the stored shading term of shape_from_shading
with its gradient image.  optical_flow and bundle_adjustment have neither and are checked only against the product's
independent symbolic AD.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this.
"""
import numpy as np
import scipy.sparse as sp


class Dim:
    def __init__(self, name, idx, size, L=None):
        self.name, self.idx, self.size, self.L = name, idx, size, L

    def __call__(self):
        return IdxVar(self, 0)


class IdxVar:
    def __init__(self, dim, off):
        self.dim, self.off = dim, off

    def __add__(self, k):
        return IdxVar(self.dim, self.off + int(k))

    __radd__ = __add__

    def __sub__(self, k):
        return IdxVar(self.dim, self.off - int(k))

    def asvalue(self):
        return self.dim.L._idx_value(self)


class SparseRef:
    def __init__(self, sparse, iv):
        self.sparse, self.iv = sparse, iv


class NBool:
    def __init__(self, v):
        self.v = np.asarray(v, dtype=bool)

    def __mul__(self, o):
        if isinstance(o, NBool):
            return NBool(self.v & o.v)
        if isinstance(o, Vec):
            return Vec([self * c for c in o.c])
        return _select(self, o, 0.0)

    __rmul__ = __mul__

    def get(self, *idx):                 # a stored boolean is a 0/1 image of the solver's scalar type
        L = idx[0].dim.L
        return Dual(self.v.astype(L.dtype)).get(*idx)


class Dual:
    __array_priority__ = 1000

    def __init__(self, val, d=None):
        self.val = val
        self.d = d or {}

    # ---- arithmetic
    def _lift(self, o):
        if isinstance(o, Dual):
            return o
        if isinstance(o, NBool):
            return Dual(o.v.astype(np.result_type(self.val)))
        return Dual(np.asarray(o, dtype=np.result_type(self.val)) if not np.isscalar(o) else np.result_type(self.val).type(o))

    def __add__(self, o):
        if isinstance(o, Vec):
            return NotImplemented
        o = self._lift(o)
        d = dict(self.d)
        for k, v in o.d.items():
            d[k] = d[k] + v if k in d else v
        return Dual(self.val + o.val, d)

    __radd__ = __add__

    def __neg__(self):
        return Dual(-self.val, {k: -v for k, v in self.d.items()})

    def __sub__(self, o):
        if isinstance(o, Vec):
            return NotImplemented
        return self + (-self._lift(o))

    def __rsub__(self, o):
        return self._lift(o) + (-self)

    def __mul__(self, o):
        if isinstance(o, Vec):
            return NotImplemented
        if isinstance(o, NBool):
            return _select(o, self, 0.0)
        o = self._lift(o)
        d = {k: v * o.val for k, v in self.d.items()}
        for k, v in o.d.items():
            t = v * self.val
            d[k] = d[k] + t if k in d else t
        return Dual(self.val * o.val, d)

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Vec):
            return NotImplemented
        o = self._lift(o)
        return self * _recip(o)

    def __rtruediv__(self, o):
        return self._lift(o) * _recip(self)

    def __pow__(self, c):
        assert np.isscalar(c)
        v = self.val ** c
        dv = c * self.val ** (c - 1)
        return Dual(v, {k: g * dv for k, g in self.d.items()})

    def __getitem__(self, i):  # scalar treated as 1-vector
        assert i == 0
        return self

    def __len__(self):
        return 1

    def get(self, *idx):
        """`exp:get(x+ox, y+oy)`: the expression as a stored image over its domain (ComputedArray,
        thallo.t:1777-1822), read at an offset with zero outside the domain (thallo.t:876-882).  Its
        derivatives travel with it as the gradient image, read at the same offset and re-keyed to
        the unknown accesses they belong to, shifted by that offset (thallo.t:1551-1561)."""
        if len(idx) == 1 and isinstance(idx[0], SparseRef):
            # exp:get(v(e)) (tests/minimal_sparse_materialize/minimal_sparse_materialize.t:17): the stored
            # image and its gradient image are read at element v(e); the derivatives belong to the
            # unknowns at that element
            ref = idx[0]
            L = ref.iv.dim.L
            e = L._sparse_values(ref)
            n = ref.sparse.to[0].size

            def at(a):
                return np.ascontiguousarray(np.broadcast_to(np.asarray(a), (n,)))[e]
            d = {}
            for (iname, ikey, ch), dv in self.d.items():
                assert ikey[0] == "d" and not any(ikey[1]), \
                    "a computed array fetched through a sparse index may only read unknowns at its own element"
                d[(iname, ("s", ref.sparse.name, ref.iv.off), ch)] = at(dv)
            return Dual(at(self.val), d)
        offs = [iv.off for iv in idx]
        shape = tuple(iv.dim.size for iv in reversed(idx))

        def moved(a):
            full = np.ascontiguousarray(np.broadcast_to(np.asarray(a), shape))
            return _shift(full[..., None], offs)[..., 0]
        d = {}
        for (iname, ikey, ch), dv in self.d.items():
            assert ikey[0] == "d", "computed arrays over sparse accesses are not supported"
            d[(iname, ("d", tuple(o + s for o, s in zip(ikey[1], offs))), ch)] = moved(dv)
        return Dual(moved(self.val), d)


def _recip(o):
    with np.errstate(divide="ignore", invalid="ignore"):
        v = 1.0 / o.val if not np.isscalar(o.val) else type(o.val)(1.0) / o.val
        dv = -(v * v)
    return Dual(v, {k: g * dv for k, g in o.d.items()})


def _unary(x, f, df):
    if not isinstance(x, Dual):
        return f(x)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = f(x.val)
        dv = df(x.val, v)
    return Dual(v, {k: g * dv for k, g in x.d.items()})


def _select(c, a, b):
    if isinstance(a, Vec) or isinstance(b, Vec) or isinstance(c, Vec):
        n = max(len(v) for v in (a, b, c) if isinstance(v, Vec))
        aa = a.c if isinstance(a, Vec) else [a] * n
        bb = b.c if isinstance(b, Vec) else [b] * n
        cc = c.c if isinstance(c, Vec) else [c] * n
        return Vec([_select(z, x, y) for z, x, y in zip(cc, aa, bb)])
    cv = c.v
    ref = a if isinstance(a, Dual) else (b if isinstance(b, Dual) else None)
    dt = np.result_type(ref.val) if ref is not None else np.float64
    A = a if isinstance(a, Dual) else Dual(np.asarray(a, dtype=dt))
    B = b if isinstance(b, Dual) else Dual(np.asarray(b, dtype=dt))
    val = np.where(cv, A.val, B.val)
    d = {}
    zero = dt.type(0) if hasattr(dt, "type") else 0.0
    for k in set(A.d) | set(B.d):
        d[k] = np.where(cv, A.d.get(k, zero), B.d.get(k, zero))
    return Dual(val, d)


class Vec:
    def __init__(self, comps):
        self.c = list(comps)

    def __len__(self):
        return len(self.c)

    def __getitem__(self, i):
        return self.c[i]

    def __call__(self, i):
        return self.c[i]

    def get(self, *idx):
        return Vec([c.get(*idx) if (hasattr(c, "get") and (not isinstance(c, Dual) or c.d or np.ndim(c.val) > 0)) else c
                    for c in self.c])

    def dot(self, o):
        r = self.c[0] * o.c[0]
        for a, b in zip(self.c[1:], o.c[1:]):
            r = r + a * b
        return r

    def slice(self, a, b):
        return Vec(self.c[a:b])

    def _bin(self, o, f):
        if isinstance(o, Vec):
            assert len(o) == len(self)
            return Vec([f(x, y) for x, y in zip(self.c, o.c)])
        return Vec([f(x, o) for x in self.c])

    def __add__(self, o):
        return self._bin(o, lambda x, y: x + y)

    def __radd__(self, o):
        return self._bin(o, lambda x, y: y + x)

    def __sub__(self, o):
        return self._bin(o, lambda x, y: x - y)

    def __rsub__(self, o):
        return self._bin(o, lambda x, y: y - x)

    def __mul__(self, o):
        return self._bin(o, lambda x, y: x * y)

    def __rmul__(self, o):
        return self._bin(o, lambda x, y: y * x)

    def __truediv__(self, o):
        return self._bin(o, lambda x, y: x / y)

    def __neg__(self):
        return Vec([-x for x in self.c])


class NImage:
    def __init__(self, L, name, channels, dims, pidx, kind, np_dtype):
        self.L, self.name, self.channels, self.dims, self.pidx, self.kind = L, name, channels, dims, pidx, kind
        self.np_dtype = np_dtype
        self.exclude = None

    @property
    def shape(self):  # (.., H, W)
        return tuple(d.size for d in reversed(self.dims))

    def data(self):
        arr = np.asarray(self.L.params[self.pidx])
        if self.np_dtype is None:
            arr = arr.astype(self.L.dtype, copy=False)
        return arr.reshape(self.shape + (self.channels,))

    def cardinality(self):
        return int(np.prod(self.shape)) * self.channels

    def Exclude(self, cond):
        self.exclude = cond

    def __call__(self, *idx):
        L = self.L
        data = self.data()
        comps = []
        if len(idx) == 1 and isinstance(idx[0], SparseRef):
            ref = idx[0]
            e = L._sparse_values(ref)
            for ch in range(self.channels):
                val = data.reshape(-1, self.channels)[e, ch].astype(L.dtype)
                d = {}
                if self.kind == "unknown":
                    d[(self.name, ("s", ref.sparse.name, ref.iv.off), ch)] = np.ones_like(val)
                comps.append(Dual(val, d))
        else:
            assert len(idx) == len(self.dims)
            offs = []
            for iv, dim in zip(idx, self.dims):
                assert isinstance(iv, IdxVar) and iv.dim is dim, "image index must follow its declared dims"
                offs.append(iv.off)
            L._note_domain(self.dims)
            shifted = _shift(data, offs)
            for ch in range(self.channels):
                val = shifted[..., ch].astype(L.dtype)
                d = {}
                if self.kind == "unknown":
                    d[(self.name, ("d", tuple(offs)), ch)] = np.ones_like(val)
                comps.append(Dual(val, d))
        return comps[0] if self.channels == 1 else Vec(comps)


def _shift(data, offs):
    """data[..., y, x, c] -> array whose [.., y, x] entry is data[.., y+oy, x+ox] or 0 when OOB.
    offs are per declared dim (x first)."""
    out = np.zeros_like(data)
    nd = len(offs)
    src, dst = [], []
    for axis in range(nd):  # axis 0 of array = slowest dim = last declared dim
        o = offs[nd - 1 - axis]
        n = data.shape[axis]
        if abs(o) >= n:
            return out
        if o >= 0:
            src.append(slice(o, n)); dst.append(slice(0, n - o))
        else:
            src.append(slice(0, n + o)); dst.append(slice(-o, n))
    out[tuple(dst)] = data[tuple(src)]
    return out


class NSparse:
    def __init__(self, L, name, frm, to, pidx):
        self.L, self.name, self.frm, self.to, self.pidx = L, name, frm, to, pidx

    def set_coherent(self, b):           # scheduling hint (warp-aggregated atomics); no effect on the energy
        self.coherent = bool(b)

    def __call__(self, iv):
        assert isinstance(iv, IdxVar) and iv.dim is self.frm[0]
        return SparseRef(self, iv)


class _SampledImage:
    def __init__(self, L, im, dx, dy):
        self.L, self.im, self.dx, self.dy = L, im, dx, dy

    def _sample(self, img, x, y):
        data = img.data()[..., 0].astype(self.L.dtype)
        H, W = data.shape

        def get(ix, iy):
            inb = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
            v = data[np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)]
            return np.where(inb, v, self.L.dtype(0))

        dt = self.L.dtype
        x0, x1 = np.floor(x).astype(np.int64), np.ceil(x).astype(np.int64)
        y0, y1 = np.floor(y).astype(np.int64), np.ceil(y).astype(np.int64)
        xn, yn = (x - x0.astype(dt)).astype(dt), (y - y0.astype(dt)).astype(dt)
        one = dt(1)
        u = (one - xn) * get(x0, y0) + xn * get(x1, y0)
        b = (one - xn) * get(x0, y1) + xn * get(x1, y1)
        return (one - yn) * u + yn * b

    def __call__(self, x, y):
        L = self.L
        x, y = L._todual(x), L._todual(y)
        if L.origin:                      # absolute coordinates -> coordinates of the crop
            ox, oy = L.origin.get(self.im.dims[0].idx, 0), L.origin.get(self.im.dims[1].idx, 0)
            x, y = Dual(x.val - L.dtype(ox), x.d), Dual(y.val - L.dtype(oy), y.d)
        val = self._sample(self.im, x.val, y.val)
        d = {}
        if x.d or y.d:
            gx = self._sample(self.dx, x.val, y.val)
            gy = self._sample(self.dy, x.val, y.val)
            for k, g in x.d.items():
                d[k] = g * gx
            for k, g in y.d.items():
                d[k] = d.get(k, 0) + g * gy
        return Dual(val, d)


class _Sched:
    def __init__(self):
        self.materialize = False

    def set_materialize(self, b):
        self.materialize = bool(b)


class _Group:
    def __init__(self, name, terms):
        self.name, self.terms = name, terms
        self.J, self.JtJ, self.Jp = _Sched(), _Sched(), _Sched()


class _Residuals:
    def __init__(self, groups):
        self.groups = groups
        for g in groups:
            setattr(self, g.name, g)

    def merge(self, *groups):            # scheduling hint only
        return groups[0] if groups else None


class NumpyL:
    """DSL namespace bound to concrete dimension sizes and parameter arrays."""
    float, float2, float3, float4, float9 = ("f", 1), ("f", 2), ("f", 3), ("f", 4), ("f", 9)
    uint8, int = ("u8", 1), ("i32", 1)

    def __init__(self, dims, params, dtype=np.float32, origin=None):
        # origin: {dimension index: offset} -- the arrays are a crop of a larger problem that starts at this absolute
        # index; index VALUES (x:asvalue()) are absolute, sampled images are addressed relative to the crop
        self.origin = dict(origin or {})
        self.dim_sizes = list(dims)
        self.params = list(params)
        self.dtype = np.dtype(dtype).type
        self.images, self.sparses, self.param_defs = [], [], []
        self.usepreconditioner = False   # default false, thallo.t:115
        self._domain = None

    # --- declarations
    def Dims(self, *names):
        self.dims = [Dim(n, i, int(self.dim_sizes[i]), self) for i, n in enumerate(names)]
        return self.dims if len(names) > 1 else self.dims[0]

    def Dim(self, name, idx):            # thallo.Dim(name, idx): one dimension at a time (older energy files)
        if not hasattr(self, "dims"):
            self.dims = []
        assert idx == len(self.dims), "Dim() indices must be declared in order"
        self.dims.append(Dim(name, idx, int(self.dim_sizes[idx]), self))
        return self.dims[-1]

    def Unknown(self, t, dims, pidx):
        return ("Unknown", t, dims, pidx)

    def Array(self, t, dims, pidx):
        return ("Array", t, dims, pidx)

    def Sparse(self, frm, to, pidx):
        return ("Sparse", frm, to, pidx)

    def Param(self, t, pidx):
        return ("Param", t, pidx)

    def Inputs(self, **kw):
        class NS:
            pass
        ns = NS()
        for name, decl in kw.items():
            if decl[0] in ("Unknown", "Array"):
                _, t, dims, pidx = decl
                np_dt = None if t[0] == "f" else {"u8": np.uint8, "i32": np.int32}[t[0]]
                im = NImage(self, name, t[1], dims, pidx, decl[0].lower(), np_dt)
                self.images.append(im)
                setattr(ns, name, im)
            elif decl[0] == "Sparse":
                s = NSparse(self, name, decl[1], decl[2], decl[3])
                self.sparses.append(s)
                setattr(ns, name, s)
            else:
                v = self.dtype(np.asarray(self.params[decl[2]]).reshape(-1)[0])
                setattr(ns, name, Dual(v))
        self.unknowns = sorted([i for i in self.images if i.kind == "unknown"], key=lambda i: i.pidx)
        return ns

    def UsePreconditioner(self, b):
        self.usepreconditioner = bool(b)

    # --- helpers
    def _note_domain(self, dims):
        self._domain = dims

    def _sparse_values(self, ref):
        arr = np.asarray(self.params[ref.sparse.pidx]).astype(np.int64).reshape(-1)
        assert ref.iv.off == 0
        return arr

    def _coords(self, dims):
        shape = tuple(d.size for d in reversed(dims))
        grids = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
        return {d.name: grids[len(dims) - 1 - i] for i, d in enumerate(dims)}

    def _idx_value(self, iv):
        dims = [iv.dim]
        for im in self.images:
            if any(d is iv.dim for d in im.dims):
                dims = im.dims
                break
        g = self._coords(dims)[iv.dim.name] + iv.off + self.origin.get(iv.dim.idx, 0)
        return Dual(g.astype(self.dtype))

    def _todual(self, x):
        if isinstance(x, Dual):
            return x
        if isinstance(x, NBool):
            return Dual(x.v.astype(self.dtype))
        return Dual(self.dtype(x))

    # --- expression constructors
    def Vector(self, *c):
        return Vec(c)

    def _u(self, x, f, df):
        if isinstance(x, Vec):
            return Vec([self._u(c, f, df) for c in x.c])
        return _unary(self._todual(x), f, df)

    def sqrt(self, x):
        return self._u(x, np.sqrt, lambda v, r: 1.0 / (2.0 * r))

    def sin(self, x):
        return self._u(x, np.sin, lambda v, r: np.cos(v))

    def cos(self, x):
        return self._u(x, np.cos, lambda v, r: -np.sin(v))

    def exp(self, x):
        return self._u(x, np.exp, lambda v, r: r)

    def log(self, x):
        return self._u(x, np.log, lambda v, r: 1.0 / v)

    def abs(self, x):
        return self._u(x, np.abs, lambda v, r: np.where(v >= 0, 1.0, -1.0).astype(np.result_type(v)))

    def _cmp(self, a, b, f):
        if isinstance(a, Vec):
            bb = b.c if isinstance(b, Vec) else [b] * len(a)
            return Vec([self._cmp(x, y, f) for x, y in zip(a.c, bb)])
        a, b = self._todual(a), self._todual(b)
        return NBool(f(a.val, b.val))

    def eq(self, a, b): return self._cmp(a, b, np.equal)
    def neq(self, a, b): return self._cmp(a, b, np.not_equal)
    def less(self, a, b): return self._cmp(a, b, np.less)
    def greater(self, a, b): return self._cmp(a, b, np.greater)
    def lesseq(self, a, b): return self._cmp(a, b, np.less_equal)
    def greatereq(self, a, b): return self._cmp(a, b, np.greater_equal)

    def Not(self, b):
        return NBool(~b.v)

    def And(self, *bs):
        r = bs[0]
        for b in bs[1:]:
            r = r * b
        return r

    def Or(self, *bs):
        r = bs[0].v
        for b in bs[1:]:
            r = r | b.v
        return NBool(r)

    def Select(self, c, a, b):
        return _select(c, a, b)

    def InBounds(self, *idx):
        dims = [iv.dim for iv in idx]
        g = self._coords(dims)
        ok = None
        for iv in idx:
            c = g[iv.dim.name] + iv.off
            t = (c >= 0) & (c < iv.dim.size)
            ok = t if ok is None else ok & t
        return NBool(ok)

    def InBoundsExpanded(self, *args):
        *idx, e = args
        dims = [iv.dim for iv in idx]
        g = self._coords(dims)
        ok = None
        for iv in idx:
            c = g[iv.dim.name] + iv.off
            t = (c - e >= 0) & (c + e < iv.dim.size)
            ok = t if ok is None else ok & t
        return NBool(ok)

    def SampledImage(self, im, dx=None, dy=None):
        return _SampledImage(self, im, dx, dy)

    def Residuals(self, **kw):
        groups = []
        for name in sorted(kw):          # named residuals sorted by name, thallo.t:5780
            v = kw[name]
            terms = []
            for item in (v if isinstance(v, (list, tuple)) else [v]):
                if isinstance(item, Vec):
                    terms.extend(item.c)
                else:
                    terms.append(item)
            groups.append(_Group(name, [self._todual(t) for t in terms]))
        self.residuals = _Residuals(groups)
        return self.residuals

    # --- assembly -------------------------------------------------------
    def unknown_offsets(self):
        off, base = {}, 0
        for im in self.unknowns:
            off[im.name] = base
            base += im.cardinality()
        return off, base

    def exclude_mask(self):
        """bool per flat unknown scalar: True where the unknown is excluded (thallo.t:5618-5624)."""
        off, n = self.unknown_offsets()
        m = np.zeros(n, dtype=bool)
        for im in self.unknowns:
            if im.exclude is not None:
                e = np.broadcast_to(im.exclude.v, im.shape).reshape(-1)
                m[off[im.name]:off[im.name] + im.cardinality()] = np.repeat(e, im.channels)
        return m

    def assemble(self):
        """Return F (n_res,), J (CSR n_res x n_unknown scalars), group row ranges."""
        off, nunk = self.unknown_offsets()
        by_name = {im.name: im for im in self.unknowns}
        sparse_by_name = {s.name: s for s in self.sparses}
        Fs, rows, cols, vals, ranges = [], [], [], [], []
        row0 = 0
        for g in self.residuals.groups:
            gstart = row0
            # domain of the group = broadcast shape of its terms
            shape = np.broadcast_shapes(*[np.shape(t.val) for t in g.terms])
            n = int(np.prod(shape)) if shape else 1
            nt = len(g.terms)
            for ti, t in enumerate(g.terms):
                val = np.broadcast_to(np.asarray(t.val, dtype=self.dtype), shape).reshape(-1)
                ridx = gstart + np.arange(n) * nt + ti   # element-major, residual-minor (gauss_newton.t:401-402)
                F = np.zeros(0)
                Fs.append((ridx, val))
                for (iname, ikey, ch), dv in t.d.items():
                    im = by_name[iname]
                    dv = np.broadcast_to(np.asarray(dv, dtype=self.dtype), shape).reshape(-1)
                    if ikey[0] == "d":
                        offs = ikey[1]
                        dims = im.dims
                        g_ = self._coords(dims)
                        lin = np.zeros(shape, dtype=np.int64)
                        ok = np.ones(shape, dtype=bool)
                        stride = 1
                        for dmn, o in zip(dims, offs):
                            c = g_[dmn.name] + o
                            ok &= (c >= 0) & (c < dmn.size)
                            lin = lin + stride * c
                            stride *= dmn.size
                        lin, ok = lin.reshape(-1), ok.reshape(-1)
                    else:
                        s = sparse_by_name[ikey[1]]
                        lin = np.asarray(self.params[s.pidx]).astype(np.int64).reshape(-1)
                        ok = np.ones(n, dtype=bool)
                    col = off[iname] + im.channels * lin + ch
                    rows.append(ridx[ok]); cols.append(col[ok]); vals.append(dv[ok])
            row0 = gstart + n * nt
            ranges.append((g.name, gstart, row0))
        F = np.zeros(row0, dtype=self.dtype)
        for ridx, val in Fs:
            F[ridx] = val
        if rows:
            J = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                              shape=(row0, nunk), dtype=self.dtype)
            # squares of the partials PER UNKNOWN ACCESS, which is what the reference's PCGInit1 accumulates into the
            # preconditioner (createjtfResidualwise, thallo.t:3897-3901: one `partial*partial` scatter per access).  It
            # differs from diag(J^T J) only where two accesses of one residual hit the same unknown (self loops,
            # coincident index arrays): there J holds the sum of the partials and its square is not the sum of squares.
            v2 = np.concatenate(vals)
            self.diag_sq = np.asarray(sp.csr_matrix((v2 * v2, (np.concatenate(rows), np.concatenate(cols))),
                                                    shape=(row0, nunk), dtype=self.dtype).sum(axis=0)).reshape(-1)
        else:
            J = sp.csr_matrix((row0, nunk), dtype=self.dtype)
            self.diag_sq = np.zeros(nunk, self.dtype)
        return F, J, ranges


def evaluate(define, dims, params, dtype=np.float32, **kw):
    """Evaluate energy `define` at the unknown values currently stored in `params`.
    Returns (L, F, J)."""
    L = NumpyL(dims, params, dtype)
    define(L, **kw)
    F, J, _ = L.assemble()
    return L, F, J
