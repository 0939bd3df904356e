"""Symbolic implementation of the energy DSL namespace `L` (the product front end).

Mirrors the surface of reference API/src/lib.t and the DSL part of thallo.t
(Dims/Inputs/Unknown/Array/Sparse/Param :93-585, InBounds/InBoundsExpanded
:2091-2112, Exclude/UsePreconditioner :5618-5624,:115, SampledImage :5784-5817,
Residuals :5774-5782).  Energies in `energies/*.py` run against this namespace and
produce a `ProblemSpec`: typed inputs plus named residual groups whose terms are
`ad.Exp` DAGs over `ImageAccess / Bounds / IndexValue / Param` variables.
"""
from collections import namedtuple

from . import ad

# ---- variable keys (hashable; ad.var(key))
# index component: ("d", dim_idx, off)  or  ("s", sparse_name, dim_idx, off)
ImageAccess = namedtuple("ImageAccess", "image index channel")
Bounds = namedtuple("Bounds", "ranges")            # ranges: tuple of (dim_idx, lo, hi) sorted by dim
IndexValue = namedtuple("IndexValue", "dim off")
Param = namedtuple("Param", "name")
VecArg = namedtuple("VecArg", "arg image index channel")   # element of an unknown-shaped vector argument (P, Delta, ...)


class Dim:
    def __init__(self, name, idx, size):
        self.name, self.idx, self.size = name, idx, size

    def __call__(self):
        return IndexVar(self, 0)

    def __repr__(self):
        return self.name


class IndexVar:
    def __init__(self, dim, off):
        self.dim, self.off = dim, off

    def __add__(self, k):
        return IndexVar(self.dim, self.off + int(k))

    __radd__ = __add__

    def __sub__(self, k):
        return IndexVar(self.dim, self.off - int(k))

    def asvalue(self):
        return ad.var(IndexValue(self.dim.idx, self.off))


class SparseRef:
    def __init__(self, sparse, iv):
        self.sparse, self.iv = sparse, iv


class Vector:
    def __init__(self, comps):
        self.c = [ad.toexp(x) if not isinstance(x, Vector) else x for x in comps]

    def __len__(self): return len(self.c)
    def __getitem__(self, i): return self.c[i]
    def __call__(self, i): return self.c[i]
    def __iter__(self): return iter(self.c)
    def slice(self, a, b): return Vector(self.c[a:b])

    def _bin(self, o, f):
        if isinstance(o, Vector):
            assert len(o) == len(self), "vector size mismatch"
            return Vector([f(x, y) for x, y in zip(self.c, o.c)])
        return Vector([f(x, o) for x in self.c])

    def __add__(self, o): return self._bin(o, lambda x, y: x + y)
    def __radd__(self, o): return self._bin(o, lambda x, y: y + x)
    def __sub__(self, o): return self._bin(o, lambda x, y: x - y)
    def __rsub__(self, o): return self._bin(o, lambda x, y: y - x)
    def __mul__(self, o): return self._bin(o, lambda x, y: x * y)
    def __rmul__(self, o): return self._bin(o, lambda x, y: y * x)
    def __truediv__(self, o): return self._bin(o, lambda x, y: x / y)
    def __neg__(self): return Vector([-x for x in self.c])

    def get(self, *idx):                 # ExpVector:get: every component stored / fetched on its own (ad.t ExpVector)
        return Vector([c.get(*idx) for c in self.c])

    def dot(self, o):
        r = self.c[0] * o.c[0]
        for a, b in zip(self.c[1:], o.c[1:]):
            r = r + a * b
        return r


class Image:
    def __init__(self, name, ctype, channels, dims, pidx, kind):
        self.name, self.ctype, self.channels, self.dims, self.pidx, self.kind = name, ctype, channels, list(dims), pidx, kind
        self.exclude = None
        self.elements = 1
        for d in self.dims:
            self.elements *= d.size

    @property
    def cardinality(self):
        return self.elements * self.channels

    def Exclude(self, cond):
        self.exclude = ad.toexp(cond)

    def index_of(self, idx):
        if len(idx) == 1 and isinstance(idx[0], SparseRef):
            r = idx[0]
            assert len(self.dims) == 1 and r.sparse.to[0] is self.dims[0], \
                "sparse %s does not index into the domain of %s" % (r.sparse.name, self.name)
            return (("s", r.sparse.name, r.iv.dim.idx, r.iv.off),)
        assert len(idx) == len(self.dims), "%s expects %d indices" % (self.name, len(self.dims))
        comps = []
        for iv, dim in zip(idx, self.dims):
            assert isinstance(iv, IndexVar) and iv.dim is dim, \
                "index %d of %s must be an offset of dimension %s" % (len(comps), self.name, dim.name)
            comps.append(("d", dim.idx, iv.off))
        return tuple(comps)

    def __call__(self, *idx):
        index = self.index_of(idx)
        comps = [ad.var(ImageAccess(self.name, index, ch)) for ch in range(self.channels)]
        return comps[0] if self.channels == 1 else Vector(comps)

    def __repr__(self):
        return "%s<%s%d>" % (self.name, self.ctype, self.channels)


class ComputedArray(Image):
    """`exp:get(...)` (reference thallo.t:1777-1822,1868-1893): the expression is stored as a
    plan-owned single-channel image over its index domains, re-evaluated by the `precompute`
    pass whenever the unknowns change, together with a gradient image holding its partial
    derivatives with respect to the unknown accesses it contains (one channel per unknown;
    constant derivatives are not stored, thallo.t:1551-1561).  Residuals read both images
    like ordinary arrays (zero outside the domain); the chain rule runs through the gradient image."""

    def __init__(self, L, k, exp):
        dims = sorted(set(c[1] for v in ad.variables(exp) for c in _dense_components(v.key)))
        assert dims, "computed array without index domains"
        Image.__init__(self, "StoredExp_%d" % k, "real", 1, [L.dims[d] for d in dims], -(100 + 2 * k), "computed")
        if exp.type == ad.BOOL:
            exp = ad.select(exp, 1.0, 0.0)
        self.expression = exp
        ukeys = set(im.name for im in L.images if im.kind == "unknown")
        assert not ad.variables(exp, lambda v: isinstance(v.key, ImageAccess) and v.key.image.startswith("StoredExp_")), \
            "nested computed arrays are not supported"
        self.gunknowns = ad.variables(exp, lambda v: isinstance(v.key, ImageAccess) and v.key.image in ukeys)
        self.gradients = [ad.derivative(exp, u) for u in self.gunknowns]
        self.gchannel, n = [], 0                 # stored channel of every unknown's derivative (-1: constant, not stored)
        for g in self.gradients:
            if g.kind == "const":
                self.gchannel.append(-1)
            else:
                self.gchannel.append(n)
                n += 1
        self.gradient_image = None
        if n:
            self.gradient_image = Image(self.name + "_gradient", "real", n, self.dims, -(101 + 2 * k), "computed_gradient")

    def gradient_at(self, i, index):
        """d(stored value at `index`) / d(unknown i shifted to `index`)."""
        if self.gchannel[i] < 0:
            return self.gradients[i]
        return ad.var(ImageAccess(self.gradient_image.name, index, self.gchannel[i]))


def _dense_components(key):
    if isinstance(key, (ImageAccess, VecArg)):
        return [c for c in key.index if c[0] == "d"]
    if isinstance(key, Bounds):
        return [("d", d, 0) for (d, lo, hi) in key.ranges]
    if isinstance(key, IndexValue):
        return [("d", key.dim, 0)]
    return []


class Sparse:
    def __init__(self, name, frm, to, pidx):
        self.name, self.frm, self.to, self.pidx = name, list(frm), list(to), pidx
        self.coherent = False

    def set_coherent(self, b):
        self.coherent = bool(b)

    def __call__(self, iv):
        assert isinstance(iv, IndexVar) and iv.dim is self.frm[0]
        return SparseRef(self, iv)


class ParamDef:
    def __init__(self, name, ctype, pidx):
        self.name, self.ctype, self.pidx = name, ctype, pidx


class _SampledImage:
    def __init__(self, im, dx, dy):
        assert len(im.dims) == 2 and im.channels == 1, "sampled images must be 2-D single-channel"
        self.im, self.dx, self.dy = im, dx, dy

    def __call__(self, x, y):
        return ad.sample(self.im.name, self.dx.name if self.dx else None, self.dy.name if self.dy else None, x, y)


class _Sched:
    def __init__(self):
        self.materialize = False

    def set_materialize(self, b):
        self.materialize = bool(b)


class ResidualGroup:
    def __init__(self, name, terms):
        self.name, self.terms = name, terms
        self.J, self.JtJ, self.Jp = _Sched(), _Sched(), _Sched()     # MaterializeInfo of thallo.t:5740-5748
        self.at_output = None

    def compute_at_output(self, b):
        self.at_output = bool(b)
        return self


class Residuals:
    def __init__(self, groups):
        self.groups = groups
        for g in groups:
            setattr(self, g.name, g)

    def merge(self, *groups):
        """`r:merge(a, b)` asks the reference to evaluate two residual groups in one kernel (thallo.t:5173); it does
        not change the energy.  The schedules here fuse per unknown element already, so this is accepted and ignored."""
        return groups[0] if groups else None


class _NS:
    pass


class SymbolicL:
    """DSL namespace; `dims` are the concrete sizes given to Thallo_ProblemPlan
    (baked into the plan like the reference does, thallo.t:577-584)."""
    float, float2, float3, float4, float9 = ("real", 1), ("real", 2), ("real", 3), ("real", 4), ("real", 9)
    float6 = ("real", 6)
    uint8, int = ("uchar", 1), ("int", 1)

    def __init__(self, dim_sizes):
        self.dim_sizes = [int(d) for d in dim_sizes]
        self.dims, self.images, self.sparses, self.params = [], [], [], []
        self.usepreconditioner = False       # thallo.t:115
        self.residuals = None
        self.computed = {}                   # expression id -> ComputedArray (ComputedArrayCache, thallo.t:69,1879-1885)

    def computed_get(self, exp, idx):
        if exp.kind == "const":              # a constant needs no storage
            return exp
        ca = self.computed.get(exp.id)
        if ca is None:
            ca = ComputedArray(self, len(self.computed), exp)
            self.computed[exp.id] = ca
            self.images.append(ca)
            if ca.gradient_image is not None:
                self.images.append(ca.gradient_image)
        return ca(*idx)

    def Dims(self, *names):
        assert len(names) <= len(self.dim_sizes), "energy needs %d dimensions, got %d" % (len(names), len(self.dim_sizes))
        self.dims = [Dim(n, i, self.dim_sizes[i]) for i, n in enumerate(names)]
        return self.dims if len(names) > 1 else self.dims[0]

    def Dim(self, name, idx):            # thallo.Dim(name, idx): one dimension at a time (older energy files)
        assert idx == len(self.dims), "Dim() indices must be declared in order"
        assert idx < len(self.dim_sizes), "energy needs more dimensions than were given"
        self.dims.append(Dim(name, idx, self.dim_sizes[idx]))
        return self.dims[-1]

    def Unknown(self, t, dims, pidx): return ("Unknown", t, dims, pidx)
    def Array(self, t, dims, pidx): return ("Array", t, dims, pidx)
    def Sparse(self, frm, to, pidx): return ("Sparse", frm, to, pidx)
    def Param(self, t, pidx): return ("Param", t, pidx)

    def Inputs(self, **kw):
        ns = _NS()
        for name, decl in kw.items():
            if decl[0] in ("Unknown", "Array"):
                _, t, dims, pidx = decl
                im = Image(name, t[0], t[1], dims, pidx, decl[0].lower())
                if im.kind == "unknown":
                    assert t[0] == "real", "unknowns must have the solver's scalar type (thallo.t:1046-1052)"
                self.images.append(im)
                setattr(ns, name, im)
            elif decl[0] == "Sparse":
                s = Sparse(name, decl[1], decl[2], decl[3])
                self.sparses.append(s)
                setattr(ns, name, s)
            else:
                ctype = "float" if decl[1][0] == "real" else decl[1][0]
                self.params.append(ParamDef(name, ctype, decl[2]))
                setattr(ns, name, ad.var(Param(name)))
        return ns

    def UsePreconditioner(self, b):
        self.usepreconditioner = bool(b)

    # ---- expression constructors
    def Vector(self, *c): return Vector(c)

    def _u(self, op, x):
        if isinstance(x, Vector):
            return Vector([ad.unary(op, c) for c in x.c])
        return ad.unary(op, x)

    def sqrt(self, x): return self._u("sqrt", x)
    def sin(self, x): return self._u("sin", x)
    def cos(self, x): return self._u("cos", x)
    def tan(self, x): return self._u("tan", x)
    def exp(self, x): return self._u("exp", x)
    def log(self, x): return self._u("log", x)
    def abs(self, x): return self._u("abs", x)

    def _cmp(self, op, a, b):
        if isinstance(a, Vector):
            bb = b.c if isinstance(b, Vector) else [b] * len(a)
            return Vector([ad.cmp(op, x, y) for x, y in zip(a.c, bb)])
        return ad.cmp(op, a, b)

    def eq(self, a, b): return self._cmp("eq", a, b)
    def neq(self, a, b): return self._cmp("neq", a, b)
    def less(self, a, b): return self._cmp("less", a, b)
    def greater(self, a, b): return self._cmp("greater", a, b)
    def lesseq(self, a, b): return self._cmp("lesseq", a, b)
    def greatereq(self, a, b): return self._cmp("greatereq", a, b)
    def Not(self, b): return ad.not_(b)

    def And(self, *bs):
        r = bs[0]
        for b in bs[1:]:
            r = ad.and_(r, b)
        return r

    def Or(self, *bs):
        r = bs[0]
        for b in bs[1:]:
            r = ad.or_(r, b)
        return r

    def Select(self, c, a, b):
        if isinstance(a, Vector) or isinstance(b, Vector) or isinstance(c, Vector):
            n = max(len(v) for v in (a, b, c) if isinstance(v, Vector))
            aa = a.c if isinstance(a, Vector) else [a] * n
            bb = b.c if isinstance(b, Vector) else [b] * n
            cc = c.c if isinstance(c, Vector) else [c] * n      # per-channel conditions select per channel
            return Vector([ad.select(z, x, y) for z, x, y in zip(cc, aa, bb)])
        return ad.select(c, a, b)

    def InBounds(self, *idx):
        rng = tuple(sorted((iv.dim.idx, iv.off, iv.off) for iv in idx))
        return ad.var(Bounds(rng), ad.BOOL)

    def InBoundsExpanded(self, *args):
        *idx, e = args
        rng = tuple(sorted((iv.dim.idx, iv.off - e, iv.off + e) for iv in idx))
        return ad.var(Bounds(rng), ad.BOOL)

    def SampledImage(self, im, dx=None, dy=None):
        return _SampledImage(im, dx, dy)

    def Residuals(self, **kw):
        groups = []
        for name in sorted(kw):                      # thallo.t:5780
            v = kw[name]
            terms = []
            for item in (v if isinstance(v, (list, tuple)) else [v]):
                if isinstance(item, Vector):
                    terms.extend(item.c)
                else:
                    terms.append(ad.toexp(item))
            groups.append(ResidualGroup(name, terms))
        self.residuals = Residuals(groups)
        return self.residuals


def build_spec(define, dim_sizes, **kw):
    """Run an energy definition symbolically; returns the populated SymbolicL."""
    L = SymbolicL(dim_sizes)
    ad.Exp.get = lambda self, *idx: L.computed_get(self, idx)       # exp:get(...) binds to the problem being defined
    try:
        define(L, **kw)
    finally:
        del ad.Exp.get
    assert L.residuals is not None, "energy did not call Residuals{}"
    return L
