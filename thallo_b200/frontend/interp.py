"""NumPy interpreter for lowered expression DAGs (host-side check of the lowering logic:
shifted residual instances, bounds handling, partials) -- evaluates exactly the roots the
CUDA emitter prints, vectorised over the unknown domain.  Used by the CPU test-suite to
compare the generated unknownwise operators with the oracle's J-based products before any
GPU time is spent; it is not part of the solve path.
"""
import numpy as np

from . import ad
from .dsl import ImageAccess, Bounds, IndexValue, Param, VecArg


def _shifted(arr, offs):
    out = np.zeros_like(arr)
    nd = len(offs)
    src, dst = [], []
    for axis in range(nd):
        o = offs[nd - 1 - axis]
        n = arr.shape[axis]
        if abs(o) >= n:
            return out
        if o >= 0:
            src.append(slice(o, n)); dst.append(slice(0, n - o))
        else:
            src.append(slice(0, n + o)); dst.append(slice(-o, n))
    out[tuple(dst)] = arr[tuple(src)]
    return out


class Interp:
    shared = {}      # per gather_apply call: coefficient images of the index spaces

    def __init__(self, gen, params, dom, vec=None, dtype=np.float64):
        self.gen, self.params, self.dom, self.vec, self.dt = gen, params, list(dom), vec, dtype
        L = gen.L
        self.shape = tuple(L.dims[d].size for d in reversed(self.dom))
        grids = np.meshgrid(*[np.arange(n) for n in self.shape], indexing="ij")
        self.coord = {d: grids[len(self.dom) - 1 - i] for i, d in enumerate(self.dom)}
        self.cache = {}

    def _offs(self, index):
        o = [0] * len(self.dom)
        for c in index:
            o[self.dom.index(c[1])] = c[2]
        return o

    def _image(self, name):
        if name == "__coef":      # plan-owned image of hoisted invariants: evaluate its defining expressions in place
            if "__coef" not in self.cache:
                self.cache["__coef"] = np.stack([np.broadcast_to(v, self.shape).astype(self.dt)
                                                 for v in self.eval(self.gen.coef_exprs)], axis=-1)
            return self.cache["__coef"]
        im = self.gen.images[name]
        if im.kind in ("computed", "computed_gradient"):     # plan-owned ComputedArray images: evaluate `precompute` in place
            if name not in self.cache:
                ca = im if im.kind == "computed" else self.gen.images[name[:-len("_gradient")]]
                roots = [ca.expression] if im.kind == "computed" else [g for g, ch in zip(ca.gradients, ca.gchannel) if ch >= 0]
                self.cache[name] = np.stack([np.broadcast_to(v, self.shape).astype(self.dt) for v in self.eval(roots)], axis=-1)
            return self.cache[name]
        a = np.asarray(self.params[im.pidx]).astype(self.dt).reshape(self.shape + (im.channels,))
        return a

    def var(self, k):
        g = self.gen
        if isinstance(k, ImageAccess) and k.image.startswith("__coef_s"):
            si = int(k.image[len("__coef_s"):])
            ck = ("scoef", si)
            if ck not in Interp.shared:
                sub = Interp(g, self.params, g.spaces[si]["dims"], None, self.dt)
                Interp.shared[ck] = np.stack([np.broadcast_to(v, sub.shape).astype(self.dt).reshape(-1)
                                              for v in sub.eval(list(g.scoef[si]))], axis=-1)
            arr = Interp.shared[ck]
            if k.index[0][0] == "s":
                return arr[self._sparse(k.index[0][1]), k.channel]
            return arr[:, k.channel].reshape(self.shape)
        if isinstance(k, ImageAccess):
            if k.index[0][0] == "s":        # 1-D residual domain reaching a 1-D image through an index array
                im = g.images[k.image]
                if im.kind in ("computed", "computed_gradient"):      # plan-owned image over its own domain
                    sub = Interp(g, self.params, [d.idx for d in im.dims], None, self.dt)
                    a = sub._image(k.image).reshape(-1, im.channels)
                    return a[self._sparse(k.index[0][1]), k.channel]
                a = np.asarray(self.params[im.pidx]).astype(self.dt).reshape(-1, im.channels)
                return a[self._sparse(k.index[0][1]), k.channel]
            return _shifted(self._image(k.image), self._offs(k.index))[..., k.channel]
        if isinstance(k, VecArg):
            im = g.images[k.image]
            off = g.uoff[k.image]
            if k.index[0][0] == "s":
                a = self.vec[off:off + im.cardinality].astype(self.dt).reshape(-1, im.channels)
                return a[self._sparse(k.index[0][1]), k.channel]
            a = self.vec[off:off + im.cardinality].astype(self.dt).reshape(self.shape + (im.channels,))
            return _shifted(a, self._offs(k.index))[..., k.channel]
        if isinstance(k, Bounds):
            ok = np.ones(self.shape, bool)
            for (d, lo, hi) in k.ranges:
                n = g.L.dims[d].size
                c = self.coord[d]
                ok &= (c + lo >= 0) & (c + hi < n)
            return ok
        if isinstance(k, IndexValue):
            origin = g.slow_origin if (g.partition is not None and g.udomain and k.dim == g.udomain[-1]) else 0
            return (self.coord[k.dim] + k.off + origin).astype(self.dt)
        if isinstance(k, Param):
            pd = [p for p in g.L.params if p.name == k.name][0]
            return self.dt(np.asarray(self.params[pd.pidx]).reshape(-1)[0])
        if type(k).__name__ == "JVal":          # stored partial derivative i of every residual element
            return self.jvals[k.i]
        if type(k).__name__ == "JpVal":
            return self.jp[k.t]
        if type(k).__name__ == "JpAt":          # two-phase tile operator: J p of term t of the residual at offset s (0 outside the domain)
            nd = len(self.dom)
            return _shifted(self.jp_planes[k.t], [k.s0, k.s1, k.s2][:nd])
        raise NotImplementedError(k)

    def _sparse(self, name):
        return np.asarray(self.params[self.gen.sparses[name].pidx]).reshape(-1).astype(np.int64)

    def eval(self, roots):
        val = self.cache
        for n in ad.toposort(roots):
            if n.id in val:
                continue
            if n.kind == "const":
                val[n.id] = bool(n.value) if n.type == ad.BOOL else self.dt(n.value)
            elif n.kind == "var":
                val[n.id] = self.var(n.key)
            else:
                a = [val[x.id] for x in n.args]
                op = n.op
                with np.errstate(all="ignore"):
                    if op == "add": r = a[0] + a[1]
                    elif op == "sub": r = a[0] - a[1]
                    elif op == "mul": r = a[0] * a[1]
                    elif op == "powc": r = a[0] ** n.const if n.const > 0 else 1.0 / (a[0] ** (-n.const))
                    elif op == "pow": r = a[0] ** a[1]
                    elif op == "select": r = np.where(a[0], a[1], a[2])
                    elif op == "and": r = np.logical_and(a[0], a[1])
                    elif op == "or": r = np.logical_or(a[0], a[1])
                    elif op == "not": r = np.logical_not(a[0])
                    elif op in ("eq", "neq", "less", "greater", "lesseq", "greatereq"):
                        f = dict(eq=np.equal, neq=np.not_equal, less=np.less, greater=np.greater,
                                 lesseq=np.less_equal, greatereq=np.greater_equal)[op]
                        r = f(a[0], a[1])
                    elif op == "sample":
                        yo = self.gen.slow_origin if (self.gen.partition is not None and len(self.gen.udomain) == 2) else 0
                        r = self._sample(n.const[0], a[0], a[1] - yo)
                    else:
                        r = getattr(np, dict(abs="abs", asin="arcsin", acos="arccos", atan="arctan").get(op, op))(a[0])
                val[n.id] = r
        return [np.broadcast_to(val[r.id], self.shape) for r in roots]

    def _sample(self, name, x, y):
        data = self._image(name)[..., 0]
        H, W = data.shape

        def get(ix, iy):
            inb = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
            return np.where(inb, data[np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)], 0.0)
        x0, x1 = np.floor(x).astype(np.int64), np.ceil(x).astype(np.int64)
        y0, y1 = np.floor(y).astype(np.int64), np.ceil(y).astype(np.int64)
        xn, yn = x - x0, y - y0
        u = (1 - xn) * get(x0, y0) + xn * get(x1, y0)
        b = (1 - xn) * get(x0, y1) + xn * get(x1, y1)
        return (1 - yn) * u + yn * b


def unknownwise(gen, params, vec=None, dtype=np.float64):
    """Evaluate the generated at-output operators.  Returns (g, d, out) as flat vectors in the
    solver's unknown layout; `out` is None when no vector argument is given."""
    it = Interp(gen, params, gen.udomain, vec, dtype)
    U = gen.U
    roots = gen.uw_roots["g"] + gen.uw_roots["d"] + (gen.uw_roots["out"] if vec is not None else [])
    vals = it.eval(roots)

    def pack(lst):
        flat = np.zeros(gen.nunk, dtype)
        j = 0
        for im in gen.unknowns:
            a = np.stack([lst[j + ch] for ch in range(im.channels)], axis=-1)
            flat[gen.uoff[im.name]:gen.uoff[im.name] + im.cardinality] = a.reshape(-1)
            j += im.channels
        return flat
    g, d = pack(vals[:U]), pack(vals[U:2 * U])
    out = pack(vals[2 * U:]) if vec is not None else None
    return g, d, out


def unknownwise_two_phase(gen, params, vec, dtype=np.float64):
    """The two-phase form of the tiled operator (codegen.gen_two_phase): phase 1, J p of every term at every residual
    position; phase 2, every unknown applies its partial derivatives to the stored values.  Returns J^T J vec (flat)."""
    it = Interp(gen, params, gen.udomain, vec, dtype)
    tp = gen.two_phase_roots
    it.jp_planes = [np.array(v, dtype) for v in it.eval(tp["jp"])]
    vals = it.eval(tp["out"])
    flat = np.zeros(gen.nunk, dtype)
    j = 0
    for im in gen.unknowns:
        a = np.stack([vals[j + ch] for ch in range(im.channels)], axis=-1)
        flat[gen.uoff[im.name]:gen.uoff[im.name] + im.cardinality] = a.reshape(-1)
        j += im.channels
    return flat


def gather_apply(gen, params, vec, dtype=np.float64, materialised=False):
    """Evaluate the gather schedule's endpoint functions (codegen.gen_gather) over every residual
    element and sum them per unknown element the way the th_gather_s<i> kernels do.  Returns
    J^T J vec as a flat vector in the solver's unknown layout.  materialised: groups with a
    materialised Jacobian go through their stored-value functions (computeJv -> matJ -> jt_ep)."""
    out = np.zeros(gen.nunk, dtype)
    stored = {}
    Interp.shared = {}
    for ep in gen.endpoints:
        gi = ep["group"]
        g = gen.groups[gi]
        sp = gen.spaces[ep["space"]]
        it = Interp(gen, params, g["domain"], vec, dtype)
        js = sorted(ep["roots"])
        if materialised and g["materialize"]:
            if gi not in stored:
                jv = [np.array(v) for v in Interp(gen, params, g["domain"], vec, dtype).eval(g["jv_roots"])]
                it2 = Interp(gen, params, g["domain"], vec, dtype)
                it2.jvals = jv
                stored[gi] = (jv, [np.array(v) for v in it2.eval(g["matj_roots"])])
            it.jvals, it.jp = stored[gi]
            vals = it.eval([ep["mat_roots"][j] for j in js])
        elif g.get("storejp"):          # Jt[Jp]: J p per residual row (applyJ_g), then this endpoint's partials times it
            if gi not in stored:
                stored[gi] = [np.array(v) for v in Interp(gen, params, g["domain"], vec, dtype).eval(g["applyj_roots"])]
            it.jp = stored[gi]
            js = sorted(ep["jtp_roots"])
            vals = it.eval([ep["jtp_roots"][j] for j in js])
        else:
            vals = it.eval([ep["roots"][j] for j in js])
        if ep["kind"] == "sparse":
            tgt = it._sparse(ep["sparse"])
        else:
            # residual element e at coordinates c contributes to the unknown element at c + off
            nd = len(g["domain"])
            ok = np.ones(it.shape, bool)
            lin = np.zeros(it.shape, np.int64)
            stride = 1
            for i, d in enumerate(g["domain"]):
                c = it.coord[d] + ep["off"][i]
                n = gen.L.dims[d].size
                ok &= (c >= 0) & (c < n)
                lin += np.clip(c, 0, n - 1) * stride
                stride *= n
            tgt = lin.reshape(-1)
            vals = [np.where(ok, v, 0.0) for v in vals]
        slot_to = {}
        for im in sp["images"]:
            for ch in range(im.channels):
                slot_to[sp["slots"][(im.name, ch)]] = (gen.uoff[im.name], im.channels, ch)
        for j, v in zip(js, vals):
            base, C, ch = slot_to[j]
            np.add.at(out, base + tgt * C + ch, np.asarray(v, dtype).reshape(-1))
    return out


def gather_jtf(gen, params, dtype=np.float64):
    """Evaluate the gather schedule's PCGInit1 endpoint functions (jtf_ep<k>) summed per unknown element the way
    th_gatherjtf_s<i> does.  Returns (r, d) = (-J^T F, diag J^T J) as flat vectors in the solver's unknown layout."""
    r = np.zeros(gen.nunk, dtype)
    dg = np.zeros(gen.nunk, dtype)
    Interp.shared = {}
    for ep in gen.endpoints:
        g = gen.groups[ep["group"]]
        sp = gen.spaces[ep["space"]]
        it = Interp(gen, params, g["domain"], None, dtype)
        gacc, dacc = ep["jtf_roots"]
        js = sorted(gacc)
        vals = it.eval([gacc[j] for j in js] + [dacc[j] for j in js])
        if ep["kind"] == "sparse":
            tgt = it._sparse(ep["sparse"])
            ok = None
        else:
            ok = np.ones(it.shape, bool)
            lin = np.zeros(it.shape, np.int64)
            stride = 1
            for i, d in enumerate(g["domain"]):
                c = it.coord[d] + ep["off"][i]
                n = gen.L.dims[d].size
                ok &= (c >= 0) & (c < n)
                lin += np.clip(c, 0, n - 1) * stride
                stride *= n
            tgt = lin.reshape(-1)
            vals = [np.where(ok, v, 0.0) for v in vals]
        slot_to = {}
        for im in sp["images"]:
            for ch in range(im.channels):
                slot_to[sp["slots"][(im.name, ch)]] = (gen.uoff[im.name], im.channels, ch)
        for n_, j in enumerate(js):
            base, C, ch = slot_to[j]
            np.add.at(r, base + tgt * C + ch, np.asarray(vals[n_], dtype).reshape(-1))
            np.add.at(dg, base + tgt * C + ch, np.asarray(vals[len(js) + n_], dtype).reshape(-1))
    return r, dg


def jacobian_entries(gen, params, gi, dtype=np.float64):
    """What th_computej_g<gi> writes: (values, columns) of group gi in the export layout (element-major, per element
    the partials term-major / unknown-minor; column = flat unknown index or -1 outside the domain)."""
    g = gen.groups[gi]
    it = Interp(gen, params, g["domain"], None, dtype)
    exprs, keys = [], []
    for t in g["terms"]:
        for u, p in zip(t.unknowns, t.partials):
            exprs.append(p)
            keys.append(u.key)
    count = int(np.prod(it.shape)) if it.shape else 1
    if not exprs:
        return np.zeros(0, dtype), np.zeros(0, np.int64)
    vals = np.stack([np.asarray(v, dtype).reshape(-1) for v in it.eval(exprs)], axis=1)            # (count, nnz)
    cols = np.zeros((count, len(keys)), np.int64)
    for j, k in enumerate(keys):
        im = gen.images[k.image]
        if k.index[0][0] == "s":
            lin = it._sparse(k.index[0][1])
            ok = np.ones(count, bool)
        else:
            lin = np.zeros(it.shape, np.int64)
            ok = np.ones(it.shape, bool)
            stride = 1
            offs = it._offs(k.index)
            for i, d in enumerate(g["domain"]):
                c = it.coord[d] + offs[i]
                n = gen.L.dims[d].size
                ok &= (c >= 0) & (c < n)
                lin += c * stride
                stride *= n
            lin, ok = lin.reshape(-1), ok.reshape(-1)
        cols[:, j] = np.where(ok, gen.uoff[im.name] + lin * im.channels + k.channel, -1)
    return vals.reshape(-1), cols.reshape(-1)
