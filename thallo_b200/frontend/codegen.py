"""Lowering of a ProblemSpec to per-energy CUDA C++ device functions + a plan descriptor.

This is the replacement for the reference's CUDA-emitting backend
(API/src/thallo.t:2288-3455 `createfunction`, operator builders :3531-3949,
schedule selection :4096-4134).  Same math, different shape:

  * The reference differentiates a residual template and *shifts* it to every
    position from which it touches unknown x00 (`createjtjcentered` :3603-3667,
    `createjtfcentered` :3669-3712).  Here the unknownwise operators are emitted
    residual-centric: for every (term, shift s) whose instance at x0+s touches x0,
    compute Jp once and multiply by each partial.  Instances whose position x0+s
    lies outside the residual domain are dropped (see DESIGN.md "ghost residuals").
  * Residualwise functions follow createjtfResidualwise :3867-3908,
    createapplyjtjResidualwise :3536-3569, createmodelcostResidualwise :3845-3865,
    createcost :3939-3949, createcomputejResidualwise :3792-3805.
  * No instruction scheduling: the DAG is emitted in topological order as SSA
    temporaries and nvcc/NVRTC does the rest.

Generated functions are templates over an accessor `A` (how images / vector
arguments are read: global memory with bounds checks, or TMA-staged shared-memory
tiles) and a scatter sink `S` (how contributions are accumulated), both supplied
by the hand-written skeleton in csrc/skeleton/.
"""
import os
from collections import namedtuple

from . import ad
from .dsl import ImageAccess, Bounds, IndexValue, Param, VecArg

MAXD = 3

# variables of the materialised-Jacobian functions: stored partial derivative i of the current
# residual element, and (J p) of its term t
# (namedtuples compare as plain tuples and expression nodes are hash-consed on their key, so every
# key type carries a distinguishing tag)
_JVal = namedtuple("JVal", "tag i")
_JpVal = namedtuple("JpVal", "tag t")


def JVal(i):
    return _JVal("jval", i)


def JpVal(t):
    return _JpVal("jpval", t)


# two-phase tiled operator: (J p) of term t of the residual at offset (s0, s1, s2) from the unknown, read from the
# shared-memory J p planes the first phase filled
_JpAt = namedtuple("JpAt", "tag t s0 s1 s2")


def JpAt(t, s):
    return _JpAt("jpat", t, s[0], s[1], s[2])


class Lowered:
    """Result of lowering: CUDA source of the per-energy functions and the descriptor."""

    def __init__(self):
        self.source = ""
        self.desc = {}


def _domain_of(L, exprs):
    dims, sparse = set(), False
    for e in exprs:
        for v in ad.variables(e):
            k = v.key
            if isinstance(k, (ImageAccess, VecArg)):
                for c in k.index:
                    if c[0] == "d":
                        dims.add(c[1])
                    else:
                        dims.add(c[2]); sparse = True
            elif isinstance(k, Bounds):
                for (d, lo, hi) in k.ranges:
                    dims.add(d)
            elif isinstance(k, IndexValue):
                dims.add(k.dim)
    return tuple(sorted(dims)), sparse


def _shift_key(k, s):
    """Shift a variable key by s: dict dim_idx -> offset."""
    if isinstance(k, (ImageAccess, VecArg)):
        comps = []
        for c in k.index:
            if c[0] == "d":
                comps.append(("d", c[1], c[2] + s.get(c[1], 0)))
            else:
                assert s.get(c[2], 0) == 0, "cannot shift a sparse access"
                comps.append(c)
        return k._replace(index=tuple(comps))
    if isinstance(k, Bounds):
        return Bounds(tuple((d, lo + s.get(d, 0), hi + s.get(d, 0)) for (d, lo, hi) in k.ranges))
    if isinstance(k, IndexValue):
        return IndexValue(k.dim, k.off + s.get(k.dim, 0))
    return k


def shift(e, s, memo=None):
    if not any(s.values()):
        return e
    return ad.substitute(e, lambda v: ad.var(_shift_key(v.key, s), v.type), memo)


def _inb(s):
    rng = tuple(sorted((d, o, o) for d, o in s.items() if o != 0))
    if not rng:
        return ad.const(True)
    return ad.var(Bounds(rng), ad.BOOL)


class _Term:
    def __init__(self, L, exp):
        self.exp = exp
        ukeys = set(im.name for im in L.images if im.kind == "unknown")
        direct = ad.variables(exp, lambda v: isinstance(v.key, ImageAccess) and v.key.image in ukeys)
        partial = {}                     # unknown access key -> d exp / d unknown
        order = []
        for u in direct:
            partial[u.key] = ad.derivative(exp, u)
            order.append(u.key)
        # chain rule through computed arrays: d exp/d C(index) * gradient image of C at the same index,
        # for every unknown access C's expression contains, shifted to `index` (thallo.t:1551-1561,1742-1747)
        computed = dict((im.name, im) for im in L.images if im.kind == "computed")
        for c in ad.variables(exp, lambda v: isinstance(v.key, ImageAccess) and v.key.image in computed):
            im = computed[c.key.image]
            dc = ad.derivative(exp, c)
            through_sparse = any(comp[0] == "s" for comp in c.key.index)
            s = {} if through_sparse else dict((comp[1], comp[2]) for comp in c.key.index)
            for i, u in enumerate(im.gunknowns):
                g = im.gradient_at(i, c.key.index)
                if g.is_const(0.0):
                    continue
                if through_sparse:
                    # C:get(v(e)) (tests/minimal_sparse_materialize): the unknowns C reads at its own element
                    # are the unknowns at element v(e)
                    assert all(comp[0] == "d" and comp[2] == 0 for comp in u.key.index), \
                        "a computed array fetched through a sparse index may only read unknowns at its own element"
                    k = u.key._replace(index=c.key.index)
                else:
                    k = _shift_key(u.key, s)
                if k not in partial:
                    partial[k] = ad.const(0.0)
                    order.append(k)
                partial[k] = partial[k] + dc * g
        # keep only structurally non-zero partials
        self.unknowns = [ad.var(k) for k in order if not partial[k].is_const(0.0)]
        self.partials = [partial[u.key] for u in self.unknowns]

    def jp(self, arg):
        r = ad.const(0.0)
        for u, p in zip(self.unknowns, self.partials):
            k = u.key
            r = r + p * ad.var(VecArg(arg, k.image, k.index, k.channel))
        return r


class Emitter:
    """Emit a set of root expressions as straight-line CUDA statements."""

    def __init__(self, gen):
        self.gen = gen
        self.lines = []
        self.names = {}

    def ref(self, e):
        if e.kind == "const":
            if e.type == ad.BOOL:
                return "true" if e.value else "false"
            return self.gen.lit(e.value)
        return self.names[e.id]

    def emit(self, roots):
        for n in ad.toposort(roots):
            if n.kind == "const" or n.id in self.names:
                continue
            # numbered within the function, not by the DAG's process-wide node ids: the same energy lowers to
            # the same text in every process (the JIT's disk cache is keyed by the source)
            name = ("b%d" if n.type == ad.BOOL else "t%d") % (len(self.names) + 1)
            ty = "bool" if n.type == ad.BOOL else "real"
            self.lines.append("const %s %s = %s;" % (ty, name, self.rhs(n)))
            self.names[n.id] = name
        return [self.ref(r) for r in roots]

    def emit_guarded(self, groups):
        """groups: list of (condition text or None, roots, outs(refs) -> statements).  Nodes needed by more than one
        group (or by an unconditional group) are emitted at function scope, the others inside `if (condition) { }`
        together with the group's output statements."""
        users = {}
        for gi, (cond, roots, _) in enumerate(groups):
            for n in ad.toposort(roots):
                users.setdefault(n.id, set()).add(gi if cond is not None else -1)
        allroots = [r for _, roots, _ in groups for r in roots]
        shared = [n for n in ad.toposort(allroots) if len(users[n.id]) > 1 or -1 in users[n.id]]
        self.emit(shared)
        for gi, (cond, roots, outs) in enumerate(groups):
            if cond is None:
                self.lines.extend(outs([self.ref(r) for r in roots]))
                continue
            outer = dict(self.names)
            mark = len(self.lines)
            refs = self.emit(roots)
            inner = self.lines[mark:]
            del self.lines[mark:]
            self.lines.append("if (%s) {" % cond)
            self.lines.extend("    " + ln for ln in inner + outs(refs))
            self.lines.append("}")
            self.names = outer          # block-scoped names are not visible to later groups

    def rhs(self, n):
        g = self.gen
        if n.kind == "var":
            return g.var_rhs(n.key)
        a = [self.ref(x) for x in n.args]
        op = n.op
        if op == "add": return "%s + %s" % (a[0], a[1])
        if op == "sub": return "%s - %s" % (a[0], a[1])
        if op == "mul": return "%s * %s" % (a[0], a[1])
        if op == "powc":
            c = n.const
            if c > 0:
                return "th_powi<%d>(%s)" % (c, a[0])
            return "%s / th_powi<%d>(%s)" % (g.lit(1.0), -c, a[0])
        if op == "pow": return "th_pow(%s, %s)" % (a[0], a[1])
        if op == "select": return "(%s ? %s : %s)" % (a[0], a[1], a[2])
        if op == "and": return "(%s && %s)" % (a[0], a[1])
        if op == "or": return "(%s || %s)" % (a[0], a[1])
        if op == "not": return "!%s" % a[0]
        cmpops = dict(eq="==", neq="!=", less="<", greater=">", lesseq="<=", greatereq=">=")
        if op in cmpops: return "(%s %s %s)" % (a[0], cmpops[op], a[1])
        if op == "sample":
            im = g.image(n.const[0])
            y = a[1]
            if g.partition is not None and g.slow_origin and len(g.udomain) == 2:
                y = "(%s - %s)" % (y, g.lit(float(g.slow_origin)))      # absolute row -> row of the rank-local slab
            return "a.template samp<%d>(P, %s, %s)" % (g.ptr_slot[im.name], a[0], y)
        return "th_%s(%s)" % (op, a[0])


class Generator:
    def __init__(self, L, name, kind, double=False, schedule="auto", lm_as_committed=False, hoist=True, tile=None,
                 partition=None):
        self.L, self.name = L, name
        # multi-GPU slab partition: (ghost_lo, ghost_hi) ghost layers of the slowest axis included in `dims`
        # or, for graph domains in the gather schedule, {dim index: (lo, hi)}: elements [lo, size - hi) of that
        # dimension are owned by this rank, the others are ghost vertices / foreign edges (distributed.graph_partition)
        self.gpartition = None
        # Special keys: "replicated" = dimensions whose unknowns every rank holds in full (the cameras of bundle
        # adjustment: every rank sees observations of every camera, so their part of J^T J p is all-reduced);
        # "owner" = whether this rank counts the replicated unknowns in the dot products.
        self.replicated_dims, self.rep_owner = (), True
        if isinstance(partition, dict):
            partition = dict(partition)
            self.replicated_dims = tuple(int(x) for x in partition.pop("replicated", ()))
            self.rep_owner = bool(partition.pop("owner", True))
            self.gpartition = dict((int(k), (int(v[0]), int(v[1]))) for k, v in partition.items())
            partition = None
        # (ghost_lo, ghost_hi[, origin]): `origin` = global index of the local extent's first layer, added to the
        # VALUE of the slowest index (x:asvalue() in camera models such as shape_from_shading's) so that
        # expressions in absolute coordinates see the same numbers on every rank
        self.slow_origin = 0
        if partition is not None and len(partition) > 2:
            self.slow_origin = int(partition[2])
            partition = tuple(partition[:2])
        self.partition = tuple(int(x) for x in partition) if partition is not None else None
        self.hoist_enabled = bool(hoist)
        self.tile_request = tile
        self.double = bool(double)
        self.lm = (kind == "levenberg_marquardt") and not lm_as_committed
        self.kind = kind
        self.images = {im.name: im for im in L.images}
        self.sparses = {s.name: s for s in L.sparses}
        self.unknowns = sorted([im for im in L.images if im.kind == "unknown"], key=lambda i: i.pidx)
        assert self.unknowns, "energy declares no unknowns"
        self.computed = [im for im in L.images if im.kind == "computed"]      # ComputedArrays, in creation order
        # parameter slots
        self.ptr_slot, self.ptr_pidx = {}, []
        for obj in sorted(list(L.images) + list(L.sparses), key=lambda o: o.pidx):
            self.ptr_slot[obj.name] = len(self.ptr_pidx)
            self.ptr_pidx.append(obj.pidx)
        self.sc_slot, self.sc_defs = {}, []
        for p in sorted(L.params, key=lambda p: p.pidx):
            self.sc_slot[p.name] = len(self.sc_defs)
            self.sc_defs.append((p.pidx, p.ctype))
        # unknown vector layout: images back to back, AoS per element (thallo.t:1102-1126)
        self.uoff, off = {}, 0
        for k, im in enumerate(self.unknowns):
            self.uoff[im.name] = off
            off += im.cardinality
        self.nunk = off
        self.uidx = {im.name: k for k, im in enumerate(self.unknowns)}
        # groups
        self.groups = []
        for g in L.residuals.groups:
            terms = [_Term(L, t) for t in g.terms]
            dom, sparse = _domain_of(L, [t.exp for t in terms])
            # jtjp schedule of the group (get_schedule, thallo.t:4100-4134): J materialised -> [Jt][[J]p]
            # (PRECOMPUTE_J); only Jp materialised -> Jt[Jp] (APPLY_SEPARATELY); neither -> JtJp (INLINE)
            self.groups.append(dict(name=g.name, terms=terms, domain=dom, sparse=sparse,
                                    materialize=g.J.materialize,
                                    storejp=bool(getattr(g, "Jp", None) and g.Jp.materialize) and not g.J.materialize))
        udoms = set(tuple(d.idx for d in im.dims) for im in self.unknowns)
        can_at_output = (len(udoms) == 1 and all((not g["sparse"]) and g["domain"] == next(iter(udoms))
                                                 for g in self.groups)
                         and not any(g["materialize"] or g["storejp"] for g in self.groups))
        self._find_endpoints()
        if schedule == "auto":
            schedule = "at_output" if can_at_output else ("gather" if self.can_gather else "residualwise")
        if schedule == "at_output":
            assert can_at_output, "compute_at_output needs every residual domain to equal the unknown domain"
        if schedule == "residualwise":
            assert not any(g["materialize"] or g["storejp"] for g in self.groups), "materialised J / Jp need the gather schedule"
        if schedule == "gather":
            assert self.can_gather, "the gather schedule needs every unknown access to be a sparse index or a dense offset of the residual's own domain"
        self.schedule = schedule
        self.udomain = next(iter(udoms)) if len(udoms) == 1 else None
        # tiled form of the unknownwise operator (shared-memory stencil tiles, TMA-staged): 2-D / 3-D image domains
        self.tiled = self.schedule == "at_output" and len(self.udomain) in (2, 3)
        self.stage_center = self.tiled and len(self.udomain) == 2    # also stage centre-only arrays (2-D: shared memory to spare)
        self.tile_pad = os.environ.get("THALLO_B200_TILE_PAD", "1") != "0"      # tuning switch
        gl = os.environ.get("THALLO_B200_GATHER_LANES")                          # tuning switch: lanes per unknown element (1..16)
        self.gather_lanes = int(gl) if gl else None
        self.gather_unroll = int(os.environ.get("THALLO_B200_GATHER_UNROLL", "1"))   # tuning switch: unroll of the adjacency walk
        # two-phase tiled operator (see gen_two_phase): "auto" = where the shifted-instance form re-forms J p of many
        # residuals per unknown (shape_from_shading: 26 instances of 6 terms, the 3-D volume: 39 of 21)
        self.two_phase_request = os.environ.get("THALLO_B200_TWO_PHASE", "auto")
        self.two_phase = False
        self.coef_exprs = []          # hoisted PCG-invariant per-element expressions (channels of the __coef image)
        self._coef_index = {}

    # ---- small helpers
    def image(self, name):
        return self.images[name]

    def lit(self, v):
        s = repr(float(v))
        if s in ("inf", "-inf", "nan"):
            return {"inf": "TH_INF", "-inf": "(-TH_INF)", "nan": "TH_NAN"}[s]
        return s if self.double else s + "f"

    def _offs(self, index, domain):
        o = [0] * MAXD
        for c in index:
            pos = domain.index(c[1])
            o[pos] = c[2]
        return o

    def var_rhs(self, k):
        dom = self._dom
        if isinstance(k, ImageAccess):
            im = self.images[k.image]
            slot = self.ptr_slot[im.name]
            if k.index[0][0] == "s":
                sp = self.ptr_slot[k.index[0][1]]
                return "a.template simg<%d, th_%s, %d, %d, %d>(P)" % (slot, im.ctype, im.channels, k.channel, sp)
            assert tuple(c[1] for c in k.index) == tuple(dom), \
                "image %s is indexed by dims %s inside a function over dims %s" % (im.name, [c[1] for c in k.index], dom)
            o = self._offs(k.index, dom)
            return "a.template img<%d, th_%s, %d, %d, %d, %d, %d>(P)" % (slot, im.ctype, im.channels, k.channel, o[0], o[1], o[2])
        if isinstance(k, VecArg):
            im = self.images[k.image]
            if k.index[0][0] == "s":
                sp = self.ptr_slot[k.index[0][1]]
                return "a.template svec<%d, %d, %d>(P)" % (self.uidx[im.name], k.channel, sp)
            o = self._offs(k.index, dom)
            return "a.template vec<%d, %d, %d, %d, %d>()" % (self.uidx[im.name], k.channel, o[0], o[1], o[2])
        if isinstance(k, Bounds):
            lo, hi = [0] * MAXD, [0] * MAXD
            for (d, l, h) in k.ranges:
                pos = dom.index(d)
                lo[pos], hi[pos] = l, h
            return "a.template inb<%d, %d, %d, %d, %d, %d>()" % (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])
        if isinstance(k, IndexValue):
            origin = self.slow_origin if (self.partition is not None and self.udomain and k.dim == self.udomain[-1]) else 0
            return "(real)(a.template coord<%d>() + (%d))" % (dom.index(k.dim), k.off + origin)
        if isinstance(k, Param):
            return "P.sc[%d]" % self.sc_slot[k.name]
        if isinstance(k, _JVal):
            return "jq%d.%s" % (k.i // 4, "xyzw"[k.i % 4])
        if isinstance(k, _JpVal):
            return "jpv[%d]" % k.t
        if isinstance(k, _JpAt):
            return "J.template at<%d, %d, %d, %d>()" % (k.t, k.s0, k.s1, k.s2)
        raise NotImplementedError(k)

    def _fn(self, sig, roots, outs, dom, pre_lines=()):
        """One device function: emit `roots`, then `outs(refs)` statements."""
        self._dom = dom
        em = Emitter(self)
        refs = em.emit(roots)
        body = list(pre_lines) + em.lines + outs(refs)
        return sig + " {\n    " + "\n    ".join(body) + "\n}\n"

    def _scatter_stmt(self, which, key, ref, dom):
        im = self.images[key.image]
        k = self.uidx[im.name]
        if key.index[0][0] == "s":
            return "s.template sadd<%d, %d, %d, %d>(P, %s);" % (which, k, key.channel, self.ptr_slot[key.index[0][1]], ref)
        o = self._offs(key.index, dom)
        return "s.template add<%d, %d, %d, %d, %d, %d>(%s);" % (which, k, key.channel, o[0], o[1], o[2], ref)

    # ---- unknownwise (at-output) functions
    def _instances(self):
        """(term, shift, [(slot j, partial shifted)], Jp shifted builder) for every residual
        instance that touches the unknown at x0."""
        out = []
        terms = [t for g in self.groups for t in g["terms"]]
        slot_of, j = {}, 0
        for im in self.unknowns:
            for ch in range(im.channels):
                slot_of[(im.name, ch)] = j
                j += 1
        self.U = j
        for t in terms:
            by_shift = {}
            for u, p in zip(t.unknowns, t.partials):
                k = u.key
                s = tuple(sorted((c[1], -c[2]) for c in k.index))
                by_shift.setdefault(s, []).append((slot_of[(k.image, k.channel)], p))
            for s, lst in by_shift.items():
                out.append((t, dict(s), lst))
        return out

    # ---- hoisting of PCG-invariant per-element subexpressions (at-output schedule)
    # J does not change during the PCG iterations of one nonlinear step, but the reference's
    # applyJTJ re-evaluates it (including sin/cos of the angles at the centre and every stencil
    # neighbour) on every CG iteration.  Sub-expressions rooted at an
    # expensive operator whose inputs are all read at the residual's own element (zero offsets)
    # are evaluated once per nonlinear iteration into a plan-owned multi-channel image
    # ("__coef"), and the operator reads that image (through the staged tile) instead.  The
    # values are bit-identical: the same device function is applied to the same inputs.
    _EXPENSIVE = ("sin", "cos", "tan", "exp", "log", "sqrt", "asin", "acos", "atan", "sinh", "cosh", "tanh", "pow", "sample")

    def _pure0(self, e, memo):
        """(pure, reads_image): every variable under e is a zero-offset dense access on the unknown
        domain, a scalar Param or the element's own coordinate."""
        r = memo.get(e.id)
        if r is not None:
            return r
        if e.kind == "const":
            r = (True, False)
        elif e.kind == "var":
            k = e.key
            if isinstance(k, ImageAccess):
                ok = (all(c[0] == "d" and c[2] == 0 for c in k.index)
                      and tuple(c[1] for c in k.index) == tuple(self.udomain))
                r = (ok, ok)
            elif isinstance(k, Param):
                r = (True, False)
            elif isinstance(k, IndexValue):
                r = (k.off == 0, False)
            else:
                r = (False, False)
        else:
            ok, img = True, False
            for a in e.args:
                o, i = self._pure0(a, memo)
                ok, img = ok and o, img or i
            if e.op == "sample":
                img = True
            r = (ok, img)
        memo[e.id] = r
        return r

    def _hoist(self, e, memo, pmemo):
        if e.id in memo:
            return memo[e.id]
        if e.kind != "apply":
            r = e
        else:
            pure, img = self._pure0(e, pmemo)
            if pure and img and e.op in self._EXPENSIVE and e.type == ad.REAL:
                ch = self._coef_index.get(e.id)
                if ch is None:
                    ch = len(self.coef_exprs)
                    self._coef_index[e.id] = ch
                    self.coef_exprs.append(e)
                index = tuple(("d", d, 0) for d in self.udomain)
                r = ad.var(ImageAccess("__coef", index, ch))
            else:
                r = ad.rebuild(e, [self._hoist(a, memo, pmemo) for a in e.args])
        memo[e.id] = r
        return r

    def _prepare_unknownwise(self):
        """Hoist invariants out of the partial derivatives used by applyJTJ and register the
        plan-owned coefficient image; must run before the header is emitted."""
        self.inst = self._instances()
        self.inst_h = self.inst
        if self.hoist_enabled:
            memo, pmemo = {}, {}
            self.inst_h = [(t, s, [(j, self._hoist(p, memo, pmemo)) for j, p in lst]) for t, s, lst in self.inst]
            # Jp of a whole term also needs the hoisted partials of *all* its unknowns
            self.jp_h = {}
            for t, s, lst in self.inst:
                if id(t) in self.jp_h:
                    continue
                r = ad.const(0.0)
                for u, p in zip(t.unknowns, t.partials):
                    k = u.key
                    r = r + self._hoist(p, memo, pmemo) * ad.var(VecArg("P", k.image, k.index, k.channel))
                self.jp_h[id(t)] = r
        if self.coef_exprs:
            from .dsl import Image
            im = Image("__coef", "real", len(self.coef_exprs), [self.L.dims[d] for d in self.udomain], -1, "plan")
            self.images["__coef"] = im
            self.ptr_slot["__coef"] = len(self.ptr_pidx)
            self.ptr_pidx.append(-1)

    def _tile_layout(self, roots):
        """Shared-memory layout of the staged tiles of the tiled operator kernel."""
        nd = len(self.udomain)
        halo = self._halos(roots)
        H = [0] * MAXD
        for tbl in (halo["img"], halo["vec"]):
            for v in tbl.values():
                H = [max(a, b) for a, b in zip(H, v)]
        tile = list(self.tile_request) if self.tile_request else ([32, 8, 1] if nd == 2 else [8, 8, 4])
        if os.environ.get("THALLO_B200_TILE"):                                   # tuning switch, e.g. "16,8,4"
            tile = [int(x) for x in os.environ["THALLO_B200_TILE"].split(",")]
        tile = tile + [1] * (MAXD - len(tile))
        es_real = 8 if self.double else 4
        esz = dict(real=es_real, float=4, uchar=1, int=4, double=8)
        ext = [tile[d] + 2 * H[d] for d in range(MAXD)]
        off = 0
        stages = []

        def conflict_degree(channels, es, row_bytes):
            """Worst shared-memory bank-conflict degree of one warp reading channel 0 of its elements
            (thread x fastest, 4-byte banks, same-word reads broadcast)."""
            worst = 1
            nthreads = tile[0] * tile[1] * tile[2]
            for w0 in range(0, nthreads, 32):
                banks = {}
                for t in range(w0, min(w0 + 32, nthreads)):
                    tx, ty, tz = t % tile[0], (t // tile[0]) % tile[1], t // (tile[0] * tile[1])
                    word = (tx * channels * es + ty * row_bytes + tz * row_bytes * ext[1]) // 4
                    banks.setdefault(word % 32, set()).add(word)
                worst = max(worst, max(len(v) for v in banks.values()))
            return worst

        def place(channels, es):
            # TMA requires the innermost start coordinate of a box to be 16-byte aligned (measured on
            # B200: an unaligned start raises "illegal instruction"), so the left halo is padded to
            # 16 bytes: a row is [padl | tile | right halo] scalars, rounded up to 16 bytes.  Rows are
            # then widened by up to 7 more 16-byte units when that lowers the bank-conflict degree of a
            # warp that spans several rows (8-wide 3-D tiles with 128-byte rows: 4-way -> none).
            nonlocal off
            padl = -(-(H[0] * channels * es) // 16) * 16 // es
            row_bytes = -(-((padl + (tile[0] + H[0]) * channels) * es) // 16) * 16
            if self.tile_pad:
                cands = [row_bytes + 16 * k for k in range(8) if (row_bytes + 16 * k) // es <= 256]
                row_bytes = min(cands, key=lambda rb: (conflict_degree(channels, es, rb), rb))
            roww = row_bytes // es
            nbytes = row_bytes * ext[1] * ext[2]
            o = off
            off = -(-(off + nbytes) // 128) * 128
            return o, roww, nbytes, padl
        def place_center(channels, es):
            # tile-only box (no halo) for arrays that are read at the element itself
            nonlocal off
            row_bytes = tile[0] * channels * es
            if row_bytes % 16:
                return None
            nbytes = row_bytes * tile[1] * tile[2]
            o = off
            off = -(-(off + nbytes) // 128) * 128
            return o, row_bytes // es, nbytes
        vt = []
        for im in self.unknowns:                     # z and p_old tiles (with halo) per unknown image
            zo, roww, nb, padl = place(im.channels, es_real)
            po, _, _, _ = place(im.channels, es_real)
            v = dict(name=im.name, channels=im.channels, roww=roww, zoff=zo, poff=po, bytes=nb, padl=padl,
                     coff=-1, croww=0, cbytes=0)
            if self.lm and self.stage_center:        # CtC tile (LM diagonal), read at the element only
                c = place_center(im.channels, es_real)
                if c is not None:
                    v["coff"], v["croww"], v["cbytes"] = c
            vt.append(v)
        if any(v["coff"] < 0 for v in vt):           # all or nothing, keeps the kernel simple
            for v in vt:
                v["coff"], v["croww"], v["cbytes"] = -1, 0, 0
        slot_stage = [-1] * len(self.ptr_pidx)
        for name, h in halo["img"].items():
            im = self.images[name]
            es = esz[im.ctype]
            if any(h):
                o, roww, nb, padl = place(im.channels, es)
                center = 0
            else:
                c = place_center(im.channels, es) if self.stage_center else None
                if c is None:
                    continue                         # read straight from global memory
                (o, roww, nb), padl, center = c, 0, 1
            slot_stage[self.ptr_slot[name]] = len(stages)
            stages.append(dict(name=name, slot=self.ptr_slot[name], ctype=im.ctype, es=es, channels=im.channels,
                               roww=roww, off=o, bytes=nb, padl=padl, center=center))
        # pipeline depth and residency of th_pcg_a (227 KB of shared memory per SM, 64 K registers):
        # two stages beside two other resident CTAs when they fit; otherwise (3-D tiles with wide
        # halos) still two stages -- a one-stage CTA sits idle for the whole load of its next tile
        # (profiles/r01i: a third of the warp samples of the 3-D operator) -- with as many CTAs as fit.
        nthreads = tile[0] * tile[1] * tile[2]
        stage = max(128, off)
        minb = 3
        if 2 * stage * 3 + 3 * 1024 <= 227 * 1024:
            pipe = 2
        else:
            pipe, minb = 2, (227 * 1024 - 2048) // (2 * stage)
            if minb < 1:
                pipe, minb = 1, max(1, min(3, (227 * 1024 - 2048) // stage))
        minb = max(1, min(minb, 2048 // nthreads))
        if os.environ.get("THALLO_B200_PIPE"):
            pipe = int(os.environ["THALLO_B200_PIPE"])
        if os.environ.get("THALLO_B200_MINB"):
            minb = int(os.environ["THALLO_B200_MINB"])
        self.tl = dict(tile=tile, halo=H, ext=ext, vt=vt, stages=stages, slot_stage=slot_stage, smem=off, pipe=pipe, minb=minb)
        return self.tl

    def gen_unknownwise(self):
        dom = self.udomain
        src = []
        inst = self.inst
        U = self.U
        zero = ad.const(0.0)
        # evalJTF: g_j = sum partial*F, d_j = sum partial^2  (createjtfcentered)
        g = [zero] * U
        d = [zero] * U
        for t, s, lst in inst:
            memo = {}
            cond = _inb(s)
            F = shift(t.exp, s, memo)
            for j, p in lst:
                ps = shift(p, s, memo)
                g[j] = g[j] + ad.select(cond, ps * F, 0.0)
                d[j] = d[j] + ad.select(cond, ps * ps, 0.0)
        src.append(self._fn(
            "template <class A> __device__ __forceinline__ void evalJTF_uw(const A& a, const Params& P, real* __restrict__ g, real* __restrict__ d)",
            g + d,
            lambda r: ["g[%d] = %s;" % (j, r[j]) for j in range(U)] + ["d[%d] = %s;" % (j, r[U + j]) for j in range(U)],
            dom))
        # applyJTJ: out_j = sum partial * Jp   (createjtjcentered, residual-centric)
        out = [zero] * U
        for t, s, lst in self.inst_h:
            memo = {}
            cond = _inb(s)
            jp = shift(self.jp_h[id(t)] if self.hoist_enabled else t.jp("P"), s, memo)
            for j, p in lst:
                ps = shift(p, s, memo)
                out[j] = out[j] + ad.select(cond, ps * jp, 0.0)
        src.append(self._fn(
            "template <class A> __device__ __forceinline__ void applyJTJ_uw(const A& a, const Params& P, real* __restrict__ out)",
            out, lambda r: ["out[%d] = %s;" % (j, r[j]) for j in range(U)], dom))
        self.uw_roots = dict(g=g, d=d, out=out)        # kept for the NumPy interpreter (frontend/interp.py)
        # halo radius per image / vector argument for the tile-staged variant
        self.halo = self._halos(out)
        if self.coef_exprs:
            nc = len(self.coef_exprs)
            src.append(self._fn(
                "template <class A> __device__ __forceinline__ void coef_uw(const A& a, const Params& P, real* __restrict__ out)",
                list(self.coef_exprs), lambda r: ["out[%d] = %s;" % (i, r[i]) for i in range(nc)], dom))
        if self.tiled:
            excl = [im.exclude for im in self.unknowns if im.exclude is not None]
            self._tile_layout(out + excl)
            nterms = len(set(id(t) for t, _, _ in self.inst))
            want = self.two_phase_request
            self.two_phase = (want == "1") or (want == "auto" and len(self.inst) >= 2 * nterms + 8)
            if self.two_phase:
                tp_src = self.gen_two_phase()
                # shared memory: pipeline stages + the J p planes must fit beside at least one resident CTA
                limit = 227 * 1024 - 2048
                stage = max(128, self.tl["smem"])
                jpb = self.tp["nt"] * self.tp["nbox"] * (8 if self.double else 4)
                nthreads = self.tl["tile"][0] * self.tl["tile"][1] * self.tl["tile"][2]
                pipe = self.tl["pipe"]
                if pipe * stage + jpb > limit:
                    pipe = 1
                if pipe * stage + jpb > limit:
                    self.two_phase = False            # does not fit: keep the shifted-instance form
                else:
                    if not os.environ.get("THALLO_B200_PIPE"):
                        self.tl["pipe"] = pipe
                    if not os.environ.get("THALLO_B200_MINB"):
                        self.tl["minb"] = max(1, min(self.tl["minb"], limit // (self.tl["pipe"] * stage + jpb), 2048 // nthreads))
                    # Two sets of planes (alternating per tile) would save the barrier at the start of a tile -- the previous
                    # tile's phase 2 may still be reading.  Built, correct, and measured on shape_from_shading 8192^2: 1.430 ms
                    # against 1.400 ms with one set (profiles/r02l_summary.txt), so one set stays the default
                    # (THALLO_B200_JP_BUFS=2 selects two where they fit).
                    bufs = 1
                    if os.environ.get("THALLO_B200_JP_BUFS") == "2" and (self.tl["pipe"] * stage + 2 * jpb) * self.tl["minb"] <= limit:
                        bufs = 2
                    self.tp["bufs"] = bufs
                    src.append(tp_src)
        return "\n".join(src)

    # ---- two-phase form of the tiled operator
    # The shifted-instance form above lets every unknown re-form J p of every residual instance that touches it: a term
    # with k unknowns in its support is evaluated (with all k of its partial derivatives) from k different unknowns.
    # Where that dominates (many instances per term), the tile kernel instead runs the reference's Jt[Jp] idea inside
    # shared memory: phase 1, every residual position of the tile and of the positions around it that reach into the
    # tile forms J p of its terms ONCE into shared-memory planes; phase 2, every unknown multiplies its own partial
    # derivative of each instance with the stored value.  Halo positions evaluate only the term classes that reach
    # into the tile from there (the volume: one direction's three terms per face).
    def gen_two_phase(self):
        dom = self.udomain
        nd = len(dom)
        U = self.U
        zero = ad.const(0.0)
        terms, tindex = [], {}
        for t, s, lst in self.inst:
            if id(t) not in tindex:
                tindex[id(t)] = len(terms)
                terms.append(t)
        NT = len(terms)

        def svec(s):
            return tuple(s.get(d, 0) for d in dom) + (0,) * (MAXD - nd)
        shifts_of = {}
        for t, s, lst in self.inst:
            shifts_of.setdefault(tindex[id(t)], set()).add(svec(s))
        # term classes: terms with the same set of non-zero shifts are needed at the same halo positions
        cls_of, classes = {}, []
        for k in range(NT):
            key = frozenset(x for x in shifts_of[k] if any(x))
            if key not in classes:
                classes.append(key)
            cls_of[k] = classes.index(key)
        assert len(classes) <= 8, "two-phase tile: more than 8 term classes"
        tile = self.tl["tile"]
        PH = [max(abs(x[d]) for k in shifts_of for x in shifts_of[k]) for d in range(MAXD)]
        assert all(PH[d] <= self.tl["halo"][d] for d in range(MAXD)) and max(PH) <= 7
        box = [tile[d] + 2 * PH[d] for d in range(MAXD)]
        # halo positions (relative to the tile origin) and the classes needed there
        pos = []
        for qz in range(-PH[2], tile[2] + PH[2]):
            for qy in range(-PH[1], tile[1] + PH[1]):
                for qx in range(-PH[0], tile[0] + PH[0]):
                    q = (qx, qy, qz)
                    if all(0 <= q[d] < tile[d] for d in range(MAXD)):
                        continue
                    need = 0
                    for ci, key in enumerate(classes):
                        if any(all(0 <= q[d] - x[d] < tile[d] for d in range(MAXD)) for x in key):
                            need |= 1 << ci
                    if need:
                        pos.append((need, q))
        pos.sort(key=lambda e: e[0])          # equal class masks next to each other: warps diverge less
        assert all(0 <= q[d] + 8 < 256 for _, q in pos for d in range(MAXD))
        packed = [(q[0] + 8) | ((q[1] + 8) << 8) | ((q[2] + 8) << 16) | (need << 24) for need, q in pos]
        jp_exprs = [self.jp_h[id(t)] if self.hoist_enabled else t.jp("P") for t in terms]
        # phase 2 roots
        out = [zero] * U
        for t, s, lst in self.inst_h:
            memo = {}
            cond = _inb(s)
            jpv = ad.var(JpAt(tindex[id(t)], svec(s)))
            for j, p in lst:
                out[j] = out[j] + ad.select(cond, shift(p, s, memo) * jpv, 0.0)
        self.two_phase_roots = dict(jp=jp_exprs, out=out, tindex=tindex, terms=terms)
        self.tp = dict(nt=NT, ph=PH, box=box, nbox=box[0] * box[1] * box[2], pos=packed, nclasses=len(classes))
        src = []
        # halo positions: guarded by class
        self._dom = dom
        em = Emitter(self)
        groups = []
        for ci in range(len(classes)):
            ks = [k for k in range(NT) if cls_of[k] == ci]
            if not classes[ci]:
                continue                  # terms that touch only their own element are never needed outside the tile
            groups.append(("need & %du" % (1 << ci), [jp_exprs[k] for k in ks],
                           lambda r, ks=ks: ["jp[%d] = %s;" % (k, x) for k, x in zip(ks, r)]))
        em.emit_guarded(groups)
        src.append("template <class A> __device__ __forceinline__ void applyJ_halo(const A& a, const Params& P, unsigned need, real* __restrict__ jp) {\n    "
                   + "\n    ".join(em.lines) + "\n}\n")
        # the element itself: J p of its own residuals -> planes, barrier, transposed products (one scope: the partial
        # derivatives of the element's own residuals are shared between the two phases)
        em = Emitter(self)
        refs = em.emit(jp_exprs)
        lines = list(em.lines) + ["J.template put<%d>(inside ? %s : (real)0);" % (k, x) for k, x in enumerate(refs)] + ["J.sync();"]
        mark = len(em.lines)
        refs2 = em.emit(out)
        lines += em.lines[mark:] + ["out[%d] = %s;" % (j, refs2[j]) for j in range(U)]
        src.append("template <class A, class JT> __device__ __forceinline__ void applyJTJ_tile(const A& a, const Params& P, JT& J, bool inside, real* __restrict__ out) {\n    "
                   + "\n    ".join(lines) + "\n}\n")
        return "\n".join(src)

    def _halos(self, roots):
        """Per accessed image (by ptr slot) and per unknown-vector image: max |offset| per dim."""
        img, vec = {}, {}
        for e in roots:
            for v in ad.variables(e):
                k = v.key
                if isinstance(k, (ImageAccess, VecArg)) and k.index[0][0] == "d":
                    tgt = vec if isinstance(k, VecArg) else img
                    cur = tgt.setdefault(k.image, [0] * MAXD)
                    for pos, c in enumerate(k.index):
                        cur[pos] = max(cur[pos], abs(c[2]))
        return dict(img=img, vec=vec)

    # ---- gather schedule (graph domains, materialised Jacobians): index spaces and endpoints
    # The reference applies J^T J residual by residual and scatters with float atomics into a
    # cleared Ap (createapplyjtjResidualwise thallo.t:3536-3569, PCGStep1 gauss_newton.t:1006-1016,
    # then PCGStep1_Finish :774-799).  Here the operator is applied unknown by unknown: every
    # distinct way a residual group reaches an unknown element (an "endpoint": through a sparse
    # index array, or at a dense offset of the group's own domain) gets a function that returns
    # that residual element's contribution to the unknowns at that endpoint, and one kernel per
    # index space walks the residual elements incident to each unknown element (adjacency lists
    # built by the plan from the index arrays) and sums them in registers: no atomics, no clear
    # of Ap, no finishing pass, deterministic.  Groups with a materialised Jacobian read their
    # stored partial derivatives instead of re-evaluating them (CSR SpMV / SpMV^T role,
    # gauss_newton.t:1448-1525, without storing column indices: they are implied by the index arrays).
    def _find_endpoints(self):
        spaces, space_of = [], {}
        for im in self.unknowns:
            dom = tuple(d.idx for d in im.dims)
            if dom not in space_of:
                space_of[dom] = len(spaces)
                spaces.append(dict(dims=dom, images=[], slots={}, nslots=0, endpoints=[]))
            sp = spaces[space_of[dom]]
            sp["images"].append(im)
            for ch in range(im.channels):
                sp["slots"][(im.name, ch)] = sp["nslots"]
                sp["nslots"] += 1
        self.spaces, self.space_of = spaces, space_of
        self.endpoints = []
        self.can_gather = True
        for gi, g in enumerate(self.groups):
            dom = tuple(g["domain"])
            seen = {}
            for ti, t in enumerate(g["terms"]):
                for u, p in zip(t.unknowns, t.partials):
                    index = u.key.index
                    if index not in seen:
                        first = index[0]
                        if first[0] == "s":
                            ok = len(dom) == 1 and first[2] == dom[0] and first[3] == 0
                            tdom = tuple(d.idx for d in self.images[u.key.image].dims)
                            ep = dict(kind="sparse", sparse=first[1], off=[0] * MAXD)
                        else:
                            tdom = tuple(c[1] for c in index)
                            ok = tdom == dom and len(dom) <= MAXD
                            ep = dict(kind="dense", sparse=None, off=self._offs(index, dom) if ok else [0] * MAXD)
                        if not ok or tdom not in space_of:
                            self.can_gather = False
                            return
                        ep.update(group=gi, index=index, space=space_of[tdom], parts=[], id=len(self.endpoints))
                        seen[index] = ep
                        self.endpoints.append(ep)
                        spaces[ep["space"]]["endpoints"].append(ep)
                    seen[index]["parts"].append((ti, u, p))
        sid = 0
        for ep in self.endpoints:               # sparse endpoints own an adjacency list (ThGather.ptr / .perm)
            ep["sid"] = -1
            if ep["kind"] == "sparse":
                ep["sid"] = sid
                sid += 1
        self.n_sparse_ep = sid

    # ---- hoisting for the gather schedule: sub-expressions rooted at a transcendental whose image
    # reads all go through ONE index (the same sparse index array, or the residual's own element)
    # depend on a single element of an unknown index space and not on the PCG iteration: they are
    # evaluated once per nonlinear iteration into a plan-owned coefficient image over that space
    # ("__coef_s<i>") and fetched through the same index (arap_mesh: sin/cos of the three angles of
    # v0, evaluated by the reference 12 times per vertex in every PCG iteration).
    def _single_index(self, e, memo):
        """("none",) no image reads below e; ("ix", index) all image reads use `index`; ("bad",) otherwise."""
        r = memo.get(e.id)
        if r is not None:
            return r
        if e.kind == "const":
            r = ("none",)
        elif e.kind == "var":
            k = e.key
            if isinstance(k, ImageAccess):
                ix = k.index
                if ix[0][0] == "s":
                    ok = ix[0][3] == 0
                else:
                    ok = all(c[2] == 0 for c in ix) and tuple(c[1] for c in ix) in self.space_of
                r = ("ix", ix) if ok and not k.image.startswith("__") else ("bad",)
            elif isinstance(k, Param):
                r = ("none",)
            else:
                r = ("bad",)
        else:
            r = ("none",)
            for a in e.args:
                ra = self._single_index(a, memo)
                if ra[0] == "bad" or (ra[0] == "ix" and r[0] == "ix" and ra[1] != r[1]):
                    r = ("bad",)
                    break
                if ra[0] == "ix":
                    r = ra
            if e.op == "sample":
                r = ("bad",)
        memo[e.id] = r
        return r

    def _space_of_index(self, ix, e):
        if ix[0][0] == "s":
            ims = ad.variables(e, lambda v: isinstance(v.key, ImageAccess))
            dom = tuple(d.idx for d in self.images[ims[0].key.image].dims)
        else:
            dom = tuple(c[1] for c in ix)
        return self.space_of.get(dom), dom

    def _hoist_gather(self, e, memo, pmemo):
        if e.id in memo:
            return memo[e.id]
        r = e
        if e.kind == "apply":
            st = self._single_index(e, pmemo)
            if st[0] == "ix" and e.op in self._EXPENSIVE and e.type == ad.REAL and self._space_of_index(st[1], e)[0] is not None:
                si, dom = self._space_of_index(st[1], e)
                dense = tuple(("d", d, 0) for d in dom)
                # the defining expression, rewritten to read the space's own element
                base = ad.substitute(e, lambda v: ad.var(v.key._replace(index=dense), v.type) if isinstance(v.key, ImageAccess) else v)
                lst = self.scoef[si]
                ch = self._scoef_index[si].get(base.id)
                if ch is None:
                    ch = len(lst)
                    self._scoef_index[si][base.id] = ch
                    lst.append(base)
                r = ad.var(ImageAccess("__coef_s%d" % si, st[1], ch))
            else:
                r = ad.rebuild(e, [self._hoist_gather(a, memo, pmemo) for a in e.args])
        memo[e.id] = r
        return r

    def _value_layout(self, gi):
        """Storage order of a materialised group's partial derivatives within one residual element:
        endpoint-major (each endpoint's transposed product reads one contiguous run), then
        term-major, padded to a multiple of four scalars for 128-bit loads."""
        pos, i = {}, 0
        for ep in self.endpoints:
            if ep["group"] != gi:
                continue
            for (ti, u, p) in ep["parts"]:
                pos[(ti, u.key)] = i
                i += 1
        return pos, i, -(-i // 4) * 4

    def _prepare_gather(self):
        """Value layouts of the materialised groups, the endpoint expressions, and hoisting of their
        per-element invariants into plan-owned coefficient images; must run before the header is emitted."""
        L = self.L
        zero = ad.const(0.0)
        for gi, g in enumerate(self.groups):
            g["vpos"], g["nnz_stored"], g["nnzp"] = self._value_layout(gi)
        self.scoef = [[] for _ in self.spaces]
        self._scoef_index = [{} for _ in self.spaces]
        hmemo, pmemo = {}, {}
        for ep in self.endpoints:
            g = self.groups[ep["group"]]
            sp = self.spaces[ep["space"]]
            acc, jp = {}, {}
            for (ti, u, p) in ep["parts"]:
                if ti not in jp:
                    jp[ti] = g["terms"][ti].jp("P")
                j = sp["slots"][(u.key.image, u.key.channel)]
                acc[j] = acc.get(j, zero) + p * jp[ti]
            if self.hoist_enabled and not g["materialize"]:
                acc = dict((j, self._hoist_gather(x, hmemo, pmemo)) for j, x in acc.items())
            ep["roots"] = acc
            # gathered PCGInit1: -J^T F and diag(J^T J) of this endpoint's unknowns
            gacc, dacc = {}, {}
            for (ti, u, p) in ep["parts"]:
                j = sp["slots"][(u.key.image, u.key.channel)]
                gacc[j] = gacc.get(j, zero) + (-1.0) * p * g["terms"][ti].exp
                dacc[j] = dacc.get(j, zero) + p * p
            if self.hoist_enabled:
                gacc = dict((j, self._hoist_gather(x, hmemo, pmemo)) for j, x in gacc.items())
                dacc = dict((j, self._hoist_gather(x, hmemo, pmemo)) for j, x in dacc.items())
            ep["jtf_roots"] = (gacc, dacc)
            if g["storejp"]:        # Jt[Jp]: this endpoint's partials times the stored J p of the residual element
                tacc = {}
                for (ti, u, p) in ep["parts"]:
                    j = sp["slots"][(u.key.image, u.key.channel)]
                    tacc[j] = tacc.get(j, zero) + p * ad.var(JpVal(ti))
                if self.hoist_enabled:
                    tacc = dict((j, self._hoist_gather(x, hmemo, pmemo)) for j, x in tacc.items())
                ep["jtp_roots"] = tacc
        for g in self.groups:
            if g["storejp"]:
                rows = [t.jp("P") for t in g["terms"]]
                if self.hoist_enabled:
                    rows = [self._hoist_gather(x, hmemo, pmemo) for x in rows]
                g["applyj_roots"] = rows
        from .dsl import Image
        for si, sp in enumerate(self.spaces):       # plan-owned coefficient images (must exist before any function is emitted)
            if self.scoef[si]:
                nm = "__coef_s%d" % si
                self.images[nm] = Image(nm, "real", len(self.scoef[si]), [L.dims[d] for d in sp["dims"]], -(2 + si), "plan")
                self.ptr_slot[nm] = len(self.ptr_pidx)
                self.ptr_pidx.append(-(2 + si))

    def gen_gather(self):
        L = self.L
        src = []
        zero = ad.const(0.0)

        def jq_lines(roots):
            used = sorted(set(v.key.i // 4 for e in roots for v in ad.variables(e, lambda v: isinstance(v.key, _JVal))))
            return ["const real4 jq%d = th_ld4(jv, %d);" % (c, c) for c in used]
        # per materialised group: store the partial derivatives; J p per residual row
        for gi, g in enumerate(self.groups):
            if not g["materialize"]:
                continue
            dom = g["domain"]
            vals = [zero] * g["nnzp"]
            for ti, t in enumerate(g["terms"]):
                for u, p in zip(t.unknowns, t.partials):
                    vals[g["vpos"][(ti, u.key)]] = p
            src.append(self._fn(
                "template <class A> __device__ __forceinline__ void computeJv_g%d(const A& a, const Params& P, real* __restrict__ jv)" % gi,
                vals, lambda r: ["jv[%d] = %s;" % (i, x) for i, x in enumerate(r)], dom))
            rows = []
            for ti, t in enumerate(g["terms"]):
                r = zero
                for u, p in zip(t.unknowns, t.partials):
                    k = u.key
                    r = r + ad.var(JVal(g["vpos"][(ti, k)])) * ad.var(VecArg("P", k.image, k.index, k.channel))
                rows.append(r)
            g["jv_roots"], g["matj_roots"] = vals, rows        # kept for the NumPy interpreter (frontend/interp.py)
            src.append(self._fn(
                "template <class A> __device__ __forceinline__ void matJ_g%d(const A& a, const Params& P, const real* __restrict__ jv, real* __restrict__ jpv)" % gi,
                rows, lambda r: ["jpv[%d] = %s;" % (i, x) for i, x in enumerate(r)], dom, jq_lines(rows)))
        # per Jt[Jp] group: J p per residual row, matrix-free (applyJ, thallo.t:3754-3790)
        for gi, g in enumerate(self.groups):
            if not g["storejp"]:
                continue
            src.append(self._fn(
                "template <class A> __device__ __forceinline__ void applyJ_g%d(const A& a, const Params& P, real* __restrict__ jpv)" % gi,
                g["applyj_roots"], lambda r: ["jpv[%d] = %s;" % (i, x) for i, x in enumerate(r)], g["domain"]))
        # per endpoint: contribution of one residual element to the unknowns at that endpoint
        for si, sp in enumerate(self.spaces):
            if self.scoef[si]:
                nc = len(self.scoef[si])
                src.append(self._fn(
                    "template <class A> __device__ __forceinline__ void scoef_s%d(const A& a, const Params& P, real* __restrict__ out)" % si,
                    list(self.scoef[si]), lambda r, nc=nc: ["out[%d] = %s;" % (i, r[i]) for i in range(nc)], sp["dims"]))
        for ep in self.endpoints:
            g = self.groups[ep["group"]]
            sp = self.spaces[ep["space"]]
            dom = g["domain"]
            acc = ep["roots"]
            js = sorted(acc)
            src.append(self._fn(
                "template <class A> __device__ __forceinline__ void jtj_ep%d(const A& a, const Params& P, real* __restrict__ acc)" % ep["id"],
                [acc[j] for j in js], lambda r, js=js: ["acc[%d] += %s;" % (j, x) for j, x in zip(js, r)], dom))
            gacc, dacc = ep["jtf_roots"]
            gjs = sorted(gacc)
            src.append(self._fn(
                "template <class A> __device__ __forceinline__ void jtf_ep%d(const A& a, const Params& P, real* __restrict__ accg, real* __restrict__ accd)" % ep["id"],
                [gacc[j] for j in gjs] + [dacc[j] for j in gjs],
                lambda r, gjs=gjs: ["accg[%d] += %s;" % (j, x) for j, x in zip(gjs, r[:len(gjs)])] +
                                   ["accd[%d] += %s;" % (j, x) for j, x in zip(gjs, r[len(gjs):])], dom))
            if g["storejp"]:          # applyJt (thallo.t:3808-3841), gathered instead of scattered
                tacc = ep["jtp_roots"]
                tjs = sorted(tacc)
                src.append(self._fn(
                    "template <class A> __device__ __forceinline__ void jtp_ep%d(const A& a, const Params& P, const real* __restrict__ jpv, real* __restrict__ acc)" % ep["id"],
                    [tacc[j] for j in tjs], lambda r, tjs=tjs: ["acc[%d] += %s;" % (j, x) for j, x in zip(tjs, r)], dom))
            if g["materialize"]:
                macc = {}
                for (ti, u, p) in ep["parts"]:
                    j = sp["slots"][(u.key.image, u.key.channel)]
                    macc[j] = macc.get(j, zero) + ad.var(JVal(g["vpos"][(ti, u.key)])) * ad.var(JpVal(ti))
                roots = [macc[j] for j in js]
                ep["mat_roots"] = dict(zip(js, roots))
                src.append(self._fn(
                    "__device__ __forceinline__ void jt_ep%d(const real* __restrict__ jv, const real* __restrict__ jpv, real* __restrict__ acc)" % ep["id"],
                    roots, lambda r, js=js: ["acc[%d] += %s;" % (j, x) for j, x in zip(js, r)], dom, jq_lines(roots)))
        # per index space: walk the residual elements incident to one unknown element
        part_a, src = "\n".join(src), []
        for si, sp in enumerate(self.spaces):
            elements = _prod(L.dims[d].size for d in sp["dims"])
            deg = 0.0
            for ep in sp["endpoints"]:
                if ep["kind"] == "sparse":
                    deg += _prod(L.dims[d].size for d in self.groups[ep["group"]]["domain"]) / float(elements)
            sp["elements"] = elements
            # lanes sharing one unknown element: a warp when many residuals meet there (the cameras of
            # bundle adjustment), else one thread per element.  Splitting a short adjacency list over
            # 2-8 lanes was measured on B200 and loses (arap_mesh 2000x2000, degree 2 x 6: th_gather_s0
            # 0.319 ms with 1 lane, 0.466 with 2, 0.786 with 4; profiles/r01j_sweep.txt): the walk is bound
            # by the per-edge arithmetic, and idle lanes of the narrower lists cost more than the
            # shorter dependent-load chains save.  THALLO_B200_GATHER_LANES overrides for experiments.
            if deg >= 64.0:
                lanes = 32
            elif self.gather_lanes is not None:
                lanes = self.gather_lanes
            else:
                lanes = 1
            sp["lanes"] = lanes
            body, jbody = [], []
            for ep in sp["endpoints"]:      # the same walk for PCGInit1 (J^T F, diag J^T J): always matrix-free
                gi = ep["group"]
                call = "jtf_ep%d(a, P, accg, accd);" % ep["id"]
                if ep["kind"] == "sparse":
                    sid = ep["sid"]
                    jbody.append("{")
                    jbody.append("    const int lo = __ldg(G.ptr[%d] + t.lin), hi = __ldg(G.ptr[%d] + t.lin + 1);" % (sid, sid))
                    jbody.append("    const int* __restrict__ perm = G.perm[%d];" % sid)
                    jbody.append("    for (int i = lo + lane; i < hi; i += LANES) {")
                    jbody.append("        const long long e = perm ? (long long)__ldg(perm + i) : (long long)i;")
                    jbody.append("        ThIdx<dom_g%d> idx; idx.from_linear(e);" % gi)
                    jbody.append("        GAcc<dom_g%d, TH_OWN_ENDPOINT ? %d : -1> a(idx, nullptr, t.lin);" % (gi, self.ptr_slot[ep["sparse"]]))
                    jbody.append("        " + call)
                    jbody.append("    }")
                    jbody.append("}")
                else:
                    o = ep["off"]
                    jbody.append("if (lane == 0) {")
                    jbody.append("    const int x = t.c[0] - (%d), y = t.c[1] - (%d), z = t.c[2] - (%d);" % (o[0], o[1], o[2]))
                    jbody.append("    ThIdx<dom_g%d> idx;" % gi)
                    jbody.append("    if (x >= 0 && y >= 0 && z >= 0 && idx.from_coords(x, y, z)) {")
                    jbody.append("        GAcc<dom_g%d> a(idx, nullptr);" % gi)
                    jbody.append("        " + call)
                    jbody.append("    }")
                    jbody.append("}")
            src.append("template <int LANES> __device__ __forceinline__ void gatherjtf_s%d(const ThIdx<dom_s%d>& t, int lane, "
                       "const Params& P, const ThGather& G, real* __restrict__ accg, real* __restrict__ accd) {\n    %s\n}\n"
                       % (si, si, "\n    ".join(jbody)))
            for ep in sp["endpoints"]:
                gi = ep["group"]
                g = self.groups[gi]
                mat = bool(g["materialize"])
                sjp = bool(g["storejp"])
                call_free = "jtj_ep%d(a, P, acc);" % ep["id"]
                call_jtp = "jtp_ep%d(a, P, G.jp[%d] + idx.lin * %d, acc);" % (ep["id"], gi, len(g["terms"]))
                call_mat = ("jt_ep%d(G.jvals[%d] + e * %d, G.jp[%d] + e * %d, acc);" % (ep["id"], gi, g["nnzp"], gi, len(g["terms"])))
                if ep["kind"] == "sparse":
                    sid = ep["sid"]
                    body.append("{   // endpoint %d: group %s through %s" % (ep["id"], g["name"], ep["sparse"]))
                    if mat or sjp:
                        body.append("    if (WHICH == 0) {")      # A*delta of the LM reset skips groups without an applyJTJ (gauss_newton.t:1058-1065)
                    body.append("    const int lo = __ldg(G.ptr[%d] + t.lin), hi = __ldg(G.ptr[%d] + t.lin + 1);" % (sid, sid))
                    body.append("    const int* __restrict__ perm = G.perm[%d];" % sid)
                    if self.gather_unroll > 1:      # lets the compiler overlap the dependent index -> neighbour loads of consecutive edges
                        body.append("    #pragma unroll %d" % self.gather_unroll)
                    body.append("    for (int i = lo + lane; i < hi; i += LANES) {")
                    body.append("        const long long e = perm ? (long long)__ldg(perm + i) : (long long)i;")
                    if mat:
                        body.append("        " + call_mat)
                    else:
                        body.append("        ThIdx<dom_g%d> idx; idx.from_linear(e);" % gi)
                        body.append("        GAcc<dom_g%d, TH_OWN_ENDPOINT ? %d : -1> a(idx, vec, t.lin);" % (gi, self.ptr_slot[ep["sparse"]]))
                        body.append("        " + (call_jtp if sjp else call_free))
                    body.append("    }")
                    if mat or sjp:
                        body.append("    }")
                    body.append("}")
                else:
                    o = ep["off"]
                    body.append("if (lane == 0%s) {   // endpoint %d: group %s at offset (%d, %d, %d)"
                                % (" && WHICH == 0" if (mat or sjp) else "", ep["id"], g["name"], o[0], o[1], o[2]))
                    body.append("    const int x = t.c[0] - (%d), y = t.c[1] - (%d), z = t.c[2] - (%d);" % (o[0], o[1], o[2]))
                    body.append("    ThIdx<dom_g%d> idx;" % gi)
                    body.append("    if (x >= 0 && y >= 0 && z >= 0 && idx.from_coords(x, y, z)) {")
                    if mat:
                        body.append("        const long long e = idx.lin;")
                        body.append("        " + call_mat)
                    else:
                        body.append("        GAcc<dom_g%d> a(idx, vec);" % gi)
                        body.append("        " + (call_jtp if sjp else call_free))
                    body.append("    }")
                    body.append("}")
            src.append("template <int WHICH, int LANES> __device__ __forceinline__ void gather_s%d(const ThIdx<dom_s%d>& t, int lane, "
                       "const Params& P, const ThGather& G, const real* __restrict__ vec, real* __restrict__ acc) {\n    %s\n}\n"
                       % (si, si, "\n    ".join(body)))
        return part_a, "\n".join(src)

    def gen_exclude(self):
        src = []
        for k, im in enumerate(self.unknowns):
            dom = tuple(d.idx for d in im.dims)
            e = ad.const(False)          # one predicate per index space (thallo.t:5536-5538,5618-5624)
            for other in self.unknowns:
                if tuple(x.idx for x in other.dims) == dom and other.exclude is not None:
                    e = ad.or_(e, other.exclude)
            src.append(self._fn(
                "template <class A> __device__ __forceinline__ bool exclude_u%d(const A& a, const Params& P)" % k,
                [e], lambda r: ["return %s;" % r[0]], dom))
            if self.schedule == "gather" and self.spaces[self.space_of[dom]]["images"][0] is im:
                src.append(self._fn(
                    "template <class A> __device__ __forceinline__ bool exclude_s%d(const A& a, const Params& P)" % self.space_of[dom],
                    [e], lambda r: ["return %s;" % r[0]], dom))
        return "\n".join(src)

    # ---- residualwise functions (per group)
    def gen_group(self, gi, g):
        dom = g["domain"]
        terms = g["terms"]
        src = []
        half = ad.const(0.5)
        # cost
        c = ad.const(0.0)
        for t in terms:
            c = c + t.exp * t.exp
        src.append(self._fn(
            "template <class A> __device__ __forceinline__ real cost_g%d(const A& a, const Params& P)" % gi,
            [half * c], lambda r: ["return %s;" % r[0]], dom))
        # residual values (for tests / debugging)
        src.append(self._fn(
            "template <class A> __device__ __forceinline__ void residuals_g%d(const A& a, const Params& P, real* __restrict__ out)" % gi,
            [t.exp for t in terms], lambda r: ["out[%d] = %s;" % (i, x) for i, x in enumerate(r)], dom))
        # model cost: 1/2 sum (F + J delta)^2
        m = ad.const(0.0)
        for t in terms:
            rm = t.exp + t.jp("Delta")
            m = m + rm * rm
        src.append(self._fn(
            "template <class A> __device__ __forceinline__ real modelcost_g%d(const A& a, const Params& P)" % gi,
            [half * m], lambda r: ["return %s;" % r[0]], dom))
        # evalJTF scatter: R[u] += -partial*F ; Pre[u] += partial^2
        tg, tgmap = [], {}
        for t in terms:
            for u, p in zip(t.unknowns, t.partials):
                key = u.key
                if key not in tgmap:
                    tgmap[key] = [ad.const(0.0), ad.const(0.0)]
                    tg.append(key)
                tgmap[key][0] = tgmap[key][0] + (-1.0) * p * t.exp
                tgmap[key][1] = tgmap[key][1] + p * p
        roots = [tgmap[k][0] for k in tg] + [tgmap[k][1] for k in tg]
        n = len(tg)
        # one `prepare` per sparse index array the targets go through (target element + warp peers)
        sps = sorted(set(self.ptr_slot[k.index[0][1]] for k in tg if k.index[0][0] == "s"))
        prep = ["s.template prepare<%d>(P);" % sp for sp in sps]
        src.append(self._fn(
            "template <class A, class S> __device__ __forceinline__ void evalJTF_g%d(const A& a, const Params& P, S& s)" % gi,
            roots,
            lambda r: prep + [self._scatter_stmt(0, tg[i], r[i], dom) for i in range(n)] +
                      [self._scatter_stmt(1, tg[i], r[n + i], dom) for i in range(n)], dom))
        # applyJTJ scatter: Ap[u] += partial*Jp
        tmap = {}
        for t in terms:
            jp = t.jp("P")
            for u, p in zip(t.unknowns, t.partials):
                tmap[u.key] = tmap.get(u.key, ad.const(0.0)) + p * jp
        roots = [tmap[k] for k in tg]
        src.append(self._fn(
            "template <class A, class S> __device__ __forceinline__ void applyJTJ_g%d(const A& a, const Params& P, S& s)" % gi,
            roots, lambda r: prep + [self._scatter_stmt(0, tg[i], r[i], dom) for i in range(n)], dom))
        # computeJ: partials term-major, unknown-minor; columns via accessor
        vals, cols = [], []
        for t in terms:
            for u, p in zip(t.unknowns, t.partials):
                vals.append(p)
                cols.append(u.key)
        g["nnz_per_elem"] = len(vals)
        g["row_nnz"] = [len(t.unknowns) for t in terms]

        def col_stmt(i, key):
            im = self.images[key.image]
            k = self.uidx[im.name]
            if key.index[0][0] == "s":
                return "cols[%d] = a.template sucol<%d, %d, %d>(P);" % (i, k, key.channel, self.ptr_slot[key.index[0][1]])
            o = self._offs(key.index, dom)
            return "cols[%d] = a.template ucol<%d, %d, %d, %d, %d>();" % (i, k, key.channel, o[0], o[1], o[2])
        src.append(self._fn(
            "template <class A> __device__ __forceinline__ void computeJ_g%d(const A& a, const Params& P, real* __restrict__ vals, long long* __restrict__ cols)" % gi,
            vals, lambda r: ["vals[%d] = %s;" % (i, x) for i, x in enumerate(r)] +
                            [col_stmt(i, k) for i, k in enumerate(cols)], dom))
        g["targets"] = tg
        return "\n".join(src)

    # ---- whole translation unit
    def generate(self):
        L = self.L
        out = Lowered()
        if self.schedule == "at_output":
            self._prepare_unknownwise()
        if self.schedule == "gather":
            self._prepare_gather()
        hdr = []
        hdr.append("// generated by thallo_b200.frontend.codegen for energy '%s' (%s)" % (self.name, self.kind))
        hdr.append("#define TH_DOUBLE %d" % int(self.double))
        hdr.append("#define TH_LM %d" % int(self.lm))
        hdr.append("#define TH_USEPRE %d" % int(L.usepreconditioner))
        hdr.append("#define TH_AT_OUTPUT %d" % int(self.schedule == "at_output"))
        hdr.append("#define TH_GATHER %d" % int(self.schedule == "gather"))
        if self.gpartition is not None:
            assert self.schedule == "gather", "a graph partition needs the gather schedule"
            assert not self.computed, "computed arrays are not supported by the multi-GPU graph partition yet"
            assert all(len(im.dims) == 1 for im in self.unknowns), "a graph partition needs 1-D unknown domains"
            ghosted = sorted(set(im.dims[0].idx for im in self.unknowns if any(self.gpartition.get(im.dims[0].idx, (0, 0)))))
            assert len(ghosted) <= 1, "at most one unknown domain can carry ghost elements"
            assert not (ghosted and ghosted[0] in self.replicated_dims)
            hdr.append("#define TH_MULTI 1")
            hdr.append("#define TH_PART_TABLE {%s}" % ", ".join("{%dLL, %dLL}" % self.gpartition.get(d.idx, (0, 0)) for d in L.dims))
            rng = []
            for im in self.unknowns:
                glo, ghi = self.gpartition.get(im.dims[0].idx, (0, 0))
                rng.append("{%dLL, %dLL}" % (self.uoff[im.name] + glo * im.channels, self.uoff[im.name] + (im.elements - ghi) * im.channels))
            hdr.append("#define TH_RANGE_TABLE {%s}" % ", ".join(rng))
            vd = ghosted[0] if ghosted else -1
            self.gpart_line = (vd, L.dims[vd].size if ghosted else 0) + (self.gpartition[vd] if ghosted else (0, 0))
            self.rep_images = [k for k, im in enumerate(self.unknowns) if im.dims[0].idx in self.replicated_dims]
            if self.rep_images:
                hdr.append("#define TH_REP_TABLE {%s}" % ", ".join("1" if k in self.rep_images else "0" for k in range(len(self.unknowns))))
                hdr.append("#define TH_REP_OWNER %d" % int(self.rep_owner))
                hdr.append("#define TH_SPACE_REP {%s}" % ", ".join(
                    "1" if sp["dims"][0] in self.replicated_dims else "0" for sp in self.spaces))
        if self.partition is not None:
            assert self.tiled, "multi-GPU partitioning needs the tiled at-output schedule (2-D / 3-D image domain)"
            assert all(tuple(g["domain"]) == tuple(self.udomain) for g in self.groups), \
                "multi-GPU partitioning needs every residual domain to equal the unknown domain"
            hdr.append("#define TH_MULTI 1")
            hdr.append("#define TH_GHOST_LO %d\n#define TH_GHOST_HI %d" % self.partition)
            hdr.append("#define TH_DSLOW %d" % L.dims[self.udomain[-1]].size)
        hdr.append("#define TH_NUM_UIMG %d" % len(self.unknowns))
        hdr.append("#define TH_NUNK %dLL" % self.nunk)
        hdr.append("#define TH_NPTR %d" % max(1, len(self.ptr_pidx)))
        hdr.append("#define TH_NSC %d" % max(1, len(self.sc_defs)))
        hdr.append("#define TH_NGROUPS %d" % len(self.groups))
        hdr.append("#define TH_NDIMS %d" % len(L.dims))
        hdr.append("#define TH_DIM_SIZES {%s}" % ", ".join(str(d.size) for d in L.dims))
        # unknown image table: channels, flat offset, ptr slot, ndim, dim indices
        rows = []
        for im in self.unknowns:
            di = [d.idx for d in im.dims] + [0] * (MAXD - len(im.dims))
            rows.append("{%d, %dLL, %d, %d, {%d, %d, %d}, %dLL}" % (im.channels, self.uoff[im.name], self.ptr_slot[im.name],
                                                                   len(im.dims), di[0], di[1], di[2], im.elements))
        hdr.append("#define TH_UIMG_TABLE {%s}" % ", ".join(rows))
        body = []
        body.append(self.gen_exclude())
        if self.schedule == "at_output":
            dom = self.udomain
            hdr.append("#define TH_UW_NDIM %d" % len(dom))
            hdr.append("#define TH_UW_DIMS {%s}" % ", ".join(str(L.dims[d].size) for d in dom))
            body.append(self.gen_unknownwise())
            hdr.append("#define TH_U %d" % self.U)
            hdr.append("#define TH_NCOEF %d" % len(self.coef_exprs))
            if self.coef_exprs:
                hdr.append("#define TH_COEF_SLOT %d" % self.ptr_slot["__coef"])
            hdr.append("#define TH_TILED %d" % int(self.tiled))
            if self.tiled:
                tl = self.tl
                hdr.append("#define TH_TW %d\n#define TH_TH %d\n#define TH_TD %d" % tuple(tl["tile"]))
                hdr.append("#define TH_HX %d\n#define TH_HY %d\n#define TH_HZ %d" % tuple(tl["halo"]))
                hdr.append("#define TH_SMEM_BYTES %d" % max(128, tl["smem"]))
                hdr.append("#define TH_PIPE %d" % tl["pipe"])
                hdr.append("#define TH_PCG_A_MINB %d" % tl["minb"])
                hdr.append("#define TH_NSTAGE %d" % len(tl["stages"]))
                hdr.append("#define TH_STAGE_TABLE {%s}" % (", ".join(
                    "{%d, %d, %d, %d, %d, %d, %d, %d}" % (st["slot"], st["es"], st["channels"], st["roww"], st["off"], st["padl"],
                                                          st["center"], st["bytes"])
                    for st in tl["stages"]) or "{0, 0, 0, 0, 0, 0, 0, 0}"))
                hdr.append("#define TH_SLOT_STAGE_TABLE {%s}" % ", ".join(map(str, tl["slot_stage"])))
                hdr.append("#define TH_VTILE_TABLE {%s}" % ", ".join(
                    "{%d, %d, %d, %d, %d, %d, %d, %d}" % (v["roww"], v["zoff"], v["poff"], v["bytes"], v["padl"], v["coff"],
                                                          v["croww"], v["cbytes"]) for v in tl["vt"]))
                hdr.append("#define TH_STAGE_CTC %d" % int(all(v["coff"] >= 0 for v in tl["vt"])))
                # reach of the bounds predicates (tiles further than this from the domain border skip them)
                R = [0] * MAXD
                roots = list(self.uw_roots["out"]) + [im.exclude for im in self.unknowns if im.exclude is not None]
                if self.two_phase:
                    roots += self.two_phase_roots["jp"] + self.two_phase_roots["out"]
                for e in roots:
                    for v in ad.variables(e, lambda v: isinstance(v.key, Bounds)):
                        for (dd, lo, hi) in v.key.ranges:
                            pos = self.udomain.index(dd)
                            R[pos] = max(R[pos], abs(lo), abs(hi))
                if self.two_phase:
                    R = [r + ph for r, ph in zip(R, self.tp["ph"])]      # predicates are also evaluated at the positions around the tile
                # Measured on B200 (profiles/r02g_*): dropping the predicates on interior tiles gains 2-4 % on optical_flow,
                # shape_from_shading and the volume and LOSES 7 % on the headline's th_pcg_a (four instantiations of the
                # tile operator in one kernel), so it is opt-in: THALLO_B200_EDGE_SPECIALIZE=1
                if not os.environ.get("THALLO_B200_EDGE_SPECIALIZE"):
                    R = [1000000] * MAXD
                hdr.append("#define TH_INB_RX %d\n#define TH_INB_RY %d\n#define TH_INB_RZ %d" % tuple(R))
                hdr.append("#define TH_TWO_PHASE %d" % int(self.two_phase))
                if self.two_phase:
                    tp = self.tp
                    hdr.append("#define TH_JP_NT %d" % tp["nt"])
                    hdr.append("#define TH_JP_BUFS %d" % tp["bufs"])
                    hdr.append("#define TH_JP_PHX %d\n#define TH_JP_PHY %d\n#define TH_JP_PHZ %d" % tuple(tp["ph"]))
                    hdr.append("#define TH_JP_NHALO %d" % len(tp["pos"]))
                    hdr.append("#define TH_JP_POS_TABLE {%s}" % (", ".join("%du" % x for x in tp["pos"]) or "0u"))
        gl = []
        for gi, g in enumerate(self.groups):
            body.append(self.gen_group(gi, g))
            gl.append("X(%d)" % gi)
        gather_b = ""
        if self.schedule == "gather":
            gather_a, gather_b = self.gen_gather()
            body.append(gather_a)
            hdr.append("#define TH_NEP_S %d" % self.n_sparse_ep)
            hdr.append("#define TH_NSPACES %d" % len(self.spaces))
            hdr.append("#define TH_SPACE_LIST(X) %s" % " ".join("X(%d)" % i for i in range(len(self.spaces))))
            hdr.append("#define TH_MAT_LIST(X) %s" % " ".join("X(%d)" % gi for gi, g in enumerate(self.groups) if g["materialize"]))
            hdr.append("#define TH_JP_LIST(X) %s" % " ".join("X(%d)" % gi for gi, g in enumerate(self.groups) if g["storejp"]))
            hdr.append("#define TH_SPACE_TABLE {%s}" % ", ".join("{%d, %d, %dLL}" % (sp["nslots"], sp["lanes"], sp["elements"])
                                                                 for sp in self.spaces))
            mx = max(sp["nslots"] for sp in self.spaces)
            hdr.append("#define TH_MAXSLOTS %d" % mx)
            rows = []
            for sp in self.spaces:
                sl = [(self.uidx[im.name], ch) for im in sp["images"] for ch in range(im.channels)]
                sl += [(0, 0)] * (mx - len(sl))
                rows.append("{%s}" % ", ".join("{%d, %d}" % x for x in sl))
            hdr.append("#define TH_SLOT_TABLE {%s}" % ", ".join(rows))
            hdr.append("#define TH_GROUP_NNZP {%s}" % ", ".join(str(g["nnzp"]) for g in self.groups))
            hdr.append("#define TH_SCOEF_N {%s}" % ", ".join(str(len(c)) for c in self.scoef))
            hdr.append("#define TH_SCOEF_SLOT {%s}" % ", ".join(str(self.ptr_slot.get("__coef_s%d" % si, 0)) for si in range(len(self.spaces))))
            hdr.append("#define TH_SCOEF_LIST(X) %s" % " ".join("X(%d)" % si for si in range(len(self.spaces)) if self.scoef[si]))
        if self.schedule != "at_output":
            hdr.append("#define TH_TILED 0\n#define TH_NCOEF 0")
        # ComputedArrays: value + gradient channels of one element (createprecomputed, thallo.t:4046-4094)
        hdr.append("#define TH_NCOMPUTED %d" % len(self.computed))
        if self.computed:
            # under the slab partition the stored images are evaluated locally on owned AND ghost layers (the unknowns'
            # ghost layers are kept current); the outermost ghost layer may read past the local extent and hold a
            # wrong value, which no owned residual reaches (halo = stencil reach of the residuals through the arrays)
            rows = []
            for k, ca in enumerate(self.computed):
                grads = [g for g, ch in zip(ca.gradients, ca.gchannel) if ch >= 0]
                ng = len(grads)
                body.append(self._fn(
                    "template <class A> __device__ __forceinline__ void precompute_c%d(const A& a, const Params& P, real* __restrict__ out)" % k,
                    [ca.expression] + grads, lambda r: ["out[%d] = %s;" % (i, x) for i, x in enumerate(r)],
                    tuple(d.idx for d in ca.dims)))
                rows.append("{%d, %d, %d}" % (self.ptr_slot[ca.name], self.ptr_slot[ca.gradient_image.name] if ng else -1, ng))
            hdr.append("#define TH_COMPUTED_TABLE {%s}" % ", ".join(rows))
            hdr.append("#define TH_COMPUTED_LIST(X) %s" % " ".join("X(%d)" % k for k in range(len(self.computed))))
        hdr.append("#define TH_GROUP_LIST(X) %s" % " ".join(gl))
        # group domain table
        rows = []
        for g in self.groups:
            dom = list(g["domain"]) + [0] * (MAXD - len(g["domain"]))
            rows.append("{%d, {%d, %d, %d}, %d, %d}" % (len(g["domain"]), dom[0], dom[1], dom[2], len(g["terms"]), g["nnz_per_elem"]))
        hdr.append("#define TH_GROUP_TABLE {%s}" % ", ".join(rows))
        doms = []

        def dom_struct(nm, dimidx):
            sz = [L.dims[x].size for x in dimidx] + [1] * (MAXD - len(dimidx))
            ix = list(dimidx) + [0] * (MAXD - len(dimidx))
            return ("struct %s { static constexpr int ND = %d; static constexpr long long D0 = %d, D1 = %d, D2 = %d; "
                    "static constexpr int I0 = %d, I1 = %d, I2 = %d; };"
                    % (nm, len(dimidx), sz[0], sz[1], sz[2], ix[0], ix[1], ix[2]))
        for k, im in enumerate(self.unknowns):
            doms.append(dom_struct("dom_u%d" % k, [x.idx for x in im.dims]))
        if self.schedule == "at_output":
            doms.append(dom_struct("dom_uw", list(self.udomain)))
        for gi, g in enumerate(self.groups):
            doms.append(dom_struct("dom_g%d" % gi, list(g["domain"])))
        if self.schedule == "gather":
            for si, sp in enumerate(self.spaces):
                doms.append(dom_struct("dom_s%d" % si, list(sp["dims"])))
        for k, ca in enumerate(self.computed):
            doms.append(dom_struct("dom_c%d" % k, [x.idx for x in ca.dims]))
        src = ("\n".join(hdr) + "\n#include \"thallo_prelude.cuh\"\nnamespace th {\n" + "\n".join(doms) + "\n"
               + "\n".join(body) + "\n} // namespace th\n")
        if gather_b:
            src += "#include \"thallo_access.cuh\"\nnamespace th {\n" + gather_b + "\n} // namespace th\n"
        src += "#include \"thallo_kernels.cuh\"\n"
        out.source = src
        # ---- descriptor
        d = dict(
            name=self.name, kind=self.kind, lm=int(self.lm), real="double" if self.double else "float",
            usepreconditioner=int(L.usepreconditioner), schedule=self.schedule,
            dims=[dd.size for dd in L.dims], dim_names=[dd.name for dd in L.dims],
            nunk=self.nunk, ptr_pidx=self.ptr_pidx, sc_defs=self.sc_defs,
            unknowns=[dict(name=im.name, channels=im.channels, offset=self.uoff[im.name], pidx=im.pidx,
                           elements=im.elements, dims=[x.idx for x in im.dims]) for im in self.unknowns],
            groups=[dict(name=g["name"], domain=list(g["domain"]), nterms=len(g["terms"]),
                         count=_prod(L.dims[x].size for x in g["domain"]), materialize=(1 if g["materialize"] else 2 if g["storejp"] else 0),
                         nnz_per_elem=g["nnz_per_elem"], row_nnz=g["row_nnz"]) for g in self.groups],
        )
        if self.gpartition is not None:
            d["gpartition"] = self.gpart_line
            d["replicated"] = [(self.uoff[self.unknowns[k].name], self.unknowns[k].cardinality) for k in self.rep_images]
        d["computed"] = [dict(elements=ca.elements, ngrad=sum(1 for ch in ca.gchannel if ch >= 0)) for ca in self.computed]
        if self.schedule == "gather":
            d["gather"] = dict(
                spaces=[dict(elements=sp["elements"], lanes=sp["lanes"], nslots=sp["nslots"]) for sp in self.spaces],
                sparse_endpoints=[dict(sid=ep["sid"], group=ep["group"], slot=self.ptr_slot[ep["sparse"]],
                                       count=_prod(L.dims[x].size for x in self.groups[ep["group"]]["domain"]),
                                       targets=self.spaces[ep["space"]]["elements"])
                                  for ep in self.endpoints if ep["kind"] == "sparse"],
                groups=[dict(nnzp=g["nnzp"], nterms=len(g["terms"])) for g in self.groups],
                scoef=[dict(space=si, slot=self.ptr_slot["__coef_s%d" % si], channels=len(c)) for si, c in enumerate(self.scoef) if c])
        if self.schedule == "at_output":
            d["U"] = self.U
            d["uw_dims"] = [L.dims[x].size for x in self.udomain]
            d["halo_img"] = {k: v for k, v in self.halo["img"].items()}
            d["halo_vec"] = {k: v for k, v in self.halo["vec"].items()}
            d["ncoef"] = len(self.coef_exprs)
            d["partition"] = self.partition
            d["tiled"] = int(self.tiled)
            if self.tiled:
                d["tile"] = self.tl
                d["jp_bytes"] = (self.tp["nt"] * self.tp["nbox"] * (8 if self.double else 4) * self.tp["bufs"]) if self.two_phase else 0
        out.desc = d
        return out


def _prod(it):
    r = 1
    for x in it:
        r *= x
    return r


def descriptor_text(d):
    """Line-based serialisation read by csrc/plan_desc.cpp (no JSON parser in the C library)."""
    ln = []
    ln.append("name %s" % d["name"])
    ln.append("kind %s" % d["kind"])
    ln.append("lm %d" % d["lm"])
    ln.append("real %s" % d["real"])
    ln.append("usepreconditioner %d" % d["usepreconditioner"])
    ln.append("schedule %s" % d["schedule"])
    ln.append("dims %d %s" % (len(d["dims"]), " ".join(map(str, d["dims"]))))
    ln.append("nunk %d" % d["nunk"])
    ln.append("ptrs %d %s" % (len(d["ptr_pidx"]), " ".join(map(str, d["ptr_pidx"]))))
    ln.append("scalars %d %s" % (len(d["sc_defs"]), " ".join("%d:%s" % (p, t) for p, t in d["sc_defs"])))
    for u in d["unknowns"]:
        ln.append("unknown %s %d %d %d %d %d %s" % (u["name"], u["channels"], u["offset"], u["pidx"], u["elements"],
                                                   len(u["dims"]), " ".join(map(str, u["dims"]))))
    for g in d["groups"]:
        ln.append("group %s %d %d %d %d %d %s | %s" % (g["name"], g["count"], g["nterms"], g["materialize"], g["nnz_per_elem"],
                                                      len(g["domain"]), " ".join(map(str, g["domain"])),
                                                      " ".join(map(str, g["row_nnz"]))))
    for k, c in enumerate(d.get("computed", [])):
        ln.append("computed %d %d %d" % (k, c["elements"], c["ngrad"]))
    if d.get("gpartition") is not None:
        ln.append("gpartition %d %d %d %d" % tuple(d["gpartition"]))
        for off, n in d.get("replicated", []):
            ln.append("replicated %d %d" % (off, n))
    if d["schedule"] == "gather":
        ga = d["gather"]
        for sp in ga["spaces"]:
            ln.append("space %d %d %d" % (sp["elements"], sp["lanes"], sp["nslots"]))
        for ep in ga["sparse_endpoints"]:
            ln.append("sep %d %d %d %d %d" % (ep["sid"], ep["group"], ep["slot"], ep["count"], ep["targets"]))
        for gi, g in enumerate(ga["groups"]):
            ln.append("gmat %d %d %d" % (gi, g["nnzp"], g["nterms"]))
        for c in ga["scoef"]:
            ln.append("scoef %d %d %d" % (c["space"], c["slot"], c["channels"]))
    if d["schedule"] == "at_output":
        ln.append("U %d" % d["U"])
        ln.append("uw_dims %d %s" % (len(d["uw_dims"]), " ".join(map(str, d["uw_dims"]))))
        ln.append("ncoef %d" % d.get("ncoef", 0))
        if d.get("partition") is not None:
            ln.append("partition %d %d" % tuple(d["partition"]))
        if d.get("tiled"):
            tl = d["tile"]
            ln.append("tile %s %s %d %d" % (" ".join(map(str, tl["tile"])), " ".join(map(str, tl["halo"])), max(128, tl["smem"]), tl["pipe"]))
            if d.get("jp_bytes"):
                ln.append("jpbytes %d" % d["jp_bytes"])
            for v in tl["vt"]:
                ln.append("vtile %d %d %d %d %d %d %d %d" % (v["roww"], v["zoff"], v["poff"], v["bytes"], v["padl"], v["coff"],
                                                             v["croww"], v["cbytes"]))
            for st in tl["stages"]:
                ln.append("stage %d %s %d %d %d %d %d %d %d" % (st["slot"], st["ctype"], st["es"], st["channels"], st["roww"], st["off"],
                                                                st["bytes"], st["padl"], st["center"]))
    return "\n".join(ln) + "\n"


def lower(define, dims, kind="gauss_newton", name="energy", double=False, schedule="auto",
          lm_as_committed=False, hoist=True, tile=None, partition=None, jp_all=False, **define_kwargs):
    """jp_all: give every residual group the Jt[Jp] schedule (as if the energy said `r.<g>.Jp:set_materialize(true)` for
    each) -- with schedule="gather" on an image domain this is the two-pass operator: J p per residual stored once,
    then every unknown applies its transposed partials to the stored values of the residuals around it."""
    from .dsl import build_spec
    L = build_spec(define, dims, **define_kwargs)
    if jp_all:
        for g in L.residuals.groups:
            g.Jp.set_materialize(True)
    gen = Generator(L, name, kind, double, schedule, lm_as_committed, hoist, tile, partition)
    out = gen.generate()
    out.generator = gen
    return out
