"""ORACLE (test infrastructure, not product code): CPU restatement of the
reference's Gauss-Newton / Levenberg-Marquardt outer loop and PCG inner loop.

Follows reference API/src/gauss_newton.t:
  init      :1166-1198      step      :1545-1785     finalize :1200-1212
  PCGInit1 (at-output, fused) :678-710      PCGInit1_Finish (residualwise) :712-731
  PCGStep1 / _Finish :734-799 (+ CtC p in LM)   PCGStep2 :801-843   reset variant :845-886
  PCGStep3 :889-899    scalar copy :1665    zeta test :1666-1686
  LM diagonal: PCGSaveSSq :929-934, computeCtC thallo.t:3911-3937, PCGFinalizeDiagonal :936-969
  model cost thallo.t:3845-3865, accept/reject :1707-1753, guardedInvert (CERES) :638-648
  safeDivideIfNotLM :226-234, solver-parameter defaults :41-55.
J and F come from oracle.npdsl (NumPy dual numbers -> SciPy CSR), so JtF, diag(JtJ)
and JtJ p are sparse products; every vector is held in `dtype`, the dot products are
accumulated in float64 and rounded to `dtype` (the reference's own order is
atomics-nondeterministic on the GPU; see DESIGN.md "parity").

Pinned against the reference's golden outputs tests/minimal/gold.png and
tests/minimal_graph/gold.png (tests/test_oracle_golden.py) -- "parity pinned" for GN;
LM has no golden in the reference ("parity unpinned" for the LM-only blocks).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this.
"""
import math
import numpy as np

from .npdsl import NumpyL

DEFAULTS = dict(
    residual_reset_period=10, min_relative_decrease=1e-3, min_trust_region_radius=1e-32,
    max_trust_region_radius=1e16, q_tolerance=0.0001, function_tolerance=0.000001,
    trust_region_radius=1e4, radius_decrease_factor=2.0, min_lm_diagonal=1e-6,
    max_lm_diagonal=1e32, max_solver_time_in_seconds=0, nIterations=10, lIterations=10)


class OracleSolver:
    def __init__(self, define, dims, kind="gauss_newton", dtype=np.float32, mode="at_output",
                 materialized=False, define_kwargs=None, origin=None):
        self.origin = origin
        assert kind in ("gauss_newton", "levenberg_marquardt")   # thallo.t:74
        assert mode in ("at_output", "residualwise")
        self.define, self.dims, self.kind = define, list(dims), kind
        self.lm = kind == "levenberg_marquardt"
        self.dtype = np.dtype(dtype).type
        self.mode, self.materialized = mode, materialized
        self.kw = define_kwargs or {}
        self.p = dict(DEFAULTS)
        self.trace = []
        self.finalized = False
        # test hook: force_lin[i] = number of PCG iterations to run in nonlinear iteration i instead of
        # applying the zeta test (used to compare trajectories when a borderline zeta decision falls
        # differently in float32 on the GPU; every zeta is recorded in the trace either way)
        self.force_lin = None

    # ---- helpers
    def set(self, name, value):
        if name in self.p:
            self.p[name] = value

    def _dot(self, a, b):
        return self.dtype(np.dot(a.astype(np.float64), b.astype(np.float64)))

    def _eval(self, params):
        L = NumpyL(self.dims, params, self.dtype, self.origin)
        self.define(L, **self.kw)
        F, J, self.ranges = L.assemble()
        return L, F, J

    def _cost(self, params):
        _, F, _ = self._eval(params)
        return self.dtype(0.5 * np.dot(F.astype(np.float64), F.astype(np.float64)))

    def _unknown_views(self, L, params):
        views = []
        for im in L.unknowns:
            views.append(np.asarray(params[im.pidx]).reshape(-1))
        return views

    def _get_x(self, L, params):
        return np.concatenate([v.astype(self.dtype) for v in self._unknown_views(L, params)])

    def _set_x(self, L, params, x):
        o = 0
        for v in self._unknown_views(L, params):
            v[...] = x[o:o + v.size]
            o += v.size

    @staticmethod
    def _G(d, dt):
        one = dt(1)
        s = one + np.sqrt(d)
        return (one / (s * s)).astype(dt)

    def setup_vectors(self, params, radius=None):
        """The set-up of the first nonlinear iteration (PCGInit1 [+ the LM diagonal], the head of `step`) without the
        PCG loop: r = -J^T F, the preconditioner M, the LM diagonal C and the operator v -> (J^T J [+ C]) v.  Every
        quantity is local to an unknown and the residuals touching it, which is what lets a crop of a large problem be
        checked against the corresponding elements of the full-size GPU run (bench.py's parity record)."""
        dt, P = self.dtype, self.p
        L, F, J = self._eval(params)
        keep = (~L.exclude_mask()).astype(dt)
        usepre = L.usepreconditioner
        Jk = J.multiply(keep[None, :]).tocsr().astype(dt)
        JT = Jk.T.tocsr()
        r = (-(JT @ F)).astype(dt)
        dtrue = (L.diag_sq * keep).astype(dt)
        if self.mode == "at_output":
            M = self._G(dtrue if usepre else np.ones_like(dtrue), dt) * keep
        else:
            M = (self._G(dtrue, dt) if usepre else np.ones_like(dtrue)) * keep
        C = np.zeros_like(r)
        if self.lm:
            rho = dt(P["trust_region_radius"] if radius is None else radius)
            Ct = (dtrue / rho).astype(dt)
            with np.errstate(divide="ignore", invalid="ignore"):
                mult = ((dt(1) / M) / rho).astype(dt)              # SSq = M at the first nonlinear iteration
                C = np.minimum(np.maximum(Ct, dt(P["min_lm_diagonal"]) * mult), dt(P["max_lm_diagonal"]) * mult).astype(dt)
                M = (dt(1) / (C + rho * Ct)).astype(dt)
            C = np.where(keep > 0, C, 0).astype(dt)
            M = np.where(keep > 0, M, 0).astype(dt)

        def applyA(v):
            return (JT @ (Jk @ v) + C * v).astype(dt)
        return dict(r=r, M=M, C=C, applyA=applyA, cost=float(0.5 * np.dot(F.astype(np.float64), F.astype(np.float64))), L=L)

    # ---- API mirroring thallo.Plan {init, step, cost}
    def init(self, params):
        self.params = params
        self.nIter = 0
        self.finalized = False
        self.trace = []
        self.radius = self.dtype(self.p["trust_region_radius"])
        self.decrease_factor = self.dtype(self.p["radius_decrease_factor"])
        self.prevCost = self._cost(params)
        self.init_cost = self.prevCost

    def current_cost(self):
        if not self.finalized:
            self.prevCost = self._cost(self.params)
        return float(self.prevCost)

    def _finalize(self):
        self.prevCost = self._cost(self.params)
        self.finalized = True

    def step(self, params=None):
        params = self.params if params is None else params
        self.params = params
        dt, P = self.dtype, self.p
        if self.nIter >= int(P["nIterations"]):
            self._finalize()
            return 0
        L, F, J = self._eval(params)
        keep = (~L.exclude_mask()).astype(dt)
        usepre = L.usepreconditioner
        Jk = J.multiply(keep[None, :]).tocsr().astype(dt)      # excluded unknowns: columns drop out
        JT = Jk.T.tocsr()
        g = (JT @ F).astype(dt)
        dtrue = (L.diag_sq * keep).astype(dt)            # sum of squared partials per unknown access (thallo.t:3897-3901)
        r = (-g).astype(dt)
        if self.mode == "at_output":
            d = dtrue if usepre else np.ones_like(dtrue)
            M = self._G(d, dt) * keep
        else:
            M = (self._G(dtrue, dt) if usepre else np.ones_like(dtrue)) * keep
        p = (M * r).astype(dt)
        delta = np.zeros_like(r)
        aN = self._dot(r, p)
        it = dict(lin=[], nonlinear=self.nIter)
        C = None
        if self.lm:
            if self.nIter == 0:
                self.SSq = M.copy()
            rho = self.radius
            Ct = (dtrue / rho).astype(dt)
            with np.errstate(divide="ignore", invalid="ignore"):
                mult = ((dt(1) / self.SSq) / rho).astype(dt)
                lo = (dt(P["min_lm_diagonal"]) * mult).astype(dt)
                hi = (dt(P["max_lm_diagonal"]) * mult).astype(dt)
                C = np.minimum(np.maximum(Ct, lo), hi).astype(dt)
                M = (dt(1) / (C + rho * Ct)).astype(dt)
            C = np.where(keep > 0, C, 0).astype(dt)
            M = np.where(keep > 0, M, 0).astype(dt)
            b = r.copy()
            p = (M * r).astype(dt)
            aN = self._dot(r, p)
            Q0 = self._dot(dt(0.5) * delta, r + r)
        Mz = M if usepre else keep

        def applyA(v):
            out = (JT @ (Jk @ v)).astype(dt)
            if self.lm:
                out = (out + C * v).astype(dt)
            return out

        # LM residual reset (gauss_newton.t:1653-1660): computeAdelta exists only for groups that have an
        # applyJTJ function, i.e. the INLINE schedule (:1058-1065); groups whose J or J p is materialised
        # ([Jt][[J]p], Jt[Jp]) contribute nothing to A delta (SURVEY 8a quirk iii)
        rows_inline = np.ones(Jk.shape[0], dtype=dt)
        for g, (_, lo_, hi_) in zip(L.residuals.groups, self.ranges):
            if self.materialized or g.J.materialize or g.Jp.materialize:
                rows_inline[lo_:hi_] = 0

        def applyA_reset(v):
            out = (JT @ (rows_inline * (Jk @ v))).astype(dt)
            return (out + C * v).astype(dt)

        nlin = 0
        forced = None
        if self.force_lin is not None and self.nIter < len(self.force_lin):
            forced = int(self.force_lin[self.nIter])
        it["zeta"] = []
        for l in range(int(P["lIterations"])):
            Ap = applyA(p)
            aD = self._dot(p, Ap)
            with np.errstate(divide="ignore", invalid="ignore"):
                alpha = dt(aN / aD) if (self.lm or aD != 0) else dt(0)
            if self.lm and ((l + 1) % int(P["residual_reset_period"])) == 0:
                delta = (delta + alpha * p).astype(dt)
                Ad = applyA_reset(delta)
                r = (b - Ad).astype(dt)
            else:
                delta = (delta + alpha * p).astype(dt)
                r = (r - alpha * Ap).astype(dt)
            z = (Mz * r).astype(dt)
            bN = self._dot(z, r)
            q = self._dot(dt(0.5) * delta, r + b) if self.lm else dt(0)
            with np.errstate(divide="ignore", invalid="ignore"):
                beta = dt(bN / aN) if (self.lm or aN != 0) else dt(0)
            p = (z + beta * p).astype(dt)
            it["lin"].append(dict(alphaN=float(aN), alphaD=float(aD), betaN=float(bN), q=float(q)))
            aN = bN
            nlin += 1
            if self.lm:
                Q1 = q
                if not np.isfinite(Q1):
                    break
                with np.errstate(divide="ignore", invalid="ignore"):
                    zeta = dt(l + 1) * (Q1 - Q0) / Q1
                it["zeta"].append(float(zeta))
                if forced is not None:
                    if nlin >= forced:
                        break
                    Q0 = Q1
                    continue
                if not np.isfinite(zeta):
                    break
                if zeta < dt(P["q_tolerance"]):
                    break
                Q0 = Q1
        it["n_lin"] = nlin
        x = self._get_x(L, params)
        if self.lm:
            mres = (F + Jk @ delta).astype(np.float64)
            model_cost = dt(0.5 * np.dot(mres, mres))
            model_cost_change = dt(self.prevCost - model_cost)
            prevX = x.copy()
        self._set_x(L, params, (x + delta).astype(dt))
        ret = 1
        if self.lm:
            newCost = self._cost(params)
            cost_change = dt(self.prevCost - newCost)
            with np.errstate(divide="ignore", invalid="ignore"):
                rel = dt(cost_change / model_cost_change)
            it.update(cost=float(newCost), model_cost=float(model_cost), rel=float(rel))
            if cost_change >= 0 and rel > dt(P["min_relative_decrease"]):
                it["accepted"] = True
                if cost_change <= dt(self.prevCost * dt(P["function_tolerance"])):
                    self.trace.append(it)
                    self._finalize()
                    return 0
                tmp = 1.0 - math.pow(2.0 * float(rel) - 1.0, 3.0)
                self.radius = dt(float(self.radius) / max(1.0 / 3.0, tmp))
                self.radius = dt(min(float(self.radius), float(dt(P["max_trust_region_radius"]))))
                self.decrease_factor = dt(2.0)
                self.prevCost = newCost
            else:
                it["accepted"] = False
                self._set_x(L, params, prevX)
                self.radius = dt(self.radius / self.decrease_factor)
                self.decrease_factor = dt(2.0 * float(self.decrease_factor))
                if self.radius < dt(P["min_trust_region_radius"]):
                    self.p["trust_region_radius"] = 10e4       # gauss_newton.t:1741
                    self.trace.append(it)
                    self._finalize()
                    return 0
            self.p["trust_region_radius"] = float(self.radius)
            it["radius"] = float(self.radius)
        self.nIter += 1
        self.trace.append(it)
        return ret

    def solve(self, params):
        self.init(params)
        while self.step(params):
            pass
        return self.current_cost()
