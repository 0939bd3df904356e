"""Host-side mirror of what the reference's example harness writes around a solve, so that its comparison scripts and
plots keep working on this backend's numbers:

  launch_profiled_solve   examples/shared/ThalloUtils.h:75-92    Init / Step loop, one (cost, wall ms) sample per iteration
  save_solver_results     examples/shared/SolverIteration.h:30-68  results<suffix>.csv, the per-iteration comparison table
  report_final_costs      examples/shared/SolverIteration.h:70-90  finalCosts.json
  report_performance_statistics  examples/shared/CombinedSolverBase.h:10-34,63-91  perf.json from Thallo_GetPerformanceSummary

Numbers are printed like the C++ streams print them (std::scientific with 20 / 18 digits); the files parse as the same
CSV / JSON.  Multi-GPU: every rank may call these; by convention only rank 0 writes (pass write=False elsewhere)."""
import math
import time


class SolverIteration:                                   # SolverIteration.h:14-20
    def __init__(self, cost=-math.inf, time_in_ms=-math.inf):
        self.cost, self.timeInMS = cost, time_in_ms


def launch_profiled_solve(solver, params, synchronize=None):
    """The reference's profiled solve: Thallo_ProblemInit, then Thallo_ProblemStep until it returns 0, with a device
    synchronisation, a wall-clock sample and Thallo_ProblemCurrentCost after each call.  `solver` is a
    thallo_b200.api.ThalloSolver (or a partitioned wrapper of one); returns the list of SolverIteration."""
    if synchronize is None:
        import torch
        synchronize = torch.cuda.synchronize
    out = []
    t = time.perf_counter()
    solver.init(params)
    synchronize()
    ms = (time.perf_counter() - t) * 1000.0
    out.append(SolverIteration(solver.current_cost(), ms))
    t = time.perf_counter()
    while solver.step():
        synchronize()
        ms = (time.perf_counter() - t) * 1000.0
        out.append(SolverIteration(solver.current_cost(), ms))
        t = time.perf_counter()
    return out


def _sci(x, digits):
    if isinstance(x, float) and math.isinf(x):
        return "inf" if x > 0 else "-inf"
    return ("%." + str(digits) + "e") % x


def _clamped(v, i):
    return v[0] if i < 0 else v[-1] if i >= len(v) else v[i]


def save_solver_results(directory, suffix, ceres_iters, gn_iters, lm_iters, double_precision, write=True):
    """results<suffix>.csv (SolverIteration.h:30-68).  Returns the text."""
    col = " (double)" if double_precision else " (float)"
    lines = ["Iter, Ceres Error, Thallo(GN) Error%s,  Thallo(LM) Error%s, Ceres Iter Time(ms), Thallo(GN) Iter Time(ms)%s, "
             "Thallo(LM) Iter Time(ms)%s, Total Ceres Time(ms), Total Thallo(GN) Time(ms)%s, Total Thallo(LM) Time(ms)%s"
             % (col, col, col, col, col, col)]
    ceres = list(ceres_iters) or [SolverIteration(0, 0)]
    gn = list(gn_iters) or [SolverIteration(0, 0)]
    lm = list(lm_iters) or [SolverIteration(0, 0)]
    sc = sg = sl = 0.0
    for i in range(max(len(ceres), len(gn), len(lm))):
        tc = ceres[i].timeInMS if i < len(ceres) else 0.0
        tg = gn[i].timeInMS if i < len(gn) else 0.0
        tl = lm[i].timeInMS if i < len(lm) else 0.0
        sc, sg, sl = sc + tc, sg + tg, sl + tl
        vals = [_clamped(ceres, i).cost, _clamped(gn, i).cost, _clamped(lm, i).cost, tc, tg, tl, sc, sg, sl]
        lines.append("%d, " % i + ", ".join(_sci(float(v), 20) for v in vals))
    text = "\n".join(lines) + "\n"
    if write:
        with open(directory + "results" + suffix + ".csv", "w") as f:
            f.write(text)
    return text


def report_final_costs(name, gn_cost=None, lm_cost=None, path=None):
    """finalCosts.json (SolverIteration.h:70-90); costs that are None or NaN are left out.  Returns the text."""
    costs = [(k, v) for k, v in (("ThalloGN", gn_cost), ("ThalloLM", lm_cost)) if v is not None and not math.isnan(v)]
    lines = ['{  "name" : "%s",' % name, '  "costs" : {']
    for i, (k, v) in enumerate(costs):
        lines.append('    "%s" : %s%s' % (k, _sci(float(v), 20), "," if i != len(costs) - 1 else ""))
    lines += ["  }", "}"]
    text = "\n".join(lines) + "\n"
    if path:
        with open(path, "w") as f:
            f.write(text)
    return text


_ENTRIES = ("total", "nonlinearIteration", "nonlinearSetup", "linearSolve", "nonlinearResolve")


def _entry(name, e, ident, comma):
    g = lambda k: getattr(e, k) if not isinstance(e, dict) else e[k]
    num = lambda v: _sci(9999999999999999999999.0 if math.isnan(v) else float(v), 18)
    pad = ident + "  "
    return [ident + '"%s" : {' % name, pad + '"count" : %d,' % int(g("count")), pad + '"minMS" : %s,' % num(g("minMS")),
            pad + '"maxMS" : %s,' % num(g("maxMS")), pad + '"meanMS" : %s,' % num(g("meanMS")),
            pad + '"stddevMS" : %s' % num(g("stddevMS")), ident + "}" + ("," if comma else "")]


def report_performance_statistics(name, summaries, autoscheduled=0, path=None):
    """perf.json (CombinedSolverBase.h:63-91).  `summaries`: ordered mapping solver name ("ThalloGN", "ThalloLM") ->
    Thallo_PerformanceSummary (thallo_b200.api.PerformanceSummary, or dicts with the same fields).  Returns the text."""
    lines = ['{  "name" : "%s",' % name, '  "autoscheduled" : %d,' % int(autoscheduled), '  "performance" : {']
    items = list(summaries.items())
    for i, (solver, s) in enumerate(items):
        lines.append('    "%s" : {' % solver)
        for j, en in enumerate(_ENTRIES):
            e = getattr(s, en) if not isinstance(s, dict) else s[en]
            lines += _entry(en, e, "      ", j != len(_ENTRIES) - 1)
        lines.append("    }" + ("," if i != len(items) - 1 else ""))
    lines += ["  }", "}"]
    text = "\n".join(lines) + "\n"
    if path:
        with open(path, "w") as f:
            f.write(text)
    return text
