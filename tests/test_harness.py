"""The example-harness writers (thallo_b200/harness.py) produce files that parse like the reference's
(results*.csv, finalCosts.json, perf.json: examples/shared/SolverIteration.h, CombinedSolverBase.h), and the profiled
solve follows the reference's call cadence (ThalloUtils.h:75-92)."""
import csv
import io
import json
import math

from thallo_b200 import harness as hz
from thallo_b200.api import PerformanceEntry, PerformanceSummary


class FakeSolver:
    def __init__(self, costs):
        self.costs, self.i, self.calls = costs, 0, []

    def init(self, params):
        self.calls.append("init")

    def step(self):
        self.calls.append("step")
        self.i += 1
        return 1 if self.i < len(self.costs) else 0

    def current_cost(self):
        self.calls.append("cost")
        return self.costs[min(self.i, len(self.costs) - 1)]


def test_profiled_solve_samples_after_init_and_after_every_step_that_continues():
    s = FakeSolver([10.0, 4.0, 1.0])
    syncs = []
    its = hz.launch_profiled_solve(s, [], synchronize=lambda: syncs.append(1))
    assert [i.cost for i in its] == [10.0, 4.0, 1.0] and all(i.timeInMS >= 0 for i in its)
    assert s.calls == ["init", "cost", "step", "cost", "step", "cost", "step"]        # the final Step (returns 0) is not sampled
    assert len(syncs) == 3


def test_results_csv_matches_the_reference_layout(tmp_path):
    gn = [hz.SolverIteration(10.0, 1.5), hz.SolverIteration(4.0, 2.5)]
    lm = [hz.SolverIteration(10.0, 1.0), hz.SolverIteration(5.0, 2.0), hz.SolverIteration(2.0, 3.0)]
    text = hz.save_solver_results(str(tmp_path) + "/", "_x", [], gn, lm, False)
    assert (tmp_path / "results_x.csv").read_text() == text
    rows = list(csv.reader(io.StringIO(text), skipinitialspace=True))
    assert rows[0] == ["Iter", "Ceres Error", "Thallo(GN) Error (float)", "Thallo(LM) Error (float)", "Ceres Iter Time(ms)",
                       "Thallo(GN) Iter Time(ms) (float)", "Thallo(LM) Iter Time(ms) (float)", "Total Ceres Time(ms)",
                       "Total Thallo(GN) Time(ms) (float)", "Total Thallo(LM) Time(ms) (float)"]
    vals = [[float(x) for x in r] for r in rows[1:]]
    assert len(vals) == 3
    assert vals[2][:4] == [2, 0.0, 4.0, 2.0]                  # shorter series repeat their last cost (clampedRead) ...
    assert vals[2][4:7] == [0.0, 0.0, 3.0]                    # ... with zero time
    assert vals[2][7:] == [0.0, 4.0, 6.0]                     # running totals
    assert rows[1][2] == "1.00000000000000000000e+01"                          # std::scientific, precision 20


def test_final_costs_and_perf_json_parse(tmp_path):
    text = hz.report_final_costs("image_warping", gn_cost=26.5, lm_cost=float("nan"), path=str(tmp_path / "finalCosts.json"))
    assert json.loads(text) == {"name": "image_warping", "costs": {"ThalloGN": 26.5}}
    s = PerformanceSummary()
    s.total = PerformanceEntry(1, 5.0, 5.0, 5.0, 0.0)
    s.linearSolve = PerformanceEntry(8, 0.1, 0.4, 0.25, float("nan"))
    text = hz.report_performance_statistics("image_warping", {"ThalloGN": s, "ThalloLM": s}, path=str(tmp_path / "perf.json"))
    j = json.loads(text)
    assert j["name"] == "image_warping" and j["autoscheduled"] == 0 and list(j["performance"]) == ["ThalloGN", "ThalloLM"]
    p = j["performance"]["ThalloLM"]
    assert list(p) == ["total", "nonlinearIteration", "nonlinearSetup", "linearSolve", "nonlinearResolve"]
    assert p["total"] == {"count": 1, "minMS": 5.0, "maxMS": 5.0, "meanMS": 5.0, "stddevMS": 0.0}
    assert p["linearSolve"]["count"] == 8 and p["linearSolve"]["stddevMS"] == 9999999999999999999999.0       # NaN sentinel of the reference
    assert math.isclose(p["linearSolve"]["meanMS"], 0.25)
