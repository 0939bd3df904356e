#!/bin/bash
# round 2, pass b: every configured workload at its configured size on one GPU (bench.py with extras), then the full-size parity tests
OUT=gpurun_out/r02b
mkdir -p $OUT
timeout 1500 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
python - <<PY | tee -a $OUT/summary.txt
import json
try:
    l = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("headline", round(l["value"], 1), "e2e", round(l["e2e"]["value"], 1), "parity", l.get("parity"))
    print({k: (round(v["avg_launch_ms"], 4), v.get("frac")) for k, v in l["roofline"]["kernels"].items()})
    for k, c in l["configs"].items():
        if "value" not in c:
            print(k, c); continue
        print(k, c["workload"][:60], "it/s", round(c["value"], 1), "ms/it", round(c["ms_per_pcg_iteration_whole_solve"], 4), "lin ms/it", round(c["linear_solve_ms_per_pcg_iteration"], 4),
              "floor", round(c["survey_8d_floor_ms"], 4), "wall", c["wall_s"])
        print("    ", {n: (round(v["avg_launch_ms"], 4), v.get("frac")) for n, v in c["roofline"]["kernels"].items()})
        p = c.get("parity", {})
        print("     parity", {x: p.get(x) for x in ("max_rel", "operator_max_rel", "alpha_rel", "error", "checker_seconds")}, p.get("crops"))
except Exception as e:
    print("failed", e)
PY
tail -3 $OUT/bench.err
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q -rA 2>&1 | tail -30 | tee $OUT/fullsize_tests.txt
