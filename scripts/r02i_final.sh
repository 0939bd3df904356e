#!/bin/bash
# round 2, pass i (final single-GPU pass): whole GPU test-suite, smoke(), the bench line with every configured workload,
# the reference arm, the ncu launch list and --set full captures of the headline kernels and of shape_from_shading's operator
OUT=gpurun_out/r02i
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $OUT/gpu_tests_full.txt 2>&1; echo "tests rc=$?" | tee -a $OUT/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/gpu_tests_full.txt | tail -15 | tee -a $OUT/summary.txt
grep -A6 "parity tolerance audit" $OUT/gpu_tests_full.txt | head -12 | tee -a $OUT/summary.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt; tail -2 $OUT/smoke.log
timeout 450 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
python - <<PY | tee -a $OUT/summary.txt
import json
try:
    l = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    p = l.get("parity") or {}
    print("headline", round(l["value"], 1), "e2e", round(l["e2e"]["value"], 1), "frac", l["roofline"]["frac"], "iteration frac", round(l["roofline"]["pcg_iteration_frac"], 3),
          "cpu", l.get("cpu_baseline", {}).get("value"), "wall", l.get("wall_s"))
    print("  parity", {k: p.get(k) for k in ("max_rel", "pcg_counts_equal", "first_step_rel")}, "oracle f32-vs-f64", max(p.get("oracle_float32_vs_float64_accumulation_rel") or [0]))
    print(" ", {k: (round(v["avg_launch_ms"], 4), v.get("frac")) for k, v in l["roofline"]["kernels"].items()})
    for k, c in l["configs"].items():
        if "value" not in c:
            print(k, c); continue
        q = c.get("parity", {})
        print(k, "it/s %.1f ms_to_converge %.2f lin ms/it %.4f floor %.4f" % (c["value"], c["ms_to_converge"], c["linear_solve_ms_per_pcg_iteration"], c["survey_8d_floor_ms"]),
              "parity", {x: q.get(x) for x in ("max_rel", "alpha_rel", "error")}, "wall", c["wall_s"])
        print("    ", {n: (round(v["avg_launch_ms"], 4), v.get("frac")) for n, v in c["roofline"]["kernels"].items()})
except Exception as e:
    print("failed", e)
PY
export THALLO_B200_GRAPH=0
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --extras none > $OUT/ncu_launch_bench.log 2>&1; echo "ncu launch list rc=$?" | tee -a $OUT/summary.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"th_pcg_a|th_pcg_b" -s 40 -c 4 -f -o $OUT/prof_r02i_iw \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --extras none > $OUT/ncu_full_iw.log 2>&1; echo "ncu full (headline) rc=$?" | tee -a $OUT/summary.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"th_pcg_a" -s 12 -c 2 -f -o $OUT/prof_r02i_sfs \
    python bench.py --config 3b --extra-steps 1 --no-parity > $OUT/ncu_full_sfs.log 2>&1; echo "ncu full (sfs) rc=$?" | tee -a $OUT/summary.txt
unset THALLO_B200_GRAPH
timeout 150 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm rc=$?" | tee -a $OUT/summary.txt
cut -c1-300 $OUT/bench_reference.json
