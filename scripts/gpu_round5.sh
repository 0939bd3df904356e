#!/bin/bash
# usage: bash scripts/gpu_round5.sh <tag>  -- gathered PCGInit1: parity files of the gather schedule + before/after timings
TAG=${1:-r01p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for f in tests/test_gpu_graph.py tests/test_gpu_edge_cases.py tests/test_gpu_tfile.py tests/test_gpu_parity.py; do
    b=$(basename $f .py)
    timeout 400 python -m pytest $f -q --timeout 180 -rf > $OUT/$b.log 2>&1; echo "$b exit $?"
    grep -E "^(FAILED|E  )|passed|failed" $OUT/$b.log | head -30
done
run() {
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python scripts/bench_workloads.py "$@" > $OUT/sw_$name.json 2> $OUT/sw_$name.err
  python - "$OUT/sw_$name.json" "$name" <<'PY'
import json, sys
try:
    b = json.load(open(sys.argv[1]))
    ks = {k: round(v["avg_ms"], 5) for k, v in b["kernels"].items() if v["share"] > 0.02}
    print("%-22s it/s %8.1f ms/solve %.3f ms/it %.4f cost %.6g %s" % (sys.argv[2], b["pcg_iterations_per_s"], b["ms_per_solve"], b["linear_solve_ms_per_pcg_iteration"], b["final_cost"], ks))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
B="bundle_adjustment --cameras 2000 --points 1000000"
A="arap_mesh --size 2000"
run ba_gatherjtf -- $B
run ba_scatterjtf THALLO_B200_SCATTER_JTF=1 -- $B
run arap_gatherjtf -- $A
run arap_scatterjtf THALLO_B200_SCATTER_JTF=1 -- $A
run arap_unroll2 THALLO_B200_GATHER_UNROLL=2 -- $A
run arap_unroll3 THALLO_B200_GATHER_UNROLL=3 -- $A
run ba_unroll2 THALLO_B200_GATHER_UNROLL=2 -- $B
