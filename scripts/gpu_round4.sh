#!/bin/bash
# usage: bash scripts/gpu_round4.sh <tag>   -- parity suite per file + schedule comparison on arap_mesh
TAG=${1:-r01l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
bash scripts/gpu_tests.sh $TAG/tests
run() {
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python scripts/bench_workloads.py "$@" > $OUT/sw_$name.json 2> $OUT/sw_$name.err
  python - "$OUT/sw_$name.json" "$name" <<'PY'
import json, sys
try:
    b = json.load(open(sys.argv[1]))
    ks = {k: (v["avg_ms"], round(v.get("frac_of_hbm_peak", 0), 3)) for k, v in b["kernels"].items() if v["share"] > 0.03}
    print("%-28s it/s %8.1f ms/it %.4f cost %.6g %s" % (sys.argv[2], b["pcg_iterations_per_s"], b["linear_solve_ms_per_pcg_iteration"], b["final_cost"], ks))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run arap_inline -- arap_mesh --size 2000
run arap_jp -- arap_mesh --size 2000 --jp
run iw -- image_warping --size 2048 --nit 8 --lit 100
run ba -- bundle_adjustment --cameras 2000 --points 1000000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"th_gather_s0|th_applyj_g1" -s 10 -c 2 -f -o $OUT/prof_${TAG}_arapjp \
    python scripts/bench_workloads.py arap_mesh --size 2000 --jp --steps 1 > $OUT/ncu_full_arapjp.log 2>&1
echo "ncu exit $?"
