#!/bin/bash
# Short GPU pass: changed parity tests, then ncu --set full captures (with source counters) of the operator
# kernels of the non-headline workloads.  usage: bash scripts/gpu_round3.sh <tag>
TAG=${1:-r01k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for f in tests/test_gpu_tfile.py tests/test_gpu_warp.py tests/test_gpu_graph.py; do
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
    ks = {k: v["avg_ms"] for k, v in b["kernels"].items() if v["share"] > 0.04}
    print("%-28s it/s %8.1f ms/it %.4f cost %.6g %s" % (sys.argv[2], b["pcg_iterations_per_s"], b["linear_solve_ms_per_pcg_iteration"], b["final_cost"], ks))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run vol_default -- volumetric --size 160
run vol_nosleep THALLO_B200_NVRTC_OPTS=-DTH_WAIT_SLEEP_NS=0 -- volumetric --size 160
run arap_default -- arap_mesh --size 2000
run sfs_nosleep -- sfs --size 4096
run sfs_sleep THALLO_B200_NVRTC_OPTS=-DTH_WAIT_SLEEP_NS=64 -- sfs --size 4096
cap() {  # name kernel-regex skip count -- workload args
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4; shift
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o $OUT/prof_${TAG}_$name \
      python scripts/bench_workloads.py "$@" --steps 1 > $OUT/ncu_full_$name.log 2>&1
  echo "ncu $name exit $?"
}
cap vol "th_pcg_a" 10 1 -- volumetric --size 160
cap sfs "th_pcg_a" 10 1 -- sfs --size 4096
cap arap "th_gather_s0" 10 1 -- arap_mesh --size 2000
cap ba "th_gather_s0|th_gather_s1|th_matj_g0" 12 3 -- bundle_adjustment --cameras 2000 --points 1000000
ls -la $OUT | head -40
