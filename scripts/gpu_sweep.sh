#!/bin/bash
# Tuning sweep on one GPU box: tile / pipeline / residency of the tiled operator, lanes per unknown of
# the gather operator, barrier-wait backoff.  usage: bash scripts/gpu_sweep.sh <tag>
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {  # name, env..., -- args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python scripts/bench_workloads.py "$@" > $OUT/$name.json 2> $OUT/$name.err
  python - "$OUT/$name.json" "$name" <<'PY'
import json, sys
try:
    b = json.load(open(sys.argv[1]))
    ks = {k: v["avg_ms"] for k, v in b["kernels"].items() if v["share"] > 0.04}
    print("%-28s it/s %8.1f ms/it %.4f cost %.6g %s" % (sys.argv[2], b["pcg_iterations_per_s"], b["linear_solve_ms_per_pcg_iteration"], b["final_cost"], ks))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
V="volumetric --size 160"
run vol_default -- $V
run vol_884_p1m3 THALLO_B200_PIPE=1 THALLO_B200_MINB=3 -- $V
run vol_884_p1m3_nosleep THALLO_B200_PIPE=1 THALLO_B200_MINB=3 THALLO_B200_NVRTC_OPTS=-DTH_WAIT_SLEEP_NS=0 -- $V
run vol_884_p1m2 THALLO_B200_PIPE=1 THALLO_B200_MINB=2 -- $V
run vol_888_p2m1 THALLO_B200_TILE=8,8,8 -- $V
run vol_1684_p1m1 THALLO_B200_TILE=16,8,4 -- $V
run vol_1644_p2 THALLO_B200_TILE=16,4,4 -- $V
A="arap_mesh --size 2000"
run arap_l1 THALLO_B200_GATHER_LANES=1 -- $A
run arap_l2 THALLO_B200_GATHER_LANES=2 -- $A
run arap_l4 THALLO_B200_GATHER_LANES=4 -- $A
run arap_l8 THALLO_B200_GATHER_LANES=8 -- $A
run arap_rw_agg -- $A --schedule residualwise
run arap_rw_plain THALLO_B200_NVRTC_OPTS=-DTH_WARP_AGG=0 -- $A --schedule residualwise
B="bundle_adjustment --cameras 2000 --points 1000000"
run ba_l1 THALLO_B200_GATHER_LANES=1 -- $B
run ba_l2 THALLO_B200_GATHER_LANES=2 -- $B
run ba_l4 THALLO_B200_GATHER_LANES=4 -- $B
S="sfs --size 4096"
run sfs_default -- $S
run sfs_m2 THALLO_B200_MINB=2 -- $S
run sfs_6404 THALLO_B200_TILE=64,4,1 -- $S
run sfs_3216 THALLO_B200_TILE=32,16,1 THALLO_B200_MINB=1 -- $S
I="image_warping --size 2048 --nit 8 --lit 100"
run iw_default -- $I
run iw_nosleep THALLO_B200_NVRTC_OPTS=-DTH_WAIT_SLEEP_NS=0 -- $I
run iw_6404 THALLO_B200_TILE=64,4,1 -- $I
run iw_m3 THALLO_B200_MINB=3 -- $I
run iw_m4 THALLO_B200_MINB=4 -- $I
ls $OUT | wc -l
