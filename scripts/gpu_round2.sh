#!/bin/bash
# Budget-conscious GPU pass: parity suite per file, bench line, configured workloads, ncu launch list + full
# capture of the two PCG kernels, then a short tuning sweep.  usage: bash scripts/gpu_round2.sh <tag>
TAG=${1:-r01j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
bash scripts/gpu_tests.sh $TAG/tests
timeout 400 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cut -c1-600 $OUT/bench.json
for w in "arap_mesh --size 2000" "volumetric --size 160" "bundle_adjustment --cameras 2000 --points 1000000" "optical_flow --size 8192" "sfs --size 4096"; do
  n=$(echo $w | cut -d' ' -f1)
  timeout 300 python scripts/bench_workloads.py $w > $OUT/wl_$n.json 2> $OUT/wl_$n.err; echo "$n exit $?"
  cut -c1-300 $OUT/wl_$n.json
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"th_pcg_a|th_pcg_b" -s 40 -c 4 -f -o $OUT/prof_${TAG}_iw \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_iw.log 2>&1
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
A="arap_mesh --size 2000"
run arap_l1 THALLO_B200_GATHER_LANES=1 -- $A
run arap_l2 THALLO_B200_GATHER_LANES=2 -- $A
run arap_l4 THALLO_B200_GATHER_LANES=4 -- $A
V="volumetric --size 160"
run vol_884_p1m3 THALLO_B200_PIPE=1 THALLO_B200_MINB=3 -- $V
run vol_1684_p1m1 THALLO_B200_TILE=16,8,4 -- $V
I="image_warping --size 2048 --nit 8 --lit 100"
run iw_6404 THALLO_B200_TILE=64,4,1 -- $I
run iw_m3 THALLO_B200_MINB=3 -- $I
run iw_nosleep THALLO_B200_NVRTC_OPTS=-DTH_WAIT_SLEEP_NS=0 -- $I
ls $OUT | wc -l
