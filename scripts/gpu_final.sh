#!/bin/bash
# Round-end style pass on one GPU: parity suite per file, smoke(), bench line (+ reference arm), configured workloads,
# ncu launch list and full capture of the two PCG kernels.  usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-r01o}
QUICK=${2:-}     # "quick": skip the reference arm and the full ncu capture
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
bash scripts/gpu_tests.sh $TAG/tests
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
timeout 400 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cut -c1-400 $OUT/bench.json
if [ -z "$QUICK" ]; then
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm exit $?"
cut -c1-400 $OUT/bench_reference.json
fi
for w in "arap_mesh --size 2000" "volumetric --size 160" "bundle_adjustment --cameras 2000 --points 1000000" "optical_flow --size 8192" "sfs --size 4096"; do
  n=$(echo $w | cut -d' ' -f1)
  timeout 300 python scripts/bench_workloads.py $w > $OUT/wl_$n.json 2> $OUT/wl_$n.err; echo "$n exit $?"
  python - $OUT/wl_$n.json <<'PY'
import json, sys
try:
    b = json.load(open(sys.argv[1]))
    ks = {k: (round(v["avg_ms"], 5), round(v.get("frac_of_hbm_peak", 0), 3)) for k, v in b["kernels"].items() if v["share"] > 0.04}
    print("   it/s %8.1f ms/it %.4f cost %.6g %s" % (b["pcg_iterations_per_s"], b["linear_solve_ms_per_pcg_iteration"], b["final_cost"], ks))
except Exception as e:
    print("   FAILED", e)
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
if [ -z "$QUICK" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"th_pcg_a|th_pcg_b" -s 40 -c 4 -f -o $OUT/prof_${TAG}_iw \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_iw.log 2>&1
echo "ncu exit $?"
fi
