#!/bin/bash
# One GPU-box pass over everything the round measures: parity tests, the bench line, the other
# configured workloads, ncu launch list and full captures (with source counters) of the operator kernels.
# usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag> [quick]
TAG=${1:-r01}
QUICK=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$QUICK" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json
THALLO_B200_SIMPLIFY=0 timeout 300 python bench.py --no-cpu-baseline > $OUT/bench_nosimplify.json 2> $OUT/bench_nosimplify.err
cat $OUT/bench_nosimplify.json | cut -c1-200
for w in "volumetric --size 160" "arap_mesh --size 2000" "sfs --size 4096" "optical_flow --size 8192" "bundle_adjustment --cameras 2000 --points 1000000"; do
  n=$(echo $w | cut -d' ' -f1)
  timeout 400 python scripts/bench_workloads.py $w > $OUT/wl_$n.json 2> $OUT/wl_$n.err; echo "$n exit $?"
  cut -c1-400 $OUT/wl_$n.json
done
THALLO_B200_SIMPLIFY=0 timeout 300 python scripts/bench_workloads.py volumetric --size 160 > $OUT/wl_volumetric_nosimplify.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"th_pcg_a|th_pcg_b" -s 40 -c 6 -f -o $OUT/prof_${TAG}_iw \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_iw.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"th_pcg_a" -s 10 -c 2 -f -o $OUT/prof_${TAG}_vol \
    python scripts/bench_workloads.py volumetric --size 160 --steps 1 > $OUT/ncu_full_vol.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"th_gather_s0|th_step3" -s 10 -c 3 -f -o $OUT/prof_${TAG}_arap \
    python scripts/bench_workloads.py arap_mesh --size 2000 --steps 1 > $OUT/ncu_full_arap.log 2>&1
ls -la $OUT
