#!/bin/bash
# Bench line + ncu launch list + full capture of the PCG kernels.  usage: bash scripts/gpu_bench.sh <tag> [kernel-regex] [extra bench args]
TAG=${1:-bench}
KREGEX=${2:-"th_pcg_a|th_pcg_b"}
shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 400 python bench.py "$@" > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > $OUT/ncu_launch_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s 40 -c 6 -f -o $OUT/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT
