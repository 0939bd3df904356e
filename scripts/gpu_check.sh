#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list and a full capture of the top kernel.
# usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh <tag> [kernel-regex]
TAG=${1:-r01}
KREGEX=${2:-"th_pcg_a|th_pcg_b"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 40 -c 6 -f -o $OUT/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT
