#!/bin/bash
# round 2, pass f: the single-GPU test-suite after the graph / private-stream / two-phase / persistent-gather changes
OUT=gpurun_out/r02f
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $OUT/gpu_tests_full.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert " $OUT/gpu_tests_full.txt | head -60
grep -A6 "parity tolerance audit" $OUT/gpu_tests_full.txt | head -20
