#!/bin/bash
# round 2, pass g: the GPU tests that pass f did not reach, then the tile-operator variant matrix (pass c)
OUT=gpurun_out/r02g
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_warp.py tests/test_reference_programs.py -m gpu -q -rA -p no:cacheprovider > $OUT/gpu_tests_rest.txt 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/gpu_tests_rest.txt | head -30
grep -A6 "parity tolerance audit" $OUT/gpu_tests_rest.txt | head -12
bash scripts/r02c_twophase.sh 2>&1 | grep -v "^=\|passed\|warnings" | tail -30
