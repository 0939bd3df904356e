#!/bin/bash
# SFS (computed arrays) parity + throughput; tiled path parity after the pipeline change; 3-D operator ncu capture.
OUT=gpurun_out/${1:-prof4}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tiled.py -q --timeout 600 -rf > $OUT/pytest_tiled.log 2>&1; echo "tiled tests exit $?" | tee -a $OUT/pytest_tiled.log
grep -E "^(FAILED|E  )|passed|failed" $OUT/pytest_tiled.log | head -30
timeout 300 python scripts/bench_workloads.py sfs --size 4096 --lit 10 --nit 4 > $OUT/sfs_4096.json 2> $OUT/sfs_4096.err; cat $OUT/sfs_4096.json; tail -3 $OUT/sfs_4096.err
timeout 300 python scripts/bench_workloads.py volumetric --size 160 > $OUT/vol.json 2> $OUT/vol.err; cut -c1-900 $OUT/vol.json; tail -2 $OUT/vol.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cut -c1-200 $OUT/bench.json; tail -3 $OUT/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:th_pcg_a -s 20 -c 1 -f -o $OUT/prof_vol_pcg_a \
    python scripts/bench_workloads.py volumetric --size 160 --steps 1 > $OUT/ncu_vol.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"th_pcg_a|th_pcg_b" -s 10 -c 2 -f -o $OUT/prof_sfs \
    python scripts/bench_workloads.py sfs --size 4096 --lit 10 --nit 2 --steps 1 > $OUT/ncu_sfs.log 2>&1
ls -la $OUT
