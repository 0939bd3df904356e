#!/bin/bash
# Tile padding / pipeline-depth variants of the tiled operator on the 3-D workload, parity of the tiled path, config-2 bench.
OUT=gpurun_out/${1:-prof3}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_parity.py -q --timeout 600 -rf -x > $OUT/pytest_tiled.log 2>&1; echo "tiled tests exit $?" | tee -a $OUT/pytest_tiled.log
tail -3 $OUT/pytest_tiled.log
for v in "1 1" "0 2" "0 1" "1 2"; do set -- $v
  THALLO_B200_TILE_PAD=$1 THALLO_B200_PIPE=$2 timeout 300 python scripts/bench_workloads.py volumetric --size 160 > $OUT/vol_pad$1_pipe$2.json 2> $OUT/vol_pad$1_pipe$2.err
  echo "vol pad=$1 pipe=$2: $(python -c "import json,sys; d=json.load(open('$OUT/vol_pad$1_pipe$2.json')); print(d['pcg_iterations_per_s'], d['kernels']['th_pcg_a'], d['final_cost'])")"; tail -2 $OUT/vol_pad$1_pipe$2.err
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cut -c1-400 $OUT/bench.json; tail -3 $OUT/bench.err
THALLO_B200_TILE_PAD=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_nopad.json 2> $OUT/bench_nopad.err; echo "bench nopad exit $?"; cut -c1-400 $OUT/bench_nopad.json
timeout 300 python scripts/bench_workloads.py optical_flow --size 4096 > $OUT/of.json 2> $OUT/of.err; cut -c1-300 $OUT/of.json
