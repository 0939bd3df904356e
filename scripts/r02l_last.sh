#!/bin/bash
# round 2, pass l: double-buffered J p planes of the two-phase operator (correctness + A/B on shape_from_shading 8192^2),
# persistent-grid width of the gather kernel on the 4 M-vertex mesh
OUT=gpurun_out/r02l
mkdir -p $OUT
timeout 170 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_fullsize.py -m gpu -q -rA -p no:cacheprovider -k "variants or shading or interior or 3b" > $OUT/tests.txt 2>&1; echo "tests rc=$?" | tee -a $OUT/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed" $OUT/tests.txt | tail -8 | tee -a $OUT/summary.txt
for bufs in 2 1; do
  THALLO_B200_JP_BUFS=$bufs timeout 90 python bench.py --config 3b --no-parity --extra-steps 1 > $OUT/sfs_bufs$bufs.json 2> $OUT/sfs_bufs$bufs.err
  python - <<PY | tee -a $OUT/summary.txt
import json
try:
    l = json.loads(open("$OUT/sfs_bufs$bufs.json").read().strip().splitlines()[-1]); k = l["roofline"]["kernels"]
    print("shape_from_shading 8192^2 jp_bufs=$bufs: it/s %.1f th_pcg_a %.4f ms (frac %s) cost %r" % (l["value"], k["th_pcg_a"]["avg_launch_ms"], k["th_pcg_a"].get("frac"), l["final_cost"]))
except Exception as e:
    print("sfs bufs=$bufs failed", e)
PY
done
for b in 0 8 16; do
  THALLO_B200_GATHER_BLOCKS_PER_SM=$b timeout 60 python bench.py --config 4b --no-parity --extra-steps 1 > $OUT/arap_b$b.json 2> $OUT/arap_b$b.err
  python - <<PY | tee -a $OUT/summary.txt
import json
try:
    l = json.loads(open("$OUT/arap_b$b.json").read().strip().splitlines()[-1]); k = l["roofline"]["kernels"]
    print("arap_mesh 4M gather blocks/SM=$b: it/s %.1f th_gather_s0 %.4f ms (frac %s)" % (l["value"], k["th_gather_s0"]["avg_launch_ms"], k["th_gather_s0"].get("frac")))
except Exception as e:
    print("arap b=$b failed", e)
PY
done
