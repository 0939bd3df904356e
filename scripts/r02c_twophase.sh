#!/bin/bash
# round 2, pass c: tile-operator variants (shifted-instance / two-phase, edge predicates always / interior specialisation):
# correctness tests, then shape_from_shading 8192^2 and the 160^3 volume under each variant
OUT=gpurun_out/r02c
mkdir -p $OUT
true
for cfg in 3b 4a; do
  for tp in 0 1; do
    for edge in 0 1; do
      if [ $edge = 1 ]; then unset THALLO_B200_EDGE_SPECIALIZE; else export THALLO_B200_EDGE_SPECIALIZE=1; fi
      THALLO_B200_TWO_PHASE=$tp timeout 600 python bench.py --config $cfg --no-parity --extra-steps 1 > $OUT/cfg${cfg}_tp${tp}_edge${edge}.json 2> $OUT/cfg${cfg}_tp${tp}_edge${edge}.err
      python - <<PY | tee -a $OUT/summary.txt
import json
try:
    l = json.loads(open("$OUT/cfg${cfg}_tp${tp}_edge${edge}.json").read().strip().splitlines()[-1])
    k = l["roofline"]["kernels"]
    print("config $cfg two_phase=$tp edge_always=$edge: it/s %.1f  th_pcg_a %.4f ms (frac %s)  th_pcg_b %.4f ms  cost %r" % (l["value"], k["th_pcg_a"]["avg_launch_ms"], k["th_pcg_a"].get("frac"), k["th_pcg_b"]["avg_launch_ms"], l["final_cost"]))
except Exception as e:
    print("config $cfg two_phase=$tp edge_always=$edge failed:", e)
PY
    done
  done
done
unset THALLO_B200_EDGE_SPECIALIZE
# headline with the interior specialisation (default) vs without
for edge in 0 1; do
  if [ $edge = 1 ]; then unset THALLO_B200_EDGE_SPECIALIZE; else export THALLO_B200_EDGE_SPECIALIZE=1; fi
  timeout 600 python bench.py --extras 3a --no-parity --no-cpu-baseline > $OUT/headline_edge${edge}.json 2> $OUT/headline_edge${edge}.err
  python - <<PY | tee -a $OUT/summary.txt
import json
try:
    l = json.loads(open("$OUT/headline_edge${edge}.json").read().strip().splitlines()[-1])
    k = l["roofline"]["kernels"]; k3 = l["configs"]["3a"]["roofline"]["kernels"]
    print("headline edge_always=$edge: it/s %.1f th_pcg_a %.4f ms th_pcg_b %.4f ms | 3a it/s %.1f th_pcg_a %.4f" % (l["value"], k["th_pcg_a"]["avg_launch_ms"], k["th_pcg_b"]["avg_launch_ms"], l["configs"]["3a"]["value"], k3["th_pcg_a"]["avg_launch_ms"]))
except Exception as e:
    print("headline edge_always=$edge failed:", e)
PY
done
tail -3 $OUT/*.err | tail -20
