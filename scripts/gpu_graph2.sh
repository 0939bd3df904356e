#!/bin/bash
OUT=gpurun_out/${1:-graph2}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_graph.py -q --timeout 300 -rf > $OUT/test_gpu_graph.log 2>&1; echo "graph tests exit $?" | tee -a $OUT/test_gpu_graph.log
grep -E "^(FAILED|E  )|Error|passed|failed" $OUT/test_gpu_graph.log | head -40
timeout 400 python scripts/bench_workloads.py bundle_adjustment --cameras 2000 --points 1000000 > $OUT/ba_1m.json 2> $OUT/ba_1m.err; echo "ba exit $?"; cat $OUT/ba_1m.json; tail -3 $OUT/ba_1m.err
timeout 300 python scripts/bench_workloads.py optical_flow --size 4096 --nit 2 --lit 50 > $OUT/of_4096.json 2> $OUT/of_4096.err; echo "of exit $?"; cat $OUT/of_4096.json; tail -3 $OUT/of_4096.err
timeout 300 python scripts/bench_workloads.py volumetric --size 160 --nit 2 --lit 60 > $OUT/vol_160.json 2> $OUT/vol_160.err; echo "vol exit $?"; cat $OUT/vol_160.json; tail -3 $OUT/vol_160.err
