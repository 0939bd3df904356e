#!/bin/bash
# ncu --set full captures of the two weakest kernels (graph gather, 3-D tiled operator) + graph parity.
OUT=gpurun_out/${1:-prof2}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_graph.py -q --timeout 300 -rf > $OUT/test_gpu_graph.log 2>&1; echo "graph tests exit $?" | tee -a $OUT/test_gpu_graph.log
tail -3 $OUT/test_gpu_graph.log
timeout 300 python scripts/bench_workloads.py arap_mesh --size 2000 > $OUT/arap_own1.json 2> $OUT/arap_own1.err; cat $OUT/arap_own1.json; tail -2 $OUT/arap_own1.err
THALLO_B200_NVRTC_OPTS="-DTH_OWN_ENDPOINT=0" timeout 300 python scripts/bench_workloads.py arap_mesh --size 2000 > $OUT/arap_own0.json 2> $OUT/arap_own0.err; cat $OUT/arap_own0.json; tail -2 $OUT/arap_own0.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:th_gather_s0 -s 20 -c 2 -f -o $OUT/prof_arap_gather \
    python scripts/bench_workloads.py arap_mesh --size 2000 --steps 1 > $OUT/ncu_arap.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:th_pcg_a -s 20 -c 2 -f -o $OUT/prof_vol_pcg_a \
    python scripts/bench_workloads.py volumetric --size 160 --steps 1 > $OUT/ncu_vol.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"th_gather_s1|th_gather_s0|th_matj_g0" -s 12 -c 3 -f -o $OUT/prof_ba \
    python scripts/bench_workloads.py bundle_adjustment --cameras 2000 --points 1000000 --steps 1 > $OUT/ncu_ba.log 2>&1
ls -la $OUT
