#!/bin/bash
# Gather-schedule pass: graph / materialised-J parity tests, then throughput of configs 4b and 5.
OUT=gpurun_out/${1:-graph}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_graph.py -q --timeout 300 -rf > $OUT/test_gpu_graph.log 2>&1; echo "graph tests exit $?" | tee -a $OUT/test_gpu_graph.log
grep -E "^(FAILED|E  )|Error|passed|failed" $OUT/test_gpu_graph.log | head -40
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -rf --deselect tests/test_gpu_graph.py > $OUT/pytest_gpu.log 2>&1; echo "gpu suite exit $?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python scripts/bench_workloads.py arap_mesh --size 2000 > $OUT/arap_mesh_2000.json 2> $OUT/arap_mesh_2000.err; echo "arap exit $?"; cat $OUT/arap_mesh_2000.json; tail -3 $OUT/arap_mesh_2000.err
timeout 300 python scripts/bench_workloads.py arap_mesh --size 2000 --schedule residualwise > $OUT/arap_mesh_2000_rw.json 2> $OUT/arap_mesh_2000_rw.err; echo "arap rw exit $?"; cat $OUT/arap_mesh_2000_rw.json; tail -3 $OUT/arap_mesh_2000_rw.err
timeout 400 python scripts/bench_workloads.py bundle_adjustment --cameras 2000 --points 1000000 > $OUT/ba_1m.json 2> $OUT/ba_1m.err; echo "ba exit $?"; cat $OUT/ba_1m.json; tail -3 $OUT/ba_1m.err
timeout 400 python scripts/bench_workloads.py bundle_adjustment --cameras 2000 --points 1000000 --no-materialize > $OUT/ba_1m_free.json 2> $OUT/ba_1m_free.err; echo "ba free exit $?"; cat $OUT/ba_1m_free.json; tail -3 $OUT/ba_1m_free.err
