#!/bin/bash
# 2-GPU pass: vertex-partitioned graph solve vs single GPU (parity), slab regression, weak-scaling bench of both.
# usage (under gpurun --gpus 2): bash scripts/gpu_multi2.sh <tag>
TAG=${1:-r01m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 tests/mgpu_graph_check.py > $OUT/mg_graph_parity.txt 2>&1; echo "graph parity exit $?"
grep -E "mgpu|Error|error|MISMATCH" $OUT/mg_graph_parity.txt | head -20
timeout 300 $TR --master-port 29511 tests/mgpu_check.py > $OUT/mg_slab_parity.txt 2>&1; echo "slab parity exit $?"
grep -E "^mgpu" $OUT/mg_slab_parity.txt
timeout 300 $TR --master-port 29513 tests/mgpu_graph_check.py --bench 2000 > $OUT/mg_graph_bench.txt 2>&1; echo "graph bench exit $?"
grep -E "^\{" $OUT/mg_graph_bench.txt
timeout 200 python scripts/bench_workloads.py arap_mesh --size 2000 > $OUT/arap_1gpu.json 2> $OUT/arap_1gpu.err
python -c "import json;b=json.load(open('$OUT/arap_1gpu.json'));print('1 GPU arap 2000x2000', b['pcg_iterations_per_s'], b['ms_per_solve'])"
timeout 400 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench exit $?"
cut -c1-300 $OUT/bench_n$N.json
