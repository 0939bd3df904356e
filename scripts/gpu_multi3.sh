#!/bin/bash
# 2-GPU parity only (graph + bundle adjustment + slab cases).  usage (under gpurun --gpus 2): bash scripts/gpu_multi3.sh <tag>
TAG=${1:-r01q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 tests/mgpu_graph_check.py > $OUT/mg_graph_parity.txt 2>&1; echo "graph parity exit $?"
grep -E "mgpu|Error|error|MISMATCH" $OUT/mg_graph_parity.txt | head -20
timeout 300 $TR --master-port 29511 tests/mgpu_check.py > $OUT/mg_slab_parity.txt 2>&1; echo "slab parity exit $?"
grep -E "^mgpu" $OUT/mg_slab_parity.txt
