#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): slab-partition parity check, then the N-GPU bench line.
# usage: bash scripts/gpu_multi.sh <tag> <ngpus>
TAG=${1:-mg}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nvidia-smi topo -m >> $OUT/gpu.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tests/mgpu_check.py > $OUT/mgpu_check.log 2>&1; echo "mgpu_check exit $?" | tee -a $OUT/mgpu_check.log
grep -E "^mgpu|Error|error" $OUT/mgpu_check.log | head -20
NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench exit $?"
cat $OUT/bench_n$N.json; tail -5 $OUT/bench_n$N.err
