#!/bin/bash
# round 2, pass a: fused multi-GPU iteration (peer mailboxes + in-kernel pushes) -- parity and A/B against the NCCL path.
# usage (under gpurun --gpus N):  bash scripts/r02a_mg.sh N
N=${1:-2}
OUT=gpurun_out/r02a_mg$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/mgpu_check.py > $OUT/slab_parity.txt 2>&1; echo "slab parity rc=$?" | tee -a $OUT/summary.txt
timeout 600 $TR --master-port 29512 tests/mgpu_graph_check.py > $OUT/graph_parity.txt 2>&1; echo "graph parity rc=$?" | tee -a $OUT/summary.txt
THALLO_B200_MG_NCCL=1 timeout 600 $TR --master-port 29513 tests/mgpu_check.py > $OUT/slab_parity_nccl.txt 2>&1; echo "slab parity (nccl path) rc=$?" | tee -a $OUT/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench n1 rc=$?" | tee -a $OUT/summary.txt
timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_fused.json 2> $OUT/bench_fused.err; echo "bench fused rc=$?" | tee -a $OUT/summary.txt
THALLO_B200_MG_NCCL=1 timeout 600 $TR --master-port 29515 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_nccl.json 2> $OUT/bench_nccl.err; echo "bench nccl rc=$?" | tee -a $OUT/summary.txt
grep -h "mgpu" $OUT/*parity*.txt | tail -40
python - <<PY
import json
for n in ("bench_n1", "bench_fused", "bench_nccl"):
    try:
        l = json.loads(open("$OUT/%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value", round(l["value"], 1), "ms/step", round(l["ms_per_step"], 2), "its/step", l["pcg_iterations_per_step"], "cost", l["final_cost"],
              {k: (v["launches"], v["ms"]) for k, v in l["roofline"]["kernels"].items()})
    except Exception as e:
        print(n, "failed:", e)
PY
tail -5 $OUT/*.err
