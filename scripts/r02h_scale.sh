#!/bin/bash
# round 2, pass h: the bench line at N GPUs (weak-scaling headline + strong-scaling records of the other configured workloads,
# each with its parity record against the single-GPU solve).  Strict timeouts, process groups killed as a whole.
N=${1:-8}
OUT=gpurun_out/r02h_n$N
mkdir -p $OUT
run_group() {
  local secs=$1 log=$2; shift 2
  setsid "$@" > $log 2>&1 &
  local pid=$! t=0
  while kill -0 $pid 2>/dev/null; do
    sleep 1; t=$((t+1))
    if [ $t -ge $secs ]; then echo "TIMEOUT after ${secs}s: killing group $pid" >> $log; kill -KILL -- -$pid 2>/dev/null; sleep 1; return 124; fi
  done
  wait $pid; return $?
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run_group 45 $OUT/debug_gn.txt $TR --master-port 29801 tests/mgpu_debug.py levenberg_marquardt; rc=$?
echo "debug LM (push kernel) rc=$rc" | tee -a $OUT/summary.txt
grep -hE "^\[rank 0|TIMEOUT|rror" $OUT/debug_gn.txt | tail -5
if [ $rc != 0 ]; then      # fall back to the pushes by the last CTA (validated at N = 2 and 8)
  export THALLO_B200_PUSH_KERNEL=0
  run_group 45 $OUT/debug_gn_fallback.txt $TR --master-port 29803 tests/mgpu_debug.py levenberg_marquardt; rc=$?
  echo "debug LM (last-CTA push) rc=$rc" | tee -a $OUT/summary.txt
  [ $rc = 0 ] || exit 1
fi
run_group ${2:-330} $OUT/bench.txt $TR --master-port 29802 bench.py --gpus $N --steps 5 --warmup 3 --budget-s ${3:-230}; rc=$?
echo "bench rc=$rc" | tee -a $OUT/summary.txt
python - <<PY | tee -a $OUT/summary.txt
import json
try:
    l = [json.loads(x) for x in open("$OUT/bench.txt").read().splitlines() if x.startswith("{")][-1]
    json.dump(l, open("$OUT/bench.json", "w"))
    print("headline N=%d value %.1f ms/step %.2f its/step %s parity %s wall %s" % (l["n_gpus"], l["value"], l["ms_per_step"], l["pcg_iterations_per_step"],
          {k: l.get("parity", {}).get(k) for k in ("max_rel", "pcg_counts_equal")}, l.get("wall_s")))
    print({k: round(1e3 * v["avg_launch_ms"], 1) for k, v in l["roofline"]["kernels"].items()})
    for k, c in l["configs"].items():
        if "value" not in c:
            print(k, c); continue
        p = c.get("parity", {})
        print(k, "it/s %.1f ms_to_converge %.2f lin ms/it %.4f" % (c["value"], c["ms_to_converge"], c["linear_solve_ms_per_pcg_iteration"]),
              "parity", {x: p.get(x) for x in ("max_rel", "pcg_counts_equal", "error")}, "wall", c["wall_s"])
        print("    ", {n: round(1e3 * v["avg_launch_ms"], 1) for n, v in c["roofline"]["kernels"].items()})
except Exception as e:
    print("failed", e)
PY
grep -hE "TIMEOUT|rror" $OUT/bench.txt | tail -5
