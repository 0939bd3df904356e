#!/usr/bin/env python
"""Benchmark of the GN/LM + PCG hot path behind the Thallo C ABI (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S]

Workload (BASELINE.json configs[1]): examples/image_warping 2-D ARAP on a 2048x2048 pixel grid,
Levenberg-Marquardt, float32, nIterations=8, lIterations=100 (reference
examples/image_warping/src/main.cpp:131-134), synthetic inputs from thallo_b200.workloads.
A "step" is one whole Thallo_ProblemSolve of that problem from its initial state.
`value` = PCG (linear) iterations executed per second, inputs resident in HBM; `e2e` = the same
through Thallo_ProblemSolve with HOST buffers (pinned H2D of every input, D2H of the unknowns and
the final cost inside the timed region).  N > 1: one process per GPU (torchrun), ONE global problem of
2048 x (2048 N) pixels slab-partitioned along y over the N ranks (weak scaling: 2048x2048 owned pixels
per GPU): halo rows go over NVLink peer mappings, the PCG scalars over NCCL (thallo_b200/distributed.py).
`value` at N > 1 counts every global PCG iteration N times (it processes N 2048x2048 slabs), i.e. it is
the aggregate number of 2048x2048-slab PCG iterations per second.

`--impl reference` times the reference's CPU path: the plain-C restatement of its cpuOnly
simulator (oracle/iw_cpu.c; the Terra/Lua reference cannot be built in this image) on all host
cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pcg_iterations_per_second"
UNIT = "iter/s"
NIT, LIT = 8, 100

# Algorithmic bytes per pixel and launch of each kernel for image_warping/LM/at-output (DESIGN.md
# "Kernels and their algorithmic bytes"): U = 3 unknown scalars per pixel, A = 6 auxiliary scalars
# (Angle 1, UrShape 2, Mask 1, Constraints 2), 4 B each, every array counted once per launch.
U, A = 3, 7      # A: Mask 1, UrShape 2, Constraints 2, hoisted (sin, cos) 2
KERNEL_BYTES_PER_PX = {
    "th_pcg_a": 4 * (3 * U + A + 2 * U),      # read z, p_old, CtC, aux; write p_new, Ap
    "th_pcg_a_ld": 4 * (3 * U + A + 2 * U),
    "th_pcg_b": 4 * (6 * U + 3 * U),          # read delta, p, r, Ap, pre, b; write delta, r, z
    "th_step1_uw": 4 * (2 * U + 6 + U),       # untiled schedule: read p, CtC, aux (Angle instead of sin/cos); write Ap
    "th_step3": 4 * (2 * U + U),              # untiled schedule: read z, p; write p
}
PCG_ITERATION_BYTES_PER_PX = KERNEL_BYTES_PER_PX["th_pcg_a"] + KERNEL_BYTES_PER_PX["th_pcg_b"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=2048, help="image side (default: the configured 2048)")
    ap.add_argument("--cpu-sample-pcg", type=int, default=0, help="PCG iterations in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class Clocks:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference(size, sample_pcg, threads, steps=1, warmup=0):
    """Time the C restatement of the reference CPU path on a bounded sample: the first `sample_pcg`
    PCG iterations of the configured solve.  Returns (iters_per_s, info)."""
    import numpy as np  # noqa: F401
    from oracle import iw_cpu
    from thallo_b200 import workloads as wl
    L = iw_cpu.lib()
    if hasattr(L, "iw_set_threads"):
        L.iw_set_threads(int(threads))
    base = wl.image_warping_inputs(size, size)
    its, secs = 0, 0.0
    for s in range(warmup + steps):
        d = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in base.items()}
        r = iw_cpu.solve(size, size, d, "levenberg_marquardt", max_pcg=sample_pcg, nIterations=NIT, lIterations=LIT)
        if s >= warmup:
            its += r["n_pcg"]; secs += r["seconds_pcg"]
    return its / secs, dict(threads=r["threads"], n_pcg=its, seconds=secs)


def run_reference(a, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = a.cpu_sample_pcg or max(2, int(20 * (2048 / a.size) ** 2))
    t0 = time.time()
    v, info = cpu_reference(a.size, sample, cores, a.steps, a.warmup)
    ms_step = 1e3 * info["seconds"] / a.steps
    sample_txt = ("first %d PCG iterations of the image_warping %dx%d LM solve per step (PCG inner loops timed; "
                  "C restatement of the reference cpuOnly path, OpenMP over %d threads)" % (sample, a.size, a.size, info["threads"]))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(a, world),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["threads"], "kind": "port", "sample": sample_txt},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0}))


def workload_config(a, world):
    return {"workload": "examples/image_warping 2-D ARAP %dx%d, levenberg_marquardt, float32, nIterations=%d, lIterations=%d, "
                        "one Thallo_ProblemSolve per step" % (a.size, a.size * world, NIT, LIT),
            "unknowns": 3 * a.size * a.size * world, "schedule": "at_output",
            "parallelism": "single GPU" if world == 1 else
                           "one %dx%d problem slab-partitioned along y over %d GPUs (%dx%d owned pixels each): halo rows over NVLink "
                           "peer stores, PCG scalars over NCCL all-reduce; value = global PCG iterations/s x %d slabs"
                           % (a.size, a.size * world, world, a.size, a.size, world),
            "l2": "working set 13 solver vectors x %.0f MB + inputs, larger than the 126 MB L2; no flush needed" % (12.0 * a.size * a.size / 1e6)}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        return run_reference(a, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    from thallo_b200 import workloads as wl
    from thallo_b200.api import ThalloSolver
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; thallo_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    S = a.size
    d = wl.image_warping_inputs(S, S * world)
    gloo = dist.new_group(backend="gloo") if world > 1 else None
    part = None
    if world > 1:
        from thallo_b200.distributed import SlabSolver, slab_partition, local_slab, stencil_halo
        halo = stencil_halo("image_warping", [S, S * world], "levenberg_marquardt")
        part = slab_partition(S * world, world, halo)[rank]
        for k in ("Offset", "Angle", "UrShape", "Constraints", "Mask"):
            d[k] = local_slab(d[k], S, part)
    host = [torch.from_numpy(np.ascontiguousarray(d[k])).pin_memory() for k in ("Offset", "Angle", "UrShape", "Constraints", "Mask")]
    pristine = [h.cuda() for h in host]
    work = [p.clone() for p in pristine]
    out_host = [torch.empty_like(host[0]).pin_memory(), torch.empty_like(host[1]).pin_memory()]
    scal = [np.array([d["w_fitSqrt"]], np.float32), np.array([d["w_regSqrt"]], np.float32)]
    h2d = sum(h.numel() * h.element_size() for h in host)
    d2h = sum(h.numel() * h.element_size() for h in out_host) + 8

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make(timing):
        if world > 1:
            s = SlabSolver([S, S * world], "image_warping", "levenberg_marquardt", rank, world, group=gloo, timing=timing)
        else:
            s = ThalloSolver([S, S], "image_warping", "levenberg_marquardt", timing=timing)
        s.set_parameters(nIterations=NIT, lIterations=LIT)
        return s

    def step_resident(s):
        work[0].copy_(pristine[0]); work[1].copy_(pristine[1])       # reset the unknowns (device to device, timed)
        # ... and the trust-region radius, which a solve leaves behind in the solver parameters like the reference does
        # (gauss_newton.t:1751): every step is then the same solve, with the same PCG iteration counts
        return s.solve(work + scal, trust_region_radius=1e4)

    def step_e2e(s):
        for w, h in zip(work, host):
            w.copy_(h, non_blocking=True)                           # every input from pinned host memory
        c = s.solve(work + scal, trust_region_radius=1e4)
        out_host[0].copy_(work[0], non_blocking=True); out_host[1].copy_(work[1], non_blocking=True)
        torch.cuda.synchronize()
        return c

    def timed(s, fn):
        for _ in range(a.warmup):
            fn(s)
        barrier()
        it0, l0 = s.total_linear_iterations(), s.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk = Clocks(local)
        e0.record()
        for _ in range(a.steps):
            cost = fn(s)
        e1.record()
        barrier()
        clocks = clk.stop()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        its = torch.tensor([s.total_linear_iterations() - it0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(its, op=dist.ReduceOp.SUM)
        return dict(ms=float(t.item()), iters=float(its.item()), launches=s.launches() - l0, cost=cost, clocks=clocks)

    solver = make(1)
    res = timed(solver, step_resident)
    e2e = timed(solver, step_e2e)

    # per-kernel device times with event pairs around every launch (timingLevel 2), same K steps
    prof = make(2)
    for _ in range(max(1, a.warmup - 2)):
        step_resident(prof)
    prof.kernel_times()
    k0 = prof.kernel_times()
    it0 = prof.total_linear_iterations()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(a.steps):
        step_resident(prof)
    pe1.record()
    torch.cuda.synchronize()
    k1 = prof.kernel_times()
    prof_ms = pe0.elapsed_time(pe1)
    kern = {n: (k1[n][0] - k0.get(n, (0, 0.0))[0], k1[n][1] - k0.get(n, (0, 0.0))[1]) for n in k1}
    kern = {n: v for n, v in kern.items() if v[0] > 0}
    ktotal = sum(v[1] for v in kern.values())
    dom = max(kern, key=lambda n: kern[n][1])
    peak, peak_src = peaks()
    px = host[1].numel()       # pixels this rank's kernels sweep (owned + ghost rows)
    bpl = KERNEL_BYTES_PER_PX.get(dom, 0) * px
    avg_ms = kern[dom][1] / kern[dom][0]
    achieved = bpl / (avg_ms * 1e-3) / 1e9 if bpl else None
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic.json")       # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bpl, "avg_launch_ms": avg_ms,
                "launches_timed": kern[dom][0], "share_of_kernel_time": kern[dom][1] / ktotal,
                "kernels": {n: {"launches": v[0], "ms": round(v[1], 4), "share": round(v[1] / ktotal, 4)} for n, v in sorted(kern.items())},
                "profiled_pass_ms_per_step": prof_ms / a.steps,
                "pcg_iteration_bytes": PCG_ITERATION_BYTES_PER_PX * px,
                "pcg_iteration_gbs": PCG_ITERATION_BYTES_PER_PX * px * res["iters"] / world / (res["ms"] * 1e-3) / 1e9}

    line = {"metric": METRIC, "value": res["iters"] / (res["ms"] * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": res["ms"] / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a, world),
            "ms_to_converge": res["ms"] / a.steps, "pcg_iterations_per_step": res["iters"] / a.steps / world,
            "final_cost": res["cost"],
            "e2e": {"value": e2e["iters"] / (e2e["ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e["ms"] / a.steps},
            "gpu_launches": res["launches"], "clocks": res["clocks"], "roofline": roofline}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        sample = a.cpu_sample_pcg or max(2, int(5 * (2048 / S) ** 2))
        v, info = cpu_reference(S, sample, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": info["threads"], "kind": "port",
                                "sample": "first %d PCG iterations of the same %dx%d LM solve (PCG inner loops timed), plain-C restatement "
                                          "of the reference's single-threaded cpuOnly path (oracle/iw_cpu.c), 1 thread" % (sample, S, S)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
