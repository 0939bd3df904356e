#!/usr/bin/env python
"""Benchmark of the GN/LM + PCG hot path behind the Thallo C ABI (contract: DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2] [--extras all|none|3a,4b,...]

Headline (the JSON line's metric / value / e2e / roofline / cpu_baseline): BASELINE.json configs[1], examples/image_warping
2-D ARAP on a 2048x2048 pixel grid, Levenberg-Marquardt, float32, nIterations=8, lIterations=100 (reference
examples/image_warping/src/main.cpp:131-134), synthetic inputs.  A "step" is one whole Thallo_ProblemSolve of that
problem from its initial state.  `value` = PCG (linear) iterations executed per second, inputs resident in HBM; `e2e` =
the same through Thallo_ProblemSolve with HOST buffers (pinned H2D of every input, D2H of the unknowns inside the timed
region).  N > 1: one process per GPU (torchrun), ONE global problem of 2048 x (2048 N) pixels slab-partitioned along y
(weak scaling: 2048x2048 owned pixels per GPU); `value` counts every global PCG iteration N times (N slabs).

The same line carries, under "configs", one record per OTHER configured workload at its configured size (1: minimal
256^2; 3a: optical_flow 8192^2; 3b: shape_from_shading 8192^2; 4a: volumetric 160^3; 4b: arap_mesh 4 M vertices; 5:
bundle adjustment 10 k x 5 M x 25 M): PCG iterations/s, ms to converge, per-kernel roofline, and a `parity` record.  At
N > 1 these are STRONG-scaling runs of the named problem partitioned over the N GPUs (slabs / vertex ranges / point
blocks with replicated cameras); the driver can form T1 / (N TN) from the per-N lines.

`parity` (checker legs, outside every timed region; oracle/ is used only here and in the CPU baseline):
  config 2, N = 1   every cost and PCG count of the full 2048^2 LM solve against oracle/iw_cpu.c (float64 accumulation)
  others,  N = 1    r0, preconditioner, A p0 of the full-size run against the float64 oracle on three crops, and the
                    solver's alpha against a float64 recomputation from its own full-size vectors (oracle/fullsize.py)
  N > 1             every cost and PCG count of the partitioned solve against the single-GPU solve of the same problem

`--impl reference` times the reference's CPU path: the plain-C restatement of its cpuOnly simulator (oracle/iw_cpu.c;
the Terra/Lua reference cannot be built in this image) on all host cores, on a bounded sample of the headline workload.
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
EXTRA_KEYS = ["1", "3a", "4b", "5", "3b", "4a"]      # in this order; a wall-clock budget may cut the tail


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2", help="headline workload (default: 2, the configured image_warping 2048^2)")
    ap.add_argument("--extras", default="all", help="other configured workloads measured into the line's `configs`: all | none | 3a,4b,...")
    ap.add_argument("--extra-steps", type=int, default=2, help="timed solves per extra workload")
    ap.add_argument("--budget-s", type=float, default=480.0, help="wall-clock budget after which remaining extras are skipped")
    ap.add_argument("--size", type=int, default=2048, help="headline image side (default: the configured 2048)")
    ap.add_argument("--cpu-sample-pcg", type=int, default=0, help="PCG iterations in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
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


# ---------------------------------------------------------------------------------------------------- CPU legs (oracle/)
def cpu_reference(size, sample_pcg, threads, steps=1, warmup=0, nit=8, lit=100):
    """Time the C restatement of the reference CPU path on a bounded sample: the first `sample_pcg`
    PCG iterations of the configured solve.  Returns (iters_per_s, info)."""
    from oracle import iw_cpu
    from thallo_b200 import workloads as wl
    L = iw_cpu.lib()
    if hasattr(L, "iw_set_threads"):
        L.iw_set_threads(int(threads))
    base = wl.image_warping_inputs(size, size)
    its, secs = 0, 0.0
    for s in range(warmup + steps):
        d = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in base.items()}
        r = iw_cpu.solve(size, size, d, "levenberg_marquardt", max_pcg=sample_pcg, nIterations=nit, lIterations=lit)
        if s >= warmup:
            its += r["n_pcg"]; secs += r["seconds_pcg"]
    return its / secs, dict(threads=r["threads"], n_pcg=its, seconds=secs)


def headline_config(a, world, case):
    S = a.size
    return {"workload": "examples/image_warping 2-D ARAP %dx%d, levenberg_marquardt, float32, nIterations=%d, lIterations=%d, "
                        "one Thallo_ProblemSolve per step" % (S, S * world, case.nit, case.lit),
            "unknowns": 3 * S * S * world, "schedule": "at_output",
            "lm_mode": "Levenberg-Marquardt as written in gauss_newton.t (the reference snapshot runs GN for this kind string, "
                       "thallo.t:463; THALLO_LM_AS_COMMITTED=1 reproduces that)",
            "parallelism": "single GPU" if world == 1 else
                           "one %dx%d problem slab-partitioned along y over %d GPUs (%dx%d owned pixels each, weak scaling): boundary rows "
                           "stored into the neighbours' memory over NVLink by the kernel that computes them, PCG scalars all-reduced inside "
                           "the producing kernels over peer mailboxes; value = global PCG iterations/s x %d slabs"
                           % (S, S * world, world, S, S, world),
            "l2": "working set 13 solver vectors x %.0f MB + inputs, larger than the 126 MB L2; no flush needed" % (12.0 * S * S / 1e6)}


def run_reference(a, rank, world):
    if rank != 0:
        return
    from thallo_b200 import configs
    case = configs.case("2", dims=(a.size, a.size))
    cores = os.cpu_count() or 1
    sample = a.cpu_sample_pcg or max(2, int(20 * (2048 / a.size) ** 2))
    t0 = time.time()
    v, info = cpu_reference(a.size, sample, cores, a.steps, a.warmup, case.nit, case.lit)
    ms_step = 1e3 * info["seconds"] / a.steps
    sample_txt = ("first %d PCG iterations of the image_warping %dx%d LM solve per step (PCG inner loops timed; C restatement of the "
                  "reference cpuOnly path, OpenMP over %d threads -- the reference itself is single-threaded; at N > 1 this is still ONE "
                  "%dx%d slab on rank 0's host cores)" % (sample, a.size, a.size, info["threads"], a.size, a.size))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": headline_config(a, world, case),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["threads"], "kind": "port", "sample": sample_txt},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0}))


# ---------------------------------------------------------------------------------------------------- measurement helpers
class Ctx:
    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.a, self.torch, self.dist = a, torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; thallo_b200 has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.gloo = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.gloo = dist.new_group(backend="gloo")
        self.peak, self.peak_src = peaks()
        self.t0 = time.time()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, solver, fn, steps, warmup):
        """W warm-up calls, then exactly K calls between barriers, device-timed, max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        it0, l0 = solver.total_linear_iterations(), solver.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk = Clocks(self.local)
        e0.record()
        cost = None
        for _ in range(steps):
            cost = fn()
        e1.record()
        self.barrier()
        clocks = clk.stop()
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        return dict(ms=ms, iters=float(solver.total_linear_iterations() - it0), launches=solver.launches() - l0, cost=cost, clocks=clocks)

    def kernel_profile(self, solver, fn, steps, warmup):
        """Per-kernel device times from the library's own event pairs (timingLevel 2, util.t:774-790)."""
        torch = self.torch
        for _ in range(max(1, warmup)):
            fn()
        k0 = solver.kernel_times()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        k1 = solver.kernel_times()
        kern = {n: (k1[n][0] - k0.get(n, (0, 0.0))[0], k1[n][1] - k0.get(n, (0, 0.0))[1]) for n in k1}
        return {n: v for n, v in kern.items() if v[0] > 0}, e0.elapsed_time(e1)

    def roofline(self, kern, bytes_per_launch, prof_ms, steps):
        ktotal = sum(v[1] for v in kern.values()) or 1.0
        table = {}
        for n, v in sorted(kern.items()):
            e = {"launches": v[0], "ms": round(v[1], 4), "share": round(v[1] / ktotal, 4), "avg_launch_ms": v[1] / v[0]}
            if n in bytes_per_launch:
                g = bytes_per_launch[n] / (v[1] / v[0] * 1e-3) / 1e9
                e.update(algorithmic_bytes_per_launch=bytes_per_launch[n], achieved_gbs=round(g, 1), frac=round(g / self.peak, 4))
            table[n] = e
        dom = max(kern, key=lambda n: kern[n][1])
        d = table[dom]
        return {"bound": "hbm", "kernel": dom, "achieved": d.get("achieved_gbs"), "peak": self.peak, "unit": "GB/s",
                "frac": d.get("frac"), "traffic": None, "peak_source": self.peak_src,
                "algorithmic_bytes_per_launch": d.get("algorithmic_bytes_per_launch"), "avg_launch_ms": d["avg_launch_ms"],
                "launches_timed": d["launches"], "share_of_kernel_time": d["share"], "kernels": table,
                "profiled_pass_ms_per_step": prof_ms / max(1, steps)}


def trajectory(solver, params, reset=None):
    if reset:
        solver.set_parameters(**reset)
    solver.init(params)
    costs, lin = [solver.current_cost()], []
    while solver.step():
        costs.append(solver.current_cost())
        lin.append(solver.last_linear_iterations())
    costs.append(solver.current_cost())
    return costs, lin


def noise_floor(cx, one, reset, c1):
    """How far float32 rounding alone moves this trajectory: the single-GPU solve again with every unknown moved by one
    ulp.  Long truncated-PCG trajectories far from convergence amplify rounding differences (shape_from_shading's 60 x 10
    Gauss-Newton run moves by 5e-3 between two builds of the same operator that differ only in how the compiler
    contracted multiply-adds, profiles/r02g_tile_operator_variants.txt), so a partitioned solve -- whose dot products are
    summed in a different order -- is compared with the single-GPU one against this floor."""
    torch = cx.torch
    p = one.fresh()
    for i in one.unknown_slots:
        p[i].copy_(torch.nextafter(p[i], torch.full_like(p[i], float("inf"))))
    c2, _ = trajectory(one.solver, p, reset)
    n = min(len(c1), len(c2))
    return max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(c2[:n], c1[:n])) if n else None


def with_noise(rec, floor):
    rec["float32_noise_floor_rel"] = floor
    rec["within_1e-5_or_8x_noise_floor"] = bool(rec["max_rel"] <= max(1e-5, 8.0 * (floor or 0.0)))
    return rec


def compare_trajectories(c, l, cref, lref, against):
    n = min(len(c), len(cref))
    rel = max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(c[:n], cref[:n])) if n else float("inf")
    return {"against": against, "max_rel": rel, "costs_compared": n, "same_length": len(c) == len(cref),
            "pcg_counts": l, "pcg_counts_reference": lref, "pcg_counts_equal": list(l) == list(lref),
            "final_cost": c[-1] if c else None, "final_cost_reference": cref[-1] if cref else None}


# ---------------------------------------------------------------------------------------------------- headline: config 2
def run_headline(cx, a):
    torch, world, rank = cx.torch, cx.world, cx.rank
    from thallo_b200 import configs
    S = a.size
    case = configs.case("2", dims=(S, S * world))
    b = case.build(rank, world, "cuda", group=cx.gloo, timing=1)
    solver = b.solver
    tens = [i for i, p in enumerate(b.params) if hasattr(p, "numel")]
    host = [b.params[i].cpu().pin_memory() for i in tens]
    out_host = [torch.empty_like(host[tens.index(i)]).pin_memory() for i in b.unknown_slots]
    h2d = sum(h.numel() * h.element_size() for h in host)
    d2h = sum(h.numel() * h.element_size() for h in out_host) + 8
    reset = dict(trust_region_radius=1e4)      # a solve leaves its last radius in the solver parameters like the reference (gauss_newton.t:1751)

    def step_resident():
        return solver.solve(b.fresh(), **reset)

    def step_e2e():
        for i, h in zip(tens, host):
            b.params[i].copy_(h, non_blocking=True)                 # every input from pinned host memory
        c = solver.solve(b.params, **reset)
        for i, o in zip(b.unknown_slots, out_host):
            o.copy_(b.params[i], non_blocking=True)
        torch.cuda.synchronize()
        return c

    res = cx.timed(solver, step_resident, a.steps, a.warmup)
    e2e = cx.timed(solver, step_e2e, a.steps, a.warmup)
    costs, lin = trajectory(solver, b.fresh(), reset)
    solver.close()

    prof = case.build(rank, world, "cuda", group=cx.gloo, timing=2)
    kern, prof_ms = cx.kernel_profile(prof.solver, lambda: prof.solver.solve(prof.fresh(), **reset), a.steps, max(1, a.warmup - 2))
    roof = cx.roofline(kern, case.kernel_bytes(prof), prof_ms, a.steps)
    prof.solver.close()
    tf = os.path.join(ROOT, "profiles", "traffic.json")       # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tf):
        try:
            roof["traffic"] = json.load(open(tf)).get(roof["kernel"])
            roof["traffic_source"] = ("profiles/traffic.json (static: dram__bytes_read.sum + dram__bytes_write.sum of the committed "
                                      "ncu --set full capture, not measured in this run)")
        except Exception:
            pass
    px = b.local_elements
    it_bytes = sum(case.kernel_bytes(b).values())
    roof["pcg_iteration_bytes"] = it_bytes
    roof["pcg_iteration_gbs"] = it_bytes * res["iters"] / (res["ms"] * 1e-3) / 1e9
    roof["pcg_iteration_frac"] = roof["pcg_iteration_gbs"] / cx.peak
    roof["survey_8d_iteration_bytes"] = 204 * px

    line = {"metric": METRIC, "value": res["iters"] * world / (res["ms"] * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": res["ms"] / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": headline_config(a, world, case),
            "ms_to_converge": res["ms"] / a.steps, "pcg_iterations_per_step": res["iters"] / a.steps,
            "final_cost": res["cost"],
            "e2e": {"value": e2e["iters"] * world / (e2e["ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e["ms"] / a.steps},
            "gpu_launches": res["launches"], "clocks": res["clocks"], "roofline": roof}
    # ---- checker legs (outside the timed regions)
    if not a.no_parity:
        line["parity"] = headline_parity(cx, a, case, costs, lin, reset)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        sample = a.cpu_sample_pcg or max(2, int(5 * (2048 / S) ** 2))
        v, info = cpu_reference(S, sample, 1, nit=case.nit, lit=case.lit)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": info["threads"], "kind": "port",
                                "sample": "first %d PCG iterations of the same %dx%d LM solve (PCG inner loops timed), plain-C restatement "
                                          "of the reference's single-threaded cpuOnly path (oracle/iw_cpu.c), 1 thread" % (sample, S, S)}
    del b, prof
    torch.cuda.empty_cache()
    return line


def headline_parity(cx, a, case, costs, lin, reset):
    """N = 1: the full LM solve against oracle/iw_cpu.c (float64 accumulation, all host cores); N > 1: against the
    single-GPU solve of the same global problem, run on rank 0."""
    from thallo_b200 import configs, workloads as wl
    S, world = a.size, cx.world
    rec = None
    if world > 1:
        if cx.rank == 0:
            one = configs.case("2", dims=(S, S * world)).build(0, 1, "cuda", timing=1)
            c1, l1 = trajectory(one.solver, one.fresh(), reset)
            floor = noise_floor(cx, one, reset, c1)
            one.solver.close()
            rec = with_noise(compare_trajectories(costs, lin, c1, l1, "single-GPU solve of the same %dx%d problem (rank 0)" % (S, S * world)), floor)
            del one
            cx.torch.cuda.empty_cache()
        cx.barrier()
    elif cx.rank == 0:
        from oracle import iw_cpu
        d = wl.image_warping_inputs(S, S)
        t0 = time.time()
        r = iw_cpu.solve(S, S, d, "levenberg_marquardt", acc64=True, nIterations=case.nit, lIterations=case.lit)
        n = min(len(costs) - 1, len(r["costs"]))
        rec = compare_trajectories(costs[:n], lin, r["costs"][:n], r["n_lin"],
                                   "oracle/iw_cpu.c (plain-C restatement of the reference cpuOnly path, float64 accumulation), "
                                   "full %dx%d LM solve" % (S, S))
        # float32 noise floor of this truncated-PCG trajectory: the same restatement with float accumulation of the dot
        # products (the reference's own arithmetic) against its float64-accumulation twin
        r32 = iw_cpu.solve(S, S, wl.image_warping_inputs(S, S), "levenberg_marquardt", acc64=False, nIterations=case.nit, lIterations=case.lit)
        m = min(n, len(r32["costs"]))
        rec["per_cost_rel"] = [abs(a_ - b_) / abs(b_) for a_, b_ in zip(costs[:n], r["costs"][:n])]
        rec["oracle_float32_vs_float64_accumulation_rel"] = [abs(a_ - b_) / abs(b_) for a_, b_ in zip(r32["costs"][:m], r["costs"][:m])]
        rec["oracle_float32_pcg_counts"] = r32["n_lin"]
        rec["first_step_rel"] = rec["per_cost_rel"][1] if n > 1 else None
        rec["oracle_seconds"] = round(time.time() - t0, 1)
    return rec


# ---------------------------------------------------------------------------------------------------- the other configured workloads
def run_extra(cx, a, key):
    torch, world, rank = cx.torch, cx.world, cx.rank
    from thallo_b200 import configs
    case = configs.case(key)
    if world > 1 and case.partition is None:
        return {"workload": case.workload(), "skipped": "single-GPU latency configuration: not partitioned"}
    t_start = time.time()
    reset = dict(trust_region_radius=1e4)
    b = case.build(rank, world, "cuda", group=cx.gloo, timing=1)
    s = b.solver
    res = cx.timed(s, lambda: s.solve(b.fresh(), **reset), a.extra_steps, 1)
    costs, lin = (None, None)
    if world > 1 and not a.no_parity:
        costs, lin = trajectory(s, b.fresh(), reset)
    summ = s.summary()
    s.close()
    prof = case.build(rank, world, "cuda", group=cx.gloo, timing=2)
    kern, prof_ms = cx.kernel_profile(prof.solver, lambda: prof.solver.solve(prof.fresh(), **reset), 1, 1)
    kb = case.kernel_bytes(prof)
    roof = cx.roofline(kern, kb, prof_ms, 1)
    dims, desc = [int(x) for x in case.dims], prof.solver.lowered.desc
    prof.solver.close()
    its_per_solve = res["iters"] / a.extra_steps
    ms_it = res["ms"] / max(1.0, res["iters"])
    pcg_ms = sum(kern[n][1] / max(1.0, its_per_solve) for n in kern if n in kb)
    rec = {"workload": case.workload(), "dims": dims, "n_gpus": world, "scaling": "strong" if world > 1 else "single GPU",
           "partition": case.partition if world > 1 else None, "schedule": desc["schedule"],
           "value": res["iters"] / (res["ms"] * 1e-3), "unit": UNIT, "ms_to_converge": res["ms"] / a.extra_steps,
           "pcg_iterations_per_solve": its_per_solve, "ms_per_pcg_iteration_whole_solve": ms_it,
           "linear_solve_ms_per_pcg_iteration": summ.linearSolve.meanMS * summ.linearSolve.count / max(1.0, its_per_solve),
           "pcg_kernels_ms_per_iteration": pcg_ms,
           "survey_8d_iteration_bytes": case.survey_iteration_bytes(),
           "survey_8d_floor_ms": case.survey_iteration_bytes() / world / (cx.peak * 1e9) * 1e3,
           "final_cost": res["cost"], "steps": a.extra_steps, "gpu_launches": res["launches"], "clocks": res["clocks"], "roofline": roof,
           "local_elements": b.local_elements, "input_bytes_in_hbm": b.bytes_in_hbm()}
    if not a.no_parity:
        try:
            if world > 1:
                if rank == 0:
                    one = case.build(0, 1, "cuda", timing=1)
                    c1, l1 = trajectory(one.solver, one.fresh(), reset)
                    floor = noise_floor(cx, one, reset, c1)
                    one.solver.close()
                    rec["parity"] = with_noise(compare_trajectories(costs, lin, c1, l1, "single-GPU solve of the same problem (rank 0)"), floor)
                    del one
                cx.barrier()
            elif rank == 0:
                from oracle import fullsize
                one = case.build(0, 1, "cuda", timing=1)
                one.solver.close()
                t0 = time.time()
                p = fullsize.first_iteration_parity(lambda: case.make_solver(dims), one.fresh, case.energy, case.kind,
                                                    fullsize.crops_for(case, dims, desc), case.oracle_mode,
                                                    define_kwargs=case.define_kwargs, materialized=case.materialized,
                                                    solver_params=case.solver_params)
                p["against"] = ("float64 oracle (oracle/npdsl.py, oracle/solver.py) on three crops of the full-size problem: r0 = -J^T F, "
                                "preconditioner, A p0 (max |gpu - oracle| / max |oracle|); alpha against a float64 recomputation from "
                                "the solver's own full-size vectors")
                p["max_rel"] = max(p["operator_max_rel"], p["alpha_rel"])
                p["checker_seconds"] = round(time.time() - t0, 1)
                rec["parity"] = p
                del one
        except Exception as e:          # a checker failure must not lose the measurement
            if world > 1:
                raise
            rec["parity"] = {"error": repr(e)}
    del b, prof
    torch.cuda.empty_cache()
    rec["wall_s"] = round(time.time() - t_start, 1)
    return rec


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        return run_reference(a, rank, world)
    cx = Ctx(a)
    if a.config != "2":                   # one other workload as the only measurement (profiling runs)
        rec = run_extra(cx, a, a.config)
        if cx.rank == 0:
            print(json.dumps(dict(rec, metric=METRIC, higher_is_better=True, dtype="f32", data="synthetic")))
    else:
        line = run_headline(cx, a)
        keys = [] if a.extras == "none" else (EXTRA_KEYS if a.extras == "all" else [k for k in a.extras.split(",") if k])
        extras = {}
        for k in keys:
            over = cx.max_over_ranks(time.time() - cx.t0) > a.budget_s        # the same decision on every rank
            if over:
                extras[k] = {"skipped": "wall-clock budget of %.0f s exhausted" % a.budget_s}
                continue
            try:
                extras[k] = run_extra(cx, a, k)
            except Exception as e:
                if world > 1:
                    raise
                extras[k] = {"error": repr(e)}
                cx.torch.cuda.empty_cache()
        line["configs"] = extras
        line["wall_s"] = round(time.time() - cx.t0, 1)
        if cx.rank == 0:
            print(json.dumps(line))
    if cx.world > 1:
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
