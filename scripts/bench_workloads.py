#!/usr/bin/env python
"""Throughput of the other configured workloads (BASELINE.json configs 3-5; config 2 is bench.py):
PCG iterations/s through the C ABI, per-kernel device times (timingLevel 2) and achieved GB/s of
every PCG kernel against its algorithmic bytes (SURVEY.md 8d formulas, restated in DESIGN.md).

  python scripts/bench_workloads.py arap_mesh --size 2000 [--kind gauss_newton] [--nit 2 --lit 50]
  python scripts/bench_workloads.py bundle_adjustment --cameras 2000 --points 1000000
  python scripts/bench_workloads.py optical_flow --size 4096
  python scripts/bench_workloads.py volumetric --size 160
  python scripts/bench_workloads.py sfs --size 4096
Prints one JSON line per run.  Not the driver's bench contract (that is bench.py)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["arap_mesh", "bundle_adjustment", "optical_flow", "volumetric", "image_warping", "sfs"])
    ap.add_argument("--size", type=int, default=0)
    ap.add_argument("--cameras", type=int, default=2000)
    ap.add_argument("--points", type=int, default=1000000)
    ap.add_argument("--kind", default="")
    ap.add_argument("--schedule", default="auto")
    ap.add_argument("--nit", type=int, default=2)
    ap.add_argument("--lit", type=int, default=50)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--no-materialize", action="store_true")
    ap.add_argument("--jp", action="store_true", help="arap_mesh: schedule Jt[Jp] for the edge term (J p stored per edge)")
    ap.add_argument("--jp-all", action="store_true",
                    help="any workload: gather schedule with every residual group in the Jt[Jp] form (two-pass operator: J p per "
                         "residual stored, transposed partials gathered per unknown); not verified on a GPU in round 1")
    a = ap.parse_args()
    import numpy as np
    import torch
    from thallo_b200 import workloads as wl
    from thallo_b200.api import ThalloSolver
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    t0 = time.time()
    kw = {}
    if a.workload == "arap_mesh":
        n = a.size or 2000
        d = wl.arap_mesh_inputs(n, n)
        dims, energy, kind = [n * n, len(d["V0"])], "arap_mesh_deformation", a.kind or "gauss_newton"
        params = wl.arap_mesh_params(d)
        N, E, U, A = n * n, len(d["V0"]), 6, 9
        # gather schedule: index arrays (V0 offsets + V1 + permutation + V0 through it) + own reads + writes, see DESIGN.md
        bytes_iter = {"th_gather_s0": 4 * (E * 3 + N * (1 + U + A + U)), "th_pcg_b": 4 * N * 8 * U, "th_step3": 4 * N * 3 * U}
        if a.jp:    # J p (3 scalars per edge) written once by th_applyj_g1 and read once through each of the two endpoints
            kw = dict(define_kwargs=dict(jp=True))
            bytes_iter["th_applyj_g1"] = 4 * (E * (2 + 3) + N * (U + A))
            bytes_iter["th_gather_s0"] += 4 * E * 2 * 3
    elif a.workload == "bundle_adjustment":
        d = wl.bundle_adjustment_inputs(a.cameras, a.points, 5)
        O = len(d["oToC"])
        dims, energy, kind = [a.cameras, a.points, O], "bundle_adjustment", a.kind or "levenberg_marquardt"
        params = wl.bundle_adjustment_params(d)
        if a.no_materialize:
            kw = dict(define_kwargs=dict(materialize=False))
        nunk = 9 * a.cameras + 3 * a.points
        bytes_iter = {"th_matj_g0": 4 * O * (24 + 2 + 2), "th_gather_s0": 4 * O * (18 + 2 + 1), "th_gather_s1": 4 * (O * (6 + 2) + a.points * 10),
                      "th_pcg_b": 4 * nunk * 9, "th_step3": 4 * nunk * 3}
    elif a.workload == "optical_flow":
        n = a.size or 4096
        d = wl.optical_flow_inputs(n, n)
        dims, energy, kind = [n, n], "optical_flow", a.kind or "gauss_newton"
        params = wl.optical_flow_params(d)
        bytes_iter = {"th_pcg_a": 4 * n * n * (4 * 2 + 4), "th_pcg_b": 4 * n * n * 7 * 2}
    elif a.workload == "volumetric":
        n = a.size or 160
        d = wl.volumetric_inputs(n, n, n)
        dims, energy, kind = [n, n, n], "volumetric_mesh_deformation", a.kind or "gauss_newton"
        params = wl.volumetric_params(d)
        bytes_iter = {"th_pcg_a": 4 * n ** 3 * (4 * 6 + 9), "th_pcg_b": 4 * n ** 3 * 8 * 6}
    elif a.workload == "sfs":
        n = a.size or 4096
        d = wl.sfs_inputs(n, n)
        dims, energy, kind = [n, n], "shape_from_shading", a.kind or "gauss_newton"
        params = wl.sfs_params(d)
        # th_pcg_a: z, p_old, p_new, Ap (4) + gradient image (3) + validity image (1) + D_i (1) + two uint8 edge masks (0.5)
        bytes_iter = {"th_pcg_a": int(4 * n * n * 9.5), "th_pcg_b": 4 * n * n * 7}
    else:
        n = a.size or 2048
        d = wl.image_warping_inputs(n, n)
        dims, energy, kind = [n, n], "image_warping", a.kind or "levenberg_marquardt"
        params = wl.image_warping_params(d)
        bytes_iter = {"th_pcg_a": 88 * n * n, "th_pcg_b": 108 * n * n}
    gen_s = time.time() - t0
    isdev = [hasattr(p, "shape") and np.asarray(p).size > 1 for p in params]
    pristine = [dev(p) if f else p for p, f in zip(params, isdev)]

    def fresh():
        return [p.clone() if f else p for p, f in zip(pristine, isdev)]
    out = {"workload": a.workload, "energy": energy, "dims": dims, "kind": kind, "nIterations": a.nit, "lIterations": a.lit,
           "input_generation_s": round(gen_s, 1)}
    for timing in (1, 2):
        t1 = time.time()
        if a.jp_all:
            kw = dict(kw)
            kw["define_kwargs"] = dict(kw.get("define_kwargs") or {}, jp_all=True)
            a.schedule = "gather"
        s = ThalloSolver(dims, energy, kind, timing=timing, schedule=a.schedule, **kw)
        out["schedule"] = s.lowered.desc["schedule"]
        s.set_parameters(nIterations=a.nit, lIterations=a.lit)
        p = fresh()
        s.solve(p)                      # warm-up (JIT, adjacency)
        torch.cuda.synchronize()
        if timing == 1:
            out["plan_and_first_solve_s"] = round(time.time() - t1, 2)
        k0 = s.kernel_times() if timing == 2 else {}
        it0 = s.total_linear_iterations()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = 0.0
        for _ in range(a.steps):
            p = fresh()
            e0.record()
            cost = s.solve(p)
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        its = s.total_linear_iterations() - it0
        if timing == 1:
            sm = s.summary()
            out.update(pcg_iterations_per_s=its / (ms * 1e-3), ms_per_solve=ms / a.steps, pcg_iterations_per_solve=its / a.steps,
                       linear_solve_ms_per_pcg_iteration=sm.linearSolve.meanMS * sm.linearSolve.count / max(1, its / a.steps),
                       final_cost=cost)
        else:
            k1 = s.kernel_times()
            peak = 6551.0
            try:
                peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            except Exception:
                pass
            ks = {}
            tot = sum(k1[n][1] - k0.get(n, (0, 0.0))[1] for n in k1)
            for n in sorted(k1):
                c, t = k1[n][0] - k0.get(n, (0, 0.0))[0], k1[n][1] - k0.get(n, (0, 0.0))[1]
                if c <= 0:
                    continue
                e = {"launches": c, "avg_ms": round(t / c, 5), "share": round(t / tot, 4)}
                if n in bytes_iter:
                    g = bytes_iter[n] / (t / c * 1e-3) / 1e9
                    e.update(algorithmic_bytes=bytes_iter[n], achieved_gbs=round(g, 1), frac_of_hbm_peak=round(g / peak, 4))
                ks[n] = e
            out["kernels"] = ks
            out["hbm_peak_gbs"] = peak
        s.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
