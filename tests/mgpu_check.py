"""Multi-GPU parity check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Solves the same global image_warping problem (a) slab-partitioned over all ranks and (b) on one
GPU, and compares every cost of the trajectory, the LM inner iteration counts and the unknowns.
Exit code 0 = parity.  Called by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(kind, W, H, nit, lit, dist, rank, world, torch):
    from thallo_b200 import workloads as wl
    from thallo_b200.api import ThalloSolver
    from thallo_b200.distributed import SlabSolver
    d = wl.image_warping_inputs(W, H)
    names = ("Offset", "Angle", "UrShape", "Constraints", "Mask")
    scal = [np.array([d["w_fitSqrt"]], np.float32), np.array([d["w_regSqrt"]], np.float32)]
    s = SlabSolver([W, H], "image_warping", kind, rank, world)
    loc = [torch.from_numpy(s.slab(d[k])).cuda() for k in names]
    s.set_parameters(nIterations=nit, lIterations=lit)
    s.init(loc + scal)
    costs, lin = [s.current_cost()], []
    while s.step():
        costs.append(s.current_cost())
        lin.append(s.last_linear_iterations())
    costs.append(s.current_cost())
    torch.cuda.synchronize()
    own = [s.owned(t.cpu().numpy()) for t in loc[:2]]
    gathered = [None] * world
    dist.all_gather_object(gathered, own)
    ok = True
    if rank == 0:
        off = np.concatenate([g[0] for g in gathered])
        ang = np.concatenate([g[1] for g in gathered])
        d1 = wl.image_warping_inputs(W, H)
        one = [torch.from_numpy(np.ascontiguousarray(d1[k])).cuda() for k in names]
        r = ThalloSolver([W, H], "image_warping", kind)
        r.set_parameters(nIterations=nit, lIterations=lit)
        r.init(one + scal)
        c1, l1 = [r.current_cost()], []
        while r.step():
            c1.append(r.current_cost())
            l1.append(r.last_linear_iterations())
        c1.append(r.current_cost())
        rel = max(abs(a - b) / max(abs(b), 1e-3) for a, b in zip(costs, c1)) if len(costs) == len(c1) else float("inf")
        du = float(np.abs(off - one[0].cpu().numpy()).max())
        da = float(np.abs(ang - one[1].cpu().numpy()).max())
        ok = len(costs) == len(c1) and rel <= 1e-5 and lin == l1 and du < 1e-3 and da < 1e-3
        print("mgpu %s %dx%d world=%d: max rel cost diff %.3g, lin %s vs %s, max|dOffset| %.3g max|dAngle| %.3g -> %s"
              % (kind, W, H, world, rel, lin, l1, du, da, "OK" if ok else "MISMATCH"), flush=True)
        if not ok:
            print(costs, c1, flush=True)
    return ok


def run_generic(label, energy, kind, dims, params, unknown_slots, image_slots, nit, lit, dist, rank, world, torch, tol=1e-5,
                origin_param=None):
    """Any 2-D / 3-D image-domain energy: `params` is the single-GPU parameter list; `image_slots` are sliced into
    slabs, the other entries are host scalars.  origin_param = (slot, sign): scalar holding an absolute coordinate of
    the slowest axis is passed through unchanged (the lowering offsets the index VALUE by the slab origin)."""
    from thallo_b200.api import ThalloSolver
    from thallo_b200.distributed import SlabSolver
    s = SlabSolver(dims, energy, kind, rank, world)
    loc = {i: torch.from_numpy(s.slab(np.asarray(params[i]).reshape(int(np.prod(dims)), -1))).cuda() for i in image_slots}
    pl = [loc[i] if i in loc else params[i] for i in range(len(params))]
    s.set_parameters(nIterations=nit, lIterations=lit)
    s.init(pl)
    costs, lin = [s.current_cost()], []
    while s.step():
        costs.append(s.current_cost())
        lin.append(s.last_linear_iterations())
    costs.append(s.current_cost())
    torch.cuda.synchronize()
    own = [s.owned(loc[i].cpu().numpy()) for i in unknown_slots]
    gathered = [None] * world
    dist.all_gather_object(gathered, own)
    ok = True
    if rank == 0:
        one = [torch.from_numpy(np.ascontiguousarray(params[i])).cuda() if i in image_slots else params[i] for i in range(len(params))]
        r = ThalloSolver(dims, energy, kind)
        r.set_parameters(nIterations=nit, lIterations=lit)
        r.init(one)
        c1, l1 = [r.current_cost()], []
        while r.step():
            c1.append(r.current_cost())
            l1.append(r.last_linear_iterations())
        c1.append(r.current_cost())
        rel = max(abs(a - b) / max(abs(b), 1e-3) for a, b in zip(costs, c1)) if len(costs) == len(c1) else float("inf")
        du = 0.0
        for k, i in enumerate(unknown_slots):
            full = np.concatenate([g[k] for g in gathered]).reshape(-1)
            du = max(du, float(np.abs(full - one[i].cpu().numpy().reshape(-1)).max()))
        ok = len(costs) == len(c1) and rel <= tol and lin == l1 and du < 1e-3
        print("mgpu %s %s %s world=%d: max rel cost diff %.3g, lin %s vs %s, max|dX| %.3g -> %s"
              % (label, kind, "x".join(map(str, dims)), world, rel, lin, l1, du, "OK" if ok else "MISMATCH"), flush=True)
        if not ok:
            print(costs, c1, flush=True)
    return ok


def more_cases(dist, rank, world, torch):
    """Configs 3a, 3b, 4a under the slab partition: sampled images (row offset of the local slab), computed arrays
    (stored images evaluated on owned + ghost layers, index VALUES offset by the slab origin), 3-D volume."""
    from thallo_b200 import workloads as wl
    ok = True
    W, H = 64, 72
    p = wl.optical_flow_params(wl.optical_flow_inputs(W, H))
    ok = run_generic("optical_flow", "optical_flow", "gauss_newton", [W, H], p, [2], [2, 3, 4, 5, 6], 3, 20, dist, rank, world, torch) and ok
    W, H = 64, 80
    p = wl.sfs_params(wl.sfs_inputs(W, H))
    for kind in ("gauss_newton", "levenberg_marquardt"):
        ok = run_generic("shape_from_shading", "shape_from_shading", kind, [W, H], p, [16], [16, 17, 18, 19, 20], 3, 15, dist, rank, world, torch) and ok
    W, H, Dz = 16, 12, 24
    p = wl.volumetric_params(wl.volumetric_inputs(W, H, Dz))
    ok = run_generic("volumetric", "volumetric_mesh_deformation", "gauss_newton", [W, H, Dz], p, [0, 1], [0, 1, 2, 3], 3, 20, dist, rank, world, torch) and ok
    return ok


def determinism(dist, rank, world, torch, repeats=4):
    """The same LM solve repeated: every cost, every PCG iteration count and the unknowns must come out bit-identical
    (dot products are summed in a fixed order, also across ranks: rank-ordered sums of the mailbox values)."""
    from thallo_b200 import workloads as wl
    from thallo_b200.distributed import SlabSolver
    W, H, kind = 128, 96, "levenberg_marquardt"
    d = wl.image_warping_inputs(W, H)
    names = ("Offset", "Angle", "UrShape", "Constraints", "Mask")
    scal = [np.array([d["w_fitSqrt"]], np.float32), np.array([d["w_regSqrt"]], np.float32)]
    s = SlabSolver([W, H], "image_warping", kind, rank, world)
    runs = []
    for _ in range(repeats):
        loc = [torch.from_numpy(s.slab(d[k])).cuda() for k in names]
        s.set_parameters(nIterations=4, lIterations=40, trust_region_radius=1e4)     # the radius persists across solves (gauss_newton.t:1751)
        s.init(loc + scal)
        costs, lin = [s.current_cost()], []
        while s.step():
            costs.append(s.current_cost())
            lin.append(s.last_linear_iterations())
        torch.cuda.synchronize()
        runs.append((costs, lin, loc[0].cpu().numpy().tobytes(), loc[1].cpu().numpy().tobytes()))
    same = all(r == runs[0] for r in runs[1:])
    flags = [None] * world
    dist.all_gather_object(flags, same)
    ok = all(flags)
    if rank == 0:
        print("mgpu determinism world=%d: %d repeated LM solves %s (lin %s)" % (world, repeats, "deterministic" if ok else "DIFFER", runs[0][1]), flush=True)
    return ok


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ok = True
    for kind, W, H, nit, lit in [("gauss_newton", 96, 64, 3, 20), ("levenberg_marquardt", 128, 90, 5, 40)]:
        ok = run(kind, W, H, nit, lit, dist, rank, world, torch) and ok
    ok = more_cases(dist, rank, world, torch) and ok
    ok = determinism(dist, rank, world, torch) and ok
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
