"""Host-side logic of the multi-GPU path on CPU: slab partition arithmetic, slab scatter / gather
round trips across two gloo ranks, and NVRTC compilation of partitioned plans for sm_100a."""
import os
import socket

import numpy as np
import pytest

from thallo_b200 import distributed as D


def test_slab_partition_covers_extent_once():
    for extent, world, halo in [(2048, 8, 1), (83, 3, 2), (10, 1, 1), (9, 4, 1)]:
        parts = D.slab_partition(extent, world, halo)
        assert len(parts) == world
        assert parts[0]["start"] == 0 and parts[0]["ghost_lo"] == 0 and parts[-1]["ghost_hi"] == 0
        assert sum(p["count"] for p in parts) == extent
        for a, b in zip(parts, parts[1:]):
            assert a["start"] + a["count"] == b["start"]
            assert a["ghost_hi"] == halo and b["ghost_lo"] == halo


def test_local_slab_and_owned_rows_roundtrip():
    W, H, halo, world = 12, 29, 2, 3
    g = np.arange(W * H * 2, dtype=np.float32).reshape(W * H, 2)
    parts = D.slab_partition(H, world, halo)
    back = np.concatenate([D.owned_rows(D.local_slab(g, W, p), W, p) for p in parts])
    assert np.array_equal(back, g)
    loc = D.local_slab(g, W, parts[1])
    assert loc.shape[0] == (parts[1]["count"] + 2 * halo) * W
    assert np.array_equal(loc[:W * halo], g[(parts[1]["start"] - halo) * W:parts[1]["start"] * W])


def test_partitioned_plans_compile_for_sm100a():
    import energies
    from thallo_b200 import api
    from thallo_b200.frontend import codegen
    api.build_library()
    assert D.stencil_halo("image_warping", [64, 64], "levenberg_marquardt") == 1
    for part in [(0, 1), (1, 1), (1, 0)]:
        low = codegen.lower(energies.load("image_warping"), [64, 34], "levenberg_marquardt", "image_warping", partition=part)
        assert "partition %d %d" % part in codegen.descriptor_text(low.desc)
        ok, log, size = api.compile_only(low.source)
        assert ok, log[-3000:]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H = 16, 23
    g = np.arange(W * H, dtype=np.float32)
    parts = D.slab_partition(H, world, 1)
    loc = D.local_slab(g, W, parts[rank])
    # what the library's halo push does between neighbours, done here with gloo: send my first /
    # last owned row to the neighbour and check it equals the ghost row I hold of it
    import torch
    own = D.owned_rows(loc, W, parts[rank])
    if rank == 0:
        dist.send(torch.from_numpy(own[-W:].copy()), dst=1)
        got = torch.empty(W)
        dist.recv(got, src=1)
        ok = np.array_equal(got.numpy(), loc[-W:])
    else:
        got = torch.empty(W)
        dist.recv(got, src=0)
        dist.send(torch.from_numpy(own[:W].copy()), dst=0)
        ok = np.array_equal(got.numpy(), loc[:W])
    gathered = [None] * world
    dist.all_gather_object(gathered, own.tolist())
    q.put((rank, bool(ok), np.array_equal(np.concatenate([np.array(x, np.float32) for x in gathered]), g)))
    dist.destroy_process_group()


def test_two_rank_gloo_halo_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert res == [(0, True, True), (1, True, True)]


# ---------------------------------------------------------------- graph domains: vertex partition with ghost vertices
def _arap_problem(nx, ny, seed=3):
    from thallo_b200 import workloads as wl
    d = wl.arap_mesh_inputs(nx, ny)
    rng = np.random.RandomState(seed)
    d["Position"] = d["Position"] + 0.2 * rng.randn(*d["Position"].shape).astype(np.float32)
    d["Angle"] = d["Angle"] + 0.3 * rng.randn(*d["Angle"].shape).astype(np.float32)
    return d, wl.arap_mesh_params(d)


def _local_params(params, part):
    """arap_mesh parameter list (w_fit, w_reg, Position, Angle, Original, Constraints, V0, V1) of one rank."""
    out = list(params[:2])
    out += [D.local_vertex_rows(np.asarray(p, np.float64), part) for p in params[2:6]]
    out += [D.local_index_array(p, part) for p in params[6:8]]
    return out


def test_graph_partition_owns_every_vertex_and_edge_once():
    d, params = _arap_problem(30, 20)
    N, V0, V1 = 600, params[6], params[7]
    for world in (1, 2, 3, 5):
        parts = D.graph_partition(N, [V0, V1], world)
        assert sum(p["count"] for p in parts) == N and parts[0]["ghost_lo"] == 0 and parts[-1]["ghost_hi"] == 0
        owned = np.concatenate([p["edges"][:p["owned_edges"]] for p in parts])
        assert np.array_equal(np.sort(owned), np.arange(len(V0)))
        for r, p in enumerate(parts):
            lo, hi = p["start"], p["start"] + p["count"]
            e_own, e_for = p["edges"][:p["owned_edges"]], p["edges"][p["owned_edges"]:]
            assert np.all((V0[e_own] >= lo) & (V0[e_own] < hi))                     # owner = rank of the first endpoint
            assert np.all((V1[e_for] >= lo) & (V1[e_for] < hi)) and not np.any((V0[e_for] >= lo) & (V0[e_for] < hi))
            for a in (V0, V1):                                                      # every endpoint is a local vertex
                loc = D.local_index_array(a, p)
                assert loc.min() >= 0 and loc.max() < p["ghost_lo"] + p["count"] + p["ghost_hi"]
            # every edge touching an owned vertex is local
            touching = np.nonzero(((V0 >= lo) & (V0 < hi)) | ((V1 >= lo) & (V1 < hi)))[0]
            assert set(touching.tolist()) == set(p["edges"].tolist())
        pos = np.asarray(params[2])
        assert np.array_equal(np.concatenate([D.owned_vertex_rows(D.local_vertex_rows(pos, p), p) for p in parts]), pos)


def test_graph_partition_rejects_far_reaching_edges():
    N = 100
    v0 = np.arange(N - 1, dtype=np.int32)
    v1 = np.arange(1, N, dtype=np.int32)
    D.graph_partition(N, [v0, v1], 4)
    v1b = v1.copy()
    v1b[0] = 99                              # vertex 0 (rank 0) linked to vertex 99 (rank 3)
    with pytest.raises(ValueError):
        D.graph_partition(N, [v0, v1b], 4)


def test_local_operators_and_costs_of_a_vertex_partition_add_up_to_the_global_ones():
    """What the partitioned solve relies on: a rank's local problem (owned + ghost vertices, owned + foreign edges)
    reproduces J^T J p at its owned vertices exactly, and the owned residuals of all ranks are the global residuals."""
    import energies
    from oracle.npdsl import evaluate
    from thallo_b200.frontend import codegen, interp
    nx, ny, world = 12, 10, 3
    d, params = _arap_problem(nx, ny)
    N, E = nx * ny, len(params[6])
    p64 = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in params]
    define = energies.load("arap_mesh_deformation")
    _, F, J = evaluate(define, [N, E], p64, np.float64)
    pvec = np.random.RandomState(1).randn(J.shape[1])
    want = J.T.tocsr() @ (J @ pvec)
    cost = 0.5 * float(F @ F)
    parts = D.graph_partition(N, [params[6], params[7]], world)
    got = np.zeros_like(want)
    cost_sum = 0.0
    for part in parts:
        lp = _local_params(p64, part)
        nl, el = len(lp[2]), len(lp[6])
        partition = {0: (part["ghost_lo"], part["ghost_hi"]), 1: (0, el - part["owned_edges"])}
        low = codegen.lower(define, [nl, el], "gauss_newton", "arap_mesh_deformation", True, "gather", partition=partition)
        assert "gpartition 0 %d %d %d" % (nl, part["ghost_lo"], part["ghost_hi"]) in codegen.descriptor_text(low.desc)
        # local vector: Position block then Angle block, local vertex numbering
        g0 = part["start"] - part["ghost_lo"]
        ploc = np.concatenate([pvec[3 * g0:3 * (g0 + nl)], pvec[3 * N + 3 * g0:3 * N + 3 * (g0 + nl)]])
        out = interp.gather_apply(low.generator, lp, ploc)
        a, b = part["ghost_lo"], part["ghost_lo"] + part["count"]
        got[3 * part["start"]:3 * (part["start"] + part["count"])] = out[3 * a:3 * b]
        got[3 * N + 3 * part["start"]:3 * N + 3 * (part["start"] + part["count"])] = out[3 * nl + 3 * a:3 * nl + 3 * b]
        _, Fl, _ = evaluate(define, [nl, el], lp, np.float64)
        fit, reg = Fl[:3 * nl].reshape(nl, 3), Fl[3 * nl:].reshape(el, 3)          # groups sorted by name: fit, reg
        cost_sum += 0.5 * float((fit[a:b] ** 2).sum() + (reg[:part["owned_edges"]] ** 2).sum())
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    assert abs(cost_sum - cost) <= 1e-12 * cost


def test_graph_partitioned_plans_compile_for_sm100a():
    import energies
    from thallo_b200 import api
    from thallo_b200.frontend import codegen
    api.build_library()
    d, params = _arap_problem(12, 10)
    parts = D.graph_partition(120, [params[6], params[7]], 3)
    for part in parts:
        nl = part["ghost_lo"] + part["count"] + part["ghost_hi"]
        partition = {0: (part["ghost_lo"], part["ghost_hi"]), 1: (0, len(part["edges"]) - part["owned_edges"])}
        for kind in ("gauss_newton", "levenberg_marquardt"):
            low = codegen.lower(energies.load("arap_mesh_deformation"), [nl, len(part["edges"])], kind, "arap_mesh_deformation",
                                schedule="gather", partition=partition)
            ok, log, size = api.compile_only(low.source)
            assert ok, log[-3000:]
            assert "TH_PART_TABLE" in low.source and "TH_RANGE_TABLE" in low.source


def _graph_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d, params = _arap_problem(14, 9)
    N = 14 * 9
    parts = D.graph_partition(N, [params[6], params[7]], world)
    part = parts[rank]
    g = np.arange(N * 3, dtype=np.float32).reshape(N, 3)
    loc = D.local_vertex_rows(g, part)
    own = D.owned_vertex_rows(loc, part)
    # the library's ghost push between neighbours, done here with gloo: my first / last owned vertices, as many
    # as the neighbour holds ghosts of, must equal the ghost block the neighbour sliced out of the global array
    widths = [None] * world
    dist.all_gather_object(widths, (part["ghost_lo"], part["ghost_hi"]))
    ok = True
    if rank == 0:
        w = widths[1][0]
        dist.send(torch.from_numpy(own[len(own) - w:].copy()), dst=1)
        got = torch.empty(part["ghost_hi"], 3)
        dist.recv(got, src=1)
        ok = np.array_equal(got.numpy(), loc[len(loc) - part["ghost_hi"]:])
    else:
        got = torch.empty(part["ghost_lo"], 3)
        dist.recv(got, src=0)
        w = widths[0][1]
        dist.send(torch.from_numpy(own[:w].copy()), dst=0)
        ok = np.array_equal(got.numpy(), loc[:part["ghost_lo"]])
    gathered = [None] * world
    dist.all_gather_object(gathered, own.tolist())
    q.put((rank, bool(ok), np.array_equal(np.concatenate([np.array(x, np.float32) for x in gathered]), g)))
    dist.destroy_process_group()


def test_two_rank_gloo_ghost_vertex_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_graph_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert res == [(0, True, True), (1, True, True)]


# ---------------------------------------------------------------- slab partition of energies with absolute coordinates
@pytest.mark.parametrize("name", ["shape_from_shading", "optical_flow"])
def test_slab_local_operators_equal_the_global_ones(name):
    """Computed arrays (evaluated on owned + ghost layers), index VALUES (offset by the slab origin) and sampled
    images (row offset of the local slab): the generated at-output operators of every rank's local problem equal
    the global problem's on the owned rows."""
    import energies
    from thallo_b200 import workloads as wl
    from thallo_b200.frontend import codegen, interp
    W, H, world = 24, 30, 3
    if name == "shape_from_shading":
        params = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.sfs_params(wl.sfs_inputs(W, H))]
        images, uslot, kind = [16, 17, 18, 19, 20], 16, "gauss_newton"
    else:
        d = wl.optical_flow_inputs(W, H)
        d["X"] = (0.4 * np.random.RandomState(2).randn(W * H, 2)).astype(np.float32)        # |flow| < ghost width
        params = [np.asarray(p, np.float64) for p in wl.optical_flow_params(d)]
        images, uslot, kind = [2, 3, 4, 5, 6], 2, "gauss_newton"
    define = energies.load(name)
    low = codegen.lower(define, [W, H], kind, name, True, "at_output")
    U = low.generator.U
    pvec = np.random.RandomState(4).randn(W * H * U)
    g0, d0, o0 = interp.unknownwise(low.generator, params, pvec)
    halo = D.stencil_halo(name, [W, H], kind)
    parts = D.slab_partition(H, world, halo)
    for part in parts:
        ext = part["count"] + part["ghost_lo"] + part["ghost_hi"]
        lp = [D.local_slab(np.asarray(p).reshape(W * H, -1), W, part).reshape((-1,) + np.asarray(p).shape[1:]) if i in images else p
              for i, p in enumerate(params)]
        origin = part["start"] - part["ghost_lo"]
        lowl = codegen.lower(define, [W, ext], kind, name, True, "at_output", partition=(part["ghost_lo"], part["ghost_hi"], origin))
        pl = D.local_slab(pvec.reshape(W * H, U), W, part).reshape(-1)
        g, d, o = interp.unknownwise(lowl.generator, lp, pl)
        a, b = part["ghost_lo"] * W * U, (part["ghost_lo"] + part["count"]) * W * U
        ga, gb = part["start"] * W * U, (part["start"] + part["count"]) * W * U
        for got, want in ((g, g0), (d, d0), (o, o0)):
            assert np.abs(got[a:b] - want[ga:gb]).max() <= 1e-10 * max(1.0, np.abs(want).max()), (name, part)


# ---------------------------------------------------------------- bundle adjustment: points partitioned, cameras replicated
def test_point_partition_and_replicated_camera_sums():
    """Every observation belongs to exactly one rank; the ranks' local J^T J p agree with the global product on
    their own points, and their camera blocks ADD UP to the global camera block (what the all-reduce computes)."""
    import energies
    from oracle.npdsl import evaluate
    from thallo_b200 import workloads as wl
    from thallo_b200.frontend import codegen, interp
    Cn, Pn, world = 5, 37, 3
    d = wl.bundle_adjustment_inputs(Cn, Pn, 3)
    params = [np.asarray(p, np.float64) if np.asarray(p).dtype == np.float32 else p for p in wl.bundle_adjustment_params(d)]
    On = len(d["oToC"])
    define = energies.load("bundle_adjustment")
    _, F, J = evaluate(define, [Cn, Pn, On], params, np.float64, materialize=False)
    pvec = np.random.RandomState(2).randn(J.shape[1])
    want = J.T.tocsr() @ (J @ pvec)
    parts = D.point_partition(Pn, d["oToP"], world)
    assert np.array_equal(np.sort(np.concatenate([p["observations"] for p in parts])), np.arange(On))
    cam = np.zeros(9 * Cn)
    for r, part in enumerate(parts):
        obs = part["observations"]
        assert np.all((d["oToP"][obs] >= part["start"]) & (d["oToP"][obs] < part["start"] + part["count"]))
        lp = [params[0], params[1][part["start"]:part["start"] + part["count"]], params[2][obs], d["oToC"][obs],
              (d["oToP"][obs] - part["start"]).astype(np.int32)]
        dims = [Cn, part["count"], len(obs)]
        for mat in (False, True):
            low = codegen.lower(define, dims, "gauss_newton", "bundle_adjustment", True, "gather",
                                partition=dict(replicated=(0,), owner=(r == 0)), materialize=mat)
            assert "replicated 0 %d" % (9 * Cn) in codegen.descriptor_text(low.desc)
            assert ("#define TH_REP_OWNER %d" % int(r == 0)) in low.source
        ploc = np.concatenate([pvec[:9 * Cn], pvec[9 * Cn + 3 * part["start"]:9 * Cn + 3 * (part["start"] + part["count"])]])
        out = interp.gather_apply(low.generator, lp, ploc, materialised=True)
        cam += out[:9 * Cn]
        pts = want[9 * Cn + 3 * part["start"]:9 * Cn + 3 * (part["start"] + part["count"])]
        assert np.abs(out[9 * Cn:] - pts).max() <= 1e-10 * np.abs(want).max()
    assert np.abs(cam - want[:9 * Cn]).max() <= 1e-10 * np.abs(want).max()


def test_replicated_plans_compile_for_sm100a():
    import energies
    from thallo_b200 import api
    from thallo_b200.frontend import codegen
    api.build_library()
    for owner in (True, False):
        low = codegen.lower(energies.load("bundle_adjustment"), [6, 40, 150], "levenberg_marquardt", "bundle_adjustment",
                            schedule="gather", partition=dict(replicated=(0,), owner=owner))
        ok, log, size = api.compile_only(low.source)
        assert ok, log[-3000:]
        assert "th_rep_finish" in open(os.path.join(os.path.dirname(api.LIB_PATH), "..", "csrc", "skeleton", "thallo_kernels.cuh")).read()
