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
