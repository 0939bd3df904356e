"""Multi-GPU host side: slab partition of an image / volume domain along its slowest axis, one
process per GPU.  torch.distributed is only the plumbing that carries the NCCL
unique id and the CUDA IPC handles between the ranks; the solver's own exchanges (halo layers over
NVLink peer mappings, PCG scalars over NCCL) are issued by libThallo.so on its stream.

The reference has no multi-device path; this module is the host mirror of what a distributed
caller of Thallo.h would do: every rank holds a slab of every image (ghost layers included),
creates its own State / Problem / Plan for the local extent, and calls Thallo_ProblemSolve.
"""
import ctypes as C

import numpy as np


def slab_partition(extent, world, halo):
    """Split `extent` layers of the slowest axis over `world` ranks.
    Returns a list of dicts(start, count, ghost_lo, ghost_hi): rank r owns layers
    [start, start+count) and additionally holds ghost_lo / ghost_hi neighbouring layers."""
    extent, world, halo = int(extent), int(world), int(halo)
    assert world >= 1 and extent >= world, "fewer layers than ranks"
    base, rem = divmod(extent, world)
    out, start = [], 0
    for r in range(world):
        count = base + (1 if r < rem else 0)
        assert world == 1 or count >= halo, "slab thinner than the stencil halo"
        out.append(dict(start=start, count=count, ghost_lo=halo if r > 0 else 0, ghost_hi=halo if r < world - 1 else 0))
        start += count
    return out


def local_slab(array, layer_elems, part):
    """Rows of a global per-element array (first axis = elements, slowest axis outermost) that
    rank `part` holds: its owned layers plus ghost layers, as a contiguous copy."""
    a = np.asarray(array)
    lo = (part["start"] - part["ghost_lo"]) * layer_elems
    hi = (part["start"] + part["count"] + part["ghost_hi"]) * layer_elems
    return np.ascontiguousarray(a[lo:hi])


def owned_rows(local, layer_elems, part):
    """View of the owned layers inside a local slab (ghost layers stripped)."""
    lo = part["ghost_lo"] * layer_elems
    return local[lo:lo + part["count"] * layer_elems]


def stencil_halo(energy, global_dims, kind="gauss_newton", double=False):
    """Halo radius of the operator along the slowest axis (from the lowering, no compilation)."""
    import energies
    from .frontend import codegen
    mod = energies.resolve(energy) or energy
    low = codegen.lower(energies.load(mod), list(global_dims), kind, mod, double)
    assert low.desc.get("tiled"), "slab partitioning needs a 2-D / 3-D image-domain energy"
    return int(low.desc["tile"]["halo"][len(global_dims) - 1])


class SlabSolver:
    """One rank of a slab-partitioned solve.  `group` is a torch.distributed process group (any
    backend that can move small Python objects: gloo or nccl)."""

    def __init__(self, global_dims, energy, kind, rank, world, double=False, group=None, **kw):
        import torch.distributed as dist
        from .api import ThalloSolver, lib
        self.rank, self.world = rank, world
        self.global_dims = [int(d) for d in global_dims]
        self.halo = stencil_halo(energy, self.global_dims, kind, double)
        self.parts = slab_partition(self.global_dims[-1], world, self.halo)
        self.part = self.parts[rank]
        self.layer = int(np.prod(self.global_dims[:-1]))
        ext = self.part["count"] + self.part["ghost_lo"] + self.part["ghost_hi"]
        self.local_dims = self.global_dims[:-1] + [ext]
        partition = (self.part["ghost_lo"], self.part["ghost_hi"]) if world > 1 else None
        self.solver = ThalloSolver(self.local_dims, energy, kind, double=double, partition=partition, **kw)
        if world > 1:
            L = lib()
            s = self.solver
            # NCCL id from rank 0
            idbuf = C.create_string_buffer(128)
            if rank == 0:
                assert L.ThalloB200_NcclUniqueId(idbuf, 128) == 0, "ncclGetUniqueId failed"
            box = [bytes(idbuf.raw)]
            dist.broadcast_object_list(box, src=0, group=group)
            assert L.ThalloB200_PlanInitComm(s.state, s.plan, box[0], rank, world) == 0, L.ThalloB200_LastError().decode()
            # CUDA IPC handles of the neighbours' solver vectors
            h = C.create_string_buffer(64)
            extent = C.c_longlong(0)
            assert L.ThalloB200_PlanIpcHandle(s.state, s.plan, h, C.byref(extent)) == 0
            mine = (bytes(h.raw), int(extent.value))
            everyone = [None] * world
            dist.all_gather_object(everyone, mine, group=group)
            lo = everyone[rank - 1] if rank > 0 else (None, 0)
            hi = everyone[rank + 1] if rank < world - 1 else (None, 0)
            assert L.ThalloB200_PlanConnect(s.state, s.plan, lo[0], lo[1], hi[0], hi[1]) == 0
            self._keep = (box, everyone)

    def slab(self, global_array):
        return local_slab(global_array, self.layer, self.part)

    def owned(self, local_array):
        return owned_rows(local_array, self.layer, self.part)

    def __getattr__(self, name):          # solve / init / step / current_cost / set_parameters / ...
        return getattr(self.solver, name)
