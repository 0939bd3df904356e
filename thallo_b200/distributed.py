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


def connect_ranks(solver, rank, world, group=None):
    """Collective over `group`: NCCL communicator (id from rank 0) and the all-to-all CUDA IPC mapping of the ranks'
    solver-vector blocks (ThalloB200_PlanConnectAll): mailboxes of the in-kernel all-reduce + the neighbours' ghost
    layers.  Returns the objects that must stay alive as long as the plan."""
    import torch.distributed as dist
    from .api import lib
    L = lib()
    s = solver
    idbuf = C.create_string_buffer(128)
    if rank == 0:
        assert L.ThalloB200_NcclUniqueId(idbuf, 128) == 0, "ncclGetUniqueId failed"
    box = [bytes(idbuf.raw)]
    dist.broadcast_object_list(box, src=0, group=group)
    assert L.ThalloB200_PlanInitComm(s.state, s.plan, box[0], rank, world) == 0, L.ThalloB200_LastError().decode()
    h = C.create_string_buffer(64)
    extent = C.c_longlong(0)
    assert L.ThalloB200_PlanIpcHandle(s.state, s.plan, h, C.byref(extent)) == 0
    info = (C.c_longlong * 4)()
    assert L.ThalloB200_PlanPeerInfo(s.state, s.plan, info) == 0
    mine = (bytes(h.raw), [int(x) for x in info])
    everyone = [None] * world
    dist.all_gather_object(everyone, mine, group=group)
    handles = C.create_string_buffer(b"".join(e[0] for e in everyone), 64 * world)
    infos = (C.c_longlong * (4 * world))(*[x for e in everyone for x in e[1]])
    rc = L.ThalloB200_PlanConnectAll(s.state, s.plan, world, handles, infos)
    assert rc == 0, L.ThalloB200_LastError().decode()
    return (box, everyone, handles, infos)


class SlabSolver:
    """One rank of a slab-partitioned solve.  `group` is a torch.distributed process group (any
    backend that can move small Python objects: gloo or nccl)."""

    def __init__(self, global_dims, energy, kind, rank, world, double=False, group=None, **kw):
        from .api import ThalloSolver
        self.rank, self.world = rank, world
        self.global_dims = [int(d) for d in global_dims]
        self.halo = stencil_halo(energy, self.global_dims, kind, double)
        self.parts = slab_partition(self.global_dims[-1], world, self.halo)
        self.part = self.parts[rank]
        self.layer = int(np.prod(self.global_dims[:-1]))
        ext = self.part["count"] + self.part["ghost_lo"] + self.part["ghost_hi"]
        self.local_dims = self.global_dims[:-1] + [ext]
        partition = (self.part["ghost_lo"], self.part["ghost_hi"], self.part["start"] - self.part["ghost_lo"]) if world > 1 else None
        self.solver = ThalloSolver(self.local_dims, energy, kind, double=double, partition=partition, **kw)
        if world > 1:
            self._keep = connect_ranks(self.solver, rank, world, group)

    def slab(self, global_array):
        return local_slab(global_array, self.layer, self.part)

    def owned(self, local_array):
        return owned_rows(local_array, self.layer, self.part)

    def __getattr__(self, name):          # solve / init / step / current_cost / set_parameters / ...
        return getattr(self.solver, name)


# ---------------------------------------------------------------------------------------------------
# Graph domains: contiguous vertex ranges, one per rank.  A residual over the
# edge domain belongs to the rank that owns the vertex its FIRST index array points at; a rank also holds
# the foreign edges that touch its vertices through another index array (their J^T J p contributions to
# its own vertices are gathered locally, so no scatter ever crosses ranks), and copies ("ghosts") of the
# neighbouring ranks' vertices those edges reach.  With a locality-preserving vertex order (any banded
# ordering: grid order, RCM) the ghosts of a rank are the tail of the previous rank's range and the head
# of the next one's, which makes the exchange the 1-D case of the slab exchange: two contiguous blocks.
def graph_partition(num_vertices, index_arrays, world):
    """Partition `num_vertices` vertices and the edges described by `index_arrays` (list of int arrays of
    equal length E; entry e of array k is the vertex edge e reaches through its k-th index) over `world` ranks.

    Returns a list of dicts, one per rank:
      start, count          owned vertex range [start, start+count)
      ghost_lo, ghost_hi    ghost vertices held in front of / behind the owned range
      edges                 global ids of the local edges: owned edges first (ascending), then foreign edges
      owned_edges           number of owned edges (the first `owned_edges` entries of `edges`)
    Raises ValueError when an edge reaches beyond the adjacent ranks (vertex order not banded enough)."""
    N, world = int(num_vertices), int(world)
    idx = [np.asarray(a).astype(np.int64).reshape(-1) for a in index_arrays]
    assert idx and all(len(a) == len(idx[0]) for a in idx), "index arrays must have equal length"
    assert world >= 1 and N >= world, "fewer vertices than ranks"
    base, rem = divmod(N, world)
    starts = [r * base + min(r, rem) for r in range(world + 1)]
    owner = [np.searchsorted(np.asarray(starts[1:]), a, side="right") for a in idx]      # rank owning each endpoint
    parts = []
    for r in range(world):
        s, e = starts[r], starts[r + 1]
        touches = np.zeros(len(idx[0]), bool)
        for o in owner:
            touches |= o == r
        own = owner[0] == r
        owned_e = np.nonzero(own)[0]
        foreign_e = np.nonzero(touches & ~own)[0]
        edges = np.concatenate([owned_e, foreign_e])
        lo, hi = s, e
        for a in idx:
            if len(edges):
                lo = min(lo, int(a[edges].min()))
                hi = max(hi, int(a[edges].max()) + 1)
        if lo < (starts[r - 1] if r > 0 else 0) or hi > (starts[r + 2] if r + 2 <= world else N):
            raise ValueError("rank %d: edges reach beyond the adjacent ranks; reorder the vertices (e.g. RCM) "
                             "so that edges connect nearby vertex ids" % r)
        parts.append(dict(start=s, count=e - s, ghost_lo=s - lo, ghost_hi=hi - e, edges=edges, owned_edges=len(owned_e)))
    # a rank's ghosts must not be wider than what the neighbour owns (they are pushed from owned memory)
    for r in range(world):
        if r > 0 and parts[r]["ghost_lo"] > parts[r - 1]["count"] or r < world - 1 and parts[r]["ghost_hi"] > parts[r + 1]["count"]:
            raise ValueError("rank %d: ghost block wider than the neighbouring rank's range" % r)
    return parts


def local_vertex_rows(array, part):
    """Rows (vertices) of a global per-vertex array a rank holds: ghosts, owned, ghosts."""
    a = np.asarray(array)
    return np.ascontiguousarray(a[part["start"] - part["ghost_lo"]:part["start"] + part["count"] + part["ghost_hi"]])


def local_index_array(index_array, part):
    """A global index array restricted to the rank's edges and renumbered to its local vertex ids."""
    a = np.asarray(index_array).reshape(-1)
    return np.ascontiguousarray((a[part["edges"]].astype(np.int64) - (part["start"] - part["ghost_lo"])).astype(np.int32))


def owned_vertex_rows(local, part):
    return local[part["ghost_lo"]:part["ghost_lo"] + part["count"]]


class GraphSolver:
    """One rank of a vertex-partitioned graph solve (energies over a vertex domain N and an edge domain E,
    e.g. arap_mesh_deformation).  `vertex_dim` / `edge_dim` are the positions of N and E in Dims();
    `index_arrays` the global index arrays in the order the energy declares them (the first one decides
    which rank owns an edge)."""

    def __init__(self, global_dims, energy, kind, rank, world, index_arrays, vertex_dim=0, edge_dim=1, double=False,
                 group=None, **kw):
        from .api import ThalloSolver
        self.rank, self.world = rank, world
        self.parts = graph_partition(global_dims[vertex_dim], index_arrays, world)
        self.part = p = self.parts[rank]
        self.local_dims = list(global_dims)
        self.local_dims[vertex_dim] = p["ghost_lo"] + p["count"] + p["ghost_hi"]
        self.local_dims[edge_dim] = len(p["edges"])
        partition = None
        if world > 1:
            partition = {vertex_dim: (p["ghost_lo"], p["ghost_hi"]), edge_dim: (0, len(p["edges"]) - p["owned_edges"])}
        self.solver = ThalloSolver(self.local_dims, energy, kind, double=double, partition=partition, schedule="gather", **kw)
        if world > 1:
            # (what a neighbour needs to address this rank's ghost blocks -- the handle, the local vertex count and the
            # widths of the ghost blocks -- travels in ThalloB200_PlanPeerInfo)
            self._keep = connect_ranks(self.solver, rank, world, group)

    def vertex_rows(self, global_array):
        return local_vertex_rows(global_array, self.part)

    def index_array(self, global_index_array):
        return local_index_array(global_index_array, self.part)

    def owned(self, local_array):
        return owned_vertex_rows(local_array, self.part)

    def __getattr__(self, name):
        return getattr(self.solver, name)


# ---------------------------------------------------------------------------------------------------
# Bundle adjustment: points and their observations are partitioned (an observation
# touches exactly one point, so every point's residuals are local to its rank); the cameras are replicated --
# every rank sees observations of every camera -- and their part of J^T F, diag(J^T J) and J^T J p is summed over
# the ranks with one NCCL all-reduce of the camera block each (9 C scalars).
def point_partition(num_points, obs_to_point, world):
    """Contiguous point ranges and the observations that belong to them.
    Returns per rank dict(start, count, observations = global ids of its observations, ascending)."""
    P, world = int(num_points), int(world)
    o2p = np.asarray(obs_to_point).astype(np.int64).reshape(-1)
    base, rem = divmod(P, world)
    starts = [r * base + min(r, rem) for r in range(world + 1)]
    owner = np.searchsorted(np.asarray(starts[1:]), o2p, side="right")
    return [dict(start=starts[r], count=starts[r + 1] - starts[r], observations=np.nonzero(owner == r)[0]) for r in range(world)]


class ReplicatedSolver:
    """One rank of a solve with one replicated unknown domain (cameras), one partitioned unknown domain (points)
    and a residual domain (observations) indexed into both, e.g. bundle_adjustment: Dims(C, P, O).
    `rep_dim`, `part_dim`, `res_dim` are the positions of those domains in Dims()."""

    def __init__(self, global_dims, energy, kind, rank, world, obs_to_point, rep_dim=0, part_dim=1, res_dim=2, double=False,
                 group=None, **kw):
        from .api import ThalloSolver
        self.rank, self.world = rank, world
        self.parts = point_partition(global_dims[part_dim], obs_to_point, world)
        self.part = p = self.parts[rank]
        self.local_dims = list(global_dims)
        self.local_dims[part_dim] = p["count"]
        self.local_dims[res_dim] = len(p["observations"])
        partition = dict(replicated=(rep_dim,), owner=(rank == 0)) if world > 1 else None
        self.solver = ThalloSolver(self.local_dims, energy, kind, double=double, partition=partition, schedule="gather", **kw)
        if world > 1:
            self._keep = connect_ranks(self.solver, rank, world, group)

    def point_rows(self, global_array):
        a = np.asarray(global_array)
        return np.ascontiguousarray(a[self.part["start"]:self.part["start"] + self.part["count"]])

    def observation_rows(self, global_array):
        return np.ascontiguousarray(np.asarray(global_array)[self.part["observations"]])

    def point_index(self, obs_to_point):
        a = np.asarray(obs_to_point).reshape(-1)[self.part["observations"]].astype(np.int64) - self.part["start"]
        return np.ascontiguousarray(a.astype(np.int32))

    def __getattr__(self, name):
        return getattr(self.solver, name)
