"""The configured workloads (BASELINE.json `configs`) as problem builders: dimensions, solver
parameters as the reference's example drivers set them, synthetic inputs generated on the device, the partition over
the ranks of a multi-GPU run, and the algorithmic bytes of every PCG kernel.

Host-side mirror of the set-up code in the reference's examples (examples/*/src/main.cpp, CombinedSolver.h); used by
bench.py and the full-size GPU tests.  One process per GPU: `build(rank, world)` returns this rank's solver and its
LOCAL parameter list.
"""
import numpy as np

from . import workloads as wl


class Built:
    """One rank's share of a configured problem."""

    def __init__(self, solver, params, unknown_slots, local_elements, meta=None):
        self.solver, self.params, self.unknown_slots = solver, params, list(unknown_slots)
        self.local_elements = local_elements          # elements this rank's kernels sweep (owned + ghost)
        self.meta = meta or {}
        self._pristine = [params[i].clone() for i in self.unknown_slots]

    def fresh(self):
        """The parameter list with the unknowns back in their initial state (device-to-device copies)."""
        for i, p in zip(self.unknown_slots, self._pristine):
            self.params[i].copy_(p)
        return self.params

    def bytes_in_hbm(self):
        return sum(p.numel() * p.element_size() for p in self.params if hasattr(p, "numel"))


def _f32(x):
    return np.array([x], np.float32)


class Case:
    key = energy = kind = label = ""
    dims = ()
    nit, lit = 10, 10
    solver_params = {}
    partition = None              # "slab" | "graph" | "replicated" | None (single GPU only)
    schedule = "auto"
    define_kwargs = None
    oracle_mode, materialized = "at_output", False
    U = A = P = L = 0             # SURVEY 8d symbols: unknown scalars / aux scalars per element, preconditioner, LM

    def workload(self):
        return "%s, %s, float32, nIterations=%d, lIterations=%d" % (self.label, self.kind, self.nit, self.lit)

    def params_for_solver(self):
        return dict(self.solver_params, nIterations=self.nit, lIterations=self.lit)

    # bytes per PCG iteration under the reference's pass structure, whole problem
    def survey_iteration_bytes(self):
        n = int(np.prod(self.dims))
        return 4 * ((12 + self.P + 2 * self.L) * self.U + self.A) * n

    def make_solver(self, dims, timing=1, partition=None):
        from .api import ThalloSolver
        return ThalloSolver(dims, self.energy, self.kind, timing=timing, schedule=self.schedule, partition=partition,
                            define_kwargs=self.define_kwargs)


# ---------------------------------------------------------------------------------------------------- image / volume domains
class SlabCase(Case):
    partition = "slab"

    def inputs(self, device, rows):
        """-> (parameter list for rows [y0, y1) of the slowest axis, unknown slots)"""
        raise NotImplementedError

    def build(self, rank=0, world=1, device="cuda", group=None, timing=1):
        import torch
        from .distributed import SlabSolver
        dims = [int(d) for d in self.dims]
        if world == 1:
            s = self.make_solver(dims, timing)
            rows = (0, dims[-1])
        else:
            s = SlabSolver(dims, self.energy, self.kind, rank, world, group=group, timing=timing)
            p = s.part
            rows = (p["start"] - p["ghost_lo"], p["start"] + p["count"] + p["ghost_hi"])
        params, unknown_slots = self.inputs(device, rows)
        params = [x.to(device) if isinstance(x, torch.Tensor) else x for x in params]
        b = Built(s, params, unknown_slots, int(np.prod(dims[:-1])) * (rows[1] - rows[0]), dict(rows=rows))
        s.set_parameters(**self.params_for_solver())
        return b

    def kernel_bytes(self, built):
        n = built.local_elements
        U, A, P, L = self.U, self.A_tiled, self.P, self.L
        return {"th_pcg_a": 4 * n * (2 * U + L * U + A + 2 * U), "th_pcg_b": 4 * n * ((4 + P + L) * U + 3 * U)}


class Minimal(SlabCase):
    """configs[0]: tests/minimal (laplacian.t) 256 x 256, GN 10 x 10 -- the reference's own CPU-runnable case; launch-bound."""
    key, energy, kind = "1", "laplacian", "gauss_newton"
    label = "tests/minimal image-domain quadratic energy 256x256"
    dims, nit, lit = (256, 256), 10, 10
    U, A, A_tiled, P, L = 1, 1, 1, 0, 0
    partition = None

    def inputs(self, device, rows):
        import torch
        X, A = wl.minimal_inputs(self.dims[0], self.dims[1])
        return [torch.from_numpy(X), torch.from_numpy(A)], [0]


class ImageWarping(SlabCase):
    """configs[1] (the headline): examples/image_warping 2048 x 2048, LM 8 x 100 (examples/image_warping/src/main.cpp:131-134)."""
    key, energy, kind = "2", "image_warping", "levenberg_marquardt"
    label = "examples/image_warping 2-D ARAP 2048x2048"
    dims, nit, lit = (2048, 2048), 8, 100
    U, A, A_tiled, P, L = 3, 6, 7, 1, 1

    def inputs(self, device, rows):
        import torch
        W, H = self.dims
        d = wl.image_warping_inputs(W, H)
        sl = slice(rows[0] * W, rows[1] * W)
        t = [torch.from_numpy(np.ascontiguousarray(d[k][sl])) for k in ("Offset", "Angle", "UrShape", "Constraints", "Mask")]
        return t + [_f32(d["w_fitSqrt"]), _f32(d["w_regSqrt"])], [0, 1]


class OpticalFlow(SlabCase):
    """configs[2] a: examples/optical_flow dense image energy at 8192 x 8192, GN 3 x 50."""
    key, energy, kind = "3a", "optical_flow", "gauss_newton"
    label = "examples/optical_flow 8192x8192"
    dims, nit, lit = (8192, 8192), 3, 50
    U, A, A_tiled, P, L = 2, 4, 4, 0, 0

    def inputs(self, device, rows):
        d = wl.optical_flow_inputs_torch(self.dims[0], self.dims[1], device, rows)
        return [_f32(d["w_fitSqrt"]), _f32(d["w_regSqrt"]), d["X"], d["I"], d["I_hat_im"], d["I_hat_dx"], d["I_hat_dy"]], [2]

    def kernel_bytes(self, built):
        n = built.local_elements
        return {"th_pcg_a": 4 * n * (4 * 2 + 4), "th_pcg_b": 4 * n * 7 * 2}


class ShapeFromShading(SlabCase):
    """configs[2] b: examples/shape_from_shading at 8192 x 8192, GN 60 x 10 (default.SFSSolverParameters)."""
    key, energy, kind = "3b", "shape_from_shading", "gauss_newton"
    label = "examples/shape_from_shading 8192x8192"
    dims, nit, lit = (8192, 8192), 60, 10
    U, A, A_tiled, P, L = 1, 9, 5.5, 0, 0

    def inputs(self, device, rows):
        d = wl.sfs_inputs_torch(self.dims[0], self.dims[1], device, rows)
        sc = [d["w_p"], d["w_s"], d["w_g"], d["f_x"], d["f_y"], d["u_x"], d["u_y"]] + list(d["light"])
        return [_f32(x) for x in sc] + [d["X"], d["D_i"], d["Im"], d["edgeMaskR"], d["edgeMaskC"]], [16]

    def kernel_bytes(self, built):
        n = built.local_elements
        # th_pcg_a: z, p_old, p_new, Ap (4) + gradient image (3) + validity image (1) + D_i (1) + two uint8 edge masks (0.5)
        return {"th_pcg_a": int(4 * n * 9.5), "th_pcg_b": 4 * n * 7}


class Volumetric(SlabCase):
    """configs[3] a: examples/volumetric_mesh_deformation, 160^3 lattice (4.1 M nodes), GN 20 x 60."""
    key, energy, kind = "4a", "volumetric_mesh_deformation", "gauss_newton"
    label = "examples/volumetric_mesh_deformation 160x160x160 (4.1 M nodes)"
    dims, nit, lit = (160, 160, 160), 20, 60
    U, A, A_tiled, P, L = 6, 9, 9, 1, 0

    def inputs(self, device, rows):
        import torch
        W, H, D = self.dims
        d = wl.volumetric_inputs(W, H, D)
        sl = slice(rows[0] * W * H, rows[1] * W * H)
        t = [torch.from_numpy(np.ascontiguousarray(d[k][sl])) for k in ("Offset", "Angle", "UrShape", "Constraints")]
        return t + [_f32(d["w_fitSqrt"]), _f32(d["w_regSqrt"])], [0, 1]

    def kernel_bytes(self, built):
        n = built.local_elements
        return {"th_pcg_a": 4 * n * (4 * 6 + 9), "th_pcg_b": 4 * n * 8 * 6}


# ---------------------------------------------------------------------------------------------------- graph domain
class ArapMesh(Case):
    """configs[3] b: examples/arap_mesh_deformation on a triangulated 2000 x 2000 grid (4 M vertices, 24 M directed
    edges listed per vertex as examples/shared/ThalloGraph.h:67-79 does), GN 20 x 100."""
    key, energy, kind = "4b", "arap_mesh_deformation", "gauss_newton"
    label = "examples/arap_mesh_deformation, 2000x2000 triangulated grid (4.0 M vertices, 24 M edges)"
    n = 2000
    nit, lit = 20, 100
    partition, schedule = "graph", "gather"
    oracle_mode = "residualwise"
    U, A, P, L = 6, 9, 1, 0
    VERTEX = ("Position", "Angle", "Original", "Constraints")

    @property
    def dims(self):
        return (self.n * self.n, self._edges())

    def _edges(self):
        n = self.n        # 6-neighbourhood of a grid triangulated along one diagonal: horizontal, vertical and diagonal links, both directions
        return 2 * (n * (n - 1) + n * (n - 1) + (n - 1) * (n - 1))

    def survey_iteration_bytes(self):
        N, E = self.dims
        return 8 * E + 4 * N * ((2 * self.U + self.A) + self.U + 2 * self.U + (7 + self.P) * self.U + 3 * self.U)

    def build(self, rank=0, world=1, device="cuda", group=None, timing=1):
        import torch
        from .distributed import GraphSolver
        d = wl.arap_mesh_inputs(self.n, self.n)
        N, E = self.n * self.n, len(d["V0"])
        assert (N, E) == tuple(self.dims)
        scal = [_f32(d["w_fitSqrt"]), _f32(d["w_regSqrt"])]
        if world == 1:
            s = self.make_solver([N, E], timing)
            loc = [torch.from_numpy(np.ascontiguousarray(d[k])) for k in self.VERTEX]
            idx = [torch.from_numpy(d[k]) for k in ("V0", "V1")]
            nloc, eloc = N, E
        else:
            s = GraphSolver([N, E], self.energy, self.kind, rank, world, [d["V0"], d["V1"]], group=group, timing=timing)
            loc = [torch.from_numpy(s.vertex_rows(d[k])) for k in self.VERTEX]
            idx = [torch.from_numpy(s.index_array(d[k])) for k in ("V0", "V1")]
            nloc, eloc = s.local_dims[0], s.local_dims[1]
        params = scal + [x.to(device) for x in loc + idx]
        b = Built(s, params, [2, 3], nloc, dict(edges=eloc))
        s.set_parameters(**self.params_for_solver())
        return b

    def kernel_bytes(self, built):
        N, E, U, A = built.local_elements, built.meta["edges"], self.U, self.A
        return {"th_gather_s0": 4 * (E * 3 + N * (1 + U + A + U)), "th_pcg_b": 4 * N * 8 * U, "th_step3": 4 * N * 3 * U}


# ---------------------------------------------------------------------------------------------------- bundle adjustment
class BundleAdjustment(Case):
    """configs[4]: examples/bundle_adjustment, synthetic 10 k cameras x 5 M points x 25 M observations, sparse-materialised
    Jacobian, LM 5 x 150 with q_tolerance 0.1 and function_tolerance 0 (examples/bundle_adjustment/src/main.cpp:9-17,
    CombinedSolver.h:131-138)."""
    key, energy, kind = "5", "bundle_adjustment", "levenberg_marquardt"
    cameras, points, per_point = 10000, 5000000, 5
    nit, lit = 5, 150
    solver_params = dict(q_tolerance=0.1, function_tolerance=0.0)
    partition, schedule = "replicated", "gather"
    oracle_mode, materialized = "residualwise", True
    define_kwargs = dict(materialize=True)
    U, A, P, L = 0, 0, 1, 1

    @property
    def label(self):
        return "examples/bundle_adjustment synthetic %d cameras x %d points x %d observations, sparse-materialised J" % (
            self.cameras, self.points, self.points * self.per_point)

    @property
    def dims(self):
        return (self.cameras, self.points, self.points * self.per_point)

    def survey_iteration_bytes(self):
        """Block floor of SURVEY 8d: the stored partial derivatives once (24 scalars per observation), the two indices
        per observation, and the vector passes."""
        C_, P_, O_ = self.dims
        nunk = 9 * C_ + 3 * P_
        return 4 * 24 * O_ + 8 * O_ + 4 * nunk * (12 + self.P + 2 * self.L)

    def build(self, rank=0, world=1, device="cuda", group=None, timing=1):
        from .distributed import ReplicatedSolver
        C_, P_, O_ = self.dims
        d = wl.bundle_adjustment_inputs_torch(C_, P_, device, self.per_point)
        if world == 1:
            s = self.make_solver([C_, P_, O_], timing)
            params = [d["cameras"], d["points"], d["observations"], d["oToC"], d["oToP"]]
            pl, ol = P_, O_
        else:
            o2p = d["oToP"].cpu().numpy()
            s = ReplicatedSolver([C_, P_, O_], self.energy, self.kind, rank, world, o2p, group=group, timing=timing,
                                 define_kwargs=self.define_kwargs)
            p = s.part
            obs = p["observations"]
            oa, ob = int(obs[0]), int(obs[-1]) + 1            # oToP is sorted: a rank's observations are one contiguous run
            assert ob - oa == len(obs)
            params = [d["cameras"], d["points"][p["start"]:p["start"] + p["count"]].contiguous(), d["observations"][oa:ob].contiguous(),
                      d["oToC"][oa:ob].contiguous(), (d["oToP"][oa:ob] - p["start"]).contiguous()]
            pl, ol = p["count"], ob - oa
            del d
        b = Built(s, params, [0, 1], pl, dict(observations=ol, cameras=C_))
        s.set_parameters(**self.params_for_solver())
        return b

    def kernel_bytes(self, built):
        O_, P_, C_ = built.meta["observations"], built.local_elements, built.meta["cameras"]
        nunk = 9 * C_ + 3 * P_
        return {"th_matj_g0": 4 * O_ * (24 + 2 + 2), "th_gather_s0": 4 * O_ * (18 + 2 + 1), "th_gather_s1": 4 * (O_ * (6 + 2) + P_ * 10),
                "th_pcg_b": 4 * nunk * 9, "th_step3": 4 * nunk * 3}


CASES = {c.key: c for c in (Minimal, ImageWarping, OpticalFlow, ShapeFromShading, Volumetric, ArapMesh, BundleAdjustment)}


def case(key, **overrides):
    """The configured case `key` ("1", "2", "3a", "3b", "4a", "4b", "5"); overrides (e.g. dims=(1024, 1024), n=500,
    points=100000) give the same workload at a reduced size for tests."""
    c = CASES[key]()
    for k, v in overrides.items():
        setattr(c, k, v)
    return c
