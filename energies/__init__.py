"""Energy definitions: transcriptions of the reference's `.t` energy files into
a Python-embedded form of the same DSL (`Dims/Inputs/Unknown/Array/Sparse/
Param/Residuals/InBounds/Select/...`, reference API/src/lib.t, thallo.t:93-585).

Each module exposes `define(L)` where `L` is a DSL namespace. Two independent
implementations of `L` exist:
  * thallo_b200.frontend.dsl  -- symbolic (hash-consed AD DAG -> CUDA C++), the product
  * oracle.npdsl              -- numeric (NumPy dual numbers -> SciPy CSR J), test infrastructure
so the same workload definition drives both sides, the way the same `.t` file
drives both the reference's GPU path and its cpuOnly path.

`REGISTRY` maps the reference file a caller passes to Thallo_ProblemDefine to
the module that restates it.  When the file the caller names exists, it is read
and evaluated as written (thallo_b200/frontend/tlang.py, the `.t` front end); the
registered transcription is used only when no such file is present.
"""
import importlib

REGISTRY = {
    # basename of the reference .t file (as passed by tests/*/main.cpp and
    # examples/*/src) -> module name.  tests/minimal and tests/minimal_graph both
    # call their file "laplacian.t"; they are disambiguated by parent dir, and by
    # the explicit names below.
    "tests/minimal/laplacian.t": "laplacian",
    "tests/minimal_graph/laplacian.t": "graph_laplacian",
    "image_warping.t": "image_warping",
    "optical_flow.t": "optical_flow",
    "volumetric_mesh_deformation.t": "volumetric_mesh_deformation",
    "arap_mesh_deformation.t": "arap_mesh_deformation",
    "bundle_adjustment.t": "bundle_adjustment",
    "shape_from_shading.t": "shape_from_shading",
}


def load(name):
    """Return the `define(L)` callable for an energy module name."""
    return importlib.import_module("energies." + name).define


def resolve(path):
    """Map a reference-style path ("…/image_warping.t", "laplacian.t") or a bare
    module name to an energy module name, or None."""
    import os
    p = path.replace("\\", "/")
    if p.endswith(".py"):
        p = p[:-3]
    if p.endswith(".t"):
        parts = p.split("/")
        for k in (2, 1):
            key = "/".join(parts[-k:])
            for reg, mod in REGISTRY.items():
                if reg == key or reg.endswith("/" + key) or reg == parts[-1]:
                    return mod
        base = parts[-1][:-2]
    else:
        base = os.path.basename(p)
    try:
        importlib.import_module("energies." + base)
        return base
    except ImportError:
        return None


def define_for(energy):
    """`define(L)` and a short name for what a caller handed Thallo_ProblemDefine: an existing `.t`
    file is evaluated as written (reference thallo.t:5954-5975 loads the file the same way); otherwise
    the name is looked up among the registered transcriptions.  THALLO_B200_NO_TLANG=1 forces the
    transcriptions.  Returns (define, name) or (None, None)."""
    import os
    if energy.endswith(".t") and os.path.isfile(energy) and os.environ.get("THALLO_B200_NO_TLANG", "0") in ("", "0"):
        from thallo_b200.frontend import tlang
        return tlang.load(energy), os.path.basename(energy)[:-2]
    mod = resolve(energy)
    if mod is None:
        return None, None
    return load(mod), mod
