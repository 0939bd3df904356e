"""examples/arap_mesh_deformation/arap_mesh_deformation.t (reference :1-24):
graph-domain ARAP: N vertices (Position, Angle float3), E directed edges."""
from ._lib import Rotate3D


def define(L, jp=False):
    N, E = L.Dims("N", "E")
    I = L.Inputs(
        w_fitSqrt=L.Param(L.float, 0),
        w_regSqrt=L.Param(L.float, 1),
        Position=L.Unknown(L.float3, [N], 2),
        Angle=L.Unknown(L.float3, [N], 3),
        Original=L.Array(L.float3, [N], 4),
        Constraints=L.Array(L.float3, [N], 5),
        V0=L.Sparse([E], [N], 6),
        V1=L.Sparse([E], [N], 7),
    )
    Position, Angle, Original, Constraints = I.Position, I.Angle, I.Original, I.Constraints
    L.UsePreconditioner(True)
    n, e = N(), E()
    v0, v1 = I.V0(e), I.V1(e)
    e_fit = Position(n) - Constraints(n)
    valid = L.greatereq(Constraints(n)[0], -999999.9)
    arap = (Position(v0) - Position(v1)) - Rotate3D(L, Angle(v0), Original(v0) - Original(v1))
    r = L.Residuals(fit=L.Select(valid, I.w_fitSqrt * e_fit, 0), reg=I.w_regSqrt * arap)
    if jp:          # schedule Jt[Jp] for the edge term (`r.reg.Jp:set_materialize(true)`, thallo.t:4121)
        r.reg.Jp.set_materialize(True)
    return r
