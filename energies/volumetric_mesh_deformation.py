"""examples/volumetric_mesh_deformation/volumetric_mesh_deformation.t (reference :1-26):
3-D lattice ARAP, Offset/Angle float3 per node, 6-neighbourhood."""
from ._lib import Rotate3D


def define(L):
    W, H, D = L.Dims("W", "H", "D")
    I = L.Inputs(
        Offset=L.Unknown(L.float3, [W, H, D], 0),
        Angle=L.Unknown(L.float3, [W, H, D], 1),
        UrShape=L.Array(L.float3, [W, H, D], 2),
        Constraints=L.Array(L.float3, [W, H, D], 3),
        w_fitSqrt=L.Param(L.float, 4),
        w_regSqrt=L.Param(L.float, 5),
    )
    Offset, Angle, UrShape, Constraints = I.Offset, I.Angle, I.UrShape, I.Constraints
    L.UsePreconditioner(True)
    w, h, d = W(), H(), D()
    e_fit = Offset(w, h, d) - Constraints(w, h, d)
    valid = L.greatereq(Constraints(w, h, d), -999999.9)
    reg = []
    for i, j, k in [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]:
        ow, oh, od = w + i, h + j, d + k
        arap = (Offset(w, h, d) - Offset(ow, oh, od)) - Rotate3D(L, Angle(w, h, d), UrShape(w, h, d) - UrShape(ow, oh, od))
        arapF = L.Select(L.InBounds(w, h, d), L.Select(L.InBounds(ow, oh, od), arap, 0.0), 0.0)
        reg.append(I.w_regSqrt * arapF)
    return L.Residuals(fit=L.Select(valid, I.w_fitSqrt * e_fit, 0), reg=reg)
