"""examples/image_warping/image_warping.t (reference :1-34): 2-D ARAP image warp.
Unknowns Offset (float2) + Angle (float) per pixel; 4 directional ARAP terms + fit."""
from ._lib import Rotate2D, All


def define(L):
    W, H = L.Dims("W", "H")
    I = L.Inputs(
        Offset=L.Unknown(L.float2, [W, H], 0),
        Angle=L.Unknown(L.float, [W, H], 1),
        UrShape=L.Array(L.float2, [W, H], 2),      # original mesh position
        Constraints=L.Array(L.float2, [W, H], 3),  # user constraints
        Mask=L.Array(L.float, [W, H], 4),          # validity mask for mesh
        w_fitSqrt=L.Param(L.float, 5),
        w_regSqrt=L.Param(L.float, 6),
    )
    Offset, Angle, UrShape, Constraints, Mask = I.Offset, I.Angle, I.UrShape, I.Constraints, I.Mask
    w_fitSqrt, w_regSqrt = I.w_fitSqrt, I.w_regSqrt
    L.UsePreconditioner(True)
    x, y = W(), H()
    Offset.Exclude(L.Not(L.eq(Mask(x, y), 0)))
    Angle.Exclude(L.Not(L.eq(Mask(x, y), 0)))

    regs = []
    for dx, dy in [(1, 0), (-1, 0), (0, 1), (0, -1)]:
        e_reg = w_regSqrt * ((Offset(x, y) - Offset(x + dx, y + dy))
                             - Rotate2D(L, Angle(x, y), (UrShape(x, y) - UrShape(x + dx, y + dy))))
        valid = L.InBounds(x + dx, y + dy) * L.eq(Mask(x, y), 0) * L.eq(Mask(x + dx, y + dy), 0)
        regs.append(L.Select(valid, e_reg, 0))
    e_fit = Offset(x, y) - Constraints(x, y)
    valid = All(L, L.greatereq(Constraints(x, y), 0)) * L.eq(Mask(x, y), 0)
    return L.Residuals(
        reg_px=regs[0],
        reg_nx=regs[1],
        reg_py=regs[2],
        reg_ny=regs[3],
        fit=w_fitSqrt * L.Select(valid, e_fit, 0.0),
    )
