"""examples/optical_flow/optical_flow.t (reference :1-36): dense flow, bilinear
SampledImage with analytic x/y derivative images, preconditioner off (:17)."""


def define(L):
    W, H = L.Dims("W", "H")
    I_ = L.Inputs(
        w_fitSqrt=L.Param(L.float, 0),
        w_regSqrt=L.Param(L.float, 1),
        X=L.Unknown(L.float2, [W, H], 2),
        I=L.Array(L.float, [W, H], 3),
        I_hat_im=L.Array(L.float, [W, H], 4),
        I_hat_dx=L.Array(L.float, [W, H], 5),
        I_hat_dy=L.Array(L.float, [W, H], 6),
    )
    X, I = I_.X, I_.I
    I_hat = L.SampledImage(I_.I_hat_im, I_.I_hat_dx, I_.I_hat_dy)
    x, y = W(), H()
    i, j = x.asvalue(), y.asvalue()
    L.UsePreconditioner(False)
    e_fit = I_.w_fitSqrt * (I(x, y) - I_hat(i + X(x, y)[0], j + X(x, y)[1]))
    reg = []
    for ox, oy in [(1, 0), (-1, 0), (0, 1), (0, -1)]:
        nx, ny = x + ox, y + oy
        e_reg = I_.w_regSqrt * (X(x, y) - X(nx, ny))
        reg.append(L.Select(L.InBounds(nx, ny), e_reg, 0))
    return L.Residuals(fit=e_fit, reg_px=reg[0], reg_nx=reg[1], reg_py=reg[2], reg_ny=reg[3])
