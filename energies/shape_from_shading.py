"""examples/shape_from_shading/shape_from_shading.t (reference :1-119): depth refinement with a
spherical-harmonics shading term.  Two expressions are fetched through `:get()` (:79-80, :100)
and therefore become ComputedArrays: the per-pixel shading error B_I (with a gradient image over
the three depth samples it reads) and the regulariser's validity mask."""

DEPTH_DISCONTINUITY_THRE = 0.01


def define(L):
    W, H = L.Dims("W", "H")
    I_ = L.Inputs(
        w_p=L.Param(L.float, 0), w_s=L.Param(L.float, 1), w_g=L.Param(L.float, 2),
        f_x=L.Param(L.float, 3), f_y=L.Param(L.float, 4), u_x=L.Param(L.float, 5), u_y=L.Param(L.float, 6),
        L_1=L.Param(L.float, 7), L_2=L.Param(L.float, 8), L_3=L.Param(L.float, 9), L_4=L.Param(L.float, 10),
        L_5=L.Param(L.float, 11), L_6=L.Param(L.float, 12), L_7=L.Param(L.float, 13), L_8=L.Param(L.float, 14),
        L_9=L.Param(L.float, 15),
        X=L.Unknown(L.float, [W, H], 16),
        D_i=L.Array(L.float, [W, H], 17),
        Im=L.Array(L.float, [W, H], 18),
        edgeMaskR=L.Array(L.uint8, [W, H], 19),
        edgeMaskC=L.Array(L.uint8, [W, H], 20),
    )
    X, D_i, Im, edgeMaskR, edgeMaskC = I_.X, I_.D_i, I_.Im, I_.edgeMaskR, I_.edgeMaskC
    f_x, f_y, u_x, u_y = I_.f_x, I_.f_y, I_.u_x, I_.u_y
    w_p, w_s, w_g = L.sqrt(I_.w_p), L.sqrt(I_.w_s), L.sqrt(I_.w_g)
    x, y = W(), H()
    posX, posY = x.asvalue(), y.asvalue()

    def p(offX, offY):                       # equation 8
        d = X(x + offX, y + offY)
        i = offX + posX
        j = offY + posY
        return L.Vector(((i - u_x) / f_x) * d, ((j - u_y) / f_y) * d, d)

    def normalAt(offX, offY):                # equation 10
        i = offX + posX
        j = offY + posY
        _x, _y = x + offX, y + offY
        n_x = X(_x, _y - 1) * (X(_x, _y) - X(_x - 1, _y)) / f_y
        n_y = X(_x - 1, _y) * (X(_x, _y) - X(_x, _y - 1)) / f_x
        n_z = (n_x * (u_x - i) / f_x) + (n_y * (u_y - j) / f_y) - (X(_x - 1, _y) * X(_x, _y - 1) / (f_x * f_y))
        sqLength = n_x * n_x + n_y * n_y + n_z * n_z
        inverseMagnitude = L.Select(L.greater(sqLength, 0.0), 1.0 / L.sqrt(sqLength), 1.0)
        return inverseMagnitude * L.Vector(n_x, n_y, n_z)

    def B(offX, offY):
        normal = normalAt(offX, offY)
        n_x, n_y, n_z = normal[0], normal[1], normal[2]
        return (I_.L_1 + I_.L_2 * n_y + I_.L_3 * n_z + I_.L_4 * n_x + I_.L_5 * n_x * n_y + I_.L_6 * n_y * n_z
                + I_.L_7 * (-n_x * n_x - n_y * n_y + 2 * n_z * n_z) + I_.L_8 * n_z * n_x + I_.L_9 * (n_x * n_x - n_y * n_y))

    def I(offX, offY):
        return Im(x + offX, y + offY) * 0.5 + 0.25 * (Im(x + offX - 1, y + offY) + Im(x + offX, y + offY - 1))

    def DepthValid(offX, offY):
        return L.greater(D_i(x + offX, y + offY), 0)

    bi = B(0, 0) - I(0, 0)
    B_I_comp = L.Select(DepthValid(-1, 0) * DepthValid(0, 0) * DepthValid(0, -1), bi, 0)

    def B_I(offX, offY):
        return B_I_comp.get(x + offX, y + offY)

    # fitting term
    E_p = L.Select(DepthValid(0, 0), w_p * (X(x, y) - D_i(x, y)), 0)

    # shading term
    E_g_h = (B_I(0, 0) - B_I(1, 0)) * edgeMaskR(x, y)
    E_g_v = (B_I(0, 0) - B_I(0, 1)) * edgeMaskC(x, y)
    E_g_h = L.Select(L.InBoundsExpanded(x, y, 1), w_g * E_g_h, 0)
    E_g_v = L.Select(L.InBoundsExpanded(x, y, 1), w_g * E_g_v, 0)

    # regularization term
    def Continuous(offX, offY):
        return L.less(L.abs(X(x, y) - X(x + offX, y + offY)), DEPTH_DISCONTINUITY_THRE)

    valid = (DepthValid(0, 0) * DepthValid(0, -1) * DepthValid(0, 1) * DepthValid(-1, 0) * DepthValid(1, 0)
             * Continuous(0, -1) * Continuous(0, 1) * Continuous(-1, 0) * Continuous(1, 0))
    valid = L.eq(valid.get(x, y), 1)
    E_s = 4.0 * p(0, 0) - (p(-1, 0) + p(0, -1) + p(1, 0) + p(0, 1))
    E_s = L.Select(valid, w_s * E_s, 0)

    return L.Residuals(fit=E_p, shading_h=E_g_h, shading_v=E_g_v, reg=E_s)
