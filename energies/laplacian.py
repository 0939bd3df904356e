"""tests/minimal/laplacian.t (reference tests/minimal/laplacian.t:1-20).

Variant "gold" (default) guards reg_x by InBounds(x+1,y): the revision that
produced tests/minimal/gold.png.  Variant "committed" guards
reg_x by InBounds(x+1,y+1) exactly as the file reads today (laplacian.t:11).
The reference file also asks for materialized J and JtJ (:16-20); that schedule
is selected with materialize=True.
"""


def define(L, variant="gold", materialize=False):
    W, H = L.Dims("W", "H")
    I = L.Inputs(
        X=L.Unknown(L.float, [W, H], 0),
        A=L.Array(L.float, [W, H], 1),
    )
    X, A = I.X, I.A
    w_fit = 0.2
    x, y = W(), H()
    guard_x = L.InBounds(x + 1, y) if variant == "gold" else L.InBounds(x + 1, y + 1)
    r = L.Residuals(
        fit=w_fit * (X(x, y) - A(x, y)),
        reg=[
            L.Select(guard_x, X(x, y) - X(x + 1, y), 0),
            L.Select(L.InBounds(x, y + 1), X(x, y) - X(x, y + 1), 0),
        ],
    )
    if materialize:
        r.fit.J.set_materialize(True)
        r.reg.J.set_materialize(True)
    return r
