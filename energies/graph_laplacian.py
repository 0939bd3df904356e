"""tests/minimal_graph/laplacian.t (reference tests/minimal_graph/laplacian.t:1-23)."""


def define(L, materialize=False, jp=False):
    N, E = L.Dims("N", "E")
    I = L.Inputs(
        X=L.Unknown(L.float, [N], 0),
        A=L.Array(L.float, [N], 1),
        v0=L.Sparse([E], [N], 2),
        v1=L.Sparse([E], [N], 3),
    )
    X, A, v0, v1 = I.X, I.A, I.v0, I.v1
    w_fit = 0.5
    n, e = N(), E()
    r = L.Residuals(
        fit=w_fit * (X(n) - A(n)),
        reg=X(v0(e)) - X(v1(e)),
    )
    if materialize:
        r.fit.J.set_materialize(True)
    if jp:
        r.reg.Jp.set_materialize(True)
    return r
