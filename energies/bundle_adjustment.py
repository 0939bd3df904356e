"""examples/bundle_adjustment/bundle_adjustment.t (reference :1-37): Snavely
reprojection error; two unknown index spaces (cameras float9, points float3)."""
from ._lib import AngleAxisRotatePoint, dot


def define(L, materialize=True):
    C, P, O = L.Dims("C", "P", "O")
    I = L.Inputs(
        cameras=L.Unknown(L.float9, [C], 0),
        points=L.Unknown(L.float3, [P], 1),
        observations=L.Array(L.float2, [O], 2),
        oToC=L.Sparse([O], [C], 3),
        oToP=L.Sparse([O], [P], 4),
    )
    L.UsePreconditioner(True)
    o = O()
    camera, point = I.cameras(I.oToC(o)), I.points(I.oToP(o))
    p = AngleAxisRotatePoint(L, camera.slice(0, 3), point)
    p = p + camera.slice(3, 6)
    cod = L.Vector(-p[0] / p[2], -p[1] / p[2])
    l1, l2 = camera[7], camera[8]
    r2 = dot(L, cod, cod)
    distortion = 1.0 + r2 * (l1 + l2 * r2)
    focal = camera[6]
    predicted = cod * focal * distortion
    observed = I.observations(o)
    r = L.Residuals(snavely_reprojection_error=observed - predicted)
    if materialize:
        r.snavely_reprojection_error.J.set_materialize(True)
    return r
