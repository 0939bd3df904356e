"""Math helpers of the DSL, written once against the abstract namespace `L`
(operator-overloaded scalars + `L.Vector/L.cos/L.sin/L.sqrt/L.select/L.greater`).

Restates reference API/src/lib.t: gemv :78-90, Rotate3D :122-137, Rotate2D :138-142,
cross :243-245, AngleAxisRotatePoint :514-555, All :55-61.
"""


def gemv(L, matrix, v):
    cols = len(v)
    rows = len(matrix) // cols
    out = []
    for r in range(rows):
        val = matrix[r * cols] * v[0]
        for c in range(1, cols):
            val = val + matrix[r * cols + c] * v[c]
        out.append(val)
    return L.Vector(*out)


def Rotate2D(L, angle, v):
    c, s = L.cos(angle), L.sin(angle)
    return L.Vector(c * v[0] + (-s) * v[1], s * v[0] + c * v[1])


def Rotate3D(L, a, v):
    alpha, beta, gamma = a[0], a[1], a[2]
    ca, cb, cg = L.cos(alpha), L.cos(beta), L.cos(gamma)
    sa, sb, sg = L.sin(alpha), L.sin(beta), L.sin(gamma)
    m = [cg * cb, -sg * ca + cg * sb * sa, sg * sa + cg * sb * ca,
         sg * cb, cg * ca + sg * sb * sa, -cg * sa + sg * sb * ca,
         -sb, cb * sa, cb * ca]
    return gemv(L, m, v)


def cross(L, a, b):
    return L.Vector(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def dot(L, a, b):
    r = a[0] * b[0]
    for i in range(1, len(a)):
        r = r + a[i] * b[i]
    return r


def AngleAxisRotatePoint(L, angle_axis, pt):
    theta2 = dot(L, angle_axis, angle_axis)
    large_axis = L.greater(theta2, 1e-8)
    theta = L.sqrt(theta2)
    costheta = L.cos(theta)
    sintheta = L.sin(theta)
    theta_inverse = 1.0 / theta
    w = angle_axis * theta_inverse
    w_cross_pt = cross(L, w, pt)
    tmp = dot(L, w, pt) * (1.0 - costheta)
    large_result = pt * costheta + w_cross_pt * sintheta + w * tmp
    small_result = pt + cross(L, angle_axis, pt)
    return L.Select(large_axis, large_result, small_result)


def All(L, v):
    r = v[0]
    for i in range(1, len(v)):
        r = r * v[i]
    return r
