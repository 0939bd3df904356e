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


# ---- small dense linear algebra on flat row-major vectors and rigid transforms (reference API/src/lib.t:196-512)
def _vec(L, comps):
    return L.Vector(*comps)


def SelectOnAll(L, predicates, val, default):                    # lib.t:196-205: `val` only where every predicate holds
    r = L.Select(predicates[-1], val, default)
    for p in reversed(predicates[:-1]):
        r = L.Select(p, r, default)
    return r


def Max(L, a, b):                                                # lib.t:282-285
    return L.Select(L.greater(a, b), a, b)


def matmul(L, a, b):                                             # lib.t:287-303: square matrices of equal size
    n = int(round(len(a) ** 0.5))
    assert n * n == len(a) == len(b), "matmul needs two square matrices of the same size"
    out = []
    for i in range(n):
        for j in range(n):
            c = a[i * n] * b[j]
            for k in range(1, n):
                c = c + a[i * n + k] * b[k * n + j]
            out.append(c)
    return _vec(L, out)


def transpose(L, m):                                             # lib.t:441-452
    n = int(round(len(m) ** 0.5))
    assert n * n == len(m), "transpose needs a square matrix"
    return _vec(L, [m[j * n + i] for i in range(n) for j in range(n)])


def rotationFromMat4(L, t):                                      # lib.t:429-435
    return _vec(L, [t[0], t[1], t[2], t[4], t[5], t[6], t[8], t[9], t[10]])


def translationFromMat4(L, t):                                   # lib.t:437-439
    return _vec(L, [t[3], t[7], t[11]])


def RotationMatrixAndTranslationToMat4(L, r, t):                 # lib.t:256-261
    return _vec(L, [r[0], r[1], r[2], t[0], r[3], r[4], r[5], t[1], r[6], r[7], r[8], t[2], 0.0, 0.0, 0.0, 1.0])


def Mat4ToRigidTransform(L, m):                                  # lib.t:263-267: the top three rows
    return _vec(L, [m[i] for i in range(12)])


def RigidTransformToMat4(L, m):                                  # lib.t:269-274
    return _vec(L, [m[i] for i in range(12)] + [0.0, 0.0, 0.0, 1.0])


def InvertRigidTransform(L, transform):                          # lib.t:454-464: [R t]^-1 = [R^T  -R^T t]
    Rt = transpose(L, rotationFromMat4(L, transform))
    nt = gemv(L, [-Rt[i] for i in range(9)], translationFromMat4(L, transform))
    return _vec(L, [Rt[0], Rt[1], Rt[2], nt[0], Rt[3], Rt[4], Rt[5], nt[1], Rt[6], Rt[7], Rt[8], nt[2], 0.0, 0.0, 0.0, 1.0])


def CameraToDepth(L, fx, fy, cx, cy, pos):                       # lib.t:276-280: pinhole projection
    return _vec(L, [pos[0] * fx / pos[2] + cx, pos[1] * fy / pos[2] + cy])


def RodriguesSO3Exp(L, w, A, B):                                 # lib.t:207-240: I + A [w]x + B [w]x^2
    x2, y2, z2 = w[0] * w[0], w[1] * w[1], w[2] * w[2]
    xy, xz, yz = B * (w[0] * w[1]), B * (w[0] * w[2]), B * (w[1] * w[2])
    ax, ay, az = A * w[0], A * w[1], A * w[2]
    return _vec(L, [1.0 - B * (y2 + z2), xy - az, xz + ay,
                    xy + az, 1.0 - B * (x2 + z2), yz - ax,
                    xz - ay, yz + ax, 1.0 - B * (x2 + y2)])


def PoseToMatrix(L, rot, trans):                                 # lib.t:466-502: exponential map of se(3), 4x4 matrix
    th2 = dot(L, rot, rot)
    th = L.sqrt(th2)
    cr = cross(L, rot, trans)
    small = L.less(th2, 1e-8)
    mid = L.less(th2, 1e-6)
    sixth, twentieth = 1.0 / 6.0, 1.0 / 20.0
    A_s, B_s = 1.0 - sixth * th2, 0.5                            # |w|^2 < 1e-8: first-order terms
    t_s = trans + 0.5 * cr
    C_m = sixth * (1.0 - twentieth * th2)                        # |w|^2 < 1e-6: Taylor expansions
    A_m = 1.0 - th2 * C_m
    B_m = 0.5 - (0.25 * sixth * th2)
    inv = 1.0 / th                                               # otherwise: closed forms
    A_l = L.sin(th) * inv
    B_l = (1.0 - L.cos(th)) * (inv * inv)
    C_l = (1.0 - A_l) * (inv * inv)
    wcr = cross(L, rot, cr)
    t_m = trans + B_m * cr + C_m * wcr
    t_l = trans + B_l * cr + C_l * wcr
    t = L.Select(small, t_s, L.Select(mid, t_m, t_l))
    A = L.Select(small, A_s, L.Select(mid, A_m, A_l))
    B = L.Select(small, B_s, L.Select(mid, B_m, B_l))
    return RotationMatrixAndTranslationToMat4(L, RodriguesSO3Exp(L, rot, A, B), t)


def rigid_trans(L, M, v):                                        # lib.t:508-512: first three rows of M [v 1]
    r = gemv(L, [M[i] for i in range(len(M))], L.Vector(v[0], v[1], v[2], 1.0))
    return _vec(L, [r[0], r[1], r[2]])
