"""The DSL's math library (energies/_lib.py, restating reference API/src/lib.t) evaluated on plain numbers against
NumPy / SciPy: rotations, the se(3) exponential map in all three of its branches, rigid-transform helpers."""
import math

import numpy as np
import pytest
from scipy.linalg import expm

from energies import _lib


class NumL:
    """The DSL namespace over Python floats."""

    class V(list):
        def _b(self, o, f):
            if isinstance(o, list):
                return NumL.V(f(a, b) for a, b in zip(self, o))
            return NumL.V(f(a, o) for a in self)
        def __add__(self, o): return self._b(o, lambda a, b: a + b)
        def __radd__(self, o): return self._b(o, lambda a, b: b + a)
        def __sub__(self, o): return self._b(o, lambda a, b: a - b)
        def __mul__(self, o): return self._b(o, lambda a, b: a * b)
        def __rmul__(self, o): return self._b(o, lambda a, b: b * a)
        def __neg__(self): return NumL.V(-a for a in self)

    def Vector(self, *c): return NumL.V(float(x) for x in c)
    def sqrt(self, x): return math.sqrt(x)
    def sin(self, x): return math.sin(x)
    def cos(self, x): return math.cos(x)
    def less(self, a, b): return a < b
    def greater(self, a, b): return a > b

    def Select(self, c, a, b):
        return a if c else b


L = NumL()


def _hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], float)


@pytest.mark.parametrize("scale", [3e-5, 5e-4, 0.3, 2.5])          # |w|^2 < 1e-8, < 1e-6, ordinary, large
def test_pose_to_matrix_is_the_se3_exponential(scale):
    rs = np.random.RandomState(int(scale * 1e6) % 1000)
    w = rs.randn(3); w *= scale / np.linalg.norm(w)
    t = rs.randn(3)
    M = np.array(_lib.PoseToMatrix(L, L.Vector(*w), L.Vector(*t))).reshape(4, 4)
    twist = np.zeros((4, 4)); twist[:3, :3] = _hat(w); twist[:3, 3] = t
    assert np.abs(M - expm(twist)).max() < 1e-9


def test_rotations_and_rigid_helpers():
    rs = np.random.RandomState(0)
    a = rs.randn(3)
    ca, cb, cg = np.cos(a); sa, sb, sg = np.sin(a)
    Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]]); Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
    Rz = np.array([[cg, -sg, 0], [sg, cg, 0], [0, 0, 1]])
    v = rs.randn(3)
    assert np.allclose(_lib.Rotate3D(L, L.Vector(*a), L.Vector(*v)), Rz @ Ry @ Rx @ v)
    assert np.allclose(_lib.Rotate2D(L, 0.7, L.Vector(1.0, 2.0)), [math.cos(.7) - 2 * math.sin(.7), math.sin(.7) + 2 * math.cos(.7)])
    aa = rs.randn(3)
    th = np.linalg.norm(aa)
    R = expm(_hat(aa))
    assert np.allclose(_lib.AngleAxisRotatePoint(L, L.Vector(*aa), L.Vector(*v)), R @ v) and th > 1e-4
    A, B = rs.randn(4, 4), rs.randn(4, 4)
    assert np.allclose(np.array(_lib.matmul(L, L.Vector(*A.reshape(-1)), L.Vector(*B.reshape(-1)))).reshape(4, 4), A @ B)
    assert np.allclose(np.array(_lib.transpose(L, L.Vector(*A.reshape(-1)))).reshape(4, 4), A.T)
    M = np.eye(4); M[:3, :3] = R; M[:3, 3] = rs.randn(3)
    Mi = np.array(_lib.InvertRigidTransform(L, L.Vector(*M.reshape(-1)))).reshape(4, 4)
    assert np.allclose(Mi @ M, np.eye(4))
    assert np.allclose(_lib.rigid_trans(L, L.Vector(*M.reshape(-1)), L.Vector(*v)), M[:3, :3] @ v + M[:3, 3])
    assert np.allclose(_lib.rotationFromMat4(L, L.Vector(*M.reshape(-1))), R.reshape(-1))
    assert np.allclose(_lib.translationFromMat4(L, L.Vector(*M.reshape(-1))), M[:3, 3])
    assert np.allclose(_lib.CameraToDepth(L, 500.0, 510.0, 320.0, 240.0, L.Vector(0.2, -0.1, 2.0)), [370.0, 214.5])
    assert _lib.Max(L, 2.0, 3.0) == 3.0 and _lib.SelectOnAll(L, [True, True], 5.0, -1.0) == 5.0
    assert _lib.SelectOnAll(L, [True, False, True], 5.0, -1.0) == -1.0
    assert np.allclose(_lib.cross(L, L.Vector(*a), L.Vector(*v)), np.cross(a, v)) and np.isclose(_lib.dot(L, list(a), list(v)), a @ v)
