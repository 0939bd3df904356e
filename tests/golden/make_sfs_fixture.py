"""Makes tests/golden/sfs_crop.npz: a 96x80 crop of the reference's own shape_from_shading input
(examples/data/shape_from_shading/default_*.imagedump + default.SFSSolverParameters), read the way
examples/shape_from_shading/src/SimpleBuffer.cpp:12-40 and TerraSolverParameters.h do (-inf depth
-> -10000).  The principal point is moved by the crop origin so that the geometry is unchanged.
Run in the build container (needs /root/reference):  python tests/golden/make_sfs_fixture.py"""
import os
import struct

import numpy as np

BASE = "/root/reference/examples/data/shape_from_shading/default"
X0, Y0, W, H = 480, 200, 96, 80


def load(fn):
    b = open(fn, "rb").read()
    w, h, c, t = struct.unpack("<4i", b[:16])
    a = np.frombuffer(b[16:], dtype=np.float32 if t == 0 else np.uint8).copy()
    if t == 0:
        a[np.isposinf(a)] = np.finfo(np.float32).max
        a[np.isneginf(a)] = -10000.0
    return a.reshape(h, w)


def main():
    D, Im, X = load(BASE + "_targetDepth.imagedump"), load(BASE + "_targetIntensity.imagedump"), load(BASE + "_initialUnknown.imagedump")
    M = load(BASE + "_maskEdgeMap.imagedump")
    h = D.shape[0]
    mr, mc = M[:h], M[h:]
    p = struct.unpack("<36f4I", open(BASE + ".SFSSolverParameters", "rb").read())
    crop = lambda a: np.ascontiguousarray(a[Y0:Y0 + H, X0:X0 + W]).reshape(-1)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sfs_crop.npz")
    np.savez_compressed(out, W=W, H=H, w_p=p[0], w_s=p[1], w_g=p[3], f_x=p[7], f_y=p[8], u_x=p[9] - X0, u_y=p[10] - Y0,
                        light=np.array(p[27:36], np.float32), X=crop(X), D_i=crop(D), Im=crop(Im), edgeMaskR=crop(mr), edgeMaskC=crop(mc))
    print("wrote", out, os.path.getsize(out), "bytes; valid fraction", float((crop(D) > 0).mean()))


if __name__ == "__main__":
    main()
