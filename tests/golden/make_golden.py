"""Generate the committed golden fixtures from the reference's own known-answer files.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py
Writes tests/golden/kat_minimal.npz and tests/golden/kat_minimal_graph.npz holding
the decoded 8-bit golden images (reference tests/minimal/gold.png 512x512 and
tests/minimal_graph/gold.png 512x1) as uint8 arrays.  The inputs that produced them
are regenerated at test time from the MSVC rand() LCG.
"""
import os
import numpy as np
from PIL import Image

REF = "/root/reference/tests"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    for name, sub in (("kat_minimal", "minimal"), ("kat_minimal_graph", "minimal_graph")):
        im = np.array(Image.open(os.path.join(REF, sub, "gold.png")))
        assert im.dtype == np.uint8 and im.ndim == 2, (im.dtype, im.shape)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), gold=im,
                            source=f"reference tests/{sub}/gold.png")
        print(name, im.shape, int(im.sum()))


if __name__ == "__main__":
    main()
