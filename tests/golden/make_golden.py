"""Generates tests/golden/golden.json and tests/golden/vectors.npz from the COMPILED REFERENCE
(oracle/_ref/libdwt_ref.so, built by oracle/Makefile from /root/reference).  Run in the build
container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

golden.json   sha256 digests of the reference's forward coefficients and round-trip output for every
              case of tests/cases.py (inputs: the reference's own dwt_util_test_image_fill2_* patterns)
vectors.npz   a few small complete input/forward/round-trip arrays (known-answer vectors)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from cases import DT, case_id, dense_cases, digest, inplace_cases, inplace_id, sparse_cases  # noqa: E402
from oracle.orc import Ref  # noqa: E402


def run(ref, c, keep=None):
    if len(c) == 6:
        w, t, ox, oy, j, d1 = c
        ix, iy, zp = ox, oy, 0
    else:
        w, t, ox, oy, ix, iy, j, d1, zp = c
    img = np.zeros((oy, ox), dtype=DT[t])
    ref.fill(img, t, rand=0, type_=0)
    inp = img.copy()
    J = ref.fwd2(img, w, t, j_max=j, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
    fwd = img.copy()
    ref.inv2(img, w, t, j_max=J, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
    if keep is not None:
        keep[case_id(c) + "/in"] = inp
        keep[case_id(c) + "/fwd"] = fwd
        keep[case_id(c) + "/inv"] = img.copy()
    return {"J": J, "fwd": digest(fwd), "inv": digest(img)}


def main():
    ref = Ref()
    ref.set_threads(8)
    out, keep = {}, {}
    small = {("97", "s", 31, 33, -1, 0), ("97", "d", 31, 33, -1, 0), ("53", "i", 31, 33, -1, 0),
             ("97", "s", 64, 3, -1, 1), ("53", "i", 5, 5, -1, 0), ("97", "s", 16, 16, -1, 0),
             ("97", "s", 64, 64, 50, 37, -1, 0, 1), ("53", "i", 64, 64, 50, 37, -1, 1, 0)}
    for c in dense_cases() + sparse_cases():
        out[case_id(c)] = run(ref, c, keep if c in small else None)
    # type-2 pattern (overflow free) at one mid size, int and float
    for (w, t) in (("53", "i"), ("97", "s")):
        img = np.zeros((301, 517), dtype=DT[t])
        ref.fill(img, t, rand=0, type_=2)
        J = ref.fwd2(img, w, t)
        fwd = digest(img)
        ref.inv2(img, w, t, j_max=J)
        out[f"type2-{w}-{t}-517-301"] = {"J": J, "fwd": fwd, "inv": digest(img)}
    # 3-D, one level (src/volume-dwt.c:727, 1115)
    for (nx, ny, nz) in ((16, 16, 16), (33, 20, 9), (64, 48, 40), (5, 5, 5)):
        a = np.zeros((nz, ny, nx), dtype=np.float32)
        ref.volume_fill(a)
        b = np.zeros_like(a)
        ref.fwd3(a, b)
        fwd = digest(b)
        if (nx, ny, nz) == (16, 16, 16):
            keep["vol-16-16-16/in"] = a.copy()
            keep["vol-16-16-16/fwd"] = b.copy()
        ref.inv3(b)
        out[f"vol-{nx}-{ny}-{nz}"] = {"fwd": fwd, "inv": digest(b)}
    # interleaved in-place family (src/libdwt.c:12926, 17474, 16553, 17886)
    for c in inplace_cases():
        w, ox, oy, ix, iy, j, d1 = c
        img = np.zeros((oy, ox), dtype=np.float32)
        ref.fill(img, "s", rand=0, type_=0)
        J = ref.fwd2_inplace(img, w, j_max=j, decompose_one=d1, inner=(iy, ix))
        fwd = img.copy()
        ref.inv2_inplace(img, w, j_max=J, decompose_one=d1, inner=(iy, ix))
        out[inplace_id(c)] = {"J": J, "fwd": digest(fwd), "inv": digest(img)}
        if c in (("97", 31, 33, 31, 33, -1, 0), ("97", 64, 64, 50, 37, -1, 0), ("53", 31, 33, 31, 33, -1, 0)):
            keep[inplace_id(c) + "/fwd"] = fwd
            keep[inplace_id(c) + "/inv"] = img.copy()
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "vectors.npz"), **keep)
    print(f"{len(out)} digests, {len(keep)} vectors")


if __name__ == "__main__":
    main()
