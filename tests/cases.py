"""Shared parity cases: (wavelet, type) x shape x flags, all inputs closed-form (no RNG)."""
import hashlib

import numpy as np

# the three north-star transforms, then the sibling drivers sharing their kernels
KINDS = [("97", "s"), ("97", "d"), ("53", "i"), ("53", "s"), ("53", "d"), ("97", "i")]
DT = {"s": np.float32, "d": np.float64, "i": np.int32}
UT = {"s": np.uint32, "d": np.uint64, "i": np.uint32}

# (ox, oy) outer sizes exercised densely (outer == inner)
DENSE_SHAPES = [
    (512, 512), (517, 301), (1000, 37), (64, 3), (5, 1000), (2, 2), (3, 7), (1, 1), (1, 9), (9, 1), (2, 5),
    (4, 4), (5, 5), (8, 6), (16, 16), (31, 33), (128, 128), (129, 127), (240, 240), (241, 250), (479, 33),
    (256, 256), (300, 200), (720, 486), (1024, 1024), (1025, 1023), (960, 1080), (1920, 1080),
]
# (j_max, decompose_one)
DEPTHS = [(-1, 0), (1, 0), (3, 0), (-1, 1)]

# sparse: (ox, oy, ix, iy)
SPARSE = [(64, 64, 50, 37), (128, 96, 128, 50), (100, 80, 33, 80), (517, 301, 500, 280), (40, 40, 1, 1), (33, 70, 17, 5)]


def dense_cases():
    out = []
    for k in KINDS:
        for (ox, oy) in DENSE_SHAPES:
            for (j, d1) in DEPTHS:
                if (j, d1) != (-1, 0) and max(ox, oy) > 600 and (ox, oy) not in ((1000, 37), (5, 1000)):
                    continue
                out.append((k[0], k[1], ox, oy, j, d1))
    return out


def sparse_cases():
    out = []
    for k in KINDS:
        for s in SPARSE:
            for (j, d1) in [(-1, 0), (2, 0), (-1, 1)]:
                for zp in (0, 1):
                    out.append((k[0], k[1]) + s + (j, d1, zp))
    return out


# interleaved in-place family: (wavelet, ox, oy, ix, iy, j_max, decompose_one), float32 only
INPLACE_SHAPES = [
    (512, 512), (517, 301), (1000, 37), (64, 3), (5, 1000), (2, 2), (3, 7), (1, 1), (1, 9), (9, 1), (2, 5), (4, 4), (5, 5), (6, 7), (8, 6),
    (16, 16), (31, 33), (33, 32), (40, 100), (128, 128), (129, 127), (241, 250), (256, 256), (300, 200), (720, 486), (1025, 1023), (1920, 1080),
]


def inplace_cases():
    out = []
    for w in ("97", "53"):
        for (ox, oy) in INPLACE_SHAPES:
            for (j, d1) in DEPTHS:
                if (j, d1) != (-1, 0) and max(ox, oy) > 600 and (ox, oy) not in ((1000, 37), (5, 1000)):
                    continue
                out.append((w, ox, oy, ox, oy, j, d1))
        for (ox, oy, ix, iy) in [(64, 64, 50, 37), (128, 96, 128, 50), (300, 280, 255, 270), (40, 40, 1, 1), (33, 70, 17, 5)]:
            for (j, d1) in [(-1, 0), (-1, 1)]:
                out.append((w, ox, oy, ix, iy, j, d1))
    return out


def inplace_id(c):
    return "ip-" + "-".join(str(v) for v in c)


def case_id(c):
    return "-".join(str(v) for v in c)


def bits(a, t):
    return np.ascontiguousarray(a).view(UT[t])


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def describe_mismatch(got, want, t, limit=8):
    """Human-readable summary of where two same-shape images differ bit-wise."""
    g, w = bits(got, t), bits(want, t)
    bad = np.argwhere(g != w)
    if len(bad) == 0:
        return "identical"
    ys, xs = bad[:, 0], bad[:, 1]
    lines = [f"{len(bad)} / {g.size} samples differ; y in [{ys.min()},{ys.max()}], x in [{xs.min()},{xs.max()}]"]
    for (y, x) in bad[:limit]:
        lines.append(f"  [y={y}, x={x}] got {got[y, x]!r} want {want[y, x]!r}")
    if t != "i":
        with np.errstate(all="ignore"):
            d = np.abs(got.astype(np.float64) - want.astype(np.float64))
            lines.append(f"  max abs diff {np.nanmax(d):.3e}")
    return "\n".join(lines)
