"""Ring kernel generations side by side: default (gen 2 where it applies) vs DWTB200_TUNE_RING = 3 | 5 << 4 (gen 1, 7 x 2 forced).
python cmp_gen.py KIND N FRAMES [J ...]  -- batches timed with the global timer per call, singles with marks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
name, n, frames = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
Js = [int(a) for a in sys.argv[4:]]
kind = {"97s": d.CDF97_F32, "53i": d.CDF53_I32, "97d": d.CDF97_F64, "53s": d.CDF53_F32}[name]
nimg = 3 if frames * n * n * 4 * 2.4 * 3 < 60e9 else 1
ims = [d.DeviceImage(kind, n, n, frames) for _ in range(nimg)]
for im in ims: im.fill(0, 0, 6)
Jmax = L.c.dwtb200_ceil_log2(n)
for J in (Js or [Jmax]):
    out = []
    for ring in (3, 3 | (5 << 4)):
        L.check(L.c.dwtb200_set_tuning(6, ring))
        for _ in range(2):
            for im in ims: im.fwd2(J); im.inv2(J)
        L.check(L.c.dwtb200_sync())
        reps = 6
        prev = ims[-1]
        for _ in range(reps):
            for im in ims:
                im.wait_for(prev); im.mark(); im.fwd2(J); im.mark(); prev = im
            for im in ims:
                im.wait_for(prev); im.mark(); im.inv2(J); im.mark(); prev = im
        tf = ti = 0.0
        for im in ims:
            t = im.read_marks()
            for r in range(reps):
                tf += t[4 * r]; ti += t[4 * r + 2]
        k = 1e3 / (reps * len(ims) * frames)
        out.append((tf * k, ti * k))
    print(f"{name} n={n} x{frames} J={J:2d}  gen2 fwd {out[0][0]:7.2f} inv {out[0][1]:7.2f}   gen1 fwd {out[1][0]:7.2f} inv {out[1][1]:7.2f}  us per frame", flush=True)
