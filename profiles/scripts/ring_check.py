import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import libdwt_b200 as d
from oracle.orc import Oracle
from cases import DT, KINDS, bits, describe_mismatch
L = d.lib(); L.init(0); o = Oracle()
ring = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L.check(L.c.dwtb200_set_tuning(0, 0))   # stream kernels on every level above the tail
L.check(L.c.dwtb200_set_tuning(6, ring))
bad = 0
for (w, t) in KINDS:
    for (ox, oy) in ((517, 301), (300, 200), (256, 257), (1000, 333), (2048, 1536), (1999, 1201), (241, 250), (479, 33), (64, 3), (5, 1000), (4096, 4100)):
        for rows in (0, 6, 34):
            if rows and ox * oy > 1e6: continue
            L.c.dwtb200_set_strip_rows(rows)
            a = o.fill(np.zeros((oy, ox), DT[t]), t); b = a.copy()
            Ja = o.fwd2(a, w, t); Jb = d.fwd2(b, w, t)
            okf = (bits(a, t) == bits(b, t)).all()
            b[...] = a
            o.inv2(a, w, t, j_max=Ja); d.inv2(b, w, t, j_max=Ja)
            oki = (bits(a, t) == bits(b, t)).all()
            if not (okf and oki and Ja == Jb):
                bad += 1; print("BAD", w, t, ox, oy, rows, okf, oki, flush=True)
L.c.dwtb200_set_strip_rows(0)
print("parity bad =", bad, flush=True)
# batch frames
for (w, t) in KINDS[:3]:
    img = d.DeviceImage(d.kind_of(w, t), 1300, 1260, 3); img.fill(0, 0, 6); J = img.fwd2()
    for k in range(3):
        want = o.fill(np.zeros((1260, 1300), DT[t]), t, rand=k % 6); o.fwd2(want, w, t)
        got = img.download(frame=k)
        if not (bits(got, t) == bits(want, t)).all(): print("BAD batch", w, t, k)
    img.close()
L.check(L.c.dwtb200_set_tuning(0, 2048 * 2048))
# timing
for kind, name in ((d.CDF97_F32, "97s"), (d.CDF53_I32, "53i"), (d.CDF97_F64, "97d")):
    for frames in (4, 1):
        n = 8192 if kind != d.CDF97_F64 else 4096
        ims = [d.DeviceImage(kind, n, n, frames) for _ in range(2 if frames == 4 else 4)]
        for im in ims: im.fill(0, 0, 6)
        for r in (0, ring):
            L.check(L.c.dwtb200_set_tuning(6, r))
            for J in (1, -1):
                for _ in range(2):
                    for im in ims: im.fwd2(J); im.inv2(J)
                tf = ti = 0.0; reps = 5
                for _ in range(reps):
                    L.c.dwtb200_timer_start()
                    for im in ims: im.fwd2(J)
                    tf += L.c.dwtb200_timer_stop_ms()
                    L.c.dwtb200_timer_start()
                    for im in ims: im.inv2(J)
                    ti += L.c.dwtb200_timer_stop_ms()
                tf /= reps * len(ims) * frames; ti /= reps * len(ims) * frames
                es = 8 if kind == d.CDF97_F64 else 4
                alg = 2 * es * n * n * (1 if J == 1 else 4 / 3)
                print(f"{name} frames {frames} ring {r} J {J:2d}: fwd {tf*1e3:7.1f} us/img {alg/tf/1e6:6.0f} GB/s   inv {ti*1e3:7.1f} us/img {alg/ti/1e6:6.0f} GB/s", flush=True)
        for im in ims: im.close()
