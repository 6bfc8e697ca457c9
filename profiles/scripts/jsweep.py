import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "97s"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
for a in sys.argv[4:]:
    k, v = a.split('=')
    L.check(L.c.dwtb200_set_tuning(int(k), int(v)))
kind = {"97s": d.CDF97_F32, "53i": d.CDF53_I32, "97d": d.CDF97_F64}[name]
ims = [d.DeviceImage(kind, n, n, frames) for _ in range(3)]
for im in ims: im.fill(0, 0, 6)
pf = pi = 0
Jmax = L.c.dwtb200_ceil_log2(n)
for J in range(1, Jmax + 1):
    for _ in range(2):
        for im in ims: im.fwd2(J); im.inv2(J)
    tf = ti = 0.0; reps = 6
    for _ in range(reps):   # one image at a time: independent images would overlap on their own streams
        for im in ims:
            L.c.dwtb200_timer_start(); im.fwd2(J); tf += L.c.dwtb200_timer_stop_ms()
        for im in ims:
            L.c.dwtb200_timer_start(); im.inv2(J); ti += L.c.dwtb200_timer_stop_ms()
    tf *= 1e3 / (reps * len(ims) * frames); ti *= 1e3 / (reps * len(ims) * frames)
    print(f"{name} x{frames} n={n} J={J:2d} launches={ims[0].last_launches:2d} fwd {tf:7.1f} (+{tf-pf:6.1f})  inv {ti:7.1f} (+{ti-pi:6.1f})", flush=True)
    pf, pi = tf, ti
