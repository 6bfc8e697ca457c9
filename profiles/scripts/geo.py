import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
def run(name, frames, n=8192):
    kind = {"97s": d.CDF97_F32, "53i": d.CDF53_I32}[name]
    ims = [d.DeviceImage(kind, n, n, frames) for _ in range(3 if frames == 1 else 2)]
    for im in ims: im.fill(0, 0, 6)
    res = {}
    for cfg in CFGS:
        waves, pmin, pmax = cfg
        L.check(L.c.dwtb200_set_tuning(97, waves | (pmin << 8) | (pmax << 16)))
        for _ in range(2):
            for im in ims: im.fwd2(); im.inv2(13)
        tf = ti = 0.0; reps = 5
        for _ in range(reps):
            L.c.dwtb200_timer_start()
            for im in ims: im.fwd2()
            tf += L.c.dwtb200_timer_stop_ms()
            L.c.dwtb200_timer_start()
            for im in ims: im.inv2(13)
            ti += L.c.dwtb200_timer_stop_ms()
        tf *= 1e3 / (reps * len(ims) * frames); ti *= 1e3 / (reps * len(ims) * frames)
        print(f"{name} x{frames} waves={waves} pps=[{pmin},{pmax}]: fwd {tf:6.1f} inv {ti:6.1f}", flush=True)
    for im in ims: im.close()
CFGS = [(5, 16, 32), (5, 8, 32), (5, 12, 24), (5, 24, 48), (3, 16, 32), (8, 16, 32), (8, 8, 16), (3, 8, 64), (5, 16, 16), (5, 32, 32), (10, 12, 32)]
for name in ("97s", "53i"):
    for frames in (1, 4):
        run(name, frames)
