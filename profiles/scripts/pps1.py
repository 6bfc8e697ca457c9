import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
for name, kind in (("97s", d.CDF97_F32), ("53i", d.CDF53_I32)):
    ims = [d.DeviceImage(kind, 8192, 8192, 1) for _ in range(4)]
    for im in ims: im.fill(0, 0, 6)
    for pps in (10, 12, 14, 15, 16, 17, 18, 19, 20, 22, 24, 28, 32, 37, 40):
        L.c.dwtb200_set_strip_rows(2 * pps)
        for _ in range(2):
            for im in ims: im.fwd2(1); im.inv2(1)
        tf = ti = 0.0; reps = 6
        for _ in range(reps):
            L.c.dwtb200_timer_start()
            for im in ims: im.fwd2(1)
            tf += L.c.dwtb200_timer_stop_ms()
            L.c.dwtb200_timer_start()
            for im in ims: im.inv2(1)
            ti += L.c.dwtb200_timer_stop_ms()
        tf *= 1e3 / (reps * len(ims)); ti *= 1e3 / (reps * len(ims))
        nstr = -(-4096 // pps)
        print(f"{name} pps {pps:3d} ({nstr*5} CTAs = {nstr*5/296:.2f} waves): fwd {tf:6.1f} inv {ti:6.1f}", flush=True)
    L.c.dwtb200_set_strip_rows(0)
    for im in ims: im.close()
