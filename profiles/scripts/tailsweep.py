import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
for name, kind in (("97s", d.CDF97_F32), ("53i", d.CDF53_I32)):
    for frames in (1, 4):
        ims = [d.DeviceImage(kind, 8192, 8192, frames) for _ in range(3 if frames == 1 else 2)]
        for im in ims: im.fill(0, 0, 6)
        for tail in (1024, 4096, 16384):
            L.check(L.c.dwtb200_set_tuning(1, tail))
            for _ in range(2):
                for im in ims: im.fwd2(); im.inv2(13)
            tf = ti = 0.0; reps = 6
            for _ in range(reps):
                L.c.dwtb200_timer_start()
                for im in ims: im.fwd2()
                tf += L.c.dwtb200_timer_stop_ms()
                L.c.dwtb200_timer_start()
                for im in ims: im.inv2(13)
                ti += L.c.dwtb200_timer_stop_ms()
            tf *= 1e3 / (reps * len(ims) * frames); ti *= 1e3 / (reps * len(ims) * frames)
            print(f"{name} x{frames} tail_max {tail:6d}: fwd {tf:6.1f} inv {ti:6.1f}  launches {ims[0].last_launches}", flush=True)
        for im in ims: im.close()
