import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
ring = int(sys.argv[1])
L.check(L.c.dwtb200_set_tuning(6, ring))
for name, kind in (("97s", d.CDF97_F32), ("53i", d.CDF53_I32)):
    for frames in (4, 1):
        ims = [d.DeviceImage(kind, 8192, 8192, frames) for _ in range(2 if frames == 4 else 4)]
        for im in ims: im.fill(0, 0, 6)
        for rows in (0, 32, 64, 100, 128, 140, 160, 200, 256, 400, 586):
            L.c.dwtb200_set_strip_rows(rows)
            for _ in range(2):
                for im in ims: im.fwd2(1); im.inv2(1)
            tf = ti = 0.0; reps = 4
            for _ in range(reps):
                L.c.dwtb200_timer_start()
                for im in ims: im.fwd2(1)
                tf += L.c.dwtb200_timer_stop_ms()
                L.c.dwtb200_timer_start()
                for im in ims: im.inv2(1)
                ti += L.c.dwtb200_timer_stop_ms()
            tf /= reps * len(ims) * frames; ti /= reps * len(ims) * frames
            alg = 2 * 4 * 8192 * 8192
            print(f"{name} frames {frames} ring {ring} strip_rows {rows:4d}: fwd {tf*1e3:7.1f} us/img {alg/tf/1e6:6.0f} GB/s   inv {ti*1e3:7.1f} us/img {alg/ti/1e6:6.0f} GB/s", flush=True)
        L.c.dwtb200_set_strip_rows(0)
        for im in ims: im.close()
