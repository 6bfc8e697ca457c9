import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
for (n, frames) in ((2048, 64), (2048, 16), (2560, 16), (1920, 32)):
    im = d.DeviceImage(d.CDF97_F32, n, n if n != 1920 else 1080, frames); im.fill(0, 0, 6)
    im2 = d.DeviceImage(d.CDF97_F32, n, n if n != 1920 else 1080, frames); im2.fill(0, 0, 6)
    for ring in (3 | (5 << 4), 3):
        L.check(L.c.dwtb200_set_tuning(6, ring))
        for J in (1, -1):
            for _ in range(2):
                for x in (im, im2): jj = x.fwd2(J); x.inv2(jj)
            tf = ti = 0.0; reps = 6
            for _ in range(reps):
                for x in (im, im2):
                    L.c.dwtb200_timer_start(); jj = x.fwd2(J); tf += L.c.dwtb200_timer_stop_ms()
                    L.c.dwtb200_timer_start(); x.inv2(jj); ti += L.c.dwtb200_timer_stop_ms()
            tf *= 1e3 / (2 * reps * frames); ti *= 1e3 / (2 * reps * frames)
            px = n * (n if n != 1920 else 1080)
            print(f"{n} x{frames} ring=0x{ring:x} J={J:2d}: fwd {tf:6.2f} us/frame ({px/tf/1e3:5.0f} Gpix/s) inv {ti:6.2f} ({px/ti/1e3:5.0f})", flush=True)
    im.close(); im2.close()
