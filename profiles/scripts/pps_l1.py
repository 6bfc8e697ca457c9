import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
for (n, frames) in ((4096, 4), (4096, 1), (2048, 4), (2048, 16)):
    im = d.DeviceImage(d.CDF97_F32, n, n, frames); im.fill(0, 0, 6)
    im2 = d.DeviceImage(d.CDF97_F32, n, n, frames); im2.fill(0, 0, 6)
    res = []
    for pps in (0, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 48):
        L.c.dwtb200_set_strip_rows(2 * pps)
        for _ in range(2):
            for x in (im, im2): x.fwd2(1); x.inv2(1)
        tf = ti = 0.0; reps = 8
        for _ in range(reps):
            for x in (im, im2):
                L.c.dwtb200_timer_start(); x.fwd2(1); tf += L.c.dwtb200_timer_stop_ms()
                L.c.dwtb200_timer_start(); x.inv2(1); ti += L.c.dwtb200_timer_stop_ms()
        tf *= 1e3 / (2 * reps); ti *= 1e3 / (2 * reps)
        ncg = -(-n // 240); nb = -(-ncg // 7); units = n // 2
        ctas = nb * (-(-units // pps)) * frames if pps else 0
        res.append(f"pps {pps:2d} ({ctas/296:5.2f} waves): fwd {tf:6.1f} inv {ti:6.1f}")
    print(f"{n}^2 x{frames} level 0 only:\n  " + "\n  ".join(res), flush=True)
    L.c.dwtb200_set_strip_rows(0)
    im.close(); im2.close()
