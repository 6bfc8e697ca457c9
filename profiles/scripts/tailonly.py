import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
for name, kind in (("97s", d.CDF97_F32), ("53i", d.CDF53_I32)):
    im = d.DeviceImage(kind, 64, 64, 1); im.fill(0, 0, 0)
    for J in range(1, 7):
        for _ in range(3): im.fwd2(J); im.inv2(J)
        n = 50
        L.c.dwtb200_timer_start()
        for _ in range(n): im.fwd2(J)
        tf = L.c.dwtb200_timer_stop_ms() / n * 1e3
        L.c.dwtb200_timer_start()
        for _ in range(n): im.inv2(J)
        ti = L.c.dwtb200_timer_stop_ms() / n * 1e3
        print(f"{name} 64x64 J={J}: fwd {tf:5.2f} us  inv {ti:5.2f} us per call (back to back graph launches, launches={im.last_launches})", flush=True)
    im.close()
