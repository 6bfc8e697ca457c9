"""Level-0 (j_max=1) and full-pyramid throughput for several image shapes (CUDA events, 3 cycled images)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d  # noqa: E402

kind = {"97s": d.CDF97_F32, "53i": d.CDF53_I32, "97d": d.CDF97_F64}[sys.argv[1] if len(sys.argv) > 1 else "97s"]
es = 8 if kind == d.CDF97_F64 else 4
L = d.lib()
L.init(0)
shapes = [(8192, 8192, 1), (8000, 8000, 1), (8160, 8192, 1), (8192, 8160, 1), (7919, 6007, 1), (10000, 10000, 1), (16384, 16384, 1),
          (4096, 4096, 4), (2048, 2048, 16), (2048, 2048, 64), (8192, 8192, 4)]
print(f"{'shape':>18} {'J':>2} {'fwd_us':>8} {'inv_us':>8} {'fwd GB/s':>9} {'inv GB/s':>9} {'Gpix/s f':>9} {'Gpix/s i':>9}")
for (w, h, fr) in shapes:
    M = 3 if w * h * fr * es * 2.4 * 3 < 60e9 else 1
    imgs = [d.DeviceImage(kind, w, h, fr) for _ in range(M)]
    for im in imgs:
        im.fill(0, 0, 6)
    Jfull = L.c.dwtb200_ceil_log2(min(w, h))
    for J in (1, Jfull):
        for rep in range(2):
            tf = ti = 0.0
            n = 4
            for _ in range(n):
                L.c.dwtb200_timer_start()
                for im in imgs:
                    im.fwd2(J)
                tf += L.c.dwtb200_timer_stop_ms()
                L.c.dwtb200_timer_start()
                for im in imgs:
                    im.inv2(J)
                ti += L.c.dwtb200_timer_stop_ms()
        tf, ti = tf / (n * M) * 1e-3, ti / (n * M) * 1e-3
        b = 2 * es * sum(-(-w // (1 << l)) * -(-h // (1 << l)) for l in range(J)) * fr
        print(f"{w:>6}x{h:<6}x{fr:<3} {J:>2} {tf * 1e6:8.1f} {ti * 1e6:8.1f} {b / tf / 1e9:9.0f} {b / ti / 1e9:9.0f} {w * h * fr / tf / 1e9:9.1f} {w * h * fr / ti / 1e9:9.1f}",
              flush=True)
    for im in imgs:
        im.close()
