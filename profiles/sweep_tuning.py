"""Times the full 8192^2 pyramid (device-resident, CUDA events) for each kernel-selection tuning."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
kind = {"97s": d.CDF97_F32, "53i": d.CDF53_I32, "97d": d.CDF97_F64}[sys.argv[2] if len(sys.argv) > 2 else "97s"]
L = d.lib()
L.init(0)
M = 3
imgs = [d.DeviceImage(kind, n, n) for _ in range(M)]
for im in imgs:
    im.fill(0, 0, 0)


def timeit(reps=5):
    tf = ti = 0.0
    for _ in range(reps):
        L.c.dwtb200_timer_start()
        for im in imgs:
            j = im.fwd2()
        tf += L.c.dwtb200_timer_stop_ms()
        L.c.dwtb200_timer_start()
        for im in imgs:
            im.inv2(j)
        ti += L.c.dwtb200_timer_stop_ms()
    return tf / (reps * M) * 1e3, ti / (reps * M) * 1e3, imgs[0].last_launches


print(f"{'tile_max':>10} {'tail_max':>8} {'fwd_us':>8} {'inv_us':>8} launches")
for tile in (0, 256 ** 2, 512 ** 2, 1024 ** 2, 2048 ** 2, 4096 ** 2, 8192 ** 2):
    for tail in (0, 16, 256, 1024, 4096, 16384):
        L.check(L.c.dwtb200_set_tuning(0, tile))
        L.check(L.c.dwtb200_set_tuning(1, tail))
        timeit(1)
        f, i, n_l = timeit()
        print(f"{tile:>10} {tail:>8} {f:8.1f} {i:8.1f} {n_l}", flush=True)
