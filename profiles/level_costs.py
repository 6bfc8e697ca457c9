"""Incremental cost of each pyramid level: time the transform for j_max = 1..J (CUDA events, 3 cycled images)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
kind = {"97s": d.CDF97_F32, "53i": d.CDF53_I32, "97d": d.CDF97_F64}[sys.argv[2] if len(sys.argv) > 2 else "97s"]
tile = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
tail = int(sys.argv[4]) if len(sys.argv) > 4 else 0
mid = int(sys.argv[5]) if len(sys.argv) > 5 else 0
pdl = int(sys.argv[6]) if len(sys.argv) > 6 else 1
tma = int(sys.argv[7]) if len(sys.argv) > 7 else 1
maxj = int(sys.argv[8]) if len(sys.argv) > 8 else 99
L = d.lib()
L.init(0)
L.check(L.c.dwtb200_set_tuning(0, tile))
L.check(L.c.dwtb200_set_tuning(1, tail))
L.check(L.c.dwtb200_set_tuning(2, mid))
L.check(L.c.dwtb200_set_tuning(3, pdl))
L.check(L.c.dwtb200_set_tuning(4, tma))
M = 3
imgs = [d.DeviceImage(kind, n, n) for _ in range(M)]
for im in imgs:
    im.fill(0, 0, 0)


def timeit(J, reps=5):
    tf = ti = 0.0
    for _ in range(reps):
        L.c.dwtb200_timer_start()
        for im in imgs:
            im.fwd2(J)
        tf += L.c.dwtb200_timer_stop_ms()
        L.c.dwtb200_timer_start()
        for im in imgs:
            im.inv2(J)
        ti += L.c.dwtb200_timer_stop_ms()
    return tf / (reps * M) * 1e3, ti / (reps * M) * 1e3, imgs[0].last_launches


print(f"n={n} tile_max={tile} tail_max={tail} mid_max={mid} pdl={pdl} tma={tma}")
pf = pi = 0.0
Jmax = L.c.dwtb200_ceil_log2(n)
for J in range(1, min(Jmax, maxj) + 1):
    timeit(J, 1)
    f, i, nl = timeit(J)
    print(f"J={J:2d} launches={nl:2d} fwd {f:7.1f} us (+{f - pf:6.1f})   inv {i:7.1f} us (+{i - pi:6.1f})", flush=True)
    pf, pi = f, i
# an empty-ish graph: 2x2 image
small = d.DeviceImage(kind, 64, 64)
small.fill(0, 0, 0)
for _ in range(3):
    small.fwd2(1)
L.c.dwtb200_timer_start()
for _ in range(100):
    small.fwd2(1)
print("64x64 J=1 graph launch, back to back: %.2f us each" % (L.c.dwtb200_timer_stop_ms() * 10))
