"""Single-image pyramids, one image per call, timed with CUDA events on the image's own stream (dwtb200_image_timer_*):
python single_r2.py KIND N [J ...] [key=value ...]   -- prints per-J forward / inverse microseconds (3 images cycled: none is L2-resident)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
name, n = sys.argv[1], int(sys.argv[2])
Js = [int(a) for a in sys.argv[3:] if '=' not in a]
for a in sys.argv[3:]:
    if '=' in a:
        k, v = a.split('=')
        L.check(L.c.dwtb200_set_tuning(int(k), int(v)))
kind = {"97s": d.CDF97_F32, "53i": d.CDF53_I32, "97d": d.CDF97_F64}[name]
ims = [d.DeviceImage(kind, n, n, 1) for _ in range(3)]
for im in ims: im.fill(0, 0, 0)
Jmax = L.c.dwtb200_ceil_log2(n)
for J in (Js or [Jmax]):
    for _ in range(2):
        for im in ims: im.fwd2(J); im.inv2(J)
    L.check(L.c.dwtb200_sync())
    reps = 8
    # [mark, fwd, mark] per image, each image's stream ordered behind the previous image's, nothing synchronised until the end
    prev = ims[-1]
    for _ in range(reps):
        for im in ims:
            im.wait_for(prev); im.mark(); im.fwd2(J); im.mark(); prev = im
        for im in ims:
            im.wait_for(prev); im.mark(); im.inv2(J); im.mark(); prev = im
    tf = ti = 0.0
    for im in ims:
        t = im.read_marks()          # fwd, gap, fwd, gap ... per repetition: intervals 0, 2 of every 4-mark group
        for r in range(reps):
            tf += t[4 * r]; ti += t[4 * r + 2]
    k = 1e3 / (reps * len(ims))
    tg = 0.0
    for _ in range(reps):
        for im in ims:
            L.c.dwtb200_timer_start(); im.fwd2(J); tg += L.c.dwtb200_timer_stop_ms()
            im.inv2(J)
    L.check(L.c.dwtb200_sync())
    print(f"{name} n={n} J={J:2d} launches={ims[0].last_launches:2d} fwd {tf*k:7.1f} us  inv {ti*k:7.1f} us   (fwd through the global timer: {tg*k:7.1f} us)  {' '.join(a for a in sys.argv[3:] if '=' in a)}", flush=True)
