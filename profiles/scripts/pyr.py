import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
KK = {"97s": d.CDF97_F32, "53i": d.CDF53_I32, "97d": d.CDF97_F64}
def run(name, frames, n, label):
    kind = KK[name]
    ims = [d.DeviceImage(kind, n, n, frames) for _ in range(2 if frames >= 4 else 4)]
    for im in ims: im.fill(0, 0, 6)
    out = []
    for J in (1, 2, 3, -1):
        for _ in range(2):
            for im in ims: im.fwd2(J); im.inv2(J)
        tf = ti = 0.0; reps = 5
        for _ in range(reps):
            L.c.dwtb200_timer_start()
            for im in ims: im.fwd2(J)
            tf += L.c.dwtb200_timer_stop_ms()
            L.c.dwtb200_timer_start()
            for im in ims: im.inv2(J)
            ti += L.c.dwtb200_timer_stop_ms()
        tf /= reps * len(ims) * frames; ti /= reps * len(ims) * frames
        out.append(f"J{J:2d} {tf*1e3:6.1f}/{ti*1e3:6.1f}")
    es = 8 if name == "97d" else 4
    alg = 2 * es * n * n * 4 / 3
    print(f"{label:28s} {name} x{frames}: " + "  ".join(out) + f"   full: {alg/tf/1e6:5.0f}/{alg/ti/1e6:5.0f} GB/s", flush=True)
    for im in ims: im.close()
cfgs = sys.argv[1:]   # each: ring:tile_max:waves:ppsmin:ppsmax
for c in cfgs:
    ring, tmax, waves, pmin, pmax = [int(x) for x in c.split(':')]
    L.check(L.c.dwtb200_set_tuning(6, ring)); L.check(L.c.dwtb200_set_tuning(0, tmax)); L.check(L.c.dwtb200_set_tuning(97, waves | (pmin << 8) | (pmax << 16)))
    for name in ("97s", "53i"):
        for frames in (4, 1):
            run(name, frames, 8192, c)
