import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
rings = [int(x) for x in sys.argv[1].split(',')]
kinds = sys.argv[2].split(',') if len(sys.argv) > 2 else ["97s", "53i", "97d"]
KK = {"97s": d.CDF97_F32, "53i": d.CDF53_I32, "97d": d.CDF97_F64, "53s": d.CDF53_F32, "97i": d.CDF97_I32}
for name in kinds:
    kind = KK[name]
    for frames in (4, 1):
        n = 8192 if name != "97d" else 4096
        ims = [d.DeviceImage(kind, n, n, frames) for _ in range(2 if frames == 4 else 4)]
        for im in ims: im.fill(0, 0, 6)
        for r in rings:
            L.check(L.c.dwtb200_set_tuning(6, r))
            for J in (1, -1):
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
                es = 8 if name == "97d" else 4
                alg = 2 * es * n * n * (1 if J == 1 else 4 / 3)
                print(f"{name} frames {frames} ring {r:3d} J {J:2d}: fwd {tf*1e3:7.1f} us/img {alg/tf/1e6:6.0f} GB/s   inv {ti*1e3:7.1f} us/img {alg/ti/1e6:6.0f} GB/s", flush=True)
        for im in ims: im.close()
