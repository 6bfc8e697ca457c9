"""Full-depth forward / inverse timings of the BASELINE.json configurations on one GPU (device-resident, CUDA events,
one image object at a time, three objects cycled so that no transform finds its input in L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
K = {"97s": (d.CDF97_F32, 4), "97d": (d.CDF97_F64, 8), "53i": (d.CDF53_I32, 4), "53s": (d.CDF53_F32, 4), "97i": (d.CDF97_I32, 4)}
CFG = [("53i", 4096, 4096, 1), ("97s", 8192, 8192, 1), ("97s", 7919, 6007, 1), ("97d", 8192, 8192, 1), ("97s", 2048, 2048, 64),
       ("97s", 512, 512, 1), ("53s", 8192, 8192, 1), ("97i", 8192, 8192, 1), ("97s", 8192, 8192, 4), ("53i", 8192, 8192, 4)]
for (name, w, h, frames) in CFG:
    kind, es = K[name]
    ims = [d.DeviceImage(kind, w, h, frames) for _ in range(3)]
    for im in ims: im.fill(0, 0, 6)
    for _ in range(2):
        for im in ims: j = im.fwd2(); im.inv2(j)
    tf = ti = 0.0; reps = 5
    for _ in range(reps):
        for im in ims:
            L.c.dwtb200_timer_start(); j = im.fwd2(); tf += L.c.dwtb200_timer_stop_ms()
        for im in ims:
            L.c.dwtb200_timer_start(); im.inv2(j); ti += L.c.dwtb200_timer_stop_ms()
    tf *= 1e-3 / (reps * len(ims)); ti *= 1e-3 / (reps * len(ims))
    b = 2 * es * frames * sum(-(-w // (1 << l)) * -(-h // (1 << l)) for l in range(j))
    px = w * h * frames
    print(f"{name} {w}x{h} x{frames} J={j}: fwd {tf*1e6:8.1f} us {px/tf/1e9:6.1f} Gpixel/s {b/tf/1e9/6539.9:5.2f} of roofline | inv {ti*1e6:8.1f} us {px/ti/1e9:6.1f} Gpixel/s {b/ti/1e9/6539.9:5.2f}", flush=True)
    for im in ims: im.close()
