"""A few level-0 transforms for ncu: python prof_l0.py 97s <ring> [n] [frames]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
kind = {"97s": d.CDF97_F32, "53i": d.CDF53_I32, "97d": d.CDF97_F64}[sys.argv[1]]
ring = int(sys.argv[2]); n = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 1
L.check(L.c.dwtb200_set_tuning(6, ring))
im = d.DeviceImage(kind, n, n, frames); im.fill(0, 0, 6)
for _ in range(3):
    im.fwd2(1); im.inv2(1)
L.c.dwtb200_sync()
print("done")
