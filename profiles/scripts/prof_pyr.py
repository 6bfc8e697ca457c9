"""One forward + inverse full pyramid per kind for ncu: python prof_pyr.py [frames]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for kind in (d.CDF97_F32, d.CDF53_I32):
    im = d.DeviceImage(kind, 8192, 8192, frames); im.fill(0, 0, 6)
    for _ in range(2):
        j = im.fwd2(); im.inv2(j)
    L.c.dwtb200_sync(); im.close()
print("done")
