"""Profiling driver: one warm-up and one profiled forward+inverse full pyramid per sample type on one
8192x8192 device-resident image (run under ncu, see profiles/README.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["97s", "53i"]
K = {"97s": d.CDF97_F32, "53i": d.CDF53_I32, "97d": d.CDF97_F64}
L = d.lib()
L.init(0)
if len(sys.argv) > 3:
    L.check(L.c.dwtb200_set_tuning(4, int(sys.argv[3])))   # TMA fast path on/off
for name in kinds:
    img = d.DeviceImage(K[name], n, n)
    img.fill(0, 0, 0)
    for _ in range(2):
        j = img.fwd2()
        img.inv2(j)
    L.check(L.c.dwtb200_sync())
    print(name, "J", j, "launches", img.last_launches)
    img.close()
