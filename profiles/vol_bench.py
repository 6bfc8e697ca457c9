"""3-D one-level CDF 9/7 on a device-resident volume: time forward and inverse (CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = d.lib()
L.init(0)
if len(sys.argv) > 2:
    L.check(L.c.dwtb200_set_tuning(9, int(sys.argv[2])))   # DWTB200_TUNE_VOL3: 1 tensor-copy staging, 2 cp.async staging, 0 two passes
v = d.DeviceVolume(n, n, n)
v.fill()
for _ in range(2):
    v.fwd3(); v.inv3()
L.c.dwtb200_sync()
reps = 5
tf = ti = 0.0
for _ in range(reps):
    L.c.dwtb200_timer_start(); v.fwd3(); tf += L.c.dwtb200_timer_stop_ms()
    L.c.dwtb200_timer_start(); v.inv3(); ti += L.c.dwtb200_timer_stop_ms()
tf, ti = tf / reps * 1e-3, ti / reps * 1e-3
b = 2 * 4 * n ** 3
print(f"vol3={sys.argv[2] if len(sys.argv) > 2 else 1} {n}^3 float: fwd {tf * 1e3:.3f} ms ({n ** 3 / tf / 1e9:.1f} Gvoxel/s, {b / tf / 1e9:.0f} GB/s algorithmic = {b / tf / 1e9 / 6539.9:.2f} of HBM roofline)"
      f"  inv {ti * 1e3:.3f} ms ({n ** 3 / ti / 1e9:.1f} Gvoxel/s, {b / ti / 1e9 / 6539.9:.2f})")
