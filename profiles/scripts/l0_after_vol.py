import sys
sys.path.insert(0, '.')
import libdwt_b200 as d
L = d.lib()
im = d.DeviceImage(d.CDF97_F32, 8192, 8192, 4)
im.fill(0, 0, 6)
def l0(tag):
    for _ in range(3): im.fwd2(1)
    L.c.dwtb200_sync()
    L.c.dwtb200_timer_start()
    for _ in range(10): im.fwd2(1)
    t = L.c.dwtb200_timer_stop_ms() / 10
    print(f"{tag}: level 0 of 4 frames {t*1e3:.1f} us", flush=True)
    im.fill(0, 0, 6)
l0("start")
for vol3 in (0, 1):
    L.check(L.c.dwtb200_set_tuning(9, vol3))
    v = d.DeviceVolume(1024, 1024, 1024)
    v.fill(); v.fwd3(); v.inv3()
    L.c.dwtb200_sync()
    l0(f"after volume transform vol3={vol3}, volume alive")
    v.close()
    l0(f"after volume freed")
big = d.DeviceImage(d.CDF97_F32, 32768, 32768, 1)
big.fill(0, 0, 0); j = big.fwd2(); big.inv2(j); L.c.dwtb200_sync()
l0("after 32768^2 image, alive")
big.close()
l0("after 32768^2 freed")
