"""Level-0 forward streaming kernel, 8192^2 float: where does the time go?  dbg 0 = normal, 1 = no stores,
2 = no lifting arithmetic (loads + stores only); pfd = row pairs prefetched ahead; narrow = 16 B per lane."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libdwt_b200 as d
L = d.lib(); L.init(0)
for frames in (4, 1):
    ims = [d.DeviceImage(d.CDF97_F32, 8192, 8192, frames) for _ in range(2 if frames == 4 else 4)]
    for im in ims: im.fill(0, 0, 6)
    for narrow in (0, 1):
        for pfd in (1, 2):
            for dbg in (0, 1, 2):
                L.check(L.c.dwtb200_set_tuning(4, narrow)); L.check(L.c.dwtb200_set_tuning(99, dbg)); L.check(L.c.dwtb200_set_tuning(98, pfd))
                for _ in range(2):
                    for im in ims: im.fwd2(1)
                n = 6
                L.c.dwtb200_timer_start()
                for _ in range(n):
                    for im in ims: im.fwd2(1)
                t = L.c.dwtb200_timer_stop_ms() / (n * len(ims))
                print(f"frames {frames} narrow {narrow} pfd {pfd} dbg {dbg}: {t*1e3/frames:7.1f} us per frame  {2*frames*8192*8192*4/t/1e6:6.0f} GB/s alg", flush=True)
    for im in ims: im.close()
