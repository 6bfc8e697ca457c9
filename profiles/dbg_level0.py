import sys; sys.path.insert(0,'/root/repo')
import libdwt_b200 as d
L=d.lib(); L.init(0)
im=d.DeviceImage(d.CDF97_F32,8192,8192,4); im.fill(0,0,6)
for pfd,dbg in ((1,0),(1,1),(1,2),(1,0)):
    L.check(L.c.dwtb200_set_tuning(99,dbg)); L.check(L.c.dwtb200_set_tuning(98,pfd))
    for _ in range(3): im.fwd2(1)
    L.c.dwtb200_timer_start()
    for _ in range(10): im.fwd2(1)
    t=L.c.dwtb200_timer_stop_ms()/10
    print('pfd',pfd,'dbg',dbg,'%.1f us per 4-frame launch'%(t*1e3), '%.0f GB/s alg'%(2*4*8192*8192*4/t/1e6))
