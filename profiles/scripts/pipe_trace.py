"""Timeline of one pipelined host call (DWTB200_PIPE_TRACE=1 prints the stage time stamps on stderr) and the wall clock of
dwt_cdf97_2f_s / dwt_cdf97_2i_s on a pinned host 8192 x 8192 float image.  usage: pipe_trace.py [row_stride_bytes]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import libdwt_b200 as d  # noqa: E402

W = H = 8192
row = int(sys.argv[1]) if len(sys.argv) > 1 else W * 4
L = d.lib()
L.init(0)
p = L.c.dwtb200_host_alloc(row * H)
raw = (C.c_uint8 * (row * H)).from_address(p)
arr = np.ndarray(shape=(H, W), dtype=np.float32, buffer=raw, strides=(row, 4))
arr[:] = np.random.default_rng(1).random((H, W), dtype=np.float32)
for it in range(4):
    jj = [-1]
    t0 = time.perf_counter()
    d.dwt_cdf97_2f_s(p, row, 4, W, H, W, H, jj, 0, 0)
    t1 = time.perf_counter()
    d.dwt_cdf97_2i_s(p, row, 4, W, H, W, H, jj[0], 0, 0)
    t2 = time.perf_counter()
    print(f"iteration {it}: forward {1e3 * (t1 - t0):.3f} ms, inverse {1e3 * (t2 - t1):.3f} ms (row stride {row})", file=sys.stderr)
L.c.dwtb200_host_free(p)
