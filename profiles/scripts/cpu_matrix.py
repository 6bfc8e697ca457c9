"""The reference's own CPU implementation over the matrix SURVEY.md 8(d) asks for, on the host cores of the box it runs on:
threads {1, all}, accel {0, 9, 12 with 4 workers}, row stride {packed, dwt_util_get_opt_stride}; CDF 9/7 float 8192^2, CDF 5/3 int
4096^2, and the single-threaded 3-D transform.  Uses oracle/_ref/libdwt_ref.so (the unmodified reference compiled by oracle/Makefile);
test infrastructure, not product.  python profiles/scripts/cpu_matrix.py [size]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.orc import Ref, strided_image  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ref = Ref()
L = ref.lib
allthr = os.cpu_count() or 1
print(f"host: {allthr} hardware threads; image {n} x {n}")


def run(w, t, size, threads, accel, workers, opt):
    L.dwt_util_set_num_threads(threads)
    L.dwt_util_set_num_workers(workers)
    L.dwt_util_set_accel(accel)
    es = 8 if t == "d" else 4
    row = L.dwt_util_get_opt_stride(size * es) if opt else size * es
    a = strided_image((size, size), t, row)
    ref.fill(a, t, rand=0, type_=0)
    best_f = best_i = 1e30
    for _ in range(2):
        t0 = time.perf_counter()
        J = ref.fwd2(a, w, t)
        t1 = time.perf_counter()
        ref.inv2(a, w, t, j_max=J)
        t2 = time.perf_counter()
        best_f, best_i = min(best_f, t1 - t0), min(best_i, t2 - t1)
    return best_f, best_i


for threads in (allthr, 1):
    for (accel, workers) in ((0, 1), (9, 1), (12, 4)):
        for opt in (0, 1):
            f, i = run("97", "s", n, threads, accel, workers, opt)
            print(f"cdf97 float {n}^2  threads={threads:3d} accel={accel:2d} workers={workers} stride={'opt' if opt else 'packed'}: "
                  f"fwd {f:.3f} s ({n * n / f / 1e9:.3f} Gpixel/s)  inv {i:.3f} s ({n * n / i / 1e9:.3f} Gpixel/s)", flush=True)
m = n // 2
for threads in (allthr, 1):
    f, i = run("53", "i", m, threads, 0, 1, 0)
    print(f"cdf53 int {m}^2  threads={threads:3d}: fwd {f:.3f} s ({m * m / f / 1e9:.3f} Gpixel/s)  inv {i:.3f} s ({m * m / i / 1e9:.3f} Gpixel/s)", flush=True)
v = max(64, n // 16)
a = np.zeros((v, v, v), np.float32)
ref.volume_fill(a)
b = np.zeros_like(a)
t0 = time.perf_counter()
ref.fwd3(a, b)
t1 = time.perf_counter()
print(f"3-D cdf97 float {v}^3 (single-threaded by construction): fwd {t1 - t0:.3f} s ({v ** 3 / (t1 - t0) / 1e9:.4f} Gvoxel/s)")
