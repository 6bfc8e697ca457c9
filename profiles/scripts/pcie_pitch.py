"""The reference's opt-stride rows (dwt_util_get_opt_stride: 32771 bytes for 8192 floats) against packed rows (32768): 8192 rows of 32 KiB
copied with cudaMemcpy2DAsync between pinned host memory and a pitched device plane, one direction at a time and both at once."""
import time

import torch
from cuda.bindings import runtime as rt

def ck(r):
    err, *rest = r if isinstance(r, tuple) else (r,)
    if int(err) != 0:
        raise RuntimeError(str(err))
    return rest[0] if len(rest) == 1 else rest

W, ROWS, DP = 32768, 8192, 32768
torch.zeros(1, device="cuda")
s_up = ck(rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking))
s_dn = ck(rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking))
H2D, D2H = rt.cudaMemcpyKind.cudaMemcpyHostToDevice, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost
d1 = ck(rt.cudaMalloc(DP * ROWS))
d2 = ck(rt.cudaMalloc(DP * ROWS))
for pitch in (32768, 32771, 32784, 32832):
    n = pitch * ROWS
    h1 = ck(rt.cudaHostAlloc(n, rt.cudaHostAllocPortable))
    h2 = ck(rt.cudaHostAlloc(n, rt.cudaHostAllocPortable))
    def up():
        ck(rt.cudaMemcpy2DAsync(d1, DP, h1, pitch, W, ROWS, H2D, s_up))
    def dn():
        ck(rt.cudaMemcpy2DAsync(h2, pitch, d2, DP, W, ROWS, D2H, s_dn))
    res = {}
    for name, fns in (("h2d", (up,)), ("d2h", (dn,)), ("both", (up, dn))):
        best = 1e9
        for _ in range(5):
            ck(rt.cudaDeviceSynchronize())
            t0 = time.perf_counter()
            for f in fns:
                f()
            ck(rt.cudaStreamSynchronize(s_up))
            ck(rt.cudaStreamSynchronize(s_dn))
            best = min(best, time.perf_counter() - t0)
        res[name] = best
    print(f"host row pitch {pitch}: H2D {res['h2d'] * 1e3:6.3f} ms ({W * ROWS / res['h2d'] / 1e9:5.1f} GB/s)  D2H {res['d2h'] * 1e3:6.3f} ms ({W * ROWS / res['d2h'] / 1e9:5.1f} GB/s)"
          f"  both {res['both'] * 1e3:6.3f} ms ({2 * W * ROWS / res['both'] / 1e9:5.1f} GB/s)", flush=True)
    ck(rt.cudaFreeHost(h1))
    ck(rt.cudaFreeHost(h2))
