"""Why does the pipelined host path move 84 - 87 GB/s over the link when two 256 MiB copies at once reach ~99?  Both directions at
once, 256 MiB each way, as (a) one copy per direction, (b) 16 chunks per direction back to back, (c) the same chunks through
cudaMemcpy2DAsync (row = pitch = 32 KiB), (d) 2-D copies of half rows (16 KiB of every 32 KiB), (e) chunks chained by events the way
the pipeline chains them, (f) while a kernel streams through HBM."""
import sys
import time

import torch
from cuda.bindings import runtime as rt

def ck(r):
    if isinstance(r, tuple):
        err, *rest = r
    else:
        err, rest = r, []
    if int(err) != 0:
        raise RuntimeError(str(err))
    return rest[0] if len(rest) == 1 else rest

N = 256 << 20
ROW = 32768
ROWS = N // ROW
torch.cuda.init()
torch.zeros(1, device="cuda")
hs = ck(rt.cudaHostAlloc(N, rt.cudaHostAllocPortable))
hd = ck(rt.cudaHostAlloc(N, rt.cudaHostAllocPortable))
ds = ck(rt.cudaMalloc(N))
dd = ck(rt.cudaMalloc(N))
ck(rt.cudaMemset(ds, 1, N))
s_up = ck(rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking))
s_dn = ck(rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking))
s_k = torch.cuda.Stream()
big = torch.empty(1 << 28, device="cuda", dtype=torch.float32)   # 1 GiB
H2D, D2H = rt.cudaMemcpyKind.cudaMemcpyHostToDevice, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost

def run(name, fn, reps=5):
    best = 1e9
    for _ in range(reps):
        ck(rt.cudaDeviceSynchronize())
        t0 = time.perf_counter()
        fn()
        ck(rt.cudaStreamSynchronize(s_up))
        ck(rt.cudaStreamSynchronize(s_dn))
        best = min(best, time.perf_counter() - t0)
    torch.cuda.synchronize()
    print(f"{name:70s} {best * 1e3:7.3f} ms  {2 * N / best / 1e9:6.1f} GB/s both directions", flush=True)

def one():
    ck(rt.cudaMemcpyAsync(dd, hs, N, H2D, s_up))
    ck(rt.cudaMemcpyAsync(hd, ds, N, D2H, s_dn))

def chunks(n):
    def f():
        c = N // n
        for i in range(n):
            ck(rt.cudaMemcpyAsync(dd + i * c, hs + i * c, c, H2D, s_up))
            ck(rt.cudaMemcpyAsync(hd + i * c, ds + i * c, c, D2H, s_dn))
    return f

def chunks2d(n, width):
    def f():
        r = ROWS // n
        for i in range(n):
            o = i * r * ROW
            ck(rt.cudaMemcpy2DAsync(dd + o, ROW, hs + o, ROW, width, r, H2D, s_up))
            ck(rt.cudaMemcpy2DAsync(hd + o, ROW, ds + o, ROW, width, r, D2H, s_dn))
    return f

evs = [ck(rt.cudaEventCreateWithFlags(rt.cudaEventDisableTiming)) for _ in range(64)]
def chained(n):
    def f():   # the download of chunk i waits for the upload of chunk i (as the pipeline's downloads wait for a kernel behind an upload)
        c = N // n
        for i in range(n):
            ck(rt.cudaMemcpyAsync(dd + i * c, hs + i * c, c, H2D, s_up))
            ck(rt.cudaEventRecord(evs[i], s_up))
        for i in range(n):
            ck(rt.cudaStreamWaitEvent(s_dn, evs[i], 0))
            ck(rt.cudaMemcpyAsync(hd + i * c, ds + i * c, c, D2H, s_dn))
    return f

def with_kernel(fn):
    def f():
        with torch.cuda.stream(s_k):
            for _ in range(12):
                big.add_(1.0)   # 2 GiB of HBM traffic per launch, ~0.35 ms
        fn()
    return f

run("one 256 MiB copy per direction", one)
run("16 chunks of 16 MiB per direction, cudaMemcpyAsync", chunks(16))
run("32 chunks of 8 MiB per direction, cudaMemcpyAsync", chunks(32))
run("16 chunks, cudaMemcpy2DAsync, rows of 32 KiB at a 32 KiB pitch", chunks2d(16, ROW))
run("16 chunks, cudaMemcpy2DAsync, 16 KiB of every 32 KiB row (half the bytes)", chunks2d(16, ROW // 2))
run("16 chunks, downloads chained behind the uploads by events", chained(16))
run("one copy per direction while a kernel streams through HBM", with_kernel(one))
run("16 chunks per direction while a kernel streams through HBM", with_kernel(chunks(16)))
