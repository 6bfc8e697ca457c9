"""One image row-strip partitioned over the GPUs of a box (BASELINE config 5a), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        profiles/strips_multi.py [--size 65536] [--levels 4] [--kind 97s] [--check]

Halo rows travel by torch.distributed send/recv over NCCL (NVLink P2P between the ranks' GPUs).  Prints one JSON
line on rank 0: forward / inverse time (CUDA-synchronised wall clock, max over ranks), Gpixel/s, and with --check
(sizes the oracle can hold) the number of samples differing from the oracle's single-image transform."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=65536)
    ap.add_argument("--levels", type=int, default=4, help="levels done distributed; the rest runs on rank 0")
    ap.add_argument("--kind", default="97s")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import libdwt_b200 as d
    from libdwt_b200 import strips
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d.lib().init(local)
    kind = {"97s": d.CDF97_F32, "53i": d.CDF53_I32, "97d": d.CDF97_F64}[args.kind]
    W = H = args.size
    ds = strips.DistStrips(W, H, args.levels, lambda w, h: strips.DeviceStripEngine(d, torch, kind, w, h), dist)
    wide = 1 if args.size > 16384 else 0

    def fill():
        ds.local.img.fill(0, 0, 0, y_offset=ds.ea, wide=wide)   # owned rows (and, harmlessly, the halo rows) of the global pattern
        ds.local.img.L.check(ds.local.img.L.c.dwtb200_sync())

    def sync():
        torch.cuda.synchronize()
        dist.barrier()

    fill()
    keep = ds.owned_view().clone() if args.size <= 16384 else None
    tf = ti = 1e30
    for rep in range(args.reps):
        fill()
        if rep == 0:   # poison the halo rows so that the exchange is what supplies them
            v = ds.local.view()
            if ds.a > ds.ea:
                v[:ds.a - ds.ea].fill_(7)
            if ds.eb > ds.b:
                v[ds.b - ds.ea:].fill_(7)
        sync()
        t0 = time.perf_counter()
        ds.forward()
        sync()
        t1 = time.perf_counter()
        if rep == 0 and args.check:
            mallat = ds.gather_mallat()
        sync()
        t2 = time.perf_counter()
        ds.inverse()
        sync()
        t3 = time.perf_counter()
        tf, ti = min(tf, t1 - t0), min(ti, t3 - t2)
    tt = torch.tensor([tf, ti], device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    tf, ti = tt.tolist()
    err = None
    if keep is not None:
        e = (ds.owned_view().double() - keep.double()).abs().max()
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        err = float(e)
    line = {"workload": f"{W}x{H} {args.kind} row strips over {world} GPUs, {args.levels} levels distributed + {ds.J - ds.Jd} on rank 0",
            "n_gpus": world, "fwd_ms": tf * 1e3, "inv_ms": ti * 1e3, "fwd_gpixel_s": W * H / tf / 1e9, "inv_gpixel_s": W * H / ti / 1e9,
            "halo_rows": ds.plan.halo, "strip_rows": ds.b - ds.a, "roundtrip_max_abs_err": err,
            "exchange": "torch.distributed send/recv, NCCL over NVLink", "timing": "wall clock between device syncs + barriers, min of reps, max over ranks"}
    if args.check and rank == 0:
        from oracle.orc import Oracle
        orc = Oracle()
        t = args.kind[-1]
        want = orc.fill(np.zeros((H, W), mallat.dtype), t, wrap32=0 if wide else 1)
        orc.fwd2(want, args.kind[:2], t)
        line["forward_samples_differing_from_oracle"] = int((mallat.view(np.uint32 if mallat.itemsize == 4 else np.uint64)
                                                              != want.view(np.uint32 if want.itemsize == 4 else np.uint64)).sum())
    if rank == 0:
        print(json.dumps(line))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
