#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on B200.

Metric: "CDF 9/7 + 5/3 fwd+inv Gpixel/s & % HBM roofline, 8192^2 j=max".
One STEP = the hot path over one batch of synthetic input: M (default 4) independent 8192x8192 images per
sample type, each transformed forward (all M) and then inverse (all M), full depth J=13, for CDF 9/7
float32 and for CDF 5/3 int32 -- the reference's own perf protocol (M images, all forward then all
inverse, src/libdwt.c:21391) with 4*M transforms of 64 Mpixel per step.  `value` is the number of pixels
transformed (W*H per transform) per second, data resident in HBM; every image is 256 MiB (> 126 MB L2)
and M of them are cycled, so no transform finds its input in L2.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 (torchrun, one rank per GPU): every rank runs the same batch on its own GPU (independent frames
shard with no collective -> "weak" scaling); value = pixels of all ranks / max-over-ranks device time.
`--impl reference` times the reference's own CPU code (oracle/_ref/libdwt_ref.so, OpenMP on all host
cores; the oracle port if the compiled reference is absent) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W = H = 8192
J = 13
PIX = W * H


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(w, h, j, es):
    """every level reads its LL input once and writes its four subbands once."""
    return 2 * es * sum(-(-w // (1 << l)) * -(-h // (1 << l)) for l in range(j))


class ClockSampler:
    """nvidia-smi sampled every 20 ms from before the warm-up; samples whose power draw shows the GPU under
    load (>= 60 % of the maximum seen) are the ones summarised."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        good = []
        for r in rows:
            try:
                good.append((float(r[0]), float(r[1]), float(r[2]), r[3:7]))
            except Exception:
                continue
        if not good:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        pmax = max(g[2] for g in good)
        load = [g for g in good if g[2] >= 0.6 * pmax] or good
        reasons = set()
        for g in load:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), g[3]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([g[0] for g in load])), "sm_max_mhz": float(max(g[1] for g in good)),
                "reasons": sorted(reasons), "samples": len(good), "samples_under_load": len(load), "power_w_max": pmax}


# ======================================================================================================
# CPU arm: the reference's own implementation on the host cores
# ======================================================================================================
def cpu_impl():
    from oracle.orc import Oracle, Ref
    if Ref.available():
        r = Ref()
        return r, "reference"
    return Oracle(), "port"


def cpu_time_step(impl, threads, images):
    """One bounded sample: ONE 8192^2 image per type, forward + inverse (4 transforms = 1/M of a step), on pre-filled host images."""
    impl.set_threads(threads)
    secs = {}
    for (w, t) in (("97", "s"), ("53", "i")):
        img = images[t]
        impl.fill(img, t)
        t0 = time.perf_counter()
        j = impl.fwd2(img, w, t)
        t1 = time.perf_counter()
        impl.inv2(img, w, t, j_max=j)
        t2 = time.perf_counter()
        secs[w + t + "_fwd"] = t1 - t0
        secs[w + t + "_inv"] = t2 - t1
    return secs


CPU_SAMPLE = ("1 image of 8192x8192 per type (float32 9/7 + int32 5/3), forward+inverse, J=13 = 4 transforms per step (1/M of a GPU "
              "step); rows at dwt_util_get_opt_stride (the layout of the reference's examples), all host threads")


def cpu_protocol(trials, warmup):
    """The reference's own measurement protocol (dwt_util_perf_cdf97_2_s, src/libdwt.c:21391-21507: M images, N trials, per direction
    the minimum over the trials of the mean seconds per transform) with M = 1, on the host cores, after `warmup` untimed trials.
    The same function serves `--impl reference` and the cpu_baseline leg of the GPU arm."""
    from oracle.orc import strided_image
    impl, kind = cpu_impl()
    threads = os.cpu_count() or 1
    # rows at the reference's own "optimal" (prime) stride where the compiled reference is the baseline: its examples allocate that way
    # (examples/simple/simple.c: dwt_util_get_opt_stride) and its column passes run several times faster than on packed rows
    row = impl.opt_stride(W * 4) if hasattr(impl, "opt_stride") else W * 4
    images = {t: strided_image((H, W), t, row) for t in ("s", "i")}
    for _ in range(warmup):
        cpu_time_step(impl, threads, images)
    per_trial = [cpu_time_step(impl, threads, images) for _ in range(trials)]
    keys = list(per_trial[0])
    best = {k: min(tr[k] for tr in per_trial) for k in keys}          # min over trials of the mean over M = 1 images
    mean_step = sum(sum(tr.values()) for tr in per_trial) / trials
    return {"kind": kind, "cores": threads, "trials": trials, "warmup": warmup, "mean_step_s": mean_step,
            "min_of_means_s": best, "value_min_of_means": 4 * PIX / sum(best.values()) / 1e9, "value_mean": 4 * PIX / mean_step / 1e9}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_protocol(args.steps, args.warmup)
    ms = r["mean_step_s"] * 1e3
    value = r["value_mean"]
    line = {
        "impl": "reference", "metric": "cdf97_f32+cdf53_i32 fwd+inv throughput, 8192x8192 j=max", "value": value,
        "unit": "Gpixel/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
        "config": {"workload": "8192x8192 full-depth (J=13) 2-D DWT, CDF 9/7 float32 + CDF 5/3 int32, forward+inverse",
                   "sample": CPU_SAMPLE, "pattern": "dwt_util_test_image_fill_{s,i}"},
        "cpu_baseline": {"value": value, "unit": "Gpixel/s", "cores": r["cores"], "kind": r["kind"], "sample": CPU_SAMPLE,
                         "protocol": "mean over the timed steps (= the line's value); min over the steps of the per-direction times, the "
                                     "reference's own perf protocol (src/libdwt.c:21473-21507), in value_min_of_means",
                         "value_min_of_means": r["value_min_of_means"], "seconds_min_of_means": r["min_of_means_s"]},
        "e2e": {"value": value, "unit": "Gpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ======================================================================================================
# GPU arm
# ======================================================================================================
def run_ours(args):
    import torch
    import libdwt_b200 as d

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libdwt_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = d.lib()
    L.init(local)
    M = args.images

    kinds = [(d.CDF97_F32, "97s", 4), (d.CDF53_I32, "53i", 4)]
    # the M independent images of one sample type are ONE device-resident batch (dwtb200_image with M
    # frames): every level of the pyramid is one launch over all frames
    imgs = {}
    for k, name, es in kinds:
        imgs[name] = d.DeviceImage(k, W, H, M)
        imgs[name].fill(0, 0, 6)
    L.check(L.c.dwtb200_sync())

    launches = [0]

    def step(count=False):
        for _, name, _ in kinds:
            im = imgs[name]
            j = im.fwd2()
            assert j == J
            if count:
                launches[0] += im.last_launches
            im.inv2(J)
            if count:
                launches[0] += im.last_launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        time.sleep(0.3)   # let nvidia-smi start sampling
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    L.check(L.c.dwtb200_timer_start())
    for s in range(args.steps):
        step(count=(s == 0))
    ms_total = L.c.dwtb200_timer_stop_ms()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = ms_total / args.steps
    if world > 1:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    pix_step = 4 * M * PIX
    value = world * pix_step / (ms * 1e-3) / 1e9

    # ---- per-transform breakdown (device events): the batch of M, and a single image (frames = 1) ----
    peak, peak_src = peaks()
    breakdown = {}

    def time_dirs(im, frames, reps=5):
        out = {}
        tf = ti = 0.0
        for _ in range(reps):
            L.check(L.c.dwtb200_timer_start())
            im.fwd2()
            tf += L.c.dwtb200_timer_stop_ms()
            L.check(L.c.dwtb200_timer_start())
            im.inv2(J)
            ti += L.c.dwtb200_timer_stop_ms()
        for direction, t in (("fwd", tf), ("inv", ti)):
            t = t / reps * 1e-3
            b = algorithmic_bytes(W, H, J, 4) * frames
            out[direction] = {"us_per_image": t / frames * 1e6, "gpixel_s": PIX * frames / t / 1e9, "gb_s": b / t / 1e9,
                              "roofline_frac": b / t / 1e9 / peak}
        return out

    for _, name, es in kinds:
        r = time_dirs(imgs[name], M)
        breakdown[f"{name}_fwd_batch{M}"], breakdown[f"{name}_inv_batch{M}"] = r["fwd"], r["inv"]
    for k, name, es in kinds:   # one image at a time: 3 images cycled so that none is L2-resident
        singles = [d.DeviceImage(k, W, H, 1) for _ in range(3)]
        for im in singles:
            im.fill(0, 0, 0)
            im.fwd2(); im.inv2(J)
        L.check(L.c.dwtb200_sync())
        tf = ti = 0.0
        reps = 5
        for _ in range(reps):   # one image at a time (independent images would overlap on their own streams)
            for im in singles:
                L.check(L.c.dwtb200_timer_start())
                im.fwd2()
                tf += L.c.dwtb200_timer_stop_ms()
            for im in singles:
                L.check(L.c.dwtb200_timer_start())
                im.inv2(J)
                ti += L.c.dwtb200_timer_stop_ms()
        for direction, t in (("fwd", tf), ("inv", ti)):
            t = t / (reps * 3) * 1e-3
            b = algorithmic_bytes(W, H, J, 4)
            breakdown[f"{name}_{direction}_single"] = {"us_per_image": t * 1e6, "gpixel_s": PIX / t / 1e9, "gb_s": b / t / 1e9,
                                                        "roofline_frac": b / t / 1e9 / peak, "launches": singles[0].last_launches}
        for im in singles:
            im.close()

    # ---- roofline of the dominant kernel: level 0 of the forward 9/7 float transform ----
    # one launch = one j_max=1 transform of the batch: reads M 8192^2 planes once, writes their four subbands once
    im = imgs["97s"]
    for _ in range(2):
        im.fwd2(1)
    L.check(L.c.dwtb200_sync())
    reps = 10
    L.check(L.c.dwtb200_timer_start())
    for _ in range(reps):
        im.fwd2(1)
    t_fwd0 = L.c.dwtb200_timer_stop_ms() / reps * 1e-3
    assert im.last_launches == 1
    bytes0 = 2 * 4 * PIX * M
    im.fill(0, 0, 6)
    roof = {"bound": "hbm", "kernel": f"k_fwd_ring2<W97F, 8> (level 0 of dwt_cdf97_2f_s, {M} frames of 8192x8192 per launch)",
            "achieved": bytes0 / t_fwd0 / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
            "frac": bytes0 / t_fwd0 / 1e9 / peak, "traffic": None, "algorithmic_bytes_per_launch": bytes0,
            "us_per_launch": t_fwd0 * 1e6,
            "whole_pyramid_frac": {k: v["roofline_frac"] for k, v in breakdown.items()}}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            tj = json.load(open(tr))
            roof["traffic"] = tj["k_fwd_ring2_w97f_level0_dram_bytes_per_launch_4_frames"] * M / 4
            roof["traffic_source"] = tj["source_ring2"]
        except Exception:
            pass

    # ---- end to end through the reference-facing host call: pinned host image, H2D + transform + D2H ----
    # Rows lie at dwt_util_get_opt_stride(W * 4) = the next prime (src/libdwt.c:20640: 32771 bytes for 8192 floats), the layout the
    # reference's examples allocate and the one its CPU arm is timed on; the packed layout is measured next to it.
    def next_prime(n):
        while True:
            if n > 1 and all(n % q for q in range(2, int(n ** 0.5) + 1)):
                return n
            n += 1

    e2e = None
    if rank == 0 or world > 1:
        C = __import__("ctypes")
        fwd = {"97s": d.dwt_cdf97_2f_s, "53i": d.dwt_cdf53_2f_i}
        inv = {"97s": d.dwt_cdf97_2i_s, "53i": d.dwt_cdf53_2i_i}
        res = {}

        def measure(layout, row):
            nbytes = row * H
            hp = {}
            for k, name, es in kinds:
                p = L.c.dwtb200_host_alloc(nbytes)
                if not p:
                    raise RuntimeError(L.c.dwtb200_last_error().decode())
                raw = (C.c_uint8 * nbytes).from_address(p)
                arr = np.ndarray(shape=(H, W), dtype=np.float32 if name == "97s" else np.int32, buffer=raw, strides=(row, 4))
                hp[name] = (p, arr)
            for k, name, es in kinds:   # fill host inputs by downloading the device pattern once
                imgs[name].download(hp[name][1], frame=0)

            def e2e_step():
                for _, name, _ in kinds:
                    p, arr = hp[name]
                    jj = [-1]
                    fwd[name](p, row, 4, W, H, W, H, jj, 0, 0)
                    inv[name](p, row, 4, W, H, W, H, jj[0], 0, 0)
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            n_e2e = max(2, min(args.steps, 5))
            for _ in range(n_e2e):
                e2e_step()
            torch.cuda.synchronize()
            te = (time.perf_counter() - t0) / n_e2e
            if world > 1:
                tt = torch.tensor([te], device="cuda", dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                te = float(tt.item())
            res[layout] = (te, nbytes)
            for name in hp:
                L.c.dwtb200_host_free(hp[name][0])
        measure("packed", W * 4)
        layout = "packed"
        try:   # the layout the reference arm is timed on; the packed number stands if anything goes wrong with it
            measure("opt_stride", next_prime(W * 4))
            layout = "opt_stride"
        except Exception as e:
            print(f"bench: e2e on the opt-stride layout failed ({e}); reporting packed rows", file=sys.stderr)
        te = res[layout][0]
        e2e = {"value": world * 4 * PIX / te / 1e9, "unit": "Gpixel/s", "h2d_bytes_per_step": 4 * W * H * 4, "d2h_bytes_per_step": 4 * W * H * 4,
               "ms_per_step": te * 1e3, "layout": layout, "row_stride_bytes": next_prime(W * 4) if layout == "opt_stride" else W * 4,
               "packed_rows": {"value": world * 4 * PIX / res["packed"][0] / 1e9, "ms_per_step": res["packed"][0] * 1e3},
               "what": "dwt_cdf97_2f_s+2i_s and dwt_cdf53_2f_i+2i_i on a pinned host 8192x8192 image (4 transforms), rows at "
                       + ("dwt_util_get_opt_stride" if layout == "opt_stride" else "the packed stride")}

    # (The legs below allocate and free multi-GB buffers.  Measured: right after an 8 GB cudaFree the level-0 kernel of the batch runs
    # 9 % slower -- 389 instead of 357 us -- until the next large allocation; so the roofline and end-to-end legs above run first, in
    # the state the timed steps ran in.)
    # ---- the interleaved in-place family (dwt_cdf97_2f_inplace_s / dwt_cdf53_2f_inplace_s) on the same
    # shape: one image at a time and the batch of M, same byte accounting (the layout translation is overhead, not algorithmic bytes)
    inplace = {}
    if rank == 0 and world == 1:
        for wname, kind in (("97s", d.CDF97_F32), ("53s", d.CDF53_F32)):
            for frames in (1, M):
                ims = [d.DeviceImage(kind, W, H, frames) for _ in range(3 if frames == 1 else 1)]
                for im in ims:
                    im.fill(0, 0, 6)
                    jj = im.fwd2_inplace(); im.inv2_inplace(jj)
                L.check(L.c.dwtb200_sync())
                tf = ti = 0.0
                reps = 5
                for _ in range(reps):
                    for im in ims:
                        L.check(L.c.dwtb200_timer_start())
                        im.fwd2_inplace()
                        tf += L.c.dwtb200_timer_stop_ms()
                    for im in ims:
                        L.check(L.c.dwtb200_timer_start())
                        im.inv2_inplace(jj)
                        ti += L.c.dwtb200_timer_stop_ms()
                for direction, t in (("fwd", tf), ("inv", ti)):
                    t = t / (reps * len(ims)) * 1e-3
                    b = algorithmic_bytes(W, H, jj, 4) * frames
                    inplace[f"{wname}_{direction}_{'single' if frames == 1 else 'batch%d' % frames}"] = {
                        "us_per_image": t / frames * 1e6, "gpixel_s": PIX * frames / t / 1e9, "roofline_frac": b / t / 1e9 / peak,
                        "launches": ims[0].last_launches}
                for im in ims:
                    im.close()

    # ---- 3-D, one level, 1024^3 float (BASELINE config 5b): cdf97_3f_op_sep_horizontal_s / cdf97_3i_ip_sep_horizontal_s ----
    volume = None
    if rank == 0 and world == 1:
        try:
            n = 1024
            v = d.DeviceVolume(n, n, n)
            v.fill()
            v.fwd3(); v.inv3()
            L.check(L.c.dwtb200_sync())
            reps = 5
            L.check(L.c.dwtb200_timer_start())
            for _ in range(reps):
                v.fwd3()
            tf = L.c.dwtb200_timer_stop_ms() / reps * 1e-3
            L.check(L.c.dwtb200_timer_start())
            for _ in range(reps):
                v.inv3()
            ti = L.c.dwtb200_timer_stop_ms() / reps * 1e-3
            b = 2 * 4 * n ** 3   # every voxel read once and written once
            volume = {"workload": f"{n}^3 float volume, one level, forward / inverse, device-resident", "fwd_ms": tf * 1e3, "inv_ms": ti * 1e3,
                      "fwd_gvoxel_s": n ** 3 / tf / 1e9, "inv_gvoxel_s": n ** 3 / ti / 1e9,
                      "fwd_roofline_frac": b / tf / 1e9 / peak, "inv_roofline_frac": b / ti / 1e9 / peak}
            try:   # DRAM traffic of the two launches from the committed ncu capture (profiles/ncu_vol3t_r2.txt)
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                if n == 1024:
                    volume["kernel"] = "k_vol3t (one pass over the volume, tile staged by cp.async.bulk.tensor)"
                    volume["algorithmic_bytes"] = b
                    volume["fwd_traffic"] = tj["k_vol3t_fwd_1024cubed_dram_bytes_per_launch"]
                    volume["inv_traffic"] = tj["k_vol3t_inv_1024cubed_dram_bytes_per_launch"]
            except Exception:
                pass
            # end to end through the reference-facing host entry points (cdf97_3f_op_sep_horizontal_s / cdf97_3i_ip_sep_horizontal_s ->
            # dwtb200_fwd3_host / dwtb200_inv3_host) on page-locked volumes (volume_alloc_realiably_locked of the compat layer)
            try:
                C = __import__("ctypes")
                nb = 4 * n ** 3
                hs, hd = L.c.dwtb200_host_alloc(nb), L.c.dwtb200_host_alloc(nb)
                if not hs or not hd:
                    raise RuntimeError(L.c.dwtb200_last_error().decode())
                src = np.ndarray(shape=(n, n, n), dtype=np.float32, buffer=(C.c_uint8 * nb).from_address(hs))
                dst = np.ndarray(shape=(n, n, n), dtype=np.float32, buffer=(C.c_uint8 * nb).from_address(hd))
                v.fill()
                v.download(src)
                d.fwd3(src, dst)   # warm-up (allocations)
                t0 = time.perf_counter()
                d.fwd3(src, dst)
                t1 = time.perf_counter()
                d.inv3(dst)
                t2 = time.perf_counter()
                volume["e2e"] = {"fwd_ms": (t1 - t0) * 1e3, "inv_ms": (t2 - t1) * 1e3, "fwd_gvoxel_s": n ** 3 / (t1 - t0) / 1e9,
                                 "inv_gvoxel_s": n ** 3 / (t2 - t1) / 1e9, "h2d_bytes": nb, "d2h_bytes": nb,
                                 "roundtrip_max_abs_err": float(np.abs(dst[::97, ::89] - src[::97, ::89]).max()),
                                 "what": "pinned host volumes through dwtb200_fwd3_host / dwtb200_inv3_host: z ranges uploaded, transformed and downloaded in a pipeline (wall clock per call)"}
                L.c.dwtb200_host_free(hs); L.c.dwtb200_host_free(hd)
            except Exception as e:
                volume["e2e"] = {"skipped": str(e)}
            v.close()
        except Exception as e:
            volume = {"skipped": str(e)}

    # ---- one image far larger than L2 and than the launch overheads: BASELINE config 5a on a single GPU ----
    large = None
    if rank == 0 and world == 1 and not args.no_large:
        try:
            n = 32768   # 4 GiB per float plane, 9.3 GiB per image object
            for k, name, es in kinds:
                big = d.DeviceImage(k, n, n, 1)
                big.fill(0, 0, 0)
                jb = big.fwd2()
                big.inv2(jb)
                L.check(L.c.dwtb200_sync())
                L.check(L.c.dwtb200_timer_start())
                big.fwd2()
                tf = L.c.dwtb200_timer_stop_ms() * 1e-3
                L.check(L.c.dwtb200_timer_start())
                big.inv2(jb)
                ti = L.c.dwtb200_timer_stop_ms() * 1e-3
                big.close()
                b = algorithmic_bytes(n, n, jb, 4)
                large = large or {"workload": f"one {n}x{n} image, J={jb}, forward / inverse, device-resident"}
                large[name] = {"fwd_ms": tf * 1e3, "inv_ms": ti * 1e3, "fwd_gpixel_s": n * n / tf / 1e9, "inv_gpixel_s": n * n / ti / 1e9,
                               "fwd_roofline_frac": b / tf / 1e9 / peak, "inv_roofline_frac": b / ti / 1e9 / peak}
        except Exception as e:   # not enough memory on a shared GPU: the headline numbers do not depend on this leg
            large = {"skipped": str(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_protocol(5, 1)   # the protocol of `--impl reference`: 1 warm-up, 5 trials
        cpu = {"value": r["value_mean"], "unit": "Gpixel/s", "cores": r["cores"], "kind": r["kind"], "sample": CPU_SAMPLE,
               "protocol": "1 untimed warm-up + 5 trials, mean over the trials (same function as --impl reference); the reference's own "
                           "min-of-means protocol (src/libdwt.c:21473-21507) in value_min_of_means",
               "value_min_of_means": r["value_min_of_means"], "seconds_min_of_means": r["min_of_means_s"]}

    # ---- legs that need the memory of the batches: free them first ----
    for name in list(imgs):
        imgs[name].close()
    imgs.clear()
    link = batch = strips = None
    if not args.no_extra:
        for what in ("pcie", "batch_2048", "strips"):
            try:
                if what == "pcie":
                    link = pcie_leg(L, d, torch, dist if world > 1 else None, world, barrier)
                elif what == "batch_2048":
                    batch = batch_2048_leg(L, d, torch, dist if world > 1 else None, rank, world, peak, barrier)
                elif world > 1:
                    strips = strips_leg(L, d, torch, dist, rank, world, peak, barrier, size=args.strips_size)
            except Exception as e:   # an auxiliary leg must not take the headline line down with it
                msg = {"failed": f"{type(e).__name__}: {e}"}
                if what == "pcie":
                    link = msg
                elif what == "batch_2048":
                    batch = msg
                else:
                    strips = msg
    if e2e and link and "duplex_ms" in link:
        # one transform call moves 256 MiB up and 256 MiB down; `duplex_ms` is that pair of copies with nothing else in the way
        best = min([link] + ([link["library_alloc"]] if "library_alloc" in link else []), key=lambda m: m["duplex_ms"])
        e2e["link_ceiling_gbs"] = best["duplex_gbs"]
        e2e["fraction_of_link_ceiling"] = (4 * best["duplex_ms"]) / e2e["ms_per_step"]
        e2e["numa_node_of_gpu"] = link.get("numa_node_of_gpu")

    single = None
    try:
        ts = sum(breakdown[f"{name}_{dr}_single"]["us_per_image"] for name in ("97s", "53i") for dr in ("fwd", "inv")) * 1e-6
        single = {"value": world * 4 * PIX / ts / 1e9, "unit": "Gpixel/s",
                  "what": "the same four transforms one image per call (the granularity of dwt_cdf97_2f_s), device-resident, summed device time"}
    except Exception:
        pass

    if rank == 0:
        line = {
            "metric": "cdf97_f32+cdf53_i32 fwd+inv throughput, 8192x8192 j=max", "value": value, "value_single_image": single, "unit": "Gpixel/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
            "config": {"workload": "8192x8192 full-depth (J=13) 2-D DWT, CDF 9/7 float32 + CDF 5/3 int32, forward+inverse",
                       "images_per_type_per_step": M, "transforms_per_step": 4 * M, "pixels_per_step": pix_step,
                       "batching": f"the {M} images of a type are one device-resident batch: one kernel launch per pyramid level for all of them",
                       "pattern": "dwt_util_test_image_fill_{s,i}, rand = image index % 6",
                       "l2": "each batch is M x 256 MiB per plane (>> 126 MB L2): inputs larger than L2",
                       "sharding": "independent frames per GPU, no collective"},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches[0] * args.steps, "clocks": clocks,
            "breakdown": breakdown, "inplace_family": inplace, "volume": volume, "large_image": large,
            "pcie": link, "batch_2048": batch, "strips_65536": strips,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()



# ======================================================================================================
# auxiliary legs of the GPU arm (each returns a dict that goes into the JSON line)
# ======================================================================================================
def gather_max(x, world, torch, dist):
    if world == 1:
        return x
    tt = torch.tensor(x if isinstance(x, (list, tuple)) else [x], device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return tt.tolist() if isinstance(x, (list, tuple)) else float(tt.item())


def pcie_leg(L, d, torch, dist, world, barrier, nbytes=256 << 20):
    """The host link ceiling the end-to-end number is bounded by: plain pinned cudaMemcpyAsync of 256 MiB host->device,
    device->host and both at once, all ranks concurrently (max over ranks of the time, summed bytes) -- once from torch's pinned
    allocator (cudaHostAlloc wherever the thread runs) and once from dwtb200_host_alloc (bound to the GPU's NUMA node)."""
    C = __import__("ctypes")
    du, dd = torch.empty(nbytes, dtype=torch.uint8, device="cuda"), torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def measure(hu, hd):
        def timed(up, dn, reps=5):
            best = 1e30
            for _ in range(reps + 1):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                s1.wait_event(e0); s2.wait_event(e0)
                if up:
                    with torch.cuda.stream(s1):
                        du.copy_(hu, non_blocking=True)
                if dn:
                    with torch.cuda.stream(s2):
                        hd.copy_(dd, non_blocking=True)
                torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e-3)
            return gather_max(best, world, torch, dist)
        t_up, t_dn, t_both = timed(True, False), timed(False, True), timed(True, True)
        return {"h2d_gbs": world * nbytes / t_up / 1e9, "d2h_gbs": world * nbytes / t_dn / 1e9, "duplex_gbs": 2 * world * nbytes / t_both / 1e9,
                "duplex_ms": t_both * 1e3}

    hu = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)   # cudaHostAlloc
    hd = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    hu.fill_(1)
    out = {"bytes_each_way": nbytes, "ranks": world,
           "what": "pinned cudaMemcpyAsync (torch copy_ non_blocking), best of 5, all ranks at once: aggregate GB/s over the ranks"}
    out.update(measure(hu, hd))
    del hu, hd
    node = L.c.dwtb200_host_numa_node()
    out["numa_node_of_gpu"] = node
    pu, pd = L.c.dwtb200_host_alloc(nbytes), L.c.dwtb200_host_alloc(nbytes)
    if pu and pd:
        hu = torch.frombuffer((C.c_uint8 * nbytes).from_address(pu), dtype=torch.uint8)
        hd = torch.frombuffer((C.c_uint8 * nbytes).from_address(pd), dtype=torch.uint8)
        hu.fill_(1)
        if hu.is_pinned() and hd.is_pinned():
            out["library_alloc"] = measure(hu, hd)
            out["library_alloc"]["what"] = "the same copies from dwtb200_host_alloc blocks (mmap + mbind to the GPU's NUMA node + cudaHostRegister)"
        del hu, hd
        L.c.dwtb200_host_free(pu); L.c.dwtb200_host_free(pd)
    return out


def batch_2048_leg(L, d, torch, dist, rank, world, peak, barrier, frames_total=4096, chunk=512):
    """BASELINE config 4: 4096 independent 2048x2048 CDF 9/7 float frames sharded over the ranks (frame k of the GLOBAL batch is
    filled with rand = k % 6; no collective), J = 11, forward then inverse -- the reference's perf protocol
    (dwt_util_perf_cdf97_2_s, src/libdwt.c:21391-21507) with M = frames per rank.  A rank holds its frames in chunks of at most
    `chunk` frames (one dwtb200_image each: one launch per level for the whole chunk)."""
    w = h = 2048
    per_rank = frames_total // world
    first = rank * per_rank
    nchunk = max(1, -(-per_rank // chunk))
    sizes = [per_rank // nchunk + (1 if i < per_rank % nchunk else 0) for i in range(nchunk)]
    # memory: 2.33 planes of 16 MiB per frame -> 512 frames = 18.6 GiB; all of a rank's chunks stay resident when they fit
    free_b, _ = torch.cuda.mem_get_info()
    need = per_rank * 16 * (1 << 20) * 2.4
    resident = need < free_b * 0.9
    b1 = algorithmic_bytes(w, h, 11, 4)

    def fill(im, k0):
        # frame f of this chunk is global frame k0 + f: rand = (k0 + f) % 6 -> the fill takes rand_mod = 6 and a start offset
        # through `rand` (dwtb200_image_fill: frame k uses (rand + k) % rand_mod when rand_mod > 0)
        im.fill(k0 % 6, 0, 6)

    imgs, k0s = [], []
    k = first
    for n in sizes:
        if resident or not imgs:
            imgs.append(d.DeviceImage(d.CDF97_F32, w, h, n))
        k0s.append(k)
        k += n
    tf = ti = 0.0
    reps = 3
    for rep in range(reps + 1):   # rep 0 warms up (graph capture)
        f_ms = i_ms = 0.0
        for ci, n in enumerate(sizes):
            im = imgs[ci] if resident else imgs[0]
            if not resident and n != im.frames:
                im.close()
                imgs[0] = im = d.DeviceImage(d.CDF97_F32, w, h, n)
            if not resident or rep == 0:
                fill(im, k0s[ci])
            L.check(L.c.dwtb200_sync())
            if ci == 0:
                barrier()
            L.check(L.c.dwtb200_timer_start())
            assert im.fwd2() == 11
            f_ms += L.c.dwtb200_timer_stop_ms()
            L.check(L.c.dwtb200_timer_start())
            im.inv2(11)
            i_ms += L.c.dwtb200_timer_stop_ms()
        if rep:
            tf += f_ms
            ti += i_ms
    tf, ti = gather_max([tf / reps * 1e-3, ti / reps * 1e-3], world, torch, dist)
    # bit-exact check of a sample of frames of this rank against the oracle (first, a middle one, last)
    checked, bad = [], 0
    try:
        from oracle.orc import Oracle
        orc = Oracle()
        im = imgs[0]
        fill(im, k0s[0])
        im.fwd2()
        for f in sorted({0, im.frames // 2, im.frames - 1}):
            a = orc.fill(np.zeros((h, w), np.float32), "s", rand=(k0s[0] + f) % 6)
            orc.fwd2(a, "97", "s")
            got = im.download(frame=f)
            checked.append(k0s[0] + f)
            bad += int(got.tobytes() != a.tobytes())
        im.inv2(11)
    except Exception as e:   # the oracle is the checker only; its absence must not break the measurement
        checked = [f"oracle unavailable: {e}"]
    bad = int(gather_max(float(bad), world, torch, dist))
    # one chunk end to end: pinned host frames -> device, forward, coefficients -> host
    e2e = None
    try:
        n = min(sizes[0], 64)
        nb = n * w * h * 4
        C = __import__("ctypes")
        hp = L.c.dwtb200_host_alloc(nb)
        arr = np.ndarray(shape=(n, h, w), dtype=np.float32, buffer=(C.c_uint8 * nb).from_address(hp))
        sub = d.DeviceImage(d.CDF97_F32, w, h, n)
        sub.fill(first % 6, 0, 6)
        for f in range(n):
            sub.download(arr[f], frame=f)
        sub.close()
        # the chunk as `ng` groups of frames, each a device image with a stream of its own: the upload of group i + 1 is queued before
        # group i is transformed and downloaded, so the two directions of the link overlap (uploads are asynchronous, a download
        # waits for its own image only)
        ng = 8 if n % 8 == 0 else 1
        per = n // ng
        groups = [d.DeviceImage(d.CDF97_F32, w, h, per) for _ in range(ng)]

        def up(i):
            for f in range(per):
                groups[i].upload(arr[i * per + f], frame=f)
        best = 1e30
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            up(0)
            for i in range(ng):
                if i + 1 < ng:
                    up(i + 1)
                groups[i].fwd2()
                for f in range(per):
                    groups[i].download(arr[i * per + f], frame=f)
            best = min(best, time.perf_counter() - t0)
            for gi in groups:   # back to samples for the next repetition (not timed)
                gi.inv2(11)
                for f in range(per):
                    gi.download(arr[groups.index(gi) * per + f], frame=f)
        for gi in groups:
            gi.close()
        L.c.dwtb200_host_free(hp)
        best = gather_max(best, world, torch, dist)
        e2e = {"frames_per_rank": n, "gpixel_s": world * n * w * h / best / 1e9, "ms": best * 1e3,
               "what": "pinned host frames uploaded, forward transform, coefficients downloaded, in 8 groups of frames whose copies overlap (wall clock, max over ranks)"}
    except Exception as e:
        e2e = {"skipped": str(e)}
    for im in imgs:
        im.close()
    pix = world * per_rank * w * h
    return {"workload": f"{world * per_rank} frames of 2048x2048 float32 CDF 9/7, J=11, {per_rank} per GPU in {nchunk} chunk(s) of <= {max(sizes)} "
                        f"frames ({'resident' if resident else 'one chunk resident at a time, refilled'}), frame k filled with rand = k % 6",
            "fwd_gpixel_s": pix / tf / 1e9, "inv_gpixel_s": pix / ti / 1e9, "fwd_ms": tf * 1e3, "inv_ms": ti * 1e3,
            "fwd_roofline_frac_per_gpu": b1 * per_rank / tf / 1e9 / peak, "inv_roofline_frac_per_gpu": b1 * per_rank / ti / 1e9 / peak,
            "frames_checked_against_oracle": checked, "frames_differing": bad, "e2e_one_chunk": e2e}


def strips_leg(L, d, torch, dist, rank, world, peak, barrier, size=65536, levels=0):
    """BASELINE config 5a: ONE size x size CDF 9/7 float image as row strips over the ranks (dwtb200_strips_*: peer copies over NVLink,
    device-side flags, no host synchronisation inside a call); forward and inverse, CUDA events, max over ranks; strong-scaling
    efficiency against the single-GPU transform of the same image timed on the rank's own GPU in the same run; every rank's owned rows
    compared bit for bit with that single-GPU result, and level 0 of rank 0's strip with the oracle on a row window (separability)."""
    import uuid
    obj = [uuid.uuid4().hex[:12] if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(obj, src=0)
    s = d.DeviceStrips(d.CDF97_F32, size, size, rank, world, "dwtb200-bench-" + obj[0], levels)
    s.connect()
    wide = 1 if size > 16384 else 0
    full = d.DeviceImage(d.CDF97_F32, size, size)
    # the single-GPU transform of the whole picture (on every rank: it is also the checker of the rank's rows)
    t1f = t1i = 1e30
    J = 0
    for _ in range(3):
        full.fill(0, 0, 0, 0, wide)
        L.check(L.c.dwtb200_sync())
        L.check(L.c.dwtb200_timer_start())
        J = full.fwd2()
        t1f = min(t1f, L.c.dwtb200_timer_stop_ms())
        L.check(L.c.dwtb200_timer_start())
        full.inv2(J)
        t1i = min(t1i, L.c.dwtb200_timer_stop_ms())
    t1f, t1i = gather_max([t1f, t1i], world, torch, dist)
    full.fill(0, 0, 0, 0, wide)
    full.fwd2()
    L.check(L.c.dwtb200_sync())
    a0, m = 1024, 72                      # rank 0: input rows [a0, a0 + 2 m) for the oracle, output row pairs [k0, k1)
    k0, k1 = a0 // 2 + 8, a0 // 2 + m - 8
    win = got_l = got_h = None
    tf = ti = 1e30
    diff_f = diff_i = -1
    peer_f = peer_i = 0
    for rep in range(4):
        s.fill(0, 0, wide)
        s.sync()
        if rep == 0 and rank == 0:
            win = np.empty((2 * m, size), np.float32)
            s.image.copy_rows(a0, 2 * m, win.ctypes.data, win.strides[0], False)
            L.check(L.c.dwtb200_sync())
        barrier()
        L.check(L.c.dwtb200_timer_start())
        assert s.fwd2() == J
        ms = L.c.dwtb200_timer_stop_ms()
        s.sync()
        peer_f = s.last_peer_bytes
        ms = gather_max(ms, world, torch, dist)
        if rep:
            tf = min(tf, ms)
        else:
            diff_f = s.compare_owned(full, True)
            if rank == 0:   # rows k: [.. | HL_0], rows nly + k: [LH_0 | HH_0] of the strip's Mallat plane
                nly_l = (s.plan.ext1 - s.plan.ext0 + 1) // 2
                got_l, got_h = np.empty((k1 - k0, size), np.float32), np.empty((k1 - k0, size), np.float32)
                s.image.copy_rows(k0, k1 - k0, got_l.ctypes.data, got_l.strides[0], False)
                s.image.copy_rows(nly_l + k0, k1 - k0, got_h.ctypes.data, got_h.strides[0], False)
                L.check(L.c.dwtb200_sync())
        barrier()
        L.check(L.c.dwtb200_timer_start())
        s.inv2(J)
        ms = L.c.dwtb200_timer_stop_ms()
        s.sync()
        peer_i = s.last_peer_bytes
        ms = gather_max(ms, world, torch, dist)
        if rep:
            ti = min(ti, ms)
        else:
            full.inv2(J)
            diff_i = s.compare_owned(full, False)
            L.check(L.c.dwtb200_sync())
    diffs = gather_max([float(diff_f), float(diff_i)], world, torch, dist)
    peers = [float(peer_f), float(peer_i)]
    if world > 1:
        tt = torch.tensor(peers, device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        peers = tt.tolist()
    oracle_rows = None
    if rank == 0:
        try:   # level 0 of the strip against the oracle's one-level transform of the row window (interior rows only)
            from oracle.orc import Oracle
            ref = win.copy()
            Oracle().fwd2(ref, "97", "s", j_max=1)
            nlx, i0, i1 = size // 2, k0 - a0 // 2, k1 - a0 // 2
            bad = int((got_l[:, nlx:].view(np.uint32) != ref[i0:i1, nlx:].view(np.uint32)).sum())
            bad += int((got_h.view(np.uint32) != ref[m + i0:m + i1].view(np.uint32)).sum())
            oracle_rows = {"input_rows": [a0, a0 + 2 * m], "checked": f"HL, LH, HH rows {k0}..{k1 - 1} of level 0 at full width", "differing": bad}
        except Exception as e:
            oracle_rows = {"skipped": str(e)}
    barrier()
    full.close()
    plan = [s.plan.own0, s.plan.own1, s.plan.ext0, s.plan.ext1]
    Jd = s.Jd
    s.close()
    b = algorithmic_bytes(size, size, J, 4)
    return {"workload": f"one {size}x{size} float32 CDF 9/7 image, J={J}, row strips over {world} GPUs, {Jd} levels distributed, the rest on rank 0",
            "fwd_ms": tf, "inv_ms": ti, "fwd_gpixel_s": size * size / tf / 1e6, "inv_gpixel_s": size * size / ti / 1e6,
            "single_gpu_fwd_ms": t1f, "single_gpu_inv_ms": t1i, "fwd_efficiency": t1f / (world * tf), "inv_efficiency": t1i / (world * ti),
            "fwd_roofline_frac_per_gpu": b / world / (tf * 1e-3) / 1e9 / peak, "inv_roofline_frac_per_gpu": b / world / (ti * 1e-3) / 1e9 / peak,
            "peer_bytes_fwd": peers[0], "peer_bytes_inv": peers[1], "rank0_rows": plan,
            "owned_rows_differing_from_single_gpu": {"fwd": int(diffs[0]), "inv": int(diffs[1])}, "oracle_row_window": oracle_rows,
            "transport": "cudaMemcpy2DAsync on CUDA-IPC peer mappings (NVLink P2P), device-side sequence flags, no host sync inside a call",
            "timing": "CUDA events per rank from the call's first enqueued work to the rank's last, after a host barrier; max over ranks; best of 3"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=4, help="independent 8192^2 images per sample type per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-large", action="store_true", help="skip the 32768^2 single-image leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the pcie / batch_2048 / strips legs")
    ap.add_argument("--strips-size", type=int, default=65536, help="edge of the single image of the row-strip leg (N > 1)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 3 if args.impl == "reference" else 100
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
