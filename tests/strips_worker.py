"""Worker of tests/test_gpu_strips.py (and of gpurun checks): drives dwtb200_strips_* and checks its results.

  --mode emulate --world G                 all G ranks in THIS process on one GPU (the C ABI then uses raw pointers instead of IPC mappings)
  --mode rank --rank r --world G           one process per rank; the device is LOCAL_RANK (or --rank)

Checks, per round (two rounds, so that the sequence flags of a second collective call are exercised):
  * forward: the rows every rank owns, bit for bit against the single-device transform of the whole picture on the same GPU
    (dwtb200_strips_compare_owned) and, with --oracle, against the oracle's transform of the whole picture;
  * inverse likewise.
Writes a JSON result to --out (rank 0 / the emulating process) and exits 0 when everything matched."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DT = {"s": np.float32, "d": np.float64, "i": np.int32}


def poison_halo(d, s):
    """overwrite the halo rows of the strip with 7s so that only the exchange can supply them"""
    p = s.plan
    W = s.width
    for (row0, rows) in ((0, p.own0 - p.ext0), (p.own1 - p.ext0, p.ext1 - p.own1)):
        if rows > 0:
            junk = np.full((rows, W), 7, dtype=s.image.dtype)
            s.image.copy_rows(row0, rows, junk.ctypes.data, junk.strides[0], True)
            s.L.check(s.L.c.dwtb200_sync())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="emulate", choices=["emulate", "rank"])
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--height", type=int, default=4096)
    ap.add_argument("--levels", type=int, default=0)
    ap.add_argument("--kind", default="97s")
    ap.add_argument("--session", default="dwtb200-test")
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    if args.mode == "rank" and "LOCAL_RANK" not in os.environ:
        os.environ["LOCAL_RANK"] = str(args.rank)
    os.environ.setdefault("DWTB200_STRIPS_TIMEOUT_S", "20")
    import libdwt_b200 as d
    w, t = args.kind[:2], args.kind[2]
    kind = d.kind_of(w, t)
    L = d.lib()
    L.init()
    W, H, G = args.width, args.height, args.world
    wide = 1 if max(W, H) > 16384 else 0
    ranks = list(range(G)) if args.mode == "emulate" else [args.rank]
    strips = [d.DeviceStrips(kind, W, H, r, G, args.session, args.levels) for r in ranks]   # rank 0 first
    for s in strips:
        s.connect()
    full = d.DeviceImage(kind, W, H)
    res = {"world": G, "W": W, "H": H, "kind": args.kind, "J": strips[0].J, "Jd": strips[0].Jd, "rounds": [], "mode": args.mode,
           "plans": [[s.plan.own0, s.plan.own1, s.plan.ext0, s.plan.ext1] for s in strips]}
    want_f = want_i = None
    if args.oracle:
        from oracle.orc import Oracle
        orc = Oracle()
        a = orc.fill(np.zeros((H, W), DT[t]), t)
        Jo = orc.fwd2(a, w, t)
        assert Jo == strips[0].J
        want_f = a.copy()
        orc.inv2(a, w, t, j_max=Jo)
        want_i = a
    ok = True
    for rnd in range(args.rounds):
        for s in strips:
            s.fill(0, 0, wide)
            poison_halo(d, s)
        full.fill(0, 0, 0, 0, wide)
        J = full.fwd2()
        L.check(L.c.dwtb200_sync())
        for s in strips:
            assert s.fwd2() == J
        for s in strips:
            s.sync()
        r = {"fwd_diff": [s.compare_owned(full, True) for s in strips], "fwd_peer_bytes": [s.last_peer_bytes for s in strips]}
        if args.oracle:
            if args.mode == "emulate":
                host = np.full((H, W), 0x7f, dtype=DT[t])
                for s in strips:
                    s.download_owned(host, True)
            else:
                host = want_f.copy()
                strips[0].download_owned(host, True)
            r["fwd_oracle_equal"] = bool(host.tobytes() == want_f.tobytes())
            ok = ok and r["fwd_oracle_equal"]
        full.inv2(J)
        L.check(L.c.dwtb200_sync())
        for s in strips:
            s.inv2(J)
        for s in strips:
            s.sync()
        r["inv_diff"] = [s.compare_owned(full, False) for s in strips]
        r["inv_peer_bytes"] = [s.last_peer_bytes for s in strips]
        if args.oracle:
            if args.mode == "emulate":
                host = np.full((H, W), 0x7f, dtype=DT[t])
                for s in strips:
                    s.download_owned(host, False)
            else:
                host = want_i.copy()
                strips[0].download_owned(host, False)
            r["inv_oracle_equal"] = bool(host.tobytes() == want_i.tobytes())
            ok = ok and r["inv_oracle_equal"]
        ok = ok and not any(r["fwd_diff"]) and not any(r["inv_diff"])
        res["rounds"].append(r)
    res["ok"] = bool(ok)
    full.close()
    for s in reversed(strips):   # rank 0 (the owner of the session) last
        s.close()
    if args.out:
        with open(args.out + (f".{args.rank}" if args.mode == "rank" else ""), "w") as f:
            json.dump(res, f)
    print(json.dumps(res))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
