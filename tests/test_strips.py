"""Row-strip partition (libdwt_b200/strips.py): CPU tests of the host logic.

* the scheme itself (extended strips, halo recomputation, ownership arithmetic) with all ranks emulated in
  one process and the oracle as arithmetic -> must equal the oracle's single-image transform bit for bit;
* the real multi-process driver (DistStrips) with world_size 2 and 3 over the gloo backend.
The oracle is used here only as the stand-in arithmetic and the checker (tests may do that)."""
import os
import sys

import numpy as np
import pytest

from cases import DT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_engine(oracle, w, t):
    from libdwt_b200 import strips
    return strips.NumpyEngine(lambda img, j: oracle.fwd2(img, w, t, j_max=j), lambda img, j: oracle.inv2(img, w, t, j_max=j))


@pytest.mark.parametrize("kind", [("97", "s"), ("53", "i"), ("97", "d")], ids=lambda k: k[0] + k[1])
def test_strip_scheme_is_bit_exact(oracle, kind):
    from libdwt_b200 import strips
    w, t = kind
    eng = oracle_engine(oracle, w, t)
    for (W, H) in ((256, 512), (300, 1000), (517, 777), (64, 2048)):
        for G in (2, 3, 8):
            for Jd in (1, 3):
                img = oracle.fill(np.zeros((H, W), DT[t]), t)
                want = img.copy()
                J = oracle.fwd2(want, w, t)
                got, J2 = strips.forward_strips_local(img, G, Jd, eng)
                assert J == J2 and got.tobytes() == want.tobytes(), (W, H, G, Jd)
                back = strips.inverse_strips_local(want, G, Jd, eng, J)
                ref = want.copy()
                oracle.inv2(ref, w, t, j_max=J)
                assert back.tobytes() == ref.tobytes(), (W, H, G, Jd)


def test_halo_of_two_rows_per_level_is_enough_for_cdf53(oracle):
    """The C ABI (csrc/strips.cu) holds HALO * 2^Jd rows beyond the owned ones with HALO = the lifting reach: 2 for CDF 5/3."""
    from libdwt_b200 import strips
    for (w, t) in (("53", "i"), ("53", "s")):
        eng = oracle_engine(oracle, w, t)
        for (W, H, G, Jd) in ((256, 512, 2, 3), (300, 1000, 3, 2), (64, 2048, 8, 3), (517, 777, 2, 1)):
            img = oracle.fill(np.zeros((H, W), DT[t]), t)
            want = img.copy()
            J = oracle.fwd2(want, w, t)
            got, _ = strips.forward_strips_local(img, G, Jd, eng, halo_lines=2)
            assert got.tobytes() == want.tobytes(), (w, t, W, H, G, Jd)
            back = strips.inverse_strips_local(want, G, Jd, eng, J, halo_lines=2)
            ref = want.copy()
            oracle.inv2(ref, w, t, j_max=J)
            assert back.tobytes() == ref.tobytes(), (w, t, W, H, G, Jd)


def test_c_abi_plan_equals_the_python_model():
    """dwtb200_strips_plan / dwtb200_strips_band (host arithmetic of csrc/strips.cu, no device needed) against StripPlan."""
    import libdwt_b200 as d
    from libdwt_b200.strips import StripPlan
    cases = [(65536, 65536, 8, 5, 4), (65536, 65536, 8, 4, 4), (8192, 8192, 2, 2, 4), (300, 1000, 3, 3, 4), (517, 777, 2, 1, 2),
             (64, 2048, 8, 3, 2), (7919, 6007, 4, 3, 4), (4096, 4097, 4, 4, 4), (130, 777, 3, 2, 4), (1000, 37, 1, 2, 4)]
    for (W, H, G, Jd, hl) in cases:
        p = StripPlan(W, H, G, Jd, hl)
        for r in range(G):
            c = d.strips_plan(W, H, G, Jd, hl, r)
            assert (c.own0, c.own1) == p.owned(r) and (c.ext0, c.ext1) == p.extended(r) and c.halo == p.halo, (W, H, G, Jd, r)
            own, ext, off = p.ll_rows(r)
            assert (c.ll_own0, c.ll_own1) == own and (c.ll_ext0, c.ll_ext1) == ext and c.ll_ext0 == off
            assert (c.ll_w, c.ll_h) == (-(-W >> Jd), -(-H >> Jd))
            nonempty = all(b > a for a, b in (p.owned(q) for q in range(G)))
            assert bool(c.neighbours_only) == (p.neighbours_only() and nonempty or G == 1 and nonempty)
            for j in range(Jd):
                b, m = d.strips_band(W, H, G, Jd, hl, r, j), p.bands(r, j)
                assert (b["off"], b["nly_g"], b["nly_l"]) == (m["off"], m["nly_g"], m["nly_l"])
                assert (b["extL0"], b["extL1"]) == m["extL"] and (b["extH0"], b["extH1"]) == m["extH"]
                assert (b["ownL0"], b["ownL1"]) == m["ownL"] and (b["ownH0"], b["ownH1"]) == m["ownH"]


def test_plan_geometry():
    from libdwt_b200.strips import StripPlan
    p = StripPlan(65536, 65536, 8, 4)
    assert p.halo == 64 and p.R == [8192 * r for r in range(9)] and p.neighbours_only()
    assert p.extended(0) == (0, 8192 + 64) and p.extended(7) == (7 * 8192 - 64, 65536)
    q = StripPlan(300, 1000, 3, 3)
    assert q.R[0] == 0 and q.R[-1] == 1000 and all(v % 8 == 0 for v in q.R[:-1])


def _worker(rank, world, port, W, H, Jd, w, t, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from libdwt_b200 import strips
    from oracle.orc import Oracle
    orc = Oracle()
    orc.set_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Eng:   # oracle-backed stand-in for DeviceStripEngine
        def __init__(self, width, height):
            self.t = torch.zeros((height, width), dtype={"s": torch.float32, "d": torch.float64, "i": torch.int32}[t])

        def view(self):
            return self.t

        def fwd2(self, J):
            assert orc.fwd2(self.t.numpy(), w, t, j_max=J) == J

        def inv2(self, J):
            orc.inv2(self.t.numpy(), w, t, j_max=J)

    ds = strips.DistStrips(W, H, Jd, Eng, dist)
    full = orc.fill(np.zeros((H, W), DT[t]), t)
    ds.owned_view().copy_(torch.from_numpy(full[ds.a:ds.b]))
    ds.forward()
    mallat = ds.gather_mallat()
    ds.inverse()
    rec = [None] * world
    dist.all_gather_object(rec, ds.owned_view().numpy().copy())
    if rank == 0:
        want = full.copy()
        J = orc.fwd2(want, w, t)
        ok_f = mallat.tobytes() == want.tobytes()
        orc.inv2(want, w, t, j_max=J)
        ok_i = np.concatenate(rec).tobytes() == want.tobytes()
        open(os.path.join(out_dir, "result"), "w").write(f"{int(ok_f)}{int(ok_i)}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,Jd,kind", [(2, (256, 512), 2, ("97", "s")), (2, (300, 1000), 3, ("53", "i")),
                                                  (3, (130, 777), 2, ("97", "d"))],
                         ids=["w2-97s", "w2-53i", "w3-97d"])
def test_dist_strips_gloo(tmp_path, world, shape, Jd, kind):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, shape[0], shape[1], Jd, kind[0], kind[1], str(tmp_path)), nprocs=world, join=True)
    assert open(tmp_path / "result").read() == "11"
