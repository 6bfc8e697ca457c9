"""GPU (B200): the CUDA path, called through the C ABI (libdwtb200.so), against the oracle on the same
inputs -- bit-exact for int 5/3 AND for float/double 9/7 (the kernels use explicitly rounded,
non-contracted arithmetic in the reference's operation order, so the tolerance is 0 ulp; the
reference's own tolerance is 1e-3 / 1e-6, src/libdwt.c:1604, 1513) -- and against the committed golden
digests of the compiled reference."""
import json
import os

import numpy as np
import pytest

from cases import (DENSE_SHAPES, DEPTHS, DT, KINDS, SPARSE, bits, case_id, dense_cases, describe_mismatch, digest,
                   sparse_cases)

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))
KIDS = [k[0] + k[1] for k in KINDS]


def both(dev, oracle, w, t, ox, oy, j, d1, zp=0, inner=None, row_bytes=None, rand=0, type_=0):
    """Run forward + inverse on the device path and on the oracle; returns a list of failure strings."""
    from oracle.orc import strided_image
    a = strided_image((oy, ox), t, row_bytes)
    b = strided_image((oy, ox), t, row_bytes)
    oracle.fill(a, t, rand=rand, type_=type_)
    b[...] = a
    fails = []
    tag = f"{w}{t} {ox}x{oy} j={j} d1={d1} zp={zp} inner={inner} row_bytes={row_bytes}"
    Ja = oracle.fwd2(a, w, t, j_max=j, decompose_one=d1, zero_padding=zp, inner=inner)
    Jb = dev.fwd2(b, w, t, j_max=j, decompose_one=d1, zero_padding=zp, inner=inner)
    if Ja != Jb:
        fails.append(f"{tag}: J {Jb} != {Ja}")
    if not (bits(a, t) == bits(b, t)).all():
        fails.append(f"{tag}: FORWARD " + describe_mismatch(b, a, t))
        b[...] = a   # continue from the right coefficients so the inverse is judged on its own
    oracle.inv2(a, w, t, j_max=Ja, decompose_one=d1, zero_padding=zp, inner=inner)
    dev.inv2(b, w, t, j_max=Ja, decompose_one=d1, zero_padding=zp, inner=inner)
    if not (bits(a, t) == bits(b, t)).all():
        fails.append(f"{tag}: INVERSE " + describe_mismatch(b, a, t))
    return fails


def report(fails):
    assert not fails, f"{len(fails)} failures:\n" + "\n".join(fails[:12])


def test_device_present_and_is_blackwell(dev):
    L = dev.lib()
    assert L.c.dwtb200_device_count() >= 1
    assert L.c.dwtb200_device() >= 0


@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
def test_test_pattern_matches_oracle(dev, oracle, kind):
    w, t = kind
    for (ox, oy, rnd, typ) in ((517, 301, 0, 0), (64, 64, 3, 0), (300, 200, 0, 2), (4096, 2100, 0, 0)):
        if t == "d" and typ != 0:
            continue
        img = dev.DeviceImage(dev.kind_of(w, t), ox, oy)
        img.fill(rnd, typ)
        got = img.download()
        want = oracle.fill(np.zeros((oy, ox), DT[t]), t, rand=rnd, type_=typ)
        assert (bits(got, t) == bits(want, t)).all(), describe_mismatch(got, want, t)
        img.close()


@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
@pytest.mark.parametrize("shape", [(16, 16), (128, 128), (31, 33), (64, 3)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_tail_only_shapes(dev, oracle, kind, shape):
    w, t = kind
    fails = []
    for (j, d1) in DEPTHS:
        fails += both(dev, oracle, w, t, shape[0], shape[1], j, d1)
    report(fails)


@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
def test_one_streamed_level(dev, oracle, kind):
    w, t = kind
    fails = []
    for (ox, oy) in ((256, 256), (300, 200), (241, 250), (479, 33), (129, 127)):
        fails += both(dev, oracle, w, t, ox, oy, 1, 0)
    report(fails)


@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
@pytest.mark.parametrize("shape", DENSE_SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_dense_parity(dev, oracle, kind, shape):
    w, t = kind
    ox, oy = shape
    fails = []
    for c in dense_cases():
        if c[:4] == (w, t, ox, oy):
            fails += both(dev, oracle, w, t, ox, oy, c[4], c[5])
    report(fails)


# kernel selection per level: (DWTB200_TUNE_TILE_MAX, DWTB200_TUNE_TAIL_MAX, DWTB200_TUNE_MID_MAX[, DWTB200_TUNE_RING]);
# "stream" levels take the bulk-copy ring kernels (ring = 3, default) or the register double-buffer kernels (ring = 0)
BIG = 1 << 40
FAMILIES = {"stream+bigtail": (0, 16384, 0), "tile+tail": (BIG, 1024, 0), "tile-only": (BIG, 0, 0), "stream-only": (0, 0, 0),
            "tile+tinytail": (BIG, 16, 0),
            "regstream-only": (0, 0, 0, 0), "regstream+bigtail": (0, 16384, 0, 0), "ringfwd-reginv": (0, 1024, 0, 1),
            "ring15x1": (0, 1024, 0, 3 | (1 << 4)), "ring5x3": (0, 1024, 0, 3 | (3 << 4)),
            # first-generation ring kernels (240-column windows, producer warp) forced: the default is the 256-column generation
            "ring7x2-gen1": (0, 1024, 0, 3 | (5 << 4)), "ring-gen2-forced": (0, 0, 0, 3 | (4 << 4)),
            # the same mixes without the dataflow chain between the kernels of a pyramid (DWTB200_TUNE_CHAIN = 0)
            "nochain-default": (1024 * 1024, 1024, 0, 3, 0), "nochain-stream": (0, 0, 0, 3, 0), "nochain-tile": (BIG, 16, 0, 3, 0),
            "chain-tile+bigtail": (BIG, 4096, 0, 3, 1), "chain-ring+tile": (64 * 64, 256, 0, 3, 1)}
DEFAULT_TUNING = (1024 * 1024, 4096, 0, 3, 1, 0)


def set_tuning(L, t):
    t = tuple(t) + (3, 1, 0)[len(t) - 3:] if len(t) < 6 else tuple(t)
    for key, v in zip((0, 1, 2, 6, 7, 8), t):
        L.check(L.c.dwtb200_set_tuning(key, v))


@pytest.mark.parametrize("family", list(FAMILIES), ids=list(FAMILIES))
@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
def test_every_kernel_family_gives_the_same_bits(dev, oracle, kind, family):
    """Streaming, tile and tail kernels are interchangeable per level: force each mix."""
    w, t = kind
    L = dev.lib()
    set_tuning(L, FAMILIES[family])
    fails = []
    try:
        for (ox, oy) in ((512, 512), (517, 301), (1000, 37), (5, 1000), (241, 250), (129, 127), (64, 3), (31, 33),
                         (1025, 1023), (2, 2), (3, 7), (65, 33), (66, 34), (63, 31)):
            for (j, d1) in ((-1, 0), (2, 0), (-1, 1)):
                fails += both(dev, oracle, w, t, ox, oy, j, d1)
    finally:
        set_tuning(L, DEFAULT_TUNING)
    report(fails)


@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
def test_ring_gen2_on_whole_window_widths(dev, oracle, kind):
    """kernels_ring2.cu takes rows that are a whole number of 256-column (double: 128-column) warp windows: one window, several windows
    of one band, more than one band (> 8 windows), few and many rows, with every level forced onto the streaming kernels."""
    w, t = kind
    L = dev.lib()
    set_tuning(L, (0, 0, 0))
    fails = []
    try:
        for (ox, oy) in ((256, 37), (128, 128), (512, 300), (768, 130), (1024, 77), (2304, 64), (2560, 33), (4096, 40)):
            for (j, d1) in ((-1, 0), (1, 0)):
                fails += both(dev, oracle, w, t, ox, oy, j, d1)
        img = dev.DeviceImage(dev.kind_of(w, t), 1024, 200, 3)   # a batch: frames in grid.y
        img.fill(0, 0, 6)
        J = img.fwd2()
        for k in range(3):
            want = oracle.fill(np.zeros((200, 1024), DT[t]), t, rand=k % 6)
            oracle.fwd2(want, w, t)
            got = img.download(frame=k)
            if not (bits(got, t) == bits(want, t)).all():
                fails.append(f"batch frame {k} forward: " + describe_mismatch(got, want, t))
        img.inv2(J)
        for k in range(3):
            want = oracle.fill(np.zeros((200, 1024), DT[t]), t, rand=k % 6)
            got = img.download(frame=k)
            ref = want.copy()
            oracle.fwd2(ref, w, t)
            oracle.inv2(ref, w, t, j_max=J)
            if not (bits(got, t) == bits(ref, t)).all():
                fails.append(f"batch frame {k} inverse: " + describe_mismatch(got, ref, t))
        img.close()
    finally:
        set_tuning(L, DEFAULT_TUNING)
    report(fails)


@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
def test_sparse_parity(dev, oracle, kind):
    w, t = kind
    fails = []
    for c in sparse_cases():
        if c[:2] == (w, t):
            _, _, ox, oy, ix, iy, j, d1, zp = c
            fails += both(dev, oracle, w, t, ox, oy, j, d1, zp=zp, inner=(iy, ix))
    report(fails)


@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
def test_generic_kernels_on_dense_shapes(dev, oracle, kind):
    w, t = kind
    L = dev.lib()
    L.c.dwtb200_force_generic(1)
    try:
        fails = []
        for (ox, oy) in ((517, 301), (64, 3), (5, 1000), (256, 256), (1, 9)):
            for (j, d1) in ((-1, 0), (-1, 1), (2, 0)):
                fails += both(dev, oracle, w, t, ox, oy, j, d1)
    finally:
        L.c.dwtb200_force_generic(0)
    report(fails)


@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
def test_strip_boundaries(dev, oracle, kind):
    """Every strip height must give the same bits: strips re-derive their lifting state from a halo."""
    w, t = kind
    L = dev.lib()
    fails = []
    try:
        set_tuning(L, (0, 1024, 0))   # streaming kernels on every level above the tail
        for tma in (1, 0):            # 16 or 32 bytes per lane (DWTB200_TUNE_NARROW)
            L.check(L.c.dwtb200_set_tuning(4, tma))
            for rows in (2, 4, 6, 16, 34, 128, 1000):
                L.c.dwtb200_set_strip_rows(rows)
                for (ox, oy) in ((517, 301), (300, 200), (256, 257), (1000, 333)):
                    fails += [f"narrow={tma} strip_rows={rows}: " + f for f in both(dev, oracle, w, t, ox, oy, -1, 0)]
            L.c.dwtb200_set_strip_rows(0)
            for (ox, oy) in ((2048, 1536), (1999, 1201)):
                fails += [f"narrow={tma}: " + f for f in both(dev, oracle, w, t, ox, oy, -1, 0)]
    finally:
        L.c.dwtb200_set_strip_rows(0)
        L.check(L.c.dwtb200_set_tuning(4, 0))
        set_tuning(L, DEFAULT_TUNING)
    report(fails)


@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
def test_unaligned_and_channel_strides(dev, oracle, ref_opt_stride, kind):
    """dwt_util_get_opt_stride() hands out prime row strides (2053 B for 512 floats): rows are not even
    element-aligned; cvdwt.cpp passes stride_y = pixel size of an interleaved multi-channel image."""
    w, t = kind
    es = np.dtype(DT[t]).itemsize
    fails = []
    for (ox, oy) in ((512, 512), (517, 301), (100, 77)):
        fails += both(dev, oracle, w, t, ox, oy, -1, 0, row_bytes=ref_opt_stride(ox * es))
    # 3-channel interleaved: transform channel 1 in place, the other channels must survive
    oy, ox, ch = 120, 200, 3
    base = np.zeros((oy, ox, ch), DT[t])
    for c in range(ch):
        plane = np.zeros((oy, ox), DT[t])
        oracle.fill(plane, t, rand=c)
        base[:, :, c] = plane
    a, b = base.copy(), base.copy()
    Ja = oracle.fwd2(a[:, :, 1], w, t)
    Jb = dev.fwd2(b[:, :, 1], w, t)
    if Ja != Jb or not (a.view(np.uint8) == b.view(np.uint8)).all():
        fails.append("channel-strided forward: " + describe_mismatch(b[:, :, 1], a[:, :, 1], t))
    b[...] = a
    oracle.inv2(a[:, :, 1], w, t, j_max=Ja)
    dev.inv2(b[:, :, 1], w, t, j_max=Ja)
    if not (a.view(np.uint8) == b.view(np.uint8)).all():
        fails.append("channel-strided inverse: " + describe_mismatch(b[:, :, 1], a[:, :, 1], t))
    report(fails)


@pytest.fixture(scope="module")
def ref_opt_stride():
    # dwt_util_get_opt_stride (src/libdwt.c:20655): next prime >= row bytes on x86-64
    def is_prime(n):
        if n < 2:
            return False
        i = 2
        while i * i <= n:
            if n % i == 0:
                return False
            i += 1
        return True

    def f(n):
        while not is_prime(n):
            n += 1
        return n
    assert f(2048) == 2053
    return f


def test_golden_digests_of_the_reference(dev, oracle):
    """CUDA output against digests the compiled reference produced (tests/golden/make_golden.py)."""
    bad = []
    for c in dense_cases() + sparse_cases():
        if len(c) == 6:
            w, t, ox, oy, j, d1 = c
            ix, iy, zp = ox, oy, 0
        else:
            w, t, ox, oy, ix, iy, j, d1, zp = c
        img = np.zeros((oy, ox), DT[t])
        oracle.fill(img, t)
        J = dev.fwd2(img, w, t, j_max=j, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
        g = GOLD[case_id(c)]
        f = digest(img)
        dev.inv2(img, w, t, j_max=J, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
        if (J, f, digest(img)) != (g["J"], g["fwd"], g["inv"]):
            bad.append(case_id(c))
    assert not bad, f"{len(bad)} cases differ from the reference digests: {bad[:20]}"


@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
def test_device_resident_batch(dev, oracle, kind):
    """frames > 1: one launch per level for the whole batch; frame k is filled with rand = k % 6."""
    w, t = kind
    ox, oy, frames = 300, 260, 7
    img = dev.DeviceImage(dev.kind_of(w, t), ox, oy, frames)
    img.fill(0, 0, 6)
    J = img.fwd2()
    fails = []
    wants = []
    for k in range(frames):
        want = oracle.fill(np.zeros((oy, ox), DT[t]), t, rand=k % 6)
        x0 = want.copy()
        Jo = oracle.fwd2(want, w, t)
        got = img.download(frame=k)
        if J != Jo or not (bits(got, t) == bits(want, t)).all():
            fails.append(f"frame {k} forward: " + describe_mismatch(got, want, t))
        wants.append(x0)
    assert img.last_path == 0 and 1 <= img.last_launches <= J
    img.inv2(J)
    for k in range(frames):
        got = img.download(frame=k)
        ref_rt = wants[k].copy()
        oracle.fwd2(ref_rt, w, t)
        oracle.inv2(ref_rt, w, t, j_max=J)
        if not (bits(got, t) == bits(ref_rt, t)).all():
            fails.append(f"frame {k} inverse: " + describe_mismatch(got, ref_rt, t))
    img.close()
    report(fails)


def test_chained_graph_replays(dev, oracle):
    """The captured, chained pyramid of ONE image object replayed several times with the same key: from the second
    replay on the never-reset completion counters, the (gen + 1) * need targets and the done / gen hand-off of
    chain.cuh are what orders the levels (batch of 3 so that levels 0-2 are ring levels linked into one chain)."""
    w, t = "97", "s"
    ox, oy, frames = 2600, 2300, 3
    img = dev.DeviceImage(dev.kind_of(w, t), ox, oy, frames)
    want_f, want_i = [], []
    for k in range(frames):
        a = oracle.fill(np.zeros((oy, ox), DT[t]), t, rand=k % 6)
        J = oracle.fwd2(a, w, t)
        want_f.append(a.copy())
        oracle.inv2(a, w, t, j_max=J)
        want_i.append(a)
    fails = []
    for rep in range(4):
        img.fill(0, 0, 6)
        assert img.fwd2() == J
        for k in range(frames):
            got = img.download(frame=k)
            if not (bits(got, t) == bits(want_f[k], t)).all():
                fails.append(f"replay {rep} frame {k} forward: " + describe_mismatch(got, want_f[k], t))
        img.inv2(J)
        for k in range(frames):
            got = img.download(frame=k)
            if not (bits(got, t) == bits(want_i[k], t)).all():
                fails.append(f"replay {rep} frame {k} inverse: " + describe_mismatch(got, want_i[k], t))
    img.close()
    report(fails)


# ---- BASELINE.json configs at full size ---------------------------------------------------------------
FULL = [("53", "i", 4096, 4096), ("97", "s", 8192, 8192), ("97", "s", 7919, 6007), ("97", "d", 4096, 4096),
        ("97", "d", 8192, 8192), ("97", "s", 2048, 2048)]


@pytest.mark.parametrize("cfg", FULL, ids=lambda c: f"{c[0]}{c[1]}-{c[2]}x{c[3]}")
def test_full_size_configs(dev, oracle, cfg):
    w, t, ox, oy = cfg
    a = np.zeros((oy, ox), DT[t])
    oracle.fill(a, t)
    x0 = a.copy()
    img = dev.DeviceImage(dev.kind_of(w, t), ox, oy)
    img.fill(0, 0)
    got = img.download()
    assert (bits(got, t) == bits(a, t)).all(), "pattern: " + describe_mismatch(got, a, t)
    Ja = oracle.fwd2(a, w, t)
    Jb = img.fwd2()
    assert Ja == Jb == dev.lib().c.dwtb200_ceil_log2(min(ox, oy))
    got = img.download()
    assert (bits(got, t) == bits(a, t)).all(), "forward: " + describe_mismatch(got, a, t)
    oracle.inv2(a, w, t, j_max=Ja)
    img.inv2(Jb)
    got = img.download()
    assert (bits(got, t) == bits(a, t)).all(), "inverse: " + describe_mismatch(got, a, t)
    err = np.abs(got.astype(np.float64) - x0.astype(np.float64)).max()
    # reconstruction error of forward-then-inverse (the reference's own self-test bound)
    assert err <= {"i": 0, "s": 1e-3, "d": 1e-6}[t], err
    img.close()


def test_properties_at_full_size(dev):
    """Size-independent checks on the device, no oracle: int round trip is exact, float round trip is
    within the reference's eps, and the transform of a frame does not depend on its batch position."""
    img = dev.DeviceImage(dev.CDF53_I32, 4096, 4096, 2)
    keep = dev.DeviceImage(dev.CDF53_I32, 4096, 4096, 2)
    img.fill(0, 2, 0)
    keep.copy_from(img)
    J = img.fwd2()
    assert img.diff(keep) > 0
    img.inv2(J)
    assert img.diff(keep) == 0
    a, b = img.download(frame=0), img.download(frame=1)
    assert (a == b).all()
    img.close(); keep.close()
    f = dev.DeviceImage(dev.CDF97_F32, 8192, 8192)
    k = dev.DeviceImage(dev.CDF97_F32, 8192, 8192)
    f.fill(0, 0, 0)
    k.copy_from(f)
    J = f.fwd2()
    assert J == 13
    f.inv2(J)
    assert f.maxabs(k) < 1e-3
    f.close(); k.close()


def test_concurrent_host_calls_are_serialised(dev, oracle):
    """The reference is not re-entrant; this library keeps global state too, but every entry point takes one process-wide lock
    (include/dwtb200.h): host threads calling the reference-named transforms at the same time must all get the right bits."""
    import threading
    cases = [("97", "s", 517, 301), ("53", "i", 300, 200), ("97", "s", 640, 480), ("97", "d", 129, 127)]
    want, imgs = [], []
    for (w, t, ox, oy) in cases:
        a = oracle.fill(np.zeros((oy, ox), DT[t]), t)
        imgs.append(a.copy())
        oracle.fwd2(a, w, t)
        want.append(a)
    errs = []

    def work(i):
        w, t, ox, oy = cases[i]
        try:
            for _ in range(6):
                b = imgs[i].copy()
                J = dev.fwd2(b, w, t)
                if not (bits(b, t) == bits(want[i], t)).all():
                    errs.append(f"thread {i}: forward differs")
                dev.inv2(b, w, t, j_max=J)
        except Exception as e:   # noqa: BLE001
            errs.append(f"thread {i}: {e}")
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs


# ---- 3-D ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(16, 16, 16), (33, 20, 9), (64, 48, 40), (5, 5, 5), (100, 37, 21), (520, 300, 70), (256, 257, 131)],
                         ids=lambda s: "x".join(map(str, s)))
def test_volume_parity(dev, oracle, shape):
    nx, ny, nz = shape
    a = oracle.volume_fill(np.zeros((nz, ny, nx), np.float32))
    want = np.zeros_like(a)
    oracle.fwd3(a, want)
    got = np.zeros_like(a)
    dev.fwd3(a, got)
    assert (got.view(np.uint32) == want.view(np.uint32)).all(), f"{(got != want).sum()} voxels differ (forward)"
    key = f"vol-{nx}-{ny}-{nz}"
    if key in GOLD:
        assert digest(got) == GOLD[key]["fwd"]
    oracle.inv3(want)
    dev.inv3(got)
    assert (got.view(np.uint32) == want.view(np.uint32)).all(), f"{(got != want).sum()} voxels differ (inverse)"
    assert np.abs(got - a).max() < 1e-3
    v = dev.DeviceVolume(nx, ny, nz)
    v.fill()
    assert (v.download().view(np.uint32) == a.view(np.uint32)).all()
    v.close()


@pytest.mark.parametrize("shape", [(128, 32, 16), (129, 33, 17), (300, 100, 31), (520, 300, 70), (1000, 64, 40), (131, 200, 130), (64, 32, 16), (67, 36, 19),
                                   (193, 97, 50)],
                         ids=lambda s: "x".join(map(str, s)))
def test_volume_single_pass(dev, oracle, shape):
    """the one-pass kernels (x, y and z lifting in ONE pass: tiles of 64 x 32 positions marching along z; DWTB200_TUNE_VOL3 = 1 staged by
    tensor copies, 2 by cp.async) on seeded noise against the oracle, and the two-pass kernels (0) must give the same bits"""
    nx, ny, nz = shape
    rng = np.random.default_rng(nx * 7 + nz)
    a = (rng.standard_normal((nz, ny, nx)) * 10.0 ** rng.integers(-2, 3, size=(nz, ny, nx))).astype(np.float32)
    want = np.zeros_like(a)
    oracle.fwd3(a, want)
    L = dev.lib()
    for vol3 in (1, 2, 0):
        L.check(L.c.dwtb200_set_tuning(9, vol3))
        try:
            got = np.zeros_like(a)
            dev.fwd3(a, got)
        finally:
            L.check(L.c.dwtb200_set_tuning(9, 1))
        bad = got.view(np.uint32) != want.view(np.uint32)
        assert not bad.any(), f"vol3={vol3}: {bad.sum()} voxels differ, first at {np.argwhere(bad)[0]}"
    # inverse of seeded noise (any coefficients will do): one pass, two passes, oracle
    c = (rng.standard_normal((nz, ny, nx)) * 10.0 ** rng.integers(-2, 3, size=(nz, ny, nx))).astype(np.float32)
    want = c.copy()
    oracle.inv3(want)
    for vol3 in (1, 2, 0):
        L.check(L.c.dwtb200_set_tuning(9, vol3))
        try:
            got = c.copy()
            dev.inv3(got)
        finally:
            L.check(L.c.dwtb200_set_tuning(9, 1))
        bad = got.view(np.uint32) != want.view(np.uint32)
        assert not bad.any(), f"inverse vol3={vol3}: {bad.sum()} voxels differ, first at {np.argwhere(bad)[0]}"


# ---- row strips: all ranks emulated on one GPU (the multi-process driver is covered on CPU with gloo) ----
@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
def test_row_strip_partition_on_device(dev, oracle, kind):
    from libdwt_b200 import strips
    w, t = kind
    eng = strips.NumpyEngine(lambda img, j: dev.fwd2(img, w, t, j_max=j), lambda img, j: dev.inv2(img, w, t, j_max=j))
    for (W, H, G, Jd) in ((1024, 2048, 4, 3), (1000, 1500, 2, 2)):
        img = oracle.fill(np.zeros((H, W), DT[t]), t)
        want = img.copy()
        J = oracle.fwd2(want, w, t)
        got, J2 = strips.forward_strips_local(img, G, Jd, eng)
        assert J == J2 and (bits(got, t) == bits(want, t)).all(), describe_mismatch(got, want, t)
        back = strips.inverse_strips_local(want, G, Jd, eng, J)
        oracle.inv2(want, w, t, j_max=J)
        assert (bits(back, t) == bits(want, t)).all(), describe_mismatch(back, want, t)


def test_wide_pattern_and_row_offset(dev, oracle):
    """dwtb200_image_fill_ex: a strip of the 64-bit pattern equals the same rows of the oracle's 64-bit pattern."""
    for t, k in (("s", dev.CDF97_F32), ("i", dev.CDF53_I32)):
        full = oracle.fill(np.zeros((300, 200), DT[t]), t, wrap32=0)
        img = dev.DeviceImage(k, 200, 100)
        img.fill(0, 0, 0, y_offset=150, wide=1)
        got = img.download()
        assert (bits(got, t) == bits(full[150:250], t)).all()
        img.close()


@pytest.mark.parametrize("shape", [(520, 300, 130), (256, 257, 301), (1000, 70, 257)], ids=lambda s: "x".join(map(str, s)))
def test_volume_pipelined_host_path(dev, oracle, shape):
    """volumes of 64 MiB and more through the host calls: z ranges uploaded, transformed and downloaded in a pipeline
    (host3_pipelined, dwtb200.cu); same bits as the oracle and as the plain upload -> transform -> download path"""
    nx, ny, nz = shape
    rng = np.random.default_rng(nx + 3 * nz)
    a = (rng.standard_normal((nz, ny, nx)) * 10.0 ** rng.integers(-2, 3, size=(nz, ny, nx))).astype(np.float32)
    want = np.zeros_like(a)
    oracle.fwd3(a, want)
    back = want.copy()
    oracle.inv3(back)
    L = dev.lib()
    for pipeline in (1, 0):
        L.check(L.c.dwtb200_set_tuning(5, pipeline))
        try:
            got = np.zeros_like(a)
            dev.fwd3(a, got)
            bad = got.view(np.uint32) != want.view(np.uint32)
            assert not bad.any(), f"pipeline={pipeline}: {bad.sum()} voxels differ (forward), first at {np.argwhere(bad)[0]}"
            dev.inv3(got)
            bad = got.view(np.uint32) != back.view(np.uint32)
            assert not bad.any(), f"pipeline={pipeline}: {bad.sum()} voxels differ (inverse), first at {np.argwhere(bad)[0]}"
        finally:
            L.check(L.c.dwtb200_set_tuning(5, 1))


# ---- the pipelined host path (large dense images): upload / level-0 strips / download overlapped, in place ----
@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
def test_pipelined_host_path(dev, oracle, kind):
    w, t = kind
    L = dev.lib()
    es = np.dtype(DT[t]).itemsize
    shapes = [(4096, 4096), (3001, 2999), (7919, 6007), (4100, 2050)] if t != "d" else [(3001, 2999), (2200, 2100)]
    fails = []
    for (ox, oy) in shapes:
        for j in (-1, 2, 3):   # J >= 3: level 1 is pipelined with level 0
            fails += both(dev, oracle, w, t, ox, oy, j, 0)
            assert L.c.dwtb200_last_transform_ms() > 0
    fails += both(dev, oracle, w, t, 3001, 2999, -1, 0, row_bytes=3001 * es + 13)   # unaligned row stride
    L.check(L.c.dwtb200_set_tuning(5, 0))   # and the plain upload -> transform -> download path at the same size
    try:
        fails += ["pipeline off: " + f for f in both(dev, oracle, w, t, 3001, 2999, -1, 0)]
    finally:
        L.check(L.c.dwtb200_set_tuning(5, 1))
    report(fails)


# ---- out-of-place transforms and the perf harness -----------------------------
def test_out_of_place_transforms(dev, oracle):
    from test_oracle import S2_CASES
    fails = []
    for (ox, oy, ix, iy, j, d1, zp) in S2_CASES + [(3001, 2999, 3001, 2999, -1, 0, 0)]:
        src = oracle.fill(np.zeros((oy, ox), np.float32), "s")
        da = oracle.fill(np.zeros((oy, ox), np.float32), "s", rand=2) + 3
        db = da.copy()
        Ja = oracle.fwd2_s2(src, da, j_max=j, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
        jj = [j]
        dev.dwt_cdf97_2f_s2(src, db, db.strides[0], db.strides[1], ox, oy, ix, iy, jj, d1, zp)
        if jj[0] != Ja or not (bits(da, "s") == bits(db, "s")).all():
            fails.append(f"s2 forward {(ox, oy, ix, iy, j, d1, zp)}: " + describe_mismatch(db, da, "s"))
            db[...] = da
        ea = oracle.fill(np.zeros((oy, ox), np.float32), "s", rand=1) - 2
        eb = ea.copy()
        oracle.inv2_s2(da, ea, j_max=Ja, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
        dev.dwt_cdf97_2i_s2(db, eb, eb.strides[0], eb.strides[1], ox, oy, ix, iy, Ja, d1, zp)
        if not (bits(ea, "s") == bits(eb, "s")).all():
            fails.append(f"s2 inverse {(ox, oy, ix, iy, j, d1, zp)}: " + describe_mismatch(eb, ea, "s"))
    report(fails)


def test_perf_harness_on_device(dev):
    """dwt_util_perf_cdf97_2_s / dwt_util_perf_cdf53_2_i re-pointed at the device: plausible device times."""
    for kind in (dev.CDF97_F32, dev.CDF53_I32):
        f, i = dev.perf2(kind, 1920, 1080, j_max=1, M=4, N=4)
        assert 0 < f < 5e-3 and 0 < i < 5e-3, (f, i)
        f13, i13 = dev.perf2(kind, 2048, 2048, j_max=-1, M=2, N=3)
        assert 0 < f13 < 5e-3 and 0 < i13 < 5e-3
    s, bad = dev.perf3(256, N=2)   # volume_perftest_fwd97op_s: seconds per voxel, failed round trips
    assert bad == 0 and 0 < s * 256 ** 3 < 5e-3, (s, bad)


def test_device_resident_subband_views_and_moments(dev, oracle):
    """dwt_util_subband on the device: subband rectangles of the Mallat plane and their moments,
    against numpy on the oracle's coefficients (double accumulation: tolerance 1e-12 relative)."""
    for (w, t, ox, oy) in (("97", "s", 517, 301), ("53", "i", 300, 200), ("97", "d", 130, 257)):
        a = oracle.fill(np.zeros((oy, ox), DT[t]), t)
        J = oracle.fwd2(a, w, t)
        img = dev.DeviceImage(dev.kind_of(w, t), ox, oy, 2)
        img.fill(0, 0, 0)
        assert img.fwd2() == J
        base, pitch, _ = img.devptr()
        for j in (1, 2, J):
            lx, ly, hx, hy, cx, cy = ox, oy, 0, 0, ox, oy
            for _ in range(j):
                hx, hy, lx, ly, cx, cy = lx // 2, ly // 2, (lx + 1) // 2, (ly + 1) // 2, (cx + 1) // 2, (cy + 1) // 2
            for band, (r0, c0, sy, sx) in enumerate(((0, 0, ly, lx), (0, cx, ly, hx), (cy, 0, hy, lx), (cy, cx, hy, hx))):
                p, pb, gx, gy = img.subband(j, band, frame=1)
                frame_bytes = img.devptr()[2]
                assert (gx, gy) == (sx, sy) and pb == pitch
                assert p == base + frame_bytes + r0 * pitch + c0 * a.itemsize
                s, q, m = img.subband_moments(j, band, frame=1)
                ref = a[r0:r0 + sy, c0:c0 + sx].astype(np.float64)
                want = (ref.sum(), (ref * ref).sum(), np.abs(ref).max() if ref.size else 0.0)
                for got, exp in zip((s, q, m), want):
                    assert abs(got - exp) <= 1e-12 * max(1.0, abs(exp)) * max(1, ref.size) ** 0.5, (w, t, j, band, got, exp)
        img.close()


def test_device_resident_feature_vectors(dev, oracle):
    """dwt_util_wps_s / _mean_s / _var_s / _stdev_s / _maxnorm_s / _norm_s (src/libdwt.c:23201-23786) on the device-resident Mallat
    plane against the same formulas in double on the oracle's coefficients (1e-6 relative; the reference itself sums sequentially in
    float and is compared where oracle/_ref travels: 2e-3 relative on 517 x 301, sizes it can still sum accurately)"""
    import ctypes as C
    from oracle.orc import Ref
    ox, oy = 517, 301
    a = oracle.fill(np.zeros((oy, ox), np.float32), "s")
    J = oracle.fwd2(a, "97", "s")
    img = dev.DeviceImage(dev.kind_of("97", "s"), ox, oy)
    img.fill(0, 0, 0)
    assert img.fwd2() == J
    ref = Ref() if Ref.available() else None
    names = {"wps": "dwt_util_wps_s", "mean": "dwt_util_mean_s", "var": "dwt_util_var_s", "stdev": "dwt_util_stdev_s",
             "maxnorm": "dwt_util_maxnorm_s", "norm": "dwt_util_norm_s"}
    for feat, cname in names.items():
        got = img.features(J, feat)
        want = []
        for j in range(1, J):
            lx, ly, hx, hy, cx, cy = ox, oy, 0, 0, ox, oy
            for _ in range(j):
                hx, hy, lx, ly, cx, cy = lx // 2, ly // 2, (lx + 1) // 2, (ly + 1) // 2, (cx + 1) // 2, (cy + 1) // 2
            for (r0, c0, sy, sx) in ((0, cx, ly, hx), (cy, 0, hy, lx), (cy, cx, hy, hx)):
                if not sx or not sy:
                    continue
                b = a[r0:r0 + sy, c0:c0 + sx].astype(np.float64)
                want.append({"wps": (b * b).sum() / 2 ** j, "mean": b.mean(), "var": b.var(), "stdev": b.std(), "maxnorm": np.abs(b).max(),
                             "norm": np.sqrt((b * b).sum())}[feat])
        want = np.array(want)
        assert got.shape == want.shape, (feat, got.shape, want.shape)
        assert np.allclose(got, want, rtol=1e-5, atol=1e-9), (feat, np.abs(got - want).max())
        if ref is not None:
            f = getattr(ref.lib, cname)
            f.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.POINTER(C.c_float)]
            f.restype = None
            fv = (C.c_float * (3 * J))()
            f(a.ctypes.data, a.strides[0], a.strides[1], ox, oy, ox, oy, J, fv)
            rv = np.array(fv[:len(got)], dtype=np.float64)
            assert np.allclose(got, rv, rtol=2e-3, atol=1e-6), (feat, "vs reference", np.abs(got - rv).max())
    img.close()


# ---- seeded random inputs (the reference's patterns are smooth: these are not) -------------------------------------
@pytest.mark.parametrize("kind", KINDS, ids=KIDS)
def test_random_inputs(dev, oracle, kind):
    w, t = kind
    rng = np.random.default_rng(20261018)
    fails = []
    for (ox, oy) in ((64, 64), (517, 301), (1300, 1260), (2100, 1300), (4099, 2051)):
        if t == "i":
            a = rng.integers(-(1 << 20), 1 << 20, size=(oy, ox), dtype=np.int32)
        else:
            a = (rng.standard_normal((oy, ox)) * 10.0 ** rng.integers(-3, 4, size=(oy, ox))).astype(DT[t])
        b = a.copy()
        for (j, d1) in ((-1, 0), (3, 0)):
            a2, b2 = a.copy(), b.copy()
            Ja = oracle.fwd2(a2, w, t, j_max=j, decompose_one=d1)
            Jb = dev.fwd2(b2, w, t, j_max=j, decompose_one=d1)
            if Ja != Jb or not (bits(a2, t) == bits(b2, t)).all():
                fails.append(f"random {w}{t} {ox}x{oy} j={j}: FORWARD " + describe_mismatch(b2, a2, t))
                b2[...] = a2
            oracle.inv2(a2, w, t, j_max=Ja, decompose_one=d1)
            dev.inv2(b2, w, t, j_max=Ja, decompose_one=d1)
            if not (bits(a2, t) == bits(b2, t)).all():
                fails.append(f"random {w}{t} {ox}x{oy} j={j}: INVERSE " + describe_mismatch(b2, a2, t))
    report(fails)


@pytest.mark.parametrize("kind", KINDS[:3], ids=KIDS[:3])
def test_batch_of_ring_sized_frames(dev, oracle, kind):
    """frames > 1 with frames large enough for the bulk-copy ring kernels on several levels."""
    w, t = kind
    ox, oy, frames = 2100, 1300, 3
    img = dev.DeviceImage(dev.kind_of(w, t), ox, oy, frames)
    img.fill(0, 0, 6)
    J = img.fwd2()
    fails = []
    for k in range(frames):
        want = oracle.fill(np.zeros((oy, ox), DT[t]), t, rand=k % 6)
        Jo = oracle.fwd2(want, w, t)
        got = img.download(frame=k)
        if J != Jo or not (bits(got, t) == bits(want, t)).all():
            fails.append(f"frame {k} forward: " + describe_mismatch(got, want, t))
    img.inv2(J)
    for k in range(frames):
        want = oracle.fill(np.zeros((oy, ox), DT[t]), t, rand=k % 6)
        oracle.fwd2(want, w, t)
        oracle.inv2(want, w, t, j_max=J)
        got = img.download(frame=k)
        if not (bits(got, t) == bits(want, t)).all():
            fails.append(f"frame {k} inverse: " + describe_mismatch(got, want, t))
    img.close()
    report(fails)
