"""GPU (B200): the interleaved in-place family -- dwt_cdf97_2f_inplace_s (+ _sep_s / _sdl_s / _sep_sdl_s), dwt_cdf97_2i_inplace_s,
dwt_cdf53_2f_inplace_s, dwt_cdf53_2i_inplace_s -- through the C ABI against the oracle's
restatement of the reference's sweep order (oracle/dwt_oracle.c, pinned to the compiled reference by tests/test_oracle.py)
and against the committed golden digests.  Bit-exact: tolerance 0 ulp."""
import json
import os

import numpy as np
import pytest

from cases import DEPTHS, INPLACE_SHAPES, bits, describe_mismatch, digest, inplace_cases, inplace_id

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def both(dev, oracle, w, a, j, d1, inner=None, tag=""):
    """forward + inverse of the same input on the device path and on the oracle; returns failure strings"""
    b = a.copy()
    fails = []
    tag = f"{w} {a.shape[1]}x{a.shape[0]} j={j} d1={d1} inner={inner} {tag}"
    Ja = oracle.fwd2_inplace(a, w, j_max=j, decompose_one=d1, inner=inner)
    Jb = dev.fwd2_inplace(b, w, j_max=j, decompose_one=d1, inner=inner)
    if Ja != Jb:
        fails.append(f"{tag}: J {Jb} != {Ja}")
    if not (bits(a, "s") == bits(b, "s")).all():
        fails.append(f"{tag}: FORWARD " + describe_mismatch(b, a, "s"))
        b[...] = a
    oracle.inv2_inplace(a, w, j_max=Ja, decompose_one=d1, inner=inner)
    dev.inv2_inplace(b, w, j_max=Ja, decompose_one=d1, inner=inner)
    if not (bits(a, "s") == bits(b, "s")).all():
        fails.append(f"{tag}: INVERSE " + describe_mismatch(b, a, "s"))
    return fails


def report(fails):
    assert not fails, f"{len(fails)} failures:\n" + "\n".join(fails[:12])


@pytest.mark.parametrize("wavelet", ["97", "53"])
@pytest.mark.parametrize("shape", INPLACE_SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_inplace_parity_and_golden(dev, oracle, wavelet, shape):
    ox, oy = shape
    fails = []
    for c in inplace_cases():
        if c[0] != wavelet or (c[1], c[2]) != (ox, oy) or (c[3], c[4]) != (ox, oy):
            continue
        _, _, _, _, _, j, d1 = c
        a = oracle.fill(np.zeros((oy, ox), np.float32), "s", rand=0, type_=0)
        fails += both(dev, oracle, wavelet, a.copy(), j, d1)
        J = dev.fwd2_inplace(a, wavelet, j_max=j, decompose_one=d1)
        g = GOLD[inplace_id(c)]
        if (J, digest(a)) != (g["J"], g["fwd"]):
            fails.append(f"{inplace_id(c)}: forward differs from the reference's digest")
        dev.inv2_inplace(a, wavelet, j_max=J, decompose_one=d1)
        if digest(a) != g["inv"]:
            fails.append(f"{inplace_id(c)}: round trip differs from the reference's digest")
    report(fails)


@pytest.mark.parametrize("wavelet", ["97", "53"])
def test_inplace_inner_smaller_than_outer(dev, oracle, wavelet):
    """level sizes come from the inner size, the level count from the outer one; samples outside the inner region are not touched"""
    fails = []
    for c in inplace_cases():
        w, ox, oy, ix, iy, j, d1 = c
        if w != wavelet or (ix, iy) == (ox, oy):
            continue
        a = oracle.fill(np.zeros((oy, ox), np.float32), "s", rand=0, type_=0)
        fails += both(dev, oracle, wavelet, a.copy(), j, d1, inner=(iy, ix))
        J = dev.fwd2_inplace(a, wavelet, j_max=j, decompose_one=d1, inner=(iy, ix))
        if (J, digest(a)) != (GOLD[inplace_id(c)]["J"], GOLD[inplace_id(c)]["fwd"]):
            fails.append(f"{inplace_id(c)}: forward differs from the reference's digest")
    report(fails)


@pytest.mark.parametrize("wavelet", ["97", "53"])
def test_inplace_random_inputs_every_small_shape(dev, oracle, wavelet):
    """all exception / prolog / epilog combinations (lines of 1 .. 12 samples) and the frame of larger levels, on seeded noise
    with a wide dynamic range: the order of the row and column sweeps is visible in the last bit there"""
    rng = np.random.default_rng(20261018)
    fails = []
    shapes = [(h, w) for h in range(1, 13) for w in range(1, 13)] + [(37, 53), (64, 64), (65, 63), (9, 200), (200, 9), (130, 3), (4, 77), (100, 1),
                                                                    (255, 257), (300, 517), (1023, 1025), (2100, 1300)]
    for (oy, ox) in shapes:
        for (j, d1) in ((1, 0), (-1, 0), (-1, 1)):
            if max(ox, oy) > 600 and (j, d1) == (-1, 1):
                continue
            a = (rng.standard_normal((oy, ox)) * 10.0 ** rng.integers(-2, 3, size=(oy, ox))).astype(np.float32)
            fails += both(dev, oracle, wavelet, a, j, d1, tag="random")
    report(fails)


@pytest.mark.parametrize("wavelet", ["97", "53"])
def test_inplace_strided_host_layouts(dev, oracle, wavelet):
    """prime row stride (dwt_util_get_opt_stride) and a channel-interleaved layout (stride_y = 12 bytes, cv::Mat style)"""
    from oracle.orc import strided_image
    rng = np.random.default_rng(7)
    fails = []
    for (ox, oy, row_bytes) in ((512, 512, 2053), (301, 200, 1213)):
        a = strided_image((oy, ox), "s", row_bytes)
        a[...] = rng.standard_normal((oy, ox)).astype(np.float32)
        b = strided_image((oy, ox), "s", row_bytes)
        b[...] = a
        Ja = oracle.fwd2_inplace(a, wavelet)
        Jb = dev.fwd2_inplace(b, wavelet)
        if Ja != Jb or not (bits(a, "s") == bits(b, "s")).all():
            fails.append(f"{wavelet} {ox}x{oy} row_bytes={row_bytes}: " + describe_mismatch(b, a, "s"))
    base = rng.standard_normal((120, 90, 3)).astype(np.float32)
    for ch in range(3):
        a, b = base.copy(), base.copy()
        oracle.fwd2_inplace(a[:, :, ch], wavelet)
        dev.fwd2_inplace(b[:, :, ch], wavelet)
        if not (bits(a, "s") == bits(b, "s")).all():
            fails.append(f"{wavelet} channel {ch} of an interleaved image differs")
    report(fails)


@pytest.mark.parametrize("wavelet", ["97", "53"])
def test_inplace_device_resident_batch(dev, oracle, wavelet):
    """frames > 1 on a device-resident image: one set of launches for the whole batch, captured in a CUDA graph and replayed"""
    ox, oy, frames = 300, 260, 5
    img = dev.DeviceImage(dev.kind_of(wavelet, "s"), ox, oy, frames)
    fails = []
    for rep in range(2):   # the second pass replays the captured graphs
        img.fill(0, 0, 6)
        J = img.fwd2_inplace()
        assert img.last_launches >= 1
        for k in range(frames):
            want = oracle.fill(np.zeros((oy, ox), np.float32), "s", rand=k % 6)
            Jo = oracle.fwd2_inplace(want, wavelet)
            got = img.download(frame=k)
            if J != Jo or not (bits(got, "s") == bits(want, "s")).all():
                fails.append(f"rep {rep} frame {k} forward: " + describe_mismatch(got, want, "s"))
        img.inv2_inplace(J)
        for k in range(frames):
            want = oracle.fill(np.zeros((oy, ox), np.float32), "s", rand=k % 6)
            oracle.fwd2_inplace(want, wavelet)
            oracle.inv2_inplace(want, wavelet, j_max=J)
            got = img.download(frame=k)
            if not (bits(got, "s") == bits(want, "s")).all():
                fails.append(f"rep {rep} frame {k} inverse: " + describe_mismatch(got, want, "s"))
    img.close()
    report(fails)


@pytest.mark.parametrize("wavelet", ["97", "53"])
def test_inplace_interleaved_level0_path(dev, oracle, wavelet):
    """images whose level 0 runs on the bulk-copy ring kernels: those read / write the interleaved layout directly
    (LevelParams::il), only the even rows and columns are translated; odd sizes, one / two / all levels, a batch, and the same
    shapes with that path switched off (DWTB200_TUNE_RING = 0: register kernels + full translation) must give the same bits"""
    rng = np.random.default_rng(99)
    fails = []
    L = dev.lib()
    # ring = 3: default CTA shape; 0: register kernels, whole plane translated; bits 4-6 force a shape: 15 consumer warps x 1 CTA per SM
    # (interleaved level 0 with a 217 KB five-segment inverse ring), 8 x 2 (no room for the fifth segment: falls back to full translation)
    for ring in (3, 0, 3 | (1 << 4), 3 | (2 << 4)):
        L.check(L.c.dwtb200_set_tuning(6, ring))
        try:
            for (oy, ox) in ((1301, 2101), (2047, 1031), (1500, 1501), (1056, 1000), (40, 30000)) if ring in (3, 0) else ((1301, 2101), (1056, 1000)):
                for j in (1, 2, -1):
                    a = (rng.standard_normal((oy, ox)) * 10.0 ** rng.integers(-2, 3, size=(oy, ox))).astype(np.float32)
                    fails += both(dev, oracle, wavelet, a, j, 0, tag=f"ring={ring}")
        finally:
            L.check(L.c.dwtb200_set_tuning(6, 3))
    ox, oy, frames = 1100, 1000, 3
    img = dev.DeviceImage(dev.kind_of(wavelet, "s"), ox, oy, frames)
    for rep in range(2):
        img.fill(0, 0, 6)
        J = img.fwd2_inplace()
        for k in range(frames):
            want = oracle.fill(np.zeros((oy, ox), np.float32), "s", rand=k % 6)
            oracle.fwd2_inplace(want, wavelet)
            got = img.download(frame=k)
            if not (bits(got, "s") == bits(want, "s")).all():
                fails.append(f"batch rep {rep} frame {k} forward: " + describe_mismatch(got, want, "s"))
        img.inv2_inplace(J)
        for k in range(frames):
            want = oracle.fill(np.zeros((oy, ox), np.float32), "s", rand=k % 6)
            oracle.fwd2_inplace(want, wavelet)
            oracle.inv2_inplace(want, wavelet, j_max=J)
            got = img.download(frame=k)
            if not (bits(got, "s") == bits(want, "s")).all():
                fails.append(f"batch rep {rep} frame {k} inverse: " + describe_mismatch(got, want, "s"))
    img.close()
    report(fails)


def test_inplace_compat_symbols_and_variants(dev, oracle):
    """the reference's own names in libdwt_compat.so: the four forward 9/7 variants are one transform"""
    import ctypes as C
    so = os.path.join(os.path.dirname(HERE), "libdwt_b200", "libdwt_compat.so")
    L = C.CDLL(so)
    ci, vp = C.c_int, C.c_void_p
    rng = np.random.default_rng(3)
    src = rng.standard_normal((97, 131)).astype(np.float32)
    want = src.copy()
    J = oracle.fwd2_inplace(want, "97")
    for name in ("dwt_cdf97_2f_inplace_s", "dwt_cdf97_2f_inplace_sep_s", "dwt_cdf97_2f_inplace_sdl_s", "dwt_cdf97_2f_inplace_sep_sdl_s"):
        f = getattr(L, name)
        f.argtypes = [vp, ci, ci, ci, ci, ci, ci, C.POINTER(ci), ci, ci]
        f.restype = None
        a = src.copy()
        j = ci(-1)
        f(a.ctypes.data, a.strides[0], a.strides[1], 131, 97, 131, 97, C.byref(j), 0, 0)
        assert j.value == J and (bits(a, "s") == bits(want, "s")).all(), name
    L.dwt_cdf97_2i_inplace_s.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, ci, ci]
    L.dwt_cdf97_2i_inplace_s.restype = None
    a = want.copy()
    L.dwt_cdf97_2i_inplace_s(a.ctypes.data, a.strides[0], a.strides[1], 131, 97, 131, 97, J, 0, 0)
    oracle.inv2_inplace(want, "97", j_max=J)
    assert (bits(a, "s") == bits(want, "s")).all()
    assert np.abs(a - src).max() < 1e-4
    for name, inv in (("dwt_cdf53_2f_inplace_s", "dwt_cdf53_2i_inplace_s"),):
        f, g = getattr(L, name), getattr(L, inv)
        f.argtypes = [vp, ci, ci, ci, ci, ci, ci, C.POINTER(ci), ci, ci]
        g.argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, ci, ci]
        f.restype = g.restype = None
        a, w = src.copy(), src.copy()
        j = ci(-1)
        f(a.ctypes.data, a.strides[0], a.strides[1], 131, 97, 131, 97, C.byref(j), 0, 0)
        Jo = oracle.fwd2_inplace(w, "53")
        assert j.value == Jo and (bits(a, "s") == bits(w, "s")).all()
        g(a.ctypes.data, a.strides[0], a.strides[1], 131, 97, 131, 97, j.value, 0, 0)
        oracle.inv2_inplace(w, "53", j_max=Jo)
        assert (bits(a, "s") == bits(w, "s")).all()


def test_inplace_full_size(dev, oracle):
    """8192 x 8192, J = 13 (BASELINE.json configs[1] shape): bit-exact against the oracle at 2048^2 random, and at full size the
    forward / inverse round trip within 1e-4 and the interior equal to the Mallat transform's coefficients (the sweep order
    only differs in the frame of each level)."""
    rng = np.random.default_rng(11)
    a = rng.standard_normal((2048, 2048)).astype(np.float32)
    report(both(dev, oracle, "97", a, -1, 0, tag="2048^2 random"))
    ox = oy = 8192
    img = dev.DeviceImage(dev.kind_of("97", "s"), ox, oy)
    img.fill(0, 0)
    x0 = img.download()
    J = img.fwd2_inplace(j_max=1)
    ip = img.download()
    img.upload(x0)
    img.fwd2(j_max=1)
    ml = img.download()
    inter = np.empty_like(ml)
    h = oy // 2
    inter[0::2, 0::2], inter[0::2, 1::2], inter[1::2, 0::2], inter[1::2, 1::2] = ml[:h, :h], ml[:h, h:], ml[h:, :h], ml[h:, h:]
    d = bits(inter, "s") != bits(ip, "s")
    assert not d[8:, :-8].any(), "in-place and Mallat coefficients must agree outside the top 8 rows / right 8 columns"
    assert d.any(), "... and the frame is expected to differ in the last bit somewhere"
    img.upload(x0)
    J = img.fwd2_inplace()
    assert J == 13
    img.inv2_inplace(J)
    back = img.download()
    assert np.abs(back - x0).max() < 1e-4
    img.close()


def test_inplace_perf_harness_on_device(dev):
    """dwt_util_perf_cdf97_2_inplace_s re-pointed at the device (dwtb200_perf2_inplace): M resident images, N loops, CUDA events"""
    f, i = dev.perf2_inplace(dev.CDF97_F32, 1920, 1080, j_max=-1, M=2, N=3)
    assert 0 < f < 0.01 and 0 < i < 0.01, (f, i)
    f, i = dev.perf2_inplace(dev.CDF53_F32, 1024, 768, j_max=3, M=1, N=2, inner=(700, 1000), decompose_one=1)
    assert 0 < f < 0.01 and 0 < i < 0.01, (f, i)


def test_inplace_beyond_2g_bytes(dev):
    """a 32768 x 20000 float image (2.6 GB per plane: byte offsets beyond 2^31, which the reference cannot address,
    src/inline.h:188): size-independent properties -- the interior of a two-level in-place transform equals the Mallat
    transform's coefficients bit for bit (only the frame of each level rounds differently), and the full-depth round trip
    reconstructs the input"""
    ox, oy = 32768, 20000
    img = dev.DeviceImage(dev.kind_of("97", "s"), ox, oy)
    img.fill(0, 0, 0, 0, 1)
    x0 = img.download()
    J = img.fwd2_inplace(j_max=2)
    assert J == 2
    ip = img.download()
    img.upload(x0)
    img.fwd2(j_max=2)
    ml = img.download()
    # level 1 subbands (odd rows / columns of the interleaved plane) against the Mallat quadrants
    h, w = oy // 2, ox // 2
    assert (bits(ip[1::2, 1::2], "s")[8:, :-8] == bits(ml[h:, w:], "s")[8:, :-8]).all()      # HH_1
    assert (bits(ip[0::2, 1::2], "s")[8:, :-8] == bits(ml[:h, w:], "s")[8:, :-8]).all()      # HL_1
    # level 2 subbands live at stride 4; its frame is 8 rows / 6 columns of the LL_1 grid, widened by what level 1's frame feeds into it
    h2, w2 = oy // 4, ox // 4
    assert (bits(ip[2::4, 2::4], "s")[16:, :-16] == bits(ml[h2:h, w2:w], "s")[16:, :-16]).all()   # HH_2
    assert (bits(ip[0::4, 0::4], "s")[16:, :-16] == bits(ml[:h2, :w2], "s")[16:, :-16]).all()     # LL_2
    del ip, ml
    img.upload(x0)
    J = img.fwd2_inplace()
    img.inv2_inplace(J)
    back = img.download()
    assert float(np.abs(back - x0).max()) < 1e-4
    img.close()


@pytest.mark.parametrize("wavelet", ["97", "53"])
def test_inplace_batch_of_2048_wide_frames(dev, oracle, wavelet):
    """a batch large enough for the 5-warp x 3-CTA ring shape (rows of 2048 samples split into two bands of five column groups):
    the interleaved level-0 kernels of that shape, forward and inverse, frames 0, 11 and 23 against the oracle"""
    ox, oy, frames = 2048, 1100, 24
    img = dev.DeviceImage(dev.kind_of(wavelet, "s"), ox, oy, frames)
    img.fill(0, 0, 6)
    J = img.fwd2_inplace()
    fails = []
    for k in (0, 11, 23):
        want = oracle.fill(np.zeros((oy, ox), np.float32), "s", rand=k % 6)
        oracle.fwd2_inplace(want, wavelet)
        got = img.download(frame=k)
        if not (bits(got, "s") == bits(want, "s")).all():
            fails.append(f"frame {k} forward: " + describe_mismatch(got, want, "s"))
    img.inv2_inplace(J)
    for k in (0, 11, 23):
        want = oracle.fill(np.zeros((oy, ox), np.float32), "s", rand=k % 6)
        oracle.fwd2_inplace(want, wavelet)
        oracle.inv2_inplace(want, wavelet, j_max=J)
        got = img.download(frame=k)
        if not (bits(got, "s") == bits(want, "s")).all():
            fails.append(f"frame {k} inverse: " + describe_mismatch(got, want, "s"))
    img.close()
    report(fails)
