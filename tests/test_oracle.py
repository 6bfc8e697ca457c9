"""CPU: the oracle (oracle/dwt_oracle.c) against the golden vectors produced by the compiled reference
(tests/golden/make_golden.py) and, when oracle/_ref is present, against the reference itself."""
import json
import os

import numpy as np
import pytest

from cases import DT, KINDS, bits, case_id, dense_cases, describe_mismatch, digest, sparse_cases

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))
VEC = np.load(os.path.join(HERE, "golden", "vectors.npz"))


def unpack(c):
    if len(c) == 6:
        w, t, ox, oy, j, d1 = c
        return w, t, ox, oy, ox, oy, j, d1, 0
    return c


def run_impl(impl, c, fill):
    w, t, ox, oy, ix, iy, j, d1, zp = unpack(c)
    img = np.zeros((oy, ox), dtype=DT[t])
    fill.fill(img, t, rand=0, type_=0)
    J = impl.fwd2(img, w, t, j_max=j, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
    fwd = img.copy()
    impl.inv2(img, w, t, j_max=J, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
    return J, fwd, img


def test_oracle_matches_golden_digests(oracle):
    bad = []
    for c in dense_cases() + sparse_cases():
        J, fwd, inv = run_impl(oracle, c, oracle)
        g = GOLD[case_id(c)]
        if (J, digest(fwd), digest(inv)) != (g["J"], g["fwd"], g["inv"]):
            bad.append(case_id(c))
    assert not bad, f"{len(bad)} cases differ from the reference's golden digests: {bad[:10]}"


def test_oracle_matches_golden_vectors(oracle):
    for key in VEC.files:
        if not key.endswith("/in") or key.startswith("vol"):
            continue
        name = key[:-3]
        c = next(cc for cc in dense_cases() + sparse_cases() if case_id(cc) == name)
        w, t, ox, oy, ix, iy, j, d1, zp = unpack(c)
        img = VEC[key].copy()
        J = oracle.fwd2(img, w, t, j_max=j, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
        assert (bits(img, t) == bits(VEC[name + "/fwd"], t)).all(), describe_mismatch(img, VEC[name + "/fwd"], t)
        oracle.inv2(img, w, t, j_max=J, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
        assert (bits(img, t) == bits(VEC[name + "/inv"], t)).all(), describe_mismatch(img, VEC[name + "/inv"], t)


def test_oracle_type2_and_volume_golden(oracle):
    for (w, t) in (("53", "i"), ("97", "s")):
        img = np.zeros((301, 517), dtype=DT[t])
        oracle.fill(img, t, rand=0, type_=2)
        J = oracle.fwd2(img, w, t)
        g = GOLD[f"type2-{w}-{t}-517-301"]
        assert (J, digest(img)) == (g["J"], g["fwd"])
        oracle.inv2(img, w, t, j_max=J)
        assert digest(img) == g["inv"]
    for (nx, ny, nz) in ((16, 16, 16), (33, 20, 9), (64, 48, 40), (5, 5, 5)):
        a = np.zeros((nz, ny, nx), dtype=np.float32)
        oracle.volume_fill(a)
        b = np.zeros_like(a)
        oracle.fwd3(a, b)
        g = GOLD[f"vol-{nx}-{ny}-{nz}"]
        assert digest(b) == g["fwd"]
        oracle.inv3(b)
        assert digest(b) == g["inv"]
    assert (VEC["vol-16-16-16/in"].view(np.uint32) == oracle.volume_fill(np.zeros((16, 16, 16), np.float32)).view(np.uint32)).all()


def test_int_roundtrip_is_exact_and_float_roundtrip_is_close(oracle):
    # the reference's own self-test criterion (src/libdwt.c:1548, 1604, 1513)
    for (w, t, eps) in (("53", "i", 0), ("97", "s", 1e-3), ("97", "d", 1e-6)):
        img = np.zeros((301, 517), dtype=DT[t])
        oracle.fill(img, t)
        x0 = img.copy()
        J = oracle.fwd2(img, w, t)
        oracle.inv2(img, w, t, j_max=J)
        assert np.abs(img.astype(np.float64) - x0.astype(np.float64)).max() <= eps


# ---- against the compiled reference itself (build container only) ---------------------------------
@pytest.mark.parametrize("kind", KINDS, ids=lambda k: k[0] + k[1])
def test_oracle_vs_reference_strided(oracle, ref, kind):
    from oracle.orc import strided_image
    w, t = kind
    es = np.dtype(DT[t]).itemsize
    for (ox, oy) in ((512, 512), (517, 301), (64, 3), (100, 77)):
        for row_bytes in (ox * es, ref.opt_stride(ox * es)):
            for (j, d1) in ((-1, 0), (2, 0), (-1, 1)):
                a = strided_image((oy, ox), t, row_bytes)
                b = strided_image((oy, ox), t, row_bytes)
                ref.fill(a, t)
                oracle.fill(b, t)
                assert (bits(a, t) == bits(b, t)).all(), "test pattern"
                Ja = ref.fwd2(a, w, t, j_max=j, decompose_one=d1)
                Jb = oracle.fwd2(b, w, t, j_max=j, decompose_one=d1)
                assert Ja == Jb
                assert (bits(a, t) == bits(b, t)).all(), describe_mismatch(b, a, t)
                ref.inv2(a, w, t, j_max=Ja, decompose_one=d1)
                oracle.inv2(b, w, t, j_max=Jb, decompose_one=d1)
                assert (bits(a, t) == bits(b, t)).all(), describe_mismatch(b, a, t)


def test_oracle_vs_reference_config_sizes(oracle, ref):
    # BASELINE.json configs the reference can address: 4096^2 int (wrapped pattern), 2048^2 float
    for (w, t, n) in (("53", "i", 4096), ("97", "s", 2048)):
        a = np.zeros((n, n), dtype=DT[t])
        b = np.zeros((n, n), dtype=DT[t])
        ref.fill(a, t)
        oracle.fill(b, t)
        assert (bits(a, t) == bits(b, t)).all()
        Ja, Jb = ref.fwd2(a, w, t), oracle.fwd2(b, w, t)
        assert Ja == Jb and (bits(a, t) == bits(b, t)).all()
        ref.inv2(a, w, t, j_max=Ja)
        oracle.inv2(b, w, t, j_max=Jb)
        assert (bits(a, t) == bits(b, t)).all()


S2_CASES = [(64, 64, 64, 64, -1, 0, 0), (517, 301, 517, 301, -1, 0, 0), (64, 64, 50, 37, -1, 0, 1), (128, 96, 128, 50, 2, 0, 0),
            (100, 80, 33, 80, -1, 1, 0), (1, 9, 1, 9, -1, 1, 0), (9, 1, 9, 1, -1, 1, 0), (33, 70, 17, 5, -1, 0, 1), (16, 16, 16, 16, 0, 0, 0)]


def test_out_of_place_semantics_match_reference(oracle, ref):
    """dwt_cdf97_2f_s2 / 2i_s2: the oracle's restatement (overlay + in place) against the compiled reference, including
    what is left of dst's previous content in sparse layouts."""
    for (ox, oy, ix, iy, j, d1, zp) in S2_CASES:
        src = oracle.fill(np.zeros((oy, ox), np.float32), "s")
        da = oracle.fill(np.zeros((oy, ox), np.float32), "s", rand=2) + 3
        db = da.copy()
        Ja = ref.fwd2_s2(src, da, j_max=j, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
        Jb = oracle.fwd2_s2(src, db, j_max=j, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
        assert Ja == Jb and (bits(da, "s") == bits(db, "s")).all(), (ox, oy, ix, iy, j, d1, zp, describe_mismatch(db, da, "s"))
        ea = oracle.fill(np.zeros((oy, ox), np.float32), "s", rand=1) - 2
        eb = ea.copy()
        ref.inv2_s2(da, ea, j_max=Ja, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
        oracle.inv2_s2(db, eb, j_max=Ja, decompose_one=d1, zero_padding=zp, inner=(iy, ix))
        assert (bits(ea, "s") == bits(eb, "s")).all(), (ox, oy, ix, iy, j, d1, zp, describe_mismatch(eb, ea, "s"))


@pytest.mark.parametrize("kind", KINDS, ids=lambda k: k[0] + k[1])
def test_oracle_vs_reference_random_inputs(oracle, ref, kind):
    """The reference's test patterns are smooth; seeded random data (wide dynamic range for floats) must agree bit for bit too."""
    w, t = kind
    rng = np.random.default_rng(20261018)
    for (ox, oy) in ((64, 64), (517, 301), (1025, 130)):
        if t == "i":
            a = rng.integers(-(1 << 20), 1 << 20, size=(oy, ox), dtype=np.int32)
        else:
            a = (rng.standard_normal((oy, ox)) * 10.0 ** rng.integers(-3, 4, size=(oy, ox))).astype(DT[t])
        b = a.copy()
        Ja, Jb = ref.fwd2(a, w, t), oracle.fwd2(b, w, t)
        assert Ja == Jb and (bits(a, t) == bits(b, t)).all(), describe_mismatch(b, a, t)
        ref.inv2(a, w, t, j_max=Ja)
        oracle.inv2(b, w, t, j_max=Jb)
        assert (bits(a, t) == bits(b, t)).all(), describe_mismatch(b, a, t)


# ---- interleaved in-place family ------------------------------------------------------
def test_inplace_oracle_matches_golden_digests(oracle):
    """dwt_cdf97_2f_inplace_s / 2i_inplace_s and dwt_cdf53_2f_inplace_s / 2i_inplace_s restated against digests of the
    compiled reference's output: the 9/7 pair reproduces the prolog / core / epilog sweep order bit for bit."""
    from cases import inplace_cases, inplace_id
    bad = []
    for c in inplace_cases():
        w, ox, oy, ix, iy, j, d1 = c
        img = np.zeros((oy, ox), dtype=np.float32)
        oracle.fill(img, "s", rand=0, type_=0)
        J = oracle.fwd2_inplace(img, w, j_max=j, decompose_one=d1, inner=(iy, ix))
        fwd = digest(img)
        oracle.inv2_inplace(img, w, j_max=J, decompose_one=d1, inner=(iy, ix))
        g = GOLD[inplace_id(c)]
        if (J, fwd, digest(img)) != (g["J"], g["fwd"], g["inv"]):
            bad.append(inplace_id(c))
    assert not bad, f"{len(bad)} in-place cases differ from the reference's golden digests: {bad[:10]}"
    for name in ("ip-97-31-33-31-33--1-0", "ip-97-64-64-50-37--1-0", "ip-53-31-33-31-33--1-0"):
        w, ox, oy, ix, iy = name.split("-")[1], *(int(v) for v in name.split("-")[2:6])
        img = np.zeros((oy, ox), dtype=np.float32)
        oracle.fill(img, "s", rand=0, type_=0)
        J = oracle.fwd2_inplace(img, w, inner=(iy, ix))
        assert (bits(img, "s") == bits(VEC[name + "/fwd"], "s")).all(), describe_mismatch(img, VEC[name + "/fwd"], "s")
        oracle.inv2_inplace(img, w, j_max=J, inner=(iy, ix))
        assert (bits(img, "s") == bits(VEC[name + "/inv"], "s")).all(), describe_mismatch(img, VEC[name + "/inv"], "s")


@pytest.mark.parametrize("wavelet", ["97", "53"])
def test_inplace_oracle_vs_reference_random_inputs(oracle, ref, wavelet):
    """Every small shape (all the exception / prolog / epilog combinations of src/libdwt.c:12975-13451, 17512-17598) and a few
    larger ones on seeded random data; the four forward 9/7 variants of the reference must agree with each other too."""
    rng = np.random.default_rng(20261018)
    shapes = [(h, w) for h in range(1, 12) for w in range(1, 12)] + [(64, 64), (37, 53), (9, 200), (130, 3), (255, 257), (100, 1), (4, 77)]
    for (oy, ox) in shapes:
        for (j, d1) in ((1, 0), (-1, 0), (-1, 1)):
            a = (rng.standard_normal((oy, ox)) * 10.0 ** rng.integers(-2, 3, size=(oy, ox))).astype(np.float32)
            b = a.copy()
            Ja, Jb = ref.fwd2_inplace(a, wavelet, j, d1), oracle.fwd2_inplace(b, wavelet, j, d1)
            assert Ja == Jb and (bits(a, "s") == bits(b, "s")).all(), (oy, ox, j, d1, describe_mismatch(b, a, "s"))
            if wavelet == "97" and (oy, ox) in ((64, 64), (37, 53), (7, 9), (5, 5)):
                src = (rng.standard_normal((oy, ox))).astype(np.float32)
                outs = []
                for variant in ("", "sep_", "sdl_", "sep_sdl_"):
                    c = src.copy()
                    ref.fwd2_inplace(c, "97", j, d1, variant=variant)
                    outs.append(c)
                assert all((bits(o, "s") == bits(outs[0], "s")).all() for o in outs[1:])
            ref.inv2_inplace(a, wavelet, Ja, d1)
            oracle.inv2_inplace(b, wavelet, Jb, d1)
            assert (bits(a, "s") == bits(b, "s")).all(), (oy, ox, j, d1, describe_mismatch(b, a, "s"))
    for (oy, ox, iy, ix) in ((64, 64, 40, 50), (33, 70, 20, 69), (16, 16, 5, 3)):
        a = rng.standard_normal((oy, ox)).astype(np.float32)
        b = a.copy()
        Ja, Jb = ref.fwd2_inplace(a, wavelet, -1, 1, inner=(iy, ix)), oracle.fwd2_inplace(b, wavelet, -1, 1, inner=(iy, ix))
        assert Ja == Jb and (bits(a, "s") == bits(b, "s")).all()
        ref.inv2_inplace(a, wavelet, Ja, 1, inner=(iy, ix))
        oracle.inv2_inplace(b, wavelet, Jb, 1, inner=(iy, ix))
        assert (bits(a, "s") == bits(b, "s")).all()


def test_example_file_digests_are_those_of_the_compiled_reference():
    """tests/golden/examples_md5.json (what the GPU tests of the reference's subbands / load examples compare with) is what those
    programs write when linked against the compiled reference alone; re-made here when the reference tree is present"""
    import importlib.util
    import shutil
    ROOT = os.path.dirname(HERE)
    if not os.path.isdir("/root/reference/examples") or not shutil.which("gcc") or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libdwt_ref.so")):
        pytest.skip("needs /root/reference, gcc and oracle/_ref/libdwt_ref.so")
    spec = importlib.util.spec_from_file_location("make_examples_md5", os.path.join(ROOT, "tests", "golden", "make_examples_md5.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.digests() == json.load(open(os.path.join(ROOT, "tests", "golden", "examples_md5.json")))
