"""CPU: the C-ABI library loads without a GPU, exports every symbol include/dwtb200.h declares, and
its compute entry points fail loudly (no fallback) when no CUDA device is present."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols(path):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dwtb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes
    import libdwt_b200 as d
    L = d.lib()
    declared = header_symbols(os.path.join(ROOT, "include", "dwtb200.h"))
    assert len(declared) >= 35
    bound = {n for n, _, _ in d.api.SYMBOLS}
    for name in declared:
        assert hasattr(L.c, name), f"{name} declared in include/dwtb200.h but not exported"
        assert name in bound, f"{name} has no ctypes prototype"
    assert isinstance(L.c, ctypes.CDLL)


def test_level_arithmetic_matches_oracle(oracle):
    import libdwt_b200 as d
    L = d.lib()
    for x in list(range(1, 70)) + [127, 128, 129, 4095, 4096, 4097, 65536]:
        assert L.c.dwtb200_ceil_log2(x) == oracle.ceil_log2(x)
    # src/libdwt.c:12807-12810
    assert L.c.dwtb200_clamp_j(-1, 512, 512, 0) == 9
    assert L.c.dwtb200_clamp_j(-1, 7919, 6007, 0) == 13
    assert L.c.dwtb200_clamp_j(99, 64, 3, 0) == 2
    assert L.c.dwtb200_clamp_j(-1, 64, 3, 1) == 6
    assert L.c.dwtb200_clamp_j(3, 64, 64, 0) == 3


def test_compute_fails_loudly_without_gpu():
    import numpy as np
    import libdwt_b200 as d
    L = d.lib()
    if L.c.dwtb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    img = np.zeros((8, 8), np.float32)
    with pytest.raises(d.DwtError):
        d.fwd2(img, "97", "s")
    with pytest.raises(d.DwtError):
        d.DeviceImage(d.CDF97_F32, 8, 8)


def test_compat_library_exports_the_reference_symbols():
    import ctypes
    src = open(os.path.join(ROOT, "include", "libdwt_compat.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b((?:dwt_|cdf97_3|volume_)[a-z0-9_]+)\s*\(", src)))
    assert {"dwt_cdf97_2f_s", "dwt_cdf97_2i_s", "dwt_cdf97_2f_d", "dwt_cdf97_2i_d", "dwt_cdf53_2f_i", "dwt_cdf53_2i_i",
            "dwt_util_alloc_image", "dwt_util_free_image", "cdf97_3f_op_sep_horizontal_s",
            "cdf97_3i_ip_sep_horizontal_s", "volume_alloc_realiably_locked", "volume_free", "volume_measure_fwd97op_s",
            "volume_perftest_fwd97op_s"} <= set(names)
    so = ctypes.CDLL(os.path.join(ROOT, "libdwt_b200", "libdwt_compat.so"))
    for n in names:
        assert hasattr(so, n), n


def test_strips_fail_loudly_without_gpu_and_validate_arguments():
    import ctypes as C
    import libdwt_b200 as d
    L = d.lib()
    p = d.api.StripPlanC()
    assert L.c.dwtb200_strips_plan(0, 100, 2, 1, 4, 0, C.byref(p)) == -3          # DWTB200_EINVAL
    assert L.c.dwtb200_strips_plan(100, 100, 2, 1, 4, 2, C.byref(p)) == -3        # rank out of range
    assert L.c.dwtb200_strips_plan(100, 100, 17, 1, 4, 0, C.byref(p)) == -3       # more ranks than a box has GPUs
    assert L.c.dwtb200_strips_plan(64, 64, 8, 3, 4, 0, C.byref(p)) == 0 and p.neighbours_only == 0   # strips shorter than the halo
    if L.c.dwtb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    assert not L.c.dwtb200_strips_create(d.CDF97_F32, 4096, 4096, 0, 0, 1, b"dwtb200-abi-test")
    assert b"CUDA" in L.c.dwtb200_last_error() or b"device" in L.c.dwtb200_last_error()
