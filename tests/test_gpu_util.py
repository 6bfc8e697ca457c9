"""GPU: the utilities around the transforms that work on device-resident data and the 3-D host
entry points of the compat layer, against the compiled reference's own functions."""
import ctypes as C
import os

import numpy as np
import pytest

from cases import DT

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ulps(a, b):
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


@pytest.mark.parametrize("t", ["s", "i", "d"])
def test_conv_show_on_the_device(dev, ref, t):
    """dwt_util_conv_show_{s,i,d} (src/libdwt.c:21075, 21020, 21120) of a device-resident Mallat plane."""
    w = "53" if t == "i" else "97"
    ox, oy = 517, 301
    img = dev.DeviceImage(dev.kind_of(w, t), ox, oy)
    out = dev.DeviceImage(dev.kind_of(w, t), ox, oy)
    img.fill(0, 0)
    img.fwd2()
    coeffs = img.download()
    img.conv_show(out)
    got = out.download()
    want = np.zeros_like(coeffs)
    fn = getattr(ref.lib, f"dwt_util_conv_show_{t}")
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    fn(coeffs.ctypes.data, want.ctypes.data, coeffs.strides[0], coeffs.strides[1], ox, oy)
    if t == "i":
        assert got.tobytes() == want.tobytes()
    elif t == "s":
        # the logarithm is taken in double and rounded to float on both sides: at most a rare 1-ulp double-rounding difference
        u = ulps(got, want)
        assert u.max() <= 1 and (u != 0).mean() < 1e-3, (u.max(), (u != 0).mean())
    else:
        assert np.allclose(got, want, rtol=0, atol=1e-15)
    img.conv_show()   # in place
    assert img.download().tobytes() == got.tobytes()
    img.close()
    out.close()


def test_pgm_of_a_device_resident_plane(dev, ref, tmp_path):
    """dwt_util_save_to_pgm_s (src/libdwt.c:19794): the same file, with one byte per sample crossing PCIe."""
    ox, oy = 300, 200
    img = dev.DeviceImage(dev.CDF97_F32, ox, oy)
    img.fill(0, 0)
    img.fwd2(2)
    img.conv_show()
    a = img.download()
    f_dev, f_ref = str(tmp_path / "dev.pgm"), str(tmp_path / "ref.pgm")
    fn = ref.lib.dwt_util_save_to_pgm_s
    fn.argtypes = [C.c_char_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    for mx in (1.0, 0.3):   # 0.3: part of the samples above the maximum (clamped to 255)
        img.save_pgm(f_dev, mx)
        assert fn(f_ref.encode(), mx, a.ctypes.data, a.strides[0], a.strides[1], ox, oy) == 0
        assert open(f_dev, "rb").read() == open(f_ref, "rb").read()
    img.close()


def test_symmetric_pgm_and_mat_of_a_device_resident_plane(dev, ref, tmp_path):
    """dwt_util_save_sym_to_pgm_s (src/libdwt.c:26184) and dwt_util_save_to_mat_s (:24430): the same files."""
    ox, oy = 200, 120
    img = dev.DeviceImage(dev.CDF97_F32, ox, oy)
    img.fill(0, 0)
    img.fwd2(3)
    a = img.download()
    f_dev, f_ref = str(tmp_path / "dev.out"), str(tmp_path / "ref.out")
    sym = ref.lib.dwt_util_save_sym_to_pgm_s
    sym.argtypes = [C.c_char_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    for mx in (4.0, 0.7):   # 0.7: coefficients beyond the range on both sides
        img.save_sym_pgm(f_dev, mx)
        assert sym(f_ref.encode(), mx, a.ctypes.data, a.strides[0], a.strides[1], ox, oy) == 0
        assert open(f_dev, "rb").read() == open(f_ref, "rb").read()
    mat = ref.lib.dwt_util_save_to_mat_s
    mat.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    img.save_mat(f_dev)
    assert mat(f_ref.encode(), a.ctypes.data, ox, oy, a.strides[0], a.strides[1]) == 0
    assert open(f_dev, "rb").read() == open(f_ref, "rb").read()
    img.save_mat(f_dev, inner=(50, 70))
    assert mat(f_ref.encode(), a.ctypes.data, 70, 50, a.strides[0], a.strides[1]) == 0
    assert open(f_dev, "rb").read() == open(f_ref, "rb").read()
    img.close()


class Volume(C.Structure):
    _fields_ = [("size_x", C.c_int), ("size_y", C.c_int), ("size_z", C.c_int), ("stride_x", C.c_size_t), ("stride_y", C.c_size_t),
                ("stride_z", C.c_size_t), ("data", C.c_void_p)]


def test_compat_volumes_in_pinned_memory(dev, ref, oracle):
    """volume_alloc_realiably_locked / volume_free (src/volume.c:194, 34) of the compat layer: the reference's strides, page-locked
    memory, and the 3-D transform of such a volume through cdf97_3f_ip_sep_horizontal_s / cdf97_3i_ip_sep_horizontal_s."""
    so = C.CDLL(os.path.join(ROOT, "libdwt_b200", "libdwt_compat.so"))
    so.volume_alloc_realiably_locked.restype = C.POINTER(Volume)
    so.volume_alloc_realiably_locked.argtypes = [C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int]
    so.volume_free.argtypes = [C.POINTER(Volume)]
    so.cdf97_3f_ip_sep_horizontal_s.argtypes = [C.POINTER(Volume)]
    so.cdf97_3i_ip_sep_horizontal_s.argtypes = [C.POINTER(Volume)]
    ref.lib.dwt_util_get_stride.argtypes = [C.c_int, C.c_int]
    ref.lib.dwt_util_get_stride.restype = C.c_int
    nx, ny, nz = 70, 50, 40
    for opt in range(8):
        v = so.volume_alloc_realiably_locked(4, nx, ny, nz, opt)
        sy = ref.lib.dwt_util_get_stride(4 * nx, opt)
        assert (v.contents.stride_x, v.contents.stride_y, v.contents.stride_z) == (4, sy, ref.lib.dwt_util_get_stride(sy * ny, opt)), opt
        so.volume_free(v)
    v = so.volume_alloc_realiably_locked(4, nx, ny, nz, 1)
    c = v.contents
    ref.lib.volume_fill_s.argtypes = [C.POINTER(Volume)]
    ref.lib.volume_fill_s(v)
    raw = (C.c_uint8 * (c.stride_z * nz)).from_address(c.data)
    a = np.ndarray(shape=(nz, ny, nx), dtype=np.float32, buffer=raw, strides=(c.stride_z, c.stride_y, 4))
    x0 = a.copy()
    want = np.ascontiguousarray(x0)
    dst = np.zeros_like(want)
    oracle.fwd3(want, dst)
    so.cdf97_3f_ip_sep_horizontal_s(v)
    assert np.ascontiguousarray(a).tobytes() == dst.tobytes()
    so.cdf97_3i_ip_sep_horizontal_s(v)
    assert np.abs(np.ascontiguousarray(a) - x0).max() < 1e-3
    so.volume_free(v)


def test_volume_measure_through_the_compat_layer(dev, tmp_path):
    """volume_measure_fwd97op_s (src/volume-dwt.c:2898): cube sizes size_min, size_grow(size) ... below size_max."""
    so = C.CDLL(os.path.join(ROOT, "libdwt_b200", "libdwt_compat.so"))
    so.volume_measure_fwd97op_s.argtypes = [C.c_int] * 6
    os.makedirs(tmp_path / "data" / "perftest")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        assert so.volume_measure_fwd97op_s(64, 100, 8, 2, 0, 0) == 0
    finally:
        os.chdir(cwd)
    lines = [l.split() for l in open(tmp_path / "data" / "perftest" / "time-stride=0-approach=0.txt") if not l.startswith("#")]
    assert [int(l[0]) for l in lines] == [64 ** 3, 80 ** 3, 96 ** 3] and all(0 < float(l[1]) < 1e-6 for l in lines)
