"""ctypes bindings of libdwtb200.so (include/dwtb200.h) and the reference-named host API.

The module-level functions dwt_cdf97_2f_s ... dwt_cdf53_2i_i take exactly the arguments of the
reference prototypes (/root/reference/src/libdwt.h:526-992): a host pointer (here: anything exposing
the buffer protocol / a numpy array / an int address), byte strides, outer and inner sizes, j_max
(by reference for the forward transforms: a ctypes.c_int or a one-element list), decompose_one and
zero_padding.  They are synchronous and in place on host memory like the reference.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libdwtb200.so")

CDF97_F32, CDF97_F64, CDF53_I32, CDF53_F32, CDF53_F64, CDF97_I32 = 0, 1, 2, 3, 4, 5
_DT = {CDF97_F32: np.float32, CDF97_F64: np.float64, CDF53_I32: np.int32,
       CDF53_F32: np.float32, CDF53_F64: np.float64, CDF97_I32: np.int32}
_KIND = {("97", "s"): CDF97_F32, ("97", "d"): CDF97_F64, ("53", "i"): CDF53_I32,
         ("53", "s"): CDF53_F32, ("53", "d"): CDF53_F64, ("97", "i"): CDF97_I32}

# every symbol include/dwtb200.h declares: (name, restype, argtypes)
_i, _i64, _vp, _sz, _dbl = C.c_int, C.c_int64, C.c_void_p, C.c_size_t, C.c_double
_ip = C.POINTER(C.c_int)
SYMBOLS = [
    ("dwtb200_init", _i, [_i]), ("dwtb200_finish", None, []), ("dwtb200_last_error", C.c_char_p, []),
    ("dwtb200_device_count", _i, []), ("dwtb200_device", _i, []),
    ("dwtb200_host_alloc", _vp, [_sz]), ("dwtb200_host_free", None, [_vp]), ("dwtb200_host_numa_node", _i, []),
    ("dwtb200_ceil_log2", _i, [_i]), ("dwtb200_clamp_j", _i, [_i, _i, _i, _i]),
    ("dwtb200_fwd2_host", _i, [_i, _vp, _i64, _i64, _i, _i, _i, _i, _ip, _i, _i]),
    ("dwtb200_inv2_host", _i, [_i, _vp, _i64, _i64, _i, _i, _i, _i, _i, _i, _i]),
    ("dwtb200_fwd2_host2", _i, [_i, _vp, _vp, _i64, _i64, _i, _i, _i, _i, _ip, _i, _i]),
    ("dwtb200_inv2_host2", _i, [_i, _vp, _vp, _i64, _i64, _i, _i, _i, _i, _i, _i, _i]),
    ("dwtb200_fwd2_inplace_host", _i, [_i, _vp, _i64, _i64, _i, _i, _i, _i, _ip, _i]),
    ("dwtb200_inv2_inplace_host", _i, [_i, _vp, _i64, _i64, _i, _i, _i, _i, _i, _i]),
    ("dwtb200_image_fwd2_inplace", _i, [_vp, _ip, _i]), ("dwtb200_image_inv2_inplace", _i, [_vp, _i, _i]),
    ("dwtb200_perf2", _i, [_i, _i, _i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    ("dwtb200_perf2_inplace", _i, [_i, _i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    ("dwtb200_perf3", _i, [_i, _i, C.POINTER(_dbl), _ip]),
    ("dwtb200_last_transform_ms", _dbl, []), ("dwtb200_release_host_cache", None, []),
    ("dwtb200_image_create", _vp, [_i, _i, _i, _i]), ("dwtb200_image_destroy", None, [_vp]),
    ("dwtb200_image_upload", _i, [_vp, _i, _vp, _i64, _i64]), ("dwtb200_image_download", _i, [_vp, _i, _vp, _i64, _i64]),
    ("dwtb200_image_fill", _i, [_vp, _i, _i, _i]), ("dwtb200_image_fill_ex", _i, [_vp, _i, _i, _i, _i, _i]),
    ("dwtb200_image_copy_rows", _i, [_vp, _i, _i, _i, _vp, _i64, _i]),
    ("dwtb200_image_ipc_export", _i, [_vp, _vp]), ("dwtb200_ipc_open", _vp, [_vp]), ("dwtb200_ipc_close", _i, [_vp]),
    ("dwtb200_image_fwd2", _i, [_vp, _i, _i, _ip, _i, _i]), ("dwtb200_image_inv2", _i, [_vp, _i, _i, _i, _i, _i]),
    ("dwtb200_image_devptr", _vp, [_vp, C.POINTER(_sz), C.POINTER(_sz)]),
    ("dwtb200_image_subband", _i, [_vp, _i, _i, _i, _i, _i, C.POINTER(_vp), C.POINTER(_sz), _ip, _ip]),
    ("dwtb200_image_subband_moments", _i, [_vp, _i, _i, _i, _i, _i, C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_dbl)]),
    ("dwtb200_image_features", _i, [_vp, _i, _i, _i, _i, _i, C.POINTER(C.c_float), _ip]),
    ("dwtb200_image_diff", _i64, [_vp, _vp]), ("dwtb200_image_maxabs", _dbl, [_vp, _vp]),
    ("dwtb200_image_copy", _i, [_vp, _vp]),
    ("dwtb200_image_conv_show", _i, [_vp, _vp, _i, _i]), ("dwtb200_image_save_pgm", _i, [_vp, _i, C.c_char_p, _dbl, _i, _i]),
    ("dwtb200_image_save_sym_pgm", _i, [_vp, _i, C.c_char_p, _dbl, _i, _i]), ("dwtb200_image_save_mat", _i, [_vp, _i, C.c_char_p, _i, _i]),
    ("dwtb200_image_last_launches", _i, [_vp]), ("dwtb200_image_last_path", _i, [_vp]),
    ("dwtb200_force_generic", None, [_i]), ("dwtb200_set_strip_rows", None, [_i]),
    ("dwtb200_set_tuning", _i, [_i, C.c_longlong]),
    ("dwtb200_fwd3_host", _i, [_vp, _sz, _sz, _sz, _vp, _sz, _sz, _sz, _i, _i, _i]),
    ("dwtb200_inv3_host", _i, [_vp, _sz, _sz, _sz, _i, _i, _i]),
    ("dwtb200_volume_create", _vp, [_i, _i, _i]), ("dwtb200_volume_destroy", None, [_vp]),
    ("dwtb200_volume_upload", _i, [_vp, _vp, _sz, _sz, _sz]), ("dwtb200_volume_download", _i, [_vp, _vp, _sz, _sz, _sz]),
    ("dwtb200_volume_fill", _i, [_vp]), ("dwtb200_volume_fwd3", _i, [_vp]), ("dwtb200_volume_inv3", _i, [_vp]),
    ("dwtb200_sync", _i, []), ("dwtb200_timer_start", _i, []), ("dwtb200_timer_stop_ms", _dbl, []),
    ("dwtb200_stream", _vp, []), ("dwtb200_flush_l2", _i, [_sz]),
    ("dwtb200_image_timer_start", _i, [_vp]), ("dwtb200_image_timer_stop_ms", _dbl, [_vp]),
    ("dwtb200_image_timer_mark", _i, [_vp]), ("dwtb200_image_timer_read", _i, [_vp, C.POINTER(_dbl), _i]), ("dwtb200_image_wait", _i, [_vp, _vp]),
    # one image as row strips over the GPUs of a box
    ("dwtb200_strips_plan", _i, [_i, _i, _i, _i, _i, _i, _vp]), ("dwtb200_strips_band", _i, [_i, _i, _i, _i, _i, _i, _i, _ip]),
    ("dwtb200_strips_create", _vp, [_i, _i, _i, _i, _i, _i, C.c_char_p]), ("dwtb200_strips_connect", _i, [_vp]),
    ("dwtb200_strips_destroy", None, [_vp]), ("dwtb200_strips_image", _vp, [_vp]), ("dwtb200_strips_top", _vp, [_vp]),
    ("dwtb200_strips_levels", _i, [_vp, _ip, _ip]), ("dwtb200_strips_get_plan", _i, [_vp, _vp]),
    ("dwtb200_strips_fwd2", _i, [_vp, _ip]), ("dwtb200_strips_inv2", _i, [_vp, _i]), ("dwtb200_strips_sync", _i, [_vp]),
    ("dwtb200_strips_last_peer_bytes", C.c_ulonglong, [_vp]), ("dwtb200_strips_compare_owned", _i64, [_vp, _vp, _i]),
    ("dwtb200_strips_download_owned", _i, [_vp, _vp, _i64, _i]),
]


class DwtError(RuntimeError):
    pass


class Library:
    """The loaded C ABI.  Loading needs libcudart but no GPU; compute calls need a B200."""

    def __init__(self, path=SO):
        if not os.path.exists(path):
            raise DwtError(f"{path} is missing: run `make` (or __graft_entry__.build()); there is no fallback path")
        self.c = C.CDLL(path)
        for name, res, args in SYMBOLS:
            f = getattr(self.c, name)
            f.restype, f.argtypes = res, args

    def check(self, rc):
        if rc != 0:
            raise DwtError(f"libdwtb200 error {rc}: {self.c.dwtb200_last_error().decode()}")

    def init(self, device=-1):
        self.check(self.c.dwtb200_init(device))


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = Library()
    return _lib


def kind_of(wavelet, t):
    return _KIND[(str(wavelet), t)]


def _addr(ptr):
    if isinstance(ptr, np.ndarray):
        return ptr.ctypes.data
    if isinstance(ptr, int):
        return ptr
    if isinstance(ptr, C.c_void_p):
        return ptr.value
    return C.addressof(C.c_char.from_buffer(ptr))


# ---- the reference's prototypes (src/libdwt.h:526-537, 562-573, 686-697, 831-842, 867-878, 981-992) ----
def _fwd(kind, ptr, stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x, size_i_big_y, j_max_ptr,
         decompose_one, zero_padding):
    j = C.c_int(j_max_ptr[0] if isinstance(j_max_ptr, list) else j_max_ptr.value)
    L = lib()
    L.check(L.c.dwtb200_fwd2_host(kind, _addr(ptr), stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x,
                                  size_i_big_y, C.byref(j), decompose_one, zero_padding))
    if isinstance(j_max_ptr, list):
        j_max_ptr[0] = j.value
    else:
        j_max_ptr.value = j.value


def _inv(kind, ptr, stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x, size_i_big_y, j_max,
         decompose_one, zero_padding):
    L = lib()
    L.check(L.c.dwtb200_inv2_host(kind, _addr(ptr), stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x,
                                  size_i_big_y, j_max, decompose_one, zero_padding))


def dwt_cdf97_2f_s(*a): _fwd(CDF97_F32, *a)
def dwt_cdf97_2i_s(*a): _inv(CDF97_F32, *a)
def dwt_cdf97_2f_d(*a): _fwd(CDF97_F64, *a)
def dwt_cdf97_2i_d(*a): _inv(CDF97_F64, *a)
def dwt_cdf53_2f_i(*a): _fwd(CDF53_I32, *a)
def dwt_cdf53_2i_i(*a): _inv(CDF53_I32, *a)
def dwt_cdf53_2f_s(*a): _fwd(CDF53_F32, *a)
def dwt_cdf53_2i_s(*a): _inv(CDF53_F32, *a)
def dwt_cdf53_2f_d(*a): _fwd(CDF53_F64, *a)
def dwt_cdf53_2i_d(*a): _inv(CDF53_F64, *a)
def dwt_cdf97_2f_i(*a): _fwd(CDF97_I32, *a)
def dwt_cdf97_2i_i(*a): _inv(CDF97_I32, *a)


def dwt_cdf97_2f_s2(src, dst, stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x, size_i_big_y, j_max_ptr, decompose_one,
                    zero_padding):
    j = C.c_int(j_max_ptr[0] if isinstance(j_max_ptr, list) else j_max_ptr.value)
    L = lib()
    L.check(L.c.dwtb200_fwd2_host2(CDF97_F32, _addr(src), _addr(dst), stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x,
                                   size_i_big_y, C.byref(j), decompose_one, zero_padding))
    if isinstance(j_max_ptr, list):
        j_max_ptr[0] = j.value
    else:
        j_max_ptr.value = j.value


def dwt_cdf97_2i_s2(src, dst, stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x, size_i_big_y, j_max, decompose_one,
                    zero_padding):
    L = lib()
    L.check(L.c.dwtb200_inv2_host2(CDF97_F32, _addr(src), _addr(dst), stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x,
                                   size_i_big_y, j_max, decompose_one, zero_padding))


# ---- interleaved in-place family (src/libdwt.h:586-662, 889-900, 599-610, 944-955); zero_padding is ignored ----
def _fwd_ip(kind, ptr, stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x, size_i_big_y, j_max_ptr, decompose_one,
            zero_padding=0):
    j = C.c_int(j_max_ptr[0] if isinstance(j_max_ptr, list) else j_max_ptr.value)
    L = lib()
    L.check(L.c.dwtb200_fwd2_inplace_host(kind, _addr(ptr), stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x,
                                          size_i_big_y, C.byref(j), decompose_one))
    if isinstance(j_max_ptr, list):
        j_max_ptr[0] = j.value
    else:
        j_max_ptr.value = j.value


def _inv_ip(kind, ptr, stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x, size_i_big_y, j_max, decompose_one,
            zero_padding=0):
    L = lib()
    L.check(L.c.dwtb200_inv2_inplace_host(kind, _addr(ptr), stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x,
                                          size_i_big_y, j_max, decompose_one))


def dwt_cdf97_2f_inplace_s(*a): _fwd_ip(CDF97_F32, *a)
def dwt_cdf97_2f_inplace_sep_s(*a): _fwd_ip(CDF97_F32, *a)
def dwt_cdf97_2f_inplace_sdl_s(*a): _fwd_ip(CDF97_F32, *a)
def dwt_cdf97_2f_inplace_sep_sdl_s(*a): _fwd_ip(CDF97_F32, *a)
def dwt_cdf97_2i_inplace_s(*a): _inv_ip(CDF97_F32, *a)
def dwt_cdf53_2f_inplace_s(*a): _fwd_ip(CDF53_F32, *a)
def dwt_cdf53_2i_inplace_s(*a): _inv_ip(CDF53_F32, *a)


def fwd2_inplace(img, wavelet, j_max=-1, decompose_one=0, inner=None):
    """numpy convenience: img is a [y, x] float32 array, transformed in place (interleaved layout); returns achieved J."""
    oy, ox = img.shape
    iy, ix = inner if inner is not None else (oy, ox)
    j = [j_max]
    _fwd_ip(kind_of(wavelet, "s"), img, img.strides[0], img.strides[1], ox, oy, ix, iy, j, decompose_one)
    return j[0]


def inv2_inplace(img, wavelet, j_max=-1, decompose_one=0, inner=None):
    oy, ox = img.shape
    iy, ix = inner if inner is not None else (oy, ox)
    _inv_ip(kind_of(wavelet, "s"), img, img.strides[0], img.strides[1], ox, oy, ix, iy, j_max, decompose_one)


def perf2(kind, size_x, size_y, j_max=-1, M=1, N=1, inner=None, decompose_one=0, zero_padding=0):
    """dwt_util_perf_cdf97_2_s / dwt_util_perf_cdf53_2_i on the device: (fwd_secs, inv_secs) per transform."""
    iy, ix = inner if inner is not None else (size_y, size_x)
    f, i = C.c_float(), C.c_float()
    L = lib()
    L.check(L.c.dwtb200_perf2(kind, size_x, size_y, ix, iy, j_max, decompose_one, zero_padding, M, N, C.byref(f), C.byref(i)))
    return f.value, i.value


def perf2_inplace(kind, size_x, size_y, j_max=-1, M=1, N=1, inner=None, decompose_one=0):
    """dwt_util_perf_cdf97_2_inplace_s (and twins) on the device: (fwd_secs, inv_secs) per transform."""
    iy, ix = inner if inner is not None else (size_y, size_x)
    f, i = C.c_float(), C.c_float()
    L = lib()
    L.check(L.c.dwtb200_perf2_inplace(kind, size_x, size_y, ix, iy, j_max, decompose_one, M, N, C.byref(f), C.byref(i)))
    return f.value, i.value


def perf3(size, N=1):
    """volume_perftest_fwd97op_s on the device: (seconds per voxel of the forward transform, failed round trips)."""
    s, e = C.c_double(), C.c_int()
    L = lib()
    L.check(L.c.dwtb200_perf3(size, N, C.byref(s), C.byref(e)))
    return s.value, e.value


# ---- numpy conveniences with the calling shape of oracle/orc.py (images are [y, x] arrays) ----
def fwd2(img, wavelet, t, j_max=-1, decompose_one=0, zero_padding=0, inner=None):
    oy, ox = img.shape
    iy, ix = inner if inner is not None else (oy, ox)
    j = [j_max]
    _fwd(kind_of(wavelet, t), img, img.strides[0], img.strides[1], ox, oy, ix, iy, j, decompose_one, zero_padding)
    return j[0]


def inv2(img, wavelet, t, j_max=-1, decompose_one=0, zero_padding=0, inner=None):
    oy, ox = img.shape
    iy, ix = inner if inner is not None else (oy, ox)
    _inv(kind_of(wavelet, t), img, img.strides[0], img.strides[1], ox, oy, ix, iy, j_max, decompose_one, zero_padding)


def fwd3(src, dst):
    """cdf97_3f_op_sep_horizontal_s (src/volume-dwt.c:727): arrays are [z, y, x] float32."""
    nz, ny, nx = src.shape
    L = lib()
    L.check(L.c.dwtb200_fwd3_host(src.ctypes.data, src.strides[2], src.strides[1], src.strides[0],
                                  dst.ctypes.data, dst.strides[2], dst.strides[1], dst.strides[0], nx, ny, nz))


def inv3(vol):
    """cdf97_3i_ip_sep_horizontal_s (src/volume-dwt.c:1115)."""
    nz, ny, nx = vol.shape
    L = lib()
    L.check(L.c.dwtb200_inv3_host(vol.ctypes.data, vol.strides[2], vol.strides[1], vol.strides[0], nx, ny, nz))


class DeviceImage:
    """`frames` device-resident planes (dwtb200_image): what the roofline numbers are measured on."""

    def __init__(self, kind, size_x, size_y, frames=1):
        self.L = lib()
        self.L.init()
        self.kind, self.size_x, self.size_y, self.frames = kind, size_x, size_y, frames
        self.h = self.L.c.dwtb200_image_create(kind, size_x, size_y, frames)
        if not self.h:
            raise DwtError(self.L.c.dwtb200_last_error().decode())

    def close(self):
        if self.h:
            self.L.c.dwtb200_image_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def dtype(self):
        return _DT[self.kind]

    def upload(self, arr, frame=0):
        self.L.check(self.L.c.dwtb200_image_upload(self.h, frame, arr.ctypes.data, arr.strides[0], arr.strides[1]))

    def download(self, arr=None, frame=0):
        if arr is None:
            arr = np.empty((self.size_y, self.size_x), dtype=self.dtype)
        self.L.check(self.L.c.dwtb200_image_download(self.h, frame, arr.ctypes.data, arr.strides[0], arr.strides[1]))
        return arr

    def fill(self, rand=0, type_=0, rand_mod=0, y_offset=0, wide=0):
        self.L.check(self.L.c.dwtb200_image_fill_ex(self.h, rand, type_, rand_mod, y_offset, wide))

    def copy_rows(self, row0, rows, buf_ptr, buf_pitch_bytes, to_image, frame=0):
        """rows of the current plane <-> a dense buffer given by address (host, device or peer device)"""
        self.L.check(self.L.c.dwtb200_image_copy_rows(self.h, frame, row0, rows, buf_ptr, buf_pitch_bytes, 1 if to_image else 0))

    def devptr(self):
        pitch, frame = C.c_size_t(), C.c_size_t()
        p = self.L.c.dwtb200_image_devptr(self.h, C.byref(pitch), C.byref(frame))
        return p, pitch.value, frame.value

    def ipc_handle(self):
        buf = C.create_string_buffer(64)
        self.L.check(self.L.c.dwtb200_image_ipc_export(self.h, buf))
        return buf.raw

    def fwd2(self, j_max=-1, decompose_one=0, zero_padding=0, inner=None):
        iy, ix = inner if inner is not None else (self.size_y, self.size_x)
        j = C.c_int(j_max)
        self.L.check(self.L.c.dwtb200_image_fwd2(self.h, ix, iy, C.byref(j), decompose_one, zero_padding))
        return j.value

    def inv2(self, j_max=-1, decompose_one=0, zero_padding=0, inner=None):
        iy, ix = inner if inner is not None else (self.size_y, self.size_x)
        self.L.check(self.L.c.dwtb200_image_inv2(self.h, ix, iy, j_max, decompose_one, zero_padding))

    def fwd2_inplace(self, j_max=-1, decompose_one=0):
        """dwt_cdf97_2f_inplace_s / dwt_cdf53_2f_inplace_s on the device-resident frames (interleaved layout)."""
        j = C.c_int(j_max)
        self.L.check(self.L.c.dwtb200_image_fwd2_inplace(self.h, C.byref(j), decompose_one))
        return j.value

    def inv2_inplace(self, j_max=-1, decompose_one=0):
        self.L.check(self.L.c.dwtb200_image_inv2_inplace(self.h, j_max, decompose_one))

    def subband(self, j, band, frame=0, inner=None):
        """dwt_util_subband: (device pointer, pitch in bytes, size_x, size_y) of LL/HL/LH/HH (0..3) of level j."""
        iy, ix = inner if inner is not None else (self.size_y, self.size_x)
        p, pitch, sx, sy = C.c_void_p(), C.c_size_t(), C.c_int(), C.c_int()
        self.L.check(self.L.c.dwtb200_image_subband(self.h, frame, ix, iy, j, band, C.byref(p), C.byref(pitch), C.byref(sx), C.byref(sy)))
        return p.value, pitch.value, sx.value, sy.value

    def subband_moments(self, j, band, frame=0, inner=None):
        """(sum, sum of squares, max |x|) of a subband, accumulated in double on the device."""
        iy, ix = inner if inner is not None else (self.size_y, self.size_x)
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self.L.check(self.L.c.dwtb200_image_subband_moments(self.h, frame, ix, iy, j, band, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    FEATURES = {"wps": 0, "mean": 1, "var": 2, "stdev": 3, "maxnorm": 4, "norm": 5}

    def features(self, j_max, feature, frame=0, inner=None):
        """dwt_util_wps_s / _mean_s / _var_s / _stdev_s / _maxnorm_s / _norm_s of the device-resident Mallat plane: numpy float32 vector."""
        iy, ix = inner if inner is not None else (self.size_y, self.size_x)
        fv = (C.c_float * (3 * max(j_max, 1)))()
        n = C.c_int()
        self.L.check(self.L.c.dwtb200_image_features(self.h, frame, ix, iy, j_max, self.FEATURES[feature], fv, C.byref(n)))
        return np.array(fv[:n.value], dtype=np.float32)

    def diff(self, other):
        r = self.L.c.dwtb200_image_diff(self.h, other.h)
        if r < 0:
            raise DwtError(self.L.c.dwtb200_last_error().decode())
        return r

    def maxabs(self, other):
        r = self.L.c.dwtb200_image_maxabs(self.h, other.h)
        if r < 0:
            raise DwtError(self.L.c.dwtb200_last_error().decode())
        return r

    def timer_start(self):
        self.L.check(self.L.c.dwtb200_image_timer_start(self.h))

    def timer_stop_ms(self):
        return self.L.c.dwtb200_image_timer_stop_ms(self.h)

    def mark(self):
        self.L.check(self.L.c.dwtb200_image_timer_mark(self.h))

    def read_marks(self, capacity=4096):
        buf = (C.c_double * capacity)()
        n = self.L.c.dwtb200_image_timer_read(self.h, buf, capacity)
        if n < 0:
            raise DwtError(self.L.c.dwtb200_last_error().decode())
        return list(buf[:n])

    def wait_for(self, other):
        self.L.check(self.L.c.dwtb200_image_wait(self.h, other.h))

    def conv_show(self, dst=None, inner=None):
        """dwt_util_conv_show_* of the current plane into `dst` (default: in place)."""
        iy, ix = inner if inner is not None else (self.size_y, self.size_x)
        self.L.check(self.L.c.dwtb200_image_conv_show(self.h, (dst or self).h, ix, iy))

    def save_pgm(self, filename, max_value, frame=0, inner=None):
        iy, ix = inner if inner is not None else (self.size_y, self.size_x)
        self.L.check(self.L.c.dwtb200_image_save_pgm(self.h, frame, filename.encode(), max_value, ix, iy))

    def save_sym_pgm(self, filename, max_value, frame=0, inner=None):
        iy, ix = inner if inner is not None else (self.size_y, self.size_x)
        self.L.check(self.L.c.dwtb200_image_save_sym_pgm(self.h, frame, filename.encode(), max_value, ix, iy))

    def save_mat(self, filename, frame=0, inner=None):
        iy, ix = inner if inner is not None else (self.size_y, self.size_x)
        self.L.check(self.L.c.dwtb200_image_save_mat(self.h, frame, filename.encode(), ix, iy))

    def copy_from(self, other):
        self.L.check(self.L.c.dwtb200_image_copy(self.h, other.h))

    @property
    def last_launches(self):
        return self.L.c.dwtb200_image_last_launches(self.h)

    @property
    def last_path(self):
        return self.L.c.dwtb200_image_last_path(self.h)


class StripPlanC(C.Structure):
    """dwtb200_strip_plan (include/dwtb200.h)."""
    _fields_ = [(n, C.c_int) for n in ("halo", "own0", "own1", "ext0", "ext1", "ll_w", "ll_h", "ll_own0", "ll_own1", "ll_ext0", "ll_ext1",
                                       "neighbours_only")]


def strips_plan(width, height, world, levels_distributed, halo_lines, rank):
    p = StripPlanC()
    L = lib()
    L.check(L.c.dwtb200_strips_plan(width, height, world, levels_distributed, halo_lines, rank, C.byref(p)))
    return p


def strips_band(width, height, world, levels_distributed, halo_lines, rank, j):
    out = (C.c_int * 11)()
    L = lib()
    L.check(L.c.dwtb200_strips_band(width, height, world, levels_distributed, halo_lines, rank, j, out))
    return dict(zip(("off", "nly_g", "nly_l", "extL0", "extL1", "extH0", "extH1", "ownL0", "ownL1", "ownH0", "ownH1"), out))


class _Borrowed(DeviceImage):
    """A dwtb200_image owned by another object (the strip / top image of a DeviceStrips)."""

    def __init__(self, handle, kind, size_x, size_y):
        self.L = lib()
        self.kind, self.size_x, self.size_y, self.frames = kind, size_x, size_y, 1
        self.h = handle

    def close(self):
        self.h = None


class DeviceStrips:
    """ONE image as row strips over the GPUs of a box (dwtb200_strips_*): this process is `rank` of `world`.
    `session` must be the same string on every rank and unique to this object."""

    def __init__(self, kind, width, height, rank, world, session, levels_distributed=0):
        self.L = lib()
        self.L.init()
        self.kind, self.width, self.height, self.rank, self.world = kind, width, height, rank, world
        self.h = self.L.c.dwtb200_strips_create(kind, width, height, levels_distributed, rank, world, session.encode())
        if not self.h:
            raise DwtError(self.L.c.dwtb200_last_error().decode())
        self.plan = StripPlanC()
        self.L.check(self.L.c.dwtb200_strips_get_plan(self.h, C.byref(self.plan)))
        jt, jd = C.c_int(), C.c_int()
        self.L.check(self.L.c.dwtb200_strips_levels(self.h, C.byref(jt), C.byref(jd)))
        self.J, self.Jd = jt.value, jd.value
        self.image = _Borrowed(self.L.c.dwtb200_strips_image(self.h), kind, width, self.plan.ext1 - self.plan.ext0)
        t = self.L.c.dwtb200_strips_top(self.h)
        self.top = _Borrowed(t, kind, self.plan.ll_w, self.plan.ll_h) if t else None

    def close(self):
        if self.h:
            self.L.c.dwtb200_strips_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def connect(self):
        self.L.check(self.L.c.dwtb200_strips_connect(self.h))

    def fill(self, rand=0, type_=0, wide=0):
        """the test pattern of the whole picture on the rows this rank holds"""
        self.image.fill(rand, type_, 0, y_offset=self.plan.ext0, wide=wide)

    def fwd2(self):
        j = C.c_int(-1)
        self.L.check(self.L.c.dwtb200_strips_fwd2(self.h, C.byref(j)))
        return j.value

    def inv2(self, j_max=-1):
        self.L.check(self.L.c.dwtb200_strips_inv2(self.h, j_max))

    def sync(self):
        self.L.check(self.L.c.dwtb200_strips_sync(self.h))

    @property
    def last_peer_bytes(self):
        return int(self.L.c.dwtb200_strips_last_peer_bytes(self.h))

    def compare_owned(self, full, mallat):
        r = self.L.c.dwtb200_strips_compare_owned(self.h, full.h, 1 if mallat else 0)
        if r < 0:
            raise DwtError(self.L.c.dwtb200_last_error().decode())
        return r

    def download_owned(self, host_full, mallat):
        self.L.check(self.L.c.dwtb200_strips_download_owned(self.h, host_full.ctypes.data, host_full.strides[0], 1 if mallat else 0))


class DeviceVolume:
    def __init__(self, nx, ny, nz):
        self.L = lib()
        self.L.init()
        self.nx, self.ny, self.nz = nx, ny, nz
        self.h = self.L.c.dwtb200_volume_create(nx, ny, nz)
        if not self.h:
            raise DwtError(self.L.c.dwtb200_last_error().decode())

    def close(self):
        if self.h:
            self.L.c.dwtb200_volume_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, a):
        self.L.check(self.L.c.dwtb200_volume_upload(self.h, a.ctypes.data, a.strides[2], a.strides[1], a.strides[0]))

    def download(self, a=None):
        if a is None:
            a = np.empty((self.nz, self.ny, self.nx), dtype=np.float32)
        self.L.check(self.L.c.dwtb200_volume_download(self.h, a.ctypes.data, a.strides[2], a.strides[1], a.strides[0]))
        return a

    def fill(self):
        self.L.check(self.L.c.dwtb200_volume_fill(self.h))

    def fwd3(self):
        self.L.check(self.L.c.dwtb200_volume_fwd3(self.h))

    def inv3(self):
        self.L.check(self.L.c.dwtb200_volume_inv3(self.h))
