"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

`Oracle`  -> oracle/liborc.so          (restatement, oracle/dwt_oracle.c)
`Ref`     -> oracle/_ref/libdwt_ref.so (the unmodified reference compiled by oracle/Makefile)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under libdwt_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORC_SO = os.path.join(HERE, "liborc.so")
REF_SO = os.path.join(HERE, "_ref", "libdwt_ref.so")

_DT = {"s": np.float32, "d": np.float64, "i": np.int32}


def build():
    subprocess.check_call(["make", "-s", "-C", HERE], stdout=subprocess.DEVNULL)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """The restatement.  Images are numpy arrays indexed [y, x]; strides come from the array."""

    def __init__(self):
        if not os.path.exists(ORC_SO):
            build()
        self.lib = L = C.CDLL(ORC_SO)
        i64, ci, vp = C.c_int64, C.c_int, C.c_void_p
        for name in ("cdf97_s", "cdf97_d", "cdf53_i", "cdf53_s", "cdf53_d", "cdf97_i"):
            getattr(L, f"orc_{name}_2f").argtypes = [vp, i64, i64, ci, ci, ci, ci, C.POINTER(ci), ci, ci]
            getattr(L, f"orc_{name}_2i").argtypes = [vp, i64, i64, ci, ci, ci, ci, ci, ci, ci]
        for name in ("cdf97_s", "cdf53_s"):
            getattr(L, f"orc_{name}_2f_inplace").argtypes = [vp, i64, i64, ci, ci, ci, ci, C.POINTER(ci), ci]
            getattr(L, f"orc_{name}_2i_inplace").argtypes = [vp, i64, i64, ci, ci, ci, ci, ci, ci]
        L.orc_fill_s.argtypes = [vp, i64, i64, ci, ci, ci, ci, ci]
        L.orc_fill_d.argtypes = [vp, i64, i64, ci, ci, ci, ci]
        L.orc_fill_i.argtypes = [vp, i64, i64, ci, ci, ci, ci, ci]
        L.orc_volume_fill_s.argtypes = [vp, i64, i64, i64, ci, ci, ci]
        L.orc_cdf97_3f_s.argtypes = [vp, i64, i64, i64, vp, i64, i64, i64, ci, ci, ci]
        L.orc_cdf97_3i_s.argtypes = [vp, i64, i64, i64, ci, ci, ci]
        L.orc_ceil_log2.argtypes = [ci]
        L.orc_ceil_log2.restype = ci
        L.orc_get_max_threads.restype = ci

    name = "port"

    @staticmethod
    def _fn(wavelet, t):
        return {"97s": "cdf97_s", "97d": "cdf97_d", "53i": "cdf53_i", "53s": "cdf53_s", "53d": "cdf53_d", "97i": "cdf97_i"}[f"{wavelet}{t}"]

    def fwd2(self, img, wavelet, t, j_max=-1, decompose_one=0, zero_padding=0, inner=None):
        """In place on `img` ([oy, ox] outer array); returns achieved J."""
        oy, ox = img.shape
        iy, ix = inner if inner is not None else (oy, ox)
        j = C.c_int(j_max)
        getattr(self.lib, f"orc_{self._fn(wavelet, t)}_2f")(
            _ptr(img), img.strides[0], img.strides[1], ox, oy, ix, iy, C.byref(j), decompose_one, zero_padding)
        return j.value

    def inv2(self, img, wavelet, t, j_max=-1, decompose_one=0, zero_padding=0, inner=None):
        oy, ox = img.shape
        iy, ix = inner if inner is not None else (oy, ox)
        getattr(self.lib, f"orc_{self._fn(wavelet, t)}_2i")(
            _ptr(img), img.strides[0], img.strides[1], ox, oy, ix, iy, j_max, decompose_one, zero_padding)

    def fwd2_inplace(self, img, wavelet, j_max=-1, decompose_one=0, inner=None):
        """dwt_cdf97_2f_inplace_s (and its _sep/_sdl twins) / dwt_cdf53_2f_inplace_s: float32, interleaved layout."""
        oy, ox = img.shape
        iy, ix = inner if inner is not None else (oy, ox)
        j = C.c_int(j_max)
        getattr(self.lib, f"orc_cdf{wavelet}_s_2f_inplace")(_ptr(img), img.strides[0], img.strides[1], ox, oy, ix, iy, C.byref(j), decompose_one)
        return j.value

    def inv2_inplace(self, img, wavelet, j_max=-1, decompose_one=0, inner=None):
        oy, ox = img.shape
        iy, ix = inner if inner is not None else (oy, ox)
        getattr(self.lib, f"orc_cdf{wavelet}_s_2i_inplace")(_ptr(img), img.strides[0], img.strides[1], ox, oy, ix, iy, j_max, decompose_one)

    def fwd2_s2(self, src, dst, j_max=-1, decompose_one=0, zero_padding=0, inner=None):
        """dwt_cdf97_2f_s2 (src/libdwt.c:12619): the first pass of level 0 that runs reads its lines from src and writes
        only their L / H halves into dst (:12680-12745, then `src = dst`); everything after that is in place on dst.
        Restated with the in-place oracle: run level 0 on a scratch copy whose inner lines are src, take from it exactly
        the samples that pass wrote, then continue in place."""
        oy, ox = dst.shape
        iy, ix = inner if inner is not None else (oy, ox)
        lim = self.ceil_log2(max(ox, oy) if decompose_one else min(ox, oy))
        J = lim if (j_max < 0 or j_max > lim) else j_max
        if J == 0:
            return 0
        if (ix, iy) == (ox, oy):   # dense: every sample of dst is written by the first pass
            dst[...] = src
            return self.fwd2(dst, "97", "s", j_max=j_max, decompose_one=decompose_one, zero_padding=zero_padding, inner=inner)
        import numpy as np
        nl, nh = (ix + 1) // 2, ix // 2
        odx, ody = (ox + 1) // 2, (oy + 1) // 2
        if ox > 1:   # row pass: rows of length ix from src, as a 1-level transform of an (oy x ix)-inner, 1-row-high problem
            for y in range(oy):
                line = np.ascontiguousarray(src[y:y + 1, :ix]).copy()
                if ix >= 2:
                    tmp = np.zeros((1, ix), np.float32)
                    tmp[...] = line
                    self.fwd2(tmp, "97", "s", j_max=1, decompose_one=1)   # one level along x only (height 1: the column pass is skipped)
                    dst[y, :nl] = tmp[0, :nl]
                    dst[y, odx:odx + nh] = tmp[0, nl:nl + nh]
                else:
                    dst[y, 0] = np.float32(line[0, 0]) * np.float32(1.1496043988602)
            # column pass of level 0 in place on dst, zero padding, then the remaining levels: emulate with a 1-level
            # transform restricted to columns, i.e. transpose trick: a (ox x oy) image whose rows are dst's columns
            t = np.ascontiguousarray(dst.T).copy()
            if oy > 1:
                for x in range(ox):
                    col = np.zeros((1, iy), np.float32)
                    col[...] = t[x:x + 1, :iy]
                    if iy >= 2:
                        self.fwd2(col, "97", "s", j_max=1, decompose_one=1)
                        nly, nhy = (iy + 1) // 2, iy // 2
                        t[x, :nly] = col[0, :nly]
                        t[x, ody:ody + nhy] = col[0, nly:nly + nhy]
                    else:
                        t[x, 0] = np.float32(col[0, 0]) * np.float32(1.1496043988602)
            dst[...] = t.T
        else:
            raise NotImplementedError("width-1 sparse out-of-place case")
        if zero_padding:
            nly, nhy = (iy + 1) // 2, iy // 2
            dst[:, nl:odx] = 0
            dst[:, odx + nh:ox] = 0
            dst[nly:ody, :] = 0
            dst[ody + nhy:oy, :] = 0
        if J > 1:   # levels 1 .. J-1: the in-place transform of the LL quadrant region, i.e. of an image half the size
            sub = dst[:ody, :odx]
            self.fwd2(sub, "97", "s", j_max=J - 1, decompose_one=decompose_one, zero_padding=zero_padding,
                      inner=((iy + 1) // 2, (ix + 1) // 2))
        return J

    def inv2_s2(self, src, dst, j_max=-1, decompose_one=0, zero_padding=0, inner=None):
        """dwt_cdf97_2i_s2 (src/libdwt.c:17985): copies the inner region of src into dst (:18000), then in place."""
        oy, ox = dst.shape
        iy, ix = inner if inner is not None else (oy, ox)
        dst[:iy, :ix] = src[:iy, :ix]
        self.inv2(dst, "97", "s", j_max=j_max, decompose_one=decompose_one, zero_padding=zero_padding, inner=inner)

    def fill(self, img, t, rand=0, type_=0, wrap32=1):
        ny, nx = img.shape
        if t == "s":
            self.lib.orc_fill_s(_ptr(img), img.strides[0], img.strides[1], nx, ny, rand, type_, wrap32)
        elif t == "d":
            self.lib.orc_fill_d(_ptr(img), img.strides[0], img.strides[1], nx, ny, rand, wrap32)
        else:
            self.lib.orc_fill_i(_ptr(img), img.strides[0], img.strides[1], nx, ny, rand, type_, wrap32)
        return img

    def volume_fill(self, vol):
        nz, ny, nx = vol.shape
        self.lib.orc_volume_fill_s(_ptr(vol), vol.strides[2], vol.strides[1], vol.strides[0], nx, ny, nz)
        return vol

    def fwd3(self, src, dst):
        nz, ny, nx = src.shape
        self.lib.orc_cdf97_3f_s(_ptr(src), src.strides[2], src.strides[1], src.strides[0],
                                _ptr(dst), dst.strides[2], dst.strides[1], dst.strides[0], nx, ny, nz)

    def inv3(self, vol):
        nz, ny, nx = vol.shape
        self.lib.orc_cdf97_3i_s(_ptr(vol), vol.strides[2], vol.strides[1], vol.strides[0], nx, ny, nz)

    def ceil_log2(self, x):
        return self.lib.orc_ceil_log2(x)

    def threads(self):
        return self.lib.orc_get_max_threads()

    def set_threads(self, n):
        self.lib.orc_set_threads(n)


class _Volume(C.Structure):  # src/volume.h:14-24
    _fields_ = [("size_x", C.c_int), ("size_y", C.c_int), ("size_z", C.c_int),
                ("stride_x", C.c_size_t), ("stride_y", C.c_size_t), ("stride_z", C.c_size_t),
                ("data", C.c_void_p)]


class Ref:
    """The compiled reference, through its own prototypes (src/libdwt.h:526-992): int byte strides."""

    name = "reference"

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self):
        self.lib = L = C.CDLL(REF_SO)
        ci, vp = C.c_int, C.c_void_p
        for n in ("dwt_cdf97_2f_s", "dwt_cdf97_2f_d", "dwt_cdf53_2f_i", "dwt_cdf53_2f_s", "dwt_cdf53_2f_d", "dwt_cdf97_2f_i"):
            getattr(L, n).argtypes = [vp, ci, ci, ci, ci, ci, ci, C.POINTER(ci), ci, ci]
        for n in ("dwt_cdf97_2i_s", "dwt_cdf97_2i_d", "dwt_cdf53_2i_i", "dwt_cdf53_2i_s", "dwt_cdf53_2i_d", "dwt_cdf97_2i_i"):
            getattr(L, n).argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, ci, ci]
        for n in ("dwt_util_test_image_fill_s", "dwt_util_test_image_fill_d", "dwt_util_test_image_fill_i"):
            getattr(L, n).argtypes = [vp, ci, ci, ci, ci, ci]
        for n in ("dwt_util_test_image_fill2_s", "dwt_util_test_image_fill2_i"):
            getattr(L, n).argtypes = [vp, ci, ci, ci, ci, ci, ci]
        for n in ("dwt_cdf97_2f_inplace_s", "dwt_cdf97_2f_inplace_sep_s", "dwt_cdf97_2f_inplace_sdl_s", "dwt_cdf97_2f_inplace_sep_sdl_s",
                  "dwt_cdf53_2f_inplace_s"):
            getattr(L, n).argtypes = [vp, ci, ci, ci, ci, ci, ci, C.POINTER(ci), ci, ci]
        for n in ("dwt_cdf97_2i_inplace_s", "dwt_cdf53_2i_inplace_s"):
            getattr(L, n).argtypes = [vp, ci, ci, ci, ci, ci, ci, ci, ci, ci]
        L.dwt_cdf97_2f_s2.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, C.POINTER(ci), ci, ci]
        L.dwt_cdf97_2i_s2.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci]
        L.dwt_util_set_num_threads.argtypes = [ci]
        L.dwt_util_get_num_threads.restype = ci
        L.dwt_util_set_accel.argtypes = [ci]
        L.dwt_util_set_num_workers.argtypes = [ci]
        L.dwt_util_get_opt_stride.argtypes = [ci]
        L.dwt_util_get_opt_stride.restype = ci
        VP = C.POINTER(_Volume)
        L.volume_fill_s.argtypes = [VP]
        L.cdf97_3f_op_sep_horizontal_s.argtypes = [VP, VP]
        L.cdf97_3f_ip_sep_horizontal_s.argtypes = [VP]
        L.cdf97_3i_ip_sep_horizontal_s.argtypes = [VP]
        L.dwt_util_init()

    @staticmethod
    def _fn(wavelet, t, d):
        return {"97s": "dwt_cdf97_2%s_s", "97d": "dwt_cdf97_2%s_d", "53i": "dwt_cdf53_2%s_i",
                "53s": "dwt_cdf53_2%s_s", "53d": "dwt_cdf53_2%s_d", "97i": "dwt_cdf97_2%s_i"}[f"{wavelet}{t}"] % d

    @staticmethod
    def _check(img):
        oy, _ = img.shape
        assert abs(img.strides[0]) * oy < 2 ** 31, "reference addresses images with int (src/inline.h:188)"

    def fwd2(self, img, wavelet, t, j_max=-1, decompose_one=0, zero_padding=0, inner=None):
        self._check(img)
        oy, ox = img.shape
        iy, ix = inner if inner is not None else (oy, ox)
        j = C.c_int(j_max)
        getattr(self.lib, self._fn(wavelet, t, "f"))(
            _ptr(img), img.strides[0], img.strides[1], ox, oy, ix, iy, C.byref(j), decompose_one, zero_padding)
        return j.value

    def fwd2_inplace(self, img, wavelet, j_max=-1, decompose_one=0, inner=None, variant=""):
        """variant: "", "sep_", "sdl_", "sep_sdl_" (9/7 only).  The in-place drivers assert a single worker."""
        self._check(img)
        self.lib.dwt_util_set_num_workers(1)
        oy, ox = img.shape
        iy, ix = inner if inner is not None else (oy, ox)
        j = C.c_int(j_max)
        getattr(self.lib, f"dwt_cdf{wavelet}_2f_inplace_{variant}s")(
            _ptr(img), img.strides[0], img.strides[1], ox, oy, ix, iy, C.byref(j), decompose_one, 0)
        return j.value

    def inv2_inplace(self, img, wavelet, j_max=-1, decompose_one=0, inner=None):
        self._check(img)
        self.lib.dwt_util_set_num_workers(1)
        oy, ox = img.shape
        iy, ix = inner if inner is not None else (oy, ox)
        getattr(self.lib, f"dwt_cdf{wavelet}_2i_inplace_s")(
            _ptr(img), img.strides[0], img.strides[1], ox, oy, ix, iy, j_max, decompose_one, 0)

    def inv2(self, img, wavelet, t, j_max=-1, decompose_one=0, zero_padding=0, inner=None):
        self._check(img)
        oy, ox = img.shape
        iy, ix = inner if inner is not None else (oy, ox)
        getattr(self.lib, self._fn(wavelet, t, "i"))(
            _ptr(img), img.strides[0], img.strides[1], ox, oy, ix, iy, j_max, decompose_one, zero_padding)

    def fwd2_s2(self, src, dst, j_max=-1, decompose_one=0, zero_padding=0, inner=None):
        assert src.strides == dst.strides
        oy, ox = dst.shape
        iy, ix = inner if inner is not None else (oy, ox)
        j = C.c_int(j_max)
        self.lib.dwt_cdf97_2f_s2(_ptr(src), _ptr(dst), dst.strides[0], dst.strides[1], ox, oy, ix, iy, C.byref(j), decompose_one, zero_padding)
        return j.value

    def inv2_s2(self, src, dst, j_max=-1, decompose_one=0, zero_padding=0, inner=None):
        assert src.strides == dst.strides
        oy, ox = dst.shape
        iy, ix = inner if inner is not None else (oy, ox)
        self.lib.dwt_cdf97_2i_s2(_ptr(src), _ptr(dst), dst.strides[0], dst.strides[1], ox, oy, ix, iy, j_max, decompose_one, zero_padding)

    def fill(self, img, t, rand=0, type_=0, wrap32=1):
        self._check(img)
        ny, nx = img.shape
        if t == "d":
            assert type_ == 0
            self.lib.dwt_util_test_image_fill_d(_ptr(img), img.strides[0], img.strides[1], nx, ny, rand)
        else:
            getattr(self.lib, f"dwt_util_test_image_fill2_{t}")(
                _ptr(img), img.strides[0], img.strides[1], nx, ny, rand, type_)
        return img

    @staticmethod
    def _vol(a):
        nz, ny, nx = a.shape
        return _Volume(nx, ny, nz, a.strides[2], a.strides[1], a.strides[0], a.ctypes.data)

    def volume_fill(self, vol):
        v = self._vol(vol)
        self.lib.volume_fill_s(C.byref(v))
        return vol

    def fwd3(self, src, dst):
        s, d = self._vol(src), self._vol(dst)
        self.lib.cdf97_3f_op_sep_horizontal_s(C.byref(s), C.byref(d))

    def inv3(self, vol):
        v = self._vol(vol)
        self.lib.cdf97_3i_ip_sep_horizontal_s(C.byref(v))

    def threads(self):
        return self.lib.dwt_util_get_num_threads()

    def set_threads(self, n):
        self.lib.dwt_util_set_num_threads(n)

    def opt_stride(self, nbytes):
        return self.lib.dwt_util_get_opt_stride(nbytes)


def strided_image(shape, t, row_bytes=None):
    """A [oy, ox] view with an arbitrary (possibly unaligned, e.g. prime) row stride in bytes."""
    dt = np.dtype(_DT[t])
    oy, ox = shape
    row_bytes = row_bytes or ox * dt.itemsize
    raw = np.zeros(row_bytes * oy + 64, dtype=np.uint8)
    return np.ndarray(shape=(oy, ox), dtype=dt, buffer=raw, strides=(row_bytes, dt.itemsize))
