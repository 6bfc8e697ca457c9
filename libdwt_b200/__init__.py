"""libdwt_b200 -- B200 (sm_100a) implementation of libdwt's separable lifting DWT hot path.

The product is libdwtb200.so (hand-written CUDA kernels behind the C ABI of include/dwtb200.h) and
libdwt_compat.so (the reference's own symbol names on top of it).  This package is the thin Python
host layer used by the tests and bench.py: ctypes bindings, nothing else.  It never imports anything
from oracle/ and has no CPU fallback -- importing works without a GPU (so the symbol checks can run),
every compute call raises DwtError when no CUDA device is usable.
"""
from .api import (  # noqa: F401
    CDF97_F32, CDF97_F64, CDF53_I32, CDF53_F32, CDF53_F64, CDF97_I32, DwtError, DeviceImage, DeviceStrips, DeviceVolume, Library, lib, strips_band, strips_plan,
    dwt_cdf97_2f_s, dwt_cdf97_2i_s, dwt_cdf97_2f_d, dwt_cdf97_2i_d, dwt_cdf53_2f_i, dwt_cdf53_2i_i,
    dwt_cdf53_2f_s, dwt_cdf53_2i_s, dwt_cdf53_2f_d, dwt_cdf53_2i_d, dwt_cdf97_2f_i, dwt_cdf97_2i_i,
    dwt_cdf97_2f_s2, dwt_cdf97_2i_s2, dwt_cdf97_2f_inplace_s, dwt_cdf97_2f_inplace_sep_s, dwt_cdf97_2f_inplace_sdl_s,
    dwt_cdf97_2f_inplace_sep_sdl_s, dwt_cdf97_2i_inplace_s, dwt_cdf53_2f_inplace_s, dwt_cdf53_2i_inplace_s, fwd2_inplace, inv2_inplace,
    perf2, perf2_inplace, perf3, fwd2, inv2, fwd3, inv3, kind_of,
)
