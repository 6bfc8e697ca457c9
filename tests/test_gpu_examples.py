"""GPU: the reference's UNMODIFIED example programs (examples/simple, simple-int, simple-double), linked
against libdwt_compat.so ahead of the compiled reference (Makefile target `examples`), must print
"success": they allocate with dwt_util_alloc_image, use a prime row stride (2053 bytes for 512 floats),
call dwt_cdfXX_2f_* / 2i_* and compare the round trip with the reference's own dwt_util_compare_*.
simple-single-loop is the in-place family's example (dwt_cdf97_2f_inplace_sdl_s / dwt_cdf97_2i_inplace_s, decompose_one = 1)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["simple", "simple-int", "simple-double", "simple-perf", "simple-perf-int", "simple-single-loop", "simple-perf-single"])
def test_unmodified_example_runs_on_the_gpu(name, tmp_path):
    exe = os.path.join(ROOT, "build", "examples", name)
    assert os.path.exists(exe), f"{exe} missing: run `make examples` in the build container (it travels via gpurun)"
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-2000:]
    assert "success" in out and "images differs" not in out, out[-2000:]
    if "perf" in name:   # dwt_util_perf_cdf97_2_s / dwt_util_perf_cdf53_2_i timed on the device
        import re
        m = re.search(r"performance test: fwd=([0-9.eE+-]+) secs", out)
        assert m and 0 < float(m.group(1)) < 0.01, out[-2000:]   # 1920x1080, one level: well under 10 ms on a B200


def test_reference_test_program_runs_on_the_gpu(tmp_path):
    """examples/test/test.c (the reference's own round-trip test of the path -- float in place and out of place under 16
    acceleration settings, double, integer 9/7; 256 x 256, prime stride, decompose_one = 1), unmodified, with libdwt_compat.so in
    front of the compiled reference so that the transforms its dwt_util_test2_* helpers call are the device ones"""
    exe = os.path.join(ROOT, "build", "examples", "test")
    assert os.path.exists(exe), f"{exe} missing: run `make examples` in the build container (it travels via gpurun)"
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=180)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-2000:]
    assert out.count("success") == 34 and "fail" not in out, out[-3000:]


@pytest.mark.parametrize("name", ["subbands", "subbands-int", "load", "load-int"])
def test_subbands_examples_write_the_reference_files(name, tmp_path):
    """examples/subbands, subbands-int (unmodified): forward transform, the LH subbands erased through dwt_util_subband_*, inverse, both
    images saved as PGM; examples/load, load-int (9/7 float / 9/7 integer lifting on the default image, coefficients shown through
    dwt_util_conv_show) -- the files must be the ones the same program writes with the compiled reference alone
    (tests/golden/examples_md5.json, made by tests/golden/make_examples_md5.py)"""
    import hashlib
    import json
    exe = os.path.join(ROOT, "build", "examples", name)
    assert os.path.exists(exe), f"{exe} missing: run `make examples` in the build container (it travels via gpurun)"
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "examples_md5.json")))[name]
    got = {f: hashlib.md5(open(tmp_path / f, "rb").read()).hexdigest() for f in want}
    assert got == want


def test_newapi_example_runs_on_the_gpu(tmp_path):
    """examples/simple-newapi (unmodified): the in-place family's forward / inverse pairs, two round trips compared by the reference"""
    exe = os.path.join(ROOT, "build", "examples", "simple-newapi")
    assert os.path.exists(exe), f"{exe} missing: run `make examples` in the build container (it travels via gpurun)"
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-2000:]
    assert out.count("success") == 2 and "images differs" not in out, out[-2000:]


def test_measure_perf_harness_writes_plot_data(tmp_path):
    """dwt_util_measure_perf_cdf97_2_s / _inplace_s of libdwt_compat.so (src/libdwt.c:22559, 22646): sizes min_x .. max_x growing
    by 1.13, one 'pixels <TAB> seconds' line per size in each plot file, timed on the device"""
    import ctypes as C
    L = C.CDLL(os.path.join(ROOT, "libdwt_b200", "libdwt_compat.so"))
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    for name, arr in (("dwt_util_measure_perf_cdf97_2_s", 2), ("dwt_util_measure_perf_cdf97_2_s", 1), ("dwt_util_measure_perf_cdf97_2_inplace_s", 2)):
        f = getattr(L, name)
        f.restype = None
        f.argtypes = [C.c_int] * 10 + [C.c_void_p, C.c_void_p]
        pf, pi = str(tmp_path / "fwd.txt").encode(), str(tmp_path / "inv.txt").encode()
        ff, fi = libc.fopen(pf, b"w"), libc.fopen(pi, b"w")
        assert ff and fi
        f(arr, 100, 200, 1, -1, 0, 0, 1, 2, 0, ff, fi)
        libc.fclose(ff)
        libc.fclose(fi)
        sizes = []
        x = 100
        while x <= 200:
            sizes.append(x * x)
            t = np.float32(x) * np.float32(1.13)
            x = int(np.ceil(t))
        for path in (pf, pi):
            lines = open(path).read().split("\n")[:-1]
            assert [int(l.split("\t")[0]) for l in lines] == sizes, (name, lines)
            assert all(0 < float(l.split("\t")[1]) < 0.01 for l in lines), (name, lines)
