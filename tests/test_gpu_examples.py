"""GPU: the reference's UNMODIFIED example programs (examples/simple, simple-int, simple-double), linked
against libdwt_compat.so ahead of the compiled reference (Makefile target `examples`), must print
"success": they allocate with dwt_util_alloc_image, use a prime row stride (2053 bytes for 512 floats),
call dwt_cdfXX_2f_* / 2i_* and compare the round trip with the reference's own dwt_util_compare_*.
simple-single-loop is the in-place family's example (dwt_cdf97_2f_inplace_sdl_s / dwt_cdf97_2i_inplace_s, decompose_one = 1)."""
import os
import subprocess

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
