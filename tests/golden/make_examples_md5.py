"""Golden digests of the files the reference's `subbands` / `subbands-int` examples write (data1.pgm: the test image, data2.pgm: the
reconstruction after the LH subbands were erased) and of those of `load` / `load-int` (no Lenna.pgm in the working directory: they fall
back to the test image; data1 = round trip, data2 = original, data3 = coefficients through dwt_util_conv_show), produced by the UNMODIFIED example linked against the compiled reference alone
(oracle/_ref/libdwt_ref.so, no B200 library).  Build container only: needs /root/reference.  python tests/golden/make_examples_md5.py"""
import hashlib
import json
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("REF", "/root/reference")
EXAMPLES = (("subbands", "subbands.c"), ("subbands-int", "subbands.c"), ("load", "simple.c"), ("load-int", "simple.c"))


def digests():
    out = {}
    for name, src in EXAMPLES:
        with tempfile.TemporaryDirectory() as tmp:
            exe = os.path.join(tmp, name)
            subprocess.run(["gcc", "-std=c99", "-O2", "-D_POSIX_C_SOURCE=199309L", "-D_GNU_SOURCE", f"-I{REF}/src", f"{REF}/examples/{name}/{src}", "-o", exe,
                            f"-L{ROOT}/oracle/_ref", "-l:libdwt_ref.so", "-lm", "-lrt", "-fopenmp", f"-Wl,-rpath,{ROOT}/oracle/_ref"], check=True,
                           capture_output=True)
            subprocess.run([exe], cwd=tmp, check=True, capture_output=True)
            out[name] = {f: hashlib.md5(open(os.path.join(tmp, f), "rb").read()).hexdigest() for f in sorted(os.listdir(tmp)) if f.endswith(".pgm")}
    return out


if __name__ == "__main__":
    out = digests()
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "examples_md5.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))
