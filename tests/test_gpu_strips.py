"""GPU: ONE image as row strips over several ranks through the C ABI (dwtb200_strips_*, csrc/strips.cu).

* every rank emulated in one process on ONE GPU: the complete protocol (halo pulls, LL gather into rank 0's top image, device-side
  sequence flags, two collective rounds) with plain pointers in place of the CUDA IPC mappings;
* one process per GPU with CUDA IPC peer mappings over NVLink: needs >= 2 devices, skipped on a single-GPU box.
Each run compares the rows every rank owns bit for bit with the single-device transform of the whole picture and with the oracle
(tests/strips_worker.py)."""
import json
import os
import subprocess
import sys
import uuid

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = os.path.join(HERE, "strips_worker.py")


def run_worker(args, env_extra=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, WORKER] + [str(a) for a in args], env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("cfg", [("97s", 2048, 4096, 2, 0), ("53i", 1000, 3000, 3, 2), ("97d", 1024, 2048, 2, 1), ("97s", 4100, 8200, 4, 3),
                                 ("53s", 2048, 2048, 1, 2)],
                         ids=lambda c: f"{c[0]}-{c[1]}x{c[2]}-w{c[3]}-Jd{c[4]}")
def test_strips_all_ranks_on_one_gpu(dev, tmp_path, cfg):
    kind, W, H, G, Jd = cfg
    out = tmp_path / "res.json"
    p = run_worker(["--mode", "emulate", "--world", G, "--width", W, "--height", H, "--levels", Jd, "--kind", kind, "--oracle", "--session",
                    "dwtb200-t-" + uuid.uuid4().hex[:12], "--out", out],
                   {"CUDA_DEVICE_MAX_CONNECTIONS": "32"})
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    r = json.load(open(out))
    assert r["ok"] and len(r["rounds"]) == 2
    for rnd in r["rounds"]:
        assert not any(rnd["fwd_diff"]) and not any(rnd["inv_diff"]) and rnd["fwd_oracle_equal"] and rnd["inv_oracle_equal"]
        if G > 1:
            assert all(b > 0 for b in rnd["fwd_peer_bytes"]) and all(b > 0 for b in rnd["inv_peer_bytes"])


@pytest.mark.parametrize("cfg", [("97s", 4096, 4096, 0, True), ("53i", 4096, 4096, 2, True), ("97s", 16384, 16384, 0, False)],
                         ids=lambda c: f"{c[0]}-{c[1]}x{c[2]}-Jd{c[3]}")
def test_strips_one_process_per_gpu(dev, tmp_path, cfg):
    n = dev.lib().c.dwtb200_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (CUDA IPC peer mappings over NVLink)")
    kind, W, H, Jd, with_oracle = cfg
    G = 8 if n >= 8 else 4 if n >= 4 else 2
    session = "dwtb200-t-" + uuid.uuid4().hex[:12]
    out = tmp_path / "res.json"
    procs = []
    for r in range(G):
        env = dict(os.environ, LOCAL_RANK=str(r))
        a = [sys.executable, WORKER, "--mode", "rank", "--rank", str(r), "--world", str(G), "--width", str(W), "--height", str(H), "--levels",
             str(Jd), "--kind", kind, "--session", session, "--out", str(out)] + (["--oracle"] if with_oracle else [])
        procs.append(subprocess.Popen(a, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r}: " + outs[r][-3000:]
        res = json.load(open(f"{out}.{r}"))
        assert res["ok"], res
