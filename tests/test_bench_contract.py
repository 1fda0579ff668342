"""CPU test of bench.py's contract through the reference arm (`--impl reference` runs the
reference's own CPU step, so it works without a GPU): one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

from conftest import ROOT

REQUIRED = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"]


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600,
                       env=e)
    return r


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "cfg2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "interactions/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] > 1e6
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_gpu():
    import ctypes as C
    import traceit_b200 as tr
    n = C.c_int(0)
    if tr.lib().tr_device_count(C.byref(n)) == 0 and n.value > 0:
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "0", "--workload", "cfg2"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_library_chatter_on_stdout_is_diverted():
    """NCCL prints its version banner on stdout under NCCL_DEBUG=VERSION: the JSON line must stay alone there."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.protect_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('more chatter'); bench.emit({'a': 1})") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout == '{"a": 1}\n'
    assert "NCCL version" in r.stderr and "more chatter" in r.stderr
