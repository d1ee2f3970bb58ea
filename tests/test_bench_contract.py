"""bench.py's command-line contract, as far as it can be checked without a GPU: the reference arm prints one
JSON line with the agreed keys (the reference's own CPU decoder timed on the host cores), non-zero ranks of a
torchrun launch stay silent, and the GPU arm refuses to run (no CPU fallback) instead of printing a number."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
HAVE_DATA = os.path.exists(os.path.join(ROOT, "bench_data", "gov2_r256", "meta.json"))


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=e)


@pytest.mark.skipif(not HAVE_DATA, reason="bench_data/gov2_r256 not present")
def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "3"])
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "G ints/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["vs_baseline"] is None and line["dtype"] == "u32"


@pytest.mark.skipif(not HAVE_DATA, reason="bench_data/gov2_r256 not present")
def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--steps", "1"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not any(l.startswith("{") for l in r.stdout.splitlines())
