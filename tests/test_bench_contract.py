"""CPU: the driver-facing contract of bench.py that can be checked without a GPU.

  * `bench.py --impl reference` (the reference's own CPU path: stock torch.ao modules through oracle/vit_ref.py -- the one place
    besides tests/ and smoke() that may execute oracle/) prints ONE JSON line with the keys the driver parses, marked
    `"impl": "reference"`, with a `cpu_baseline` describing the run and an `e2e` object whose copies are zero;
  * the product arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    res = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] in baseline["metric"] and d["unit"] == "img/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert abs(d["value"] - 8 * 1e3 / d["ms_per_step"]) <= 1e-6 * d["value"]            # batch 8 per bounded step
    assert "distillation step" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "img/s" and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_arm_has_no_cpu_fallback():
    res = _run("--steps", "1", "--warmup", "1", timeout=300)
    assert res.returncode != 0
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]                 # no bench line from a CPU run
    assert "no CUDA device" in res.stderr and "no CPU fallback" in res.stderr
