"""bench.py's contract that can be checked without a GPU: the reference arm prints ONE JSON line with the keys the driver
reads (and runs only CPU code: gpu_launches 0, zero H2D / D2H bytes), and the product arm refuses to run without CUDA —
there is no CPU fallback to fall into."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")   # (default budget: one full 128-segment step fits)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train segments/sec" and d["unit"] == "segments/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["global_batch"] == 128 and d["config"]["reference_sample_per_step"] == 128   # the full batch, not a sample
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_arm_refuses_to_run_without_cuda():
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
