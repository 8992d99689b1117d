"""CPU: bench.py's reference arm prints ONE JSON line with the contract's keys (tiny frames so it runs in seconds here);
the CUDA arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys
import pytest
from conftest import ROOT


def test_reference_arm_json_contract():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so")):
        pytest.skip("oracle/_ref not built here")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--width", "320", "--height", "240", "--frames-per-gpu", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]


def test_cuda_arm_has_no_cpu_fallback():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except Exception:
        pass
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--width", "320", "--height", "240",
                          "--frames-per-gpu", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0          # no device: the product path fails loudly
    assert not [l for l in out.stdout.splitlines() if l.strip().startswith("{") and "value" in l]
