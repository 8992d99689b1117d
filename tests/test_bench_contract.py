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


def test_committed_bench_lines_carry_the_contract_keys():
    """The bench lines of record (profiles/, produced on the B200 by the committed bench.py) carry every key of the bench
    contract -- a guard against the JSON drifting from what DESIGN.md 7 and profiles/README.md describe."""
    p = os.path.join(ROOT, "profiles", "r03u_bench_n1.json")
    d = json.load(open(p))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "sustained", "parity"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 8 * 1920 * 1080 * 3 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= 1.05 * d["value"]
    assert "h2d_ceiling" in e and "from_jpeg" in e
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["algorithmic_bytes_per_launch"] == 48 * 1920 * 1080 and r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c and "mode_R" in c
    assert d["gpu_launches"] > 0 and d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] > 0
    par = d["parity"]
    assert par["nodes_equal"] is True and par["pool_symdiff_vs_reference"] == 0 and par["label_mismatch"] == 0
    # same workload wording in both arms (the driver compares them)
    ref = [json.loads(l) for l in open(os.path.join(ROOT, "profiles", "r01n_bench_ref.json")) if l.strip().startswith("{")]
    assert ref and ref[0]["metric"] == d["metric"] and ref[0]["unit"] == d["unit"]
