"""CPU: the C-ABI library loads, exports every symbol include/ertext.h declares, and refuses to
work without a CUDA device (no CPU fallback)."""
import os
import re
import pytest
from conftest import ROOT


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "ertext.h")).read()
    return sorted(set(re.findall(r"ERT_API\s+[\w\s\*]+?\b(ert_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import ertext
    L = ertext.load_library()
    names = _header_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(names) == sorted(ertext.EXPORTS)
    assert L.ert_abi_version() == 2


def test_no_cpu_fallback():
    import ertext
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(ertext.ErtError) as ei:
        ertext.ErText()
    assert "no CPU path" in str(ei.value)


def test_product_library_does_not_link_the_oracle():
    import subprocess
    import ertext
    out = subprocess.run(["ldd", ertext.LIB_PATH], capture_output=True, text=True).stdout
    assert "libert_port" not in out and "libref_oracle" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", ertext.LIB_PATH], capture_output=True, text=True).stdout
    assert "port_" not in syms and "ref_tree" not in syms
