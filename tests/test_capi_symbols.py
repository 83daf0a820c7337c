"""CPU-side checks of the C-ABI boundary: the library loads, and exports exactly what include/rk_b200.h declares."""
import os
import re

import pytest

from repkiller_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "rk_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rk_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = capi.load_library()
    for name in _declared():
        assert hasattr(L, name), name
    assert b"sm_100a" in L.rk_version()


def test_no_cpu_fallback():
    """Without a CUDA device the context must refuse to exist (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.RkError) as e:
        capi.Context(0)
    assert "no CUDA device" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "repkiller_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("# oracle-free", ""), f"{f} mentions the oracle"
