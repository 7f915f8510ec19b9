"""CPU-side boundary tests: the C-ABI library loads, exports every symbol include/tn_c_api.h
declares, and fails loudly (no CPU fallback) when no GPU is present."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tn_c_api.h")).read()
    return sorted(set(re.findall(r"\b(tn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import tnb200
    from tnb200 import _lib
    lib = tnb200.load()
    syms = _declared_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"missing export {s}"
    # every prototype the Python mirror binds is declared in the header
    for s in _lib.PROTOTYPES:
        assert s in syms
    assert lib.tn_version() >= 100


def test_no_cpu_fallback():
    import torch
    import tnb200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tnb200.TNError):
        tnb200.Context(0)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "tensornetworks.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
