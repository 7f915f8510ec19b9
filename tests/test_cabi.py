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


def test_null_handles_are_errors_not_crashes():
    """Every entry point that takes a handle rejects NULL with TN_ERR_INVALID and a message (no GPU needed: the check comes first)."""
    import ctypes as C
    import tnb200
    from tnb200 import _lib
    lib = tnb200.load()
    v = _lib.tn_cplx()
    k = C.c_int64()
    tr = _lib.tn_trunc_t(0.0, 0, 1)
    calls = [lambda: lib.tn_mps_norm(None, C.byref(v)), lambda: lib.tn_mps_normalize(None), lambda: lib.tn_mps_movecenter(None, 1, tr),
             lambda: lib.tn_env_movecenter(None, 1), lambda: lib.tn_env_calculate(None, C.byref(v)), lambda: lib.tn_envsum_movecenter(None, 1),
             lambda: lib.tn_apply_gates(None, None, tr), lambda: lib.tn_mps_maxbonddim(None, C.byref(k)), lambda: lib.tn_sync(None),
             lambda: lib.tn_mpo_compress(None, tr), lambda: lib.tn_heff_sharded_apply(None, None, None),
             lambda: lib.tn_svd_trunc_split(None, None, 4, 4, tr, 1, None, None, None, C.byref(k), None, 1, None)]
    for f in calls:
        assert f() == -1
        assert b"null" in lib.tn_last_error()
    # the free functions accept NULL like free()
    assert lib.tn_mps_free(None) == 0 and lib.tn_env_free(None) == 0 and lib.tn_gates_free(None) == 0 and lib.tn_envsum_free(None) == 0
    assert lib.tn_heff_sharded_free(None) == 0
    # the sharded matvec validates its device list before touching CUDA
    h = C.c_void_p()
    assert lib.tn_heff_sharded_create(0, None, 4, 4, 2, 3, 3, 3, None, None, None, None, _lib.tn_cplx(1.0, 0.0), C.byref(h)) != 0
