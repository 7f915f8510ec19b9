"""CPU self-check of the property checks that tests/test_gpu_zz_fullsize.py applies at chi = 1024 / n = 2048: the same functions, small
sizes, with the oracle / NumPy standing in for the device, must report errors at rounding level (so a failure on the GPU box
means the kernels, not the test)."""
import numpy as np

import oracle
from test_gpu_zz_fullsize import heff_properties, svd_properties, D, W


def test_heff_property_checks_with_numpy_product():
    def make(L, R, M1, M2):
        return lambda th: np.einsum('awb,wstx,xvuy,btuc,eyc->asve', L, M1, M2, th, R, optimize=True)
    lin, adj, sub = heff_properties(make, 24, np.random.default_rng(0))
    assert lin < 1e-13 and adj < 1e-13 and sub < 1e-13, (lin, adj, sub)


def test_svd_property_checks_with_oracle_svd():
    def svd_fn(x, **kw):
        U, S, Vh = oracle.svd(x, 2, **kw)
        return U, np.real(np.diag(S)), Vh
    for kw in (dict(maxdim=40), dict(cutoff=1e-12), dict()):
        r = svd_properties(svd_fn, 96, np.random.default_rng(1), **kw)
        assert r["rank_ok"] and r["sorted_ok"], r
        assert r["sv_err"] < 1e-13 and r["orthU"] < 1e-12 and r["orthV"] < 1e-12 and r["recon"] < 1e-12, r
