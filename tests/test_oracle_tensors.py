"""Tier-1 oracle self-tests: every contraction against a one-shot einsum, the
truncation rule on hand-made spectra, SVD split invariants."""
import numpy as np
import pytest

import oracle
from oracle.tensors import truncation_rank
from oracle.gmps import GMPS


def crandn(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def test_contract_output_order():
    rng = np.random.default_rng(0)
    x, y = crandn(rng, 3, 4, 5), crandn(rng, 5, 6, 4)
    z = oracle.contract(x, y, [2, 3], [3, 1])
    assert np.allclose(z, np.einsum('abc,cdb->ad', x, y))
    z = oracle.contract(x, y, 3, 1, True, False)
    assert np.allclose(z, np.einsum('abc,cde->abde', x.conj(), y))


def test_combine_first_index_fastest():
    rng = np.random.default_rng(1)
    x = crandn(rng, 2, 3, 4, 5)
    y, cmb = oracle.combineidxs(x, [2, 4])
    assert y.shape == (2, 4, 15)
    assert y[1, 2, 1 + 3 * 4] == x[1, 1, 2, 4]
    assert np.array_equal(oracle.uncombineidxs(y, cmb), x)


def test_moveidx():
    rng = np.random.default_rng(2)
    x = crandn(rng, 2, 3, 4, 5)
    assert oracle.moveidx(x, 2, -1).shape == (2, 4, 5, 3)
    assert np.array_equal(oracle.moveidx(x, 4, 2), np.einsum('abcd->adbc', x))


@pytest.mark.parametrize("S,kw,expect", [
    ([1.0, 0.5, 0.1, 0.0], dict(), 4),                       # zeros are kept (tensors.jl:205 is a no-op)
    ([1.0, 0.5, 0.1, 0.0], dict(cutoff=1e-12), 3),           # cutoff removes the exact zero
    ([1.0, 0.5, 0.1, 0.01], dict(maxdim=2), 2),
    ([1.0, 0.5, 0.1, 0.01], dict(maxdim=10), 4),
    ([1.0, 0.5, 0.1, 0.01], dict(cutoff=0.5, mindim=3), 3),  # mindim wins over cutoff
    ([1.0, 1e-9], dict(cutoff=1e-12), 2),                     # 1e-18 > 1e-12 ? no -> tail 1e-18 < cutoff -> keep 1
    ([1.0], dict(cutoff=0.999999), 1),
    ([3.0, 4.0], dict(cutoff=2.0), 1),                        # nothing above cutoff -> 1
])
def test_truncation_rule(S, kw, expect):
    if S == [1.0, 1e-9]:
        expect = 1
    assert truncation_rank(np.array(S), **kw) == expect


def test_svd_layout_and_reconstruction():
    rng = np.random.default_rng(3)
    x = crandn(rng, 3, 2, 4)
    for idx in (1, 2, 3, -1):
        U, S, V = oracle.svd(x, idx)
        i = 3 if idx == -1 else idx
        k = S.shape[0]
        assert U.shape[i - 1] == k and V.shape == (k, x.shape[i - 1])
        rec = np.moveaxis(np.tensordot(U, S @ V, axes=([i - 1], [0])), -1, i - 1)
        assert np.allclose(rec, x)


def test_replacesites_orthonormal_and_discarded_weight():
    rng = np.random.default_rng(4)
    chi, d = 6, 2
    psi = GMPS(1, d, [crandn(rng, 1, d, chi), crandn(rng, chi, d, chi), crandn(rng, chi, d, chi), crandn(rng, chi, d, 1)], 0)
    psi.movecenter(2)
    theta = crandn(rng, chi if False else psi[2].shape[0], d, d, psi[3].shape[2])
    full = np.linalg.svd(theta.reshape(-1, d * theta.shape[3], order='F'), compute_uv=False)
    for direction in (False, True):
        p = psi.copy()
        p.replacesites(theta, 2, direction, False, maxdim=3)
        A, B = p[2], p[3]
        assert A.shape[2] == 3 == B.shape[0]
        if not direction:   # left-orthonormal site, centre on 3
            assert p.center == 3
            m = A.reshape(-1, 3, order='F')
            assert np.allclose(m.conj().T @ m, np.eye(3))
        else:
            assert p.center == 2
            m = B.reshape(3, -1, order='F')
            assert np.allclose(m @ m.conj().T, np.eye(3))
        rec = np.tensordot(A, B, axes=([2], [0]))
        err2 = np.linalg.norm(rec - theta) ** 2
        assert np.isclose(err2, np.sum(full[3:] ** 2))


def test_gate_layout():
    rng = np.random.default_rng(5)
    A, B = crandn(rng, 3, 2, 4), crandn(rng, 4, 2, 5)
    G = crandn(rng, 2, 2, 2, 2)
    psi = GMPS(1, 2, [crandn(rng, 1, 2, 3), A, B, crandn(rng, 5, 2, 1)], 0)
    psi.center = 2
    oracle.applygate(psi, 2, G, False, error=False)
    got = np.tensordot(psi[2], psi[3], axes=([2], [0]))
    want = np.einsum('lab,bcr,xayc->lxyr', A, B, G)
    assert np.allclose(got, want)


def _rand_env(rng, chi=5, w=3, d=2, N=6):
    psi = oracle.randomMPS(d, N, chi, rng)
    psi = GMPS(1, d, [t + 1j * rng.standard_normal(t.shape) * 0.3 for t in psi.tensors], 0)
    psi.movecenter(1)
    W = [crandn(rng, 1 if i == 0 else w, d, d, 1 if i == N - 1 else w) for i in range(N)]
    H = GMPS(2, d, W, 0)
    return psi, H


def test_build_blocks_and_product_vs_einsum():
    rng = np.random.default_rng(6)
    psi, H = _rand_env(rng)
    P = oracle.ProjMPS([psi, H, psi], rank=2, center=3)
    L, R = P.block(2), P.block(5)
    # independent einsum chain for L(2)
    e = np.ones((1, 1, 1), complex)
    for s in (1, 2):
        e = np.einsum('awb,asc,wstx,btd->cxd', e, psi[s].conj(), H[s], psi[s])
    assert np.allclose(L, e)
    e = np.ones((1, 1, 1), complex)
    for s in (6, 5):
        e = np.einsum('asc,wstx,btd,cxd->awb', psi[s].conj(), H[s], psi[s], e)
    assert np.allclose(R, e)
    theta = crandn(rng, psi[3].shape[0], 2, 2, psi[4].shape[2])
    want = np.einsum('awb,wstx,xuvy,btvc,eyc->asue', L, H[3], H[4], theta, R)
    assert np.allclose(P.product(theta, False, 2), want)
    assert np.allclose(P.product_optimal(theta, False), want)
    P.movecenter(4)
    assert np.allclose(P.product(theta, True, 2), want)   # left sweep addresses the same two sites
    # calculate == <psi|H|psi>
    val = P.calculate()
    e = np.ones((1, 1, 1), complex)
    for s in range(1, 7):
        e = np.einsum('awb,asc,wstx,btd->cxd', e, psi[s].conj(), H[s], psi[s])
    assert np.isclose(val, e[0, 0, 0])


def test_heff_is_hermitian_for_hermitian_mpo():
    sh = oracle.spinhalf()
    from models import tfim
    H = oracle.MPO(sh, tfim(6))
    psi = oracle.randomMPS(2, 6, 4, np.random.default_rng(7))
    P = oracle.ProjMPS([psi, H, psi], rank=2, center=3)
    rng = np.random.default_rng(8)
    x, y = crandn(rng, *[psi[3].shape[0], 2, 2, psi[4].shape[2]]), crandn(rng, *[psi[3].shape[0], 2, 2, psi[4].shape[2]])
    assert np.isclose(np.vdot(x, P.product(y)), np.vdot(P.product(x), y))


def test_lanczos_small_matrix():
    rng = np.random.default_rng(9)
    A = crandn(rng, 30, 30)
    A = A + A.conj().T
    x0 = crandn(rng, 30)
    th, x, info = oracle.eigsolve_lowest(lambda v: A @ v, x0, krylovdim=3, maxiter=2)
    assert info["numops"] == 5
    assert np.isclose(np.linalg.norm(x), 1.0)
    assert np.isclose(th, np.real(np.vdot(x, A @ x)))
    th2, x2, info2 = oracle.eigsolve_lowest(lambda v: A @ v, x0, krylovdim=30, maxiter=1)
    assert np.isclose(th2, np.linalg.eigvalsh(A)[0])
