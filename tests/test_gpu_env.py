"""GPU parity: environment updates (K6), H_eff matvec (K1/K2), calculate, Lanczos (K4), gauge moves
and replacesites! (K5/K7) vs the oracle.  Contractions: relative Frobenius <= 1e-13."""
import numpy as np
import pytest

import oracle
from gpu_util import crandn, random_complex_mps, random_mpo, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,chi,w", [(6, 5, 3), (8, 16, 5), (5, 40, 7)])
def test_blocks_product_calculate(N, chi, w):
    import tnb200
    rng = np.random.default_rng(N * 100 + chi)
    psi = random_complex_mps(rng, N, 2, chi, center=1)
    H = random_mpo(rng, N, 2, w)
    c = 3
    P = oracle.ProjMPS([psi, H, psi], rank=2, center=c, coeff=0.7)
    gpsi, gH = tnb200.GMPS.from_host(psi), tnb200.GMPS.from_host(H)
    G = tnb200.ProjMPS(gpsi, gH, gpsi, coeff=0.7, center=c)
    for i in range(1, N + 1):
        if i != c:
            assert relerr(G.block(i), P.block(i)) < 1e-13
    theta = crandn(rng, psi[c].shape[0], 2, 2, psi[c + 1].shape[2])
    assert relerr(G.product(theta, False), P.product(theta, False, 2)) < 1e-13
    assert abs(G.calculate() - P.calculate()) < 1e-12 * abs(P.calculate())
    G.movecenter(c + 1)
    P.movecenter(c + 1)
    assert relerr(G.block(c), P.block(c)) < 1e-13
    assert relerr(G.product(theta, True), P.product(theta, True, 2)) < 1e-13
    G.movecenter(2)
    P.movecenter(2)
    assert relerr(G.block(3), P.block(3)) < 1e-13


def test_eigsolve_matches_oracle_schedule():
    import tnb200
    from models import tfim
    sh = oracle.spinhalf()
    H = oracle.MPO(sh, tfim(8))
    psi = random_complex_mps(np.random.default_rng(3), 8, 2, 8, center=4)
    P = oracle.ProjMPS([psi, H, psi], rank=2, center=4)
    gpsi, gH = tnb200.GMPS.from_host(psi), tnb200.GMPS.from_host(H)
    G = tnb200.ProjMPS(gpsi, gH, gpsi, center=4)
    A0 = np.tensordot(psi[4], psi[5], axes=([2], [0]))
    e0, v0, info = oracle.eigsolve_lowest(lambda x: P.product(x, False, 2), A0)
    e1, v1, nops = G.eigsolve(A0, False)
    assert nops == info["numops"] == 5
    assert abs(e1 - e0) < 1e-11 * abs(e0)
    assert abs(abs(np.vdot(v0, v1)) - 1.0) < 1e-9


def test_movecenter_norm_replacesites():
    import tnb200
    rng = np.random.default_rng(4)
    N, chi = 7, 12
    psi = random_complex_mps(rng, N, 2, chi, center=1)
    psi[1] = psi[1] * 1.7
    g = tnb200.GMPS.from_host(psi)
    assert abs(g.norm() - psi.norm()) < 1e-12
    g.movecenter(5)
    psi.movecenter(5)
    assert g.center == 5
    assert abs(g.norm() - psi.norm()) < 1e-12
    for i in range(1, N):
        assert g.bonddim(i) == psi.bonddim(i)
    # gauge-invariant check: same state
    def dense(p):
        v = p[1]
        for i in range(2, N + 1):
            v = np.tensordot(v, p[i], axes=([v.ndim - 1], [0]))
        return v.reshape(-1)
    assert relerr(dense(g), dense(psi)) < 1e-12
    # left-orthonormality of the sites left of the centre
    for i in range(1, 5):
        m = g[i].reshape(-1, g[i].shape[2], order='F')
        assert np.linalg.norm(m.conj().T @ m - np.eye(m.shape[1])) < 1e-11
    g.movecenter(2, maxdim=5)
    psi.movecenter(2, maxdim=5)
    assert [g.bonddim(i) for i in range(1, N)] == [psi.bonddim(i) for i in range(1, N)]
    assert np.max(np.abs(g.spectrum(3) - psi.spectrum(3))) < 1e-10
    theta = crandn(rng, g[3].shape[0], 2, 2, g[4].shape[2])
    for direction in (False, True):
        g.replacesites(theta, 3, direction, True, maxdim=6, cutoff=1e-14)
        psi.replacesites(theta, 3, direction, True, maxdim=6, cutoff=1e-14)
        assert g.center == psi.center
        assert relerr(np.tensordot(g[3], g[4], axes=([2], [0])), np.tensordot(psi[3], psi[4], axes=([2], [0]))) < 1e-11
        g.normalize()
        assert abs(g.norm() - 1.0) < 1e-13


def test_gauge_moves_that_cannot_truncate_keep_the_state():
    """movecenter! without truncation (cutoff = 0, maxdim >= bond): the library replaces the SVD of moveleft!/moveright!
    (gmps.jl:60-82) by one QR step (or by nothing when the isometry sits on the short side).  Gauge-invariant parity with
    the oracle's SVD-based moves: same state, same bond dimensions, orthonormal sites, same norm; bonds > 64 take two QR panels."""
    import tnb200
    rng = np.random.default_rng(11)
    N, chi = 14, 80
    psi = random_complex_mps(rng, N, 2, chi, center=1)
    g = tnb200.GMPS.from_host(psi)

    def dense(p):
        v = p[1]
        for i in range(2, N + 1):
            v = np.tensordot(v, p[i], axes=([v.ndim - 1], [0]))
        return v.reshape(-1)
    want = dense(psi)
    for c, kw in ((N, {}), (3, {}), (9, dict(maxdim=chi)), (1, dict(maxdim=200)), (7, {})):
        g.movecenter(c, **kw)
        psi.movecenter(c, **kw)
        assert g.center == c
        assert [g.bonddim(i) for i in range(1, N)] == [psi.bonddim(i) for i in range(1, N)]
        assert relerr(dense(g), want) < 1e-12
        assert abs(g.norm() - psi.norm()) < 1e-12
        for i in range(1, c):
            m = g[i].reshape(-1, g[i].shape[2], order='F')
            assert np.linalg.norm(m.conj().T @ m - np.eye(m.shape[1])) < 1e-11
        for i in range(c + 1, N + 1):
            m = g[i].reshape(g[i].shape[0], -1, order='F')
            assert np.linalg.norm(m @ m.conj().T - np.eye(m.shape[0])) < 1e-11
    # a truncating move still goes through the SVD and matches the oracle's spectrum
    g.movecenter(2, maxdim=20)
    psi.movecenter(2, maxdim=20)
    assert [g.bonddim(i) for i in range(1, N)] == [psi.bonddim(i) for i in range(1, N)]
    assert np.max(np.abs(g.spectrum(5) - psi.spectrum(5))) < 1e-10
