"""GPU parity: Jacobi truncated SVD (K5) vs the oracle's LAPACK zgesdd path (tensors.jl:168-227).
Tolerances: singular values <= 1e-12 relative to sigma_max, reconstruction <= 1e-12, identical
truncation rank (SURVEY.md section 4 tier 4)."""
import numpy as np
import pytest

import oracle
from gpu_util import crandn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n", [(1, 1), (2, 8), (8, 2), (64, 64), (96, 40), (40, 96), (130, 130), (257, 190), (512, 512)])
def test_full_svd_matches_lapack(m, n):
    import tnb200
    rng = np.random.default_rng(m * 1000 + n)
    x = crandn(rng, m, n)
    U, S, V = tnb200.svd(x, 2)
    Uo, So, Vo = oracle.svd(x, 2)
    s, so = np.real(np.diag(S)), np.real(np.diag(So))
    assert s.shape == so.shape
    assert np.max(np.abs(s - so)) <= 1e-12 * so[0]
    assert np.linalg.norm(U @ S @ V - x) <= 1e-12 * np.linalg.norm(x)
    k = len(s)
    assert np.linalg.norm(U.conj().T @ U - np.eye(k)) < 1e-11
    assert np.linalg.norm(V @ V.conj().T - np.eye(k)) < 1e-11


def test_truncation_rank_matches_reference_rule():
    import tnb200
    rng = np.random.default_rng(5)
    m = 96
    u, _ = np.linalg.qr(crandn(rng, m, m))
    v, _ = np.linalg.qr(crandn(rng, m, m))
    sv = np.exp(-0.6 * np.arange(m))
    x = (u * sv) @ v.conj().T
    for kw in (dict(cutoff=1e-12), dict(maxdim=10), dict(cutoff=1e-6, maxdim=50), dict(cutoff=1e-3, mindim=20), dict()):
        U, S, V = tnb200.svd(x, 2, **kw)
        Uo, So, Vo = oracle.svd(x, 2, **kw)
        assert S.shape == So.shape, kw
        s, so = np.real(np.diag(S)), np.real(np.diag(So))
        assert np.max(np.abs(s - so)) <= 1e-12 * so[0]


def test_rank_deficient_and_tensor_index():
    import tnb200
    rng = np.random.default_rng(6)
    a, b = crandn(rng, 80, 30), crandn(rng, 30, 90)
    x = a @ b                                  # rank 30
    U, S, V = tnb200.svd(x, 2, cutoff=1e-20)
    Uo, So, Vo = oracle.svd(x, 2, cutoff=1e-20)
    assert S.shape[0] >= 30 and abs(S.shape[0] - So.shape[0]) <= 2   # eps-level tail: rank may differ by rounding
    assert np.linalg.norm(U @ S @ V - x) <= 1e-12 * np.linalg.norm(x)
    t = crandn(rng, 6, 2, 7)                   # svd on a middle index: the new bond replaces it
    for idx in (1, 2, 3):
        U, S, V = tnb200.svd(t, idx)
        Uo, So, Vo = oracle.svd(t, idx)
        assert U.shape == Uo.shape and V.shape == Vo.shape
        rec = np.moveaxis(np.tensordot(U, S @ V, axes=([idx - 1], [0])), -1, idx - 1)
        assert np.linalg.norm(rec - t) <= 1e-12 * np.linalg.norm(t)


def test_preconditioned_and_plain_jacobi_agree_on_graded_matrix():
    """The two-step QR preconditioner only changes the iteration count: same singular values (1e-12 sigma_max)
    with far fewer sweeps on an MPS-like matrix whose spectrum spans 13 decades."""
    import tnb200
    rng = np.random.default_rng(9)
    n = 192
    u, _ = np.linalg.qr(crandn(rng, n, n))
    v, _ = np.linalg.qr(crandn(rng, n, n))
    x = (u * np.exp(-np.arange(n) * 30.0 / n)) @ v.conj().T
    so = np.linalg.svd(x, compute_uv=False)
    lib = tnb200.load()
    out = {}
    try:
        for mode in (1, 0):
            lib.tn_svd_set_precond(mode)
            U, S, V, sw = tnb200.svd(x, 2, return_sweeps=True)
            s = np.real(np.diag(S))
            assert np.max(np.abs(s - so)) <= 1e-12 * so[0]
            assert np.linalg.norm(U @ S @ V - x) <= 1e-12 * np.linalg.norm(x)
            out[mode] = sw
    finally:
        lib.tn_svd_set_precond(1)
    assert out[1] <= 10 and out[1] < out[0]


@pytest.mark.parametrize("m,n", [(300, 130), (130, 300), (1024, 1024)])
def test_preconditioned_shapes(m, n):
    import tnb200
    rng = np.random.default_rng(m + n)
    x = crandn(rng, m, n) * np.exp(-np.arange(n) * 12.0 / n)[None, :]
    for kw in (dict(), dict(cutoff=1e-12, maxdim=100)):
        U, S, V = tnb200.svd(x, 2, **kw)
        Uo, So, Vo = oracle.svd(x, 2, **kw)
        assert S.shape == So.shape
        s, so = np.real(np.diag(S)), np.real(np.diag(So))
        assert np.max(np.abs(s - so)) <= 1e-12 * so[0]
        k = len(s)
        assert np.linalg.norm(U.conj().T @ U - np.eye(k)) < 1e-10
        assert np.linalg.norm(V @ V.conj().T - np.eye(k)) < 1e-10
        if not kw:
            assert np.linalg.norm(U @ S @ V - x) <= 1e-12 * np.linalg.norm(x)


@pytest.mark.parametrize("m,n", [(96, 96), (130, 130), (200, 80), (80, 200), (256, 256), (512, 256), (256, 512), (512, 512)])
@pytest.mark.parametrize("side", [1, 2])
def test_split_factorisation_matches_lapack(m, n, side):
    """The isometry + weighted-factor split of replacesites! / moveleft! / moveright! (gmps.jl:60-82, 218-256); when the isometry sits
    on the long side the W-only factorisation runs (no accumulated right rotations, weighted factor = U'^H R1).  Graded spectrum
    as on an MPS bond, with and without truncation."""
    import tnb200
    rng = np.random.default_rng(7 * m + n + side)
    r = min(m, n)
    u, _ = np.linalg.qr(crandn(rng, m, r))
    v, _ = np.linalg.qr(crandn(rng, n, r))
    sv = np.exp(-np.arange(r) * 24.0 / r)
    x = (u * sv) @ v.conj().T
    for kw in (dict(cutoff=0.0), dict(cutoff=1e-12), dict(cutoff=0.0, maxdim=r // 2)):
        A, S, B, sweeps, ms = tnb200.svd_split(x, side, **kw)
        Uo, So, Vo = oracle.svd(x, 2, **kw)
        so = np.real(np.diag(So))
        k = len(S)
        assert k == len(so), (kw, k, len(so))
        assert np.max(np.abs(S - so)) <= 1e-12 * so[0]
        iso = A if side == 1 else B.conj().T
        kk = int(np.sum(so > 1e-10 * so[0]))                  # singular vectors of sigma ~ eps sigma_max are not determined
        assert np.linalg.norm(iso[:, :kk].conj().T @ iso[:, :kk] - np.eye(kk)) < 1e-11 * np.sqrt(kk)
        want = (Uo @ So) @ Vo
        assert np.linalg.norm(A @ B - want) <= 1e-12 * so[0] * np.sqrt(k)
        # the weighted factor carries the singular values: its Gram matrix is diag(S^2)
        wf = B if side == 1 else A.conj().T
        assert np.linalg.norm(wf @ wf.conj().T - np.diag(S ** 2)) <= 1e-11 * so[0] ** 2 * np.sqrt(k)


@pytest.mark.parametrize("rows,ncols,npairs,with_skip", [(64, 64, 1, False), (200, 128, 2, False), (512, 512, 8, True), (1030, 256, 3, True),
                                                          (2048, 2048, 32, False)])
def test_jacobi_pair_kernels_match_numpy(rows, ncols, npairs, with_skip):
    """The Gram-block and in-place rotation kernels of the Jacobi step (csrc/tn_jacobi.cu) on their own: G_p = P_p^H P_p (full
    Hermitian block: only the upper 8x8 tiles are computed, the mirror image is written by the epilogue) and Z(:, pair) <- Z(:, pair) J_p
    over a multi-row-block cp.async ring, for row counts that are / are not multiples of the 64-row block and the 16-row k-tile."""
    import ctypes as C
    import tnb200
    from tnb200.api import _ptr, _f, check
    rng = np.random.default_rng(rows + ncols)
    ctx = tnb200.Context.default()
    Z = _f(crandn(rng, rows, ncols))
    nb = ncols // 32
    blocks = rng.permutation(nb)[:2 * npairs].astype(np.int32)
    pairs = np.ascontiguousarray(blocks.reshape(npairs, 2))
    J = np.stack([_f(crandn(rng, 64, 64)) for _ in range(npairs)])          # any matrix: the kernel is a plain product
    Jflat = np.concatenate([j.reshape(-1, order='F') for j in J])
    skip = (rng.integers(0, 2, npairs).astype(np.int32) if with_skip else None)
    G = np.zeros(npairs * 64 * 64, dtype=np.complex128)
    Zout = np.zeros((rows, ncols), dtype=np.complex128, order='F')
    check(ctx.lib.tn_jacobi_pair_pass(ctx.h, _ptr(Z), rows, ncols, pairs.ctypes.data_as(C.c_void_p), npairs, _ptr(Jflat),
                                      skip.ctypes.data_as(C.c_void_p) if skip is not None else None, _ptr(G), _ptr(Zout)))
    want = Z.copy()
    for p in range(npairs):
        cols = np.concatenate([np.arange(32) + 32 * pairs[p, 0], np.arange(32) + 32 * pairs[p, 1]])
        P = Z[:, cols]
        Gp = np.reshape(G[p * 4096:(p + 1) * 4096], (64, 64), order='F')
        ref = P.conj().T @ P
        assert np.linalg.norm(Gp - ref) <= 1e-13 * np.linalg.norm(ref), (p, "gram")
        assert np.linalg.norm(Gp - Gp.conj().T) == 0.0 or np.linalg.norm(Gp - Gp.conj().T) <= 1e-15 * np.linalg.norm(ref)
        if skip is None or skip[p] == 0:
            want[:, cols] = P @ J[p]
    assert np.linalg.norm(Zout - want) <= 1e-13 * np.linalg.norm(want)


@pytest.mark.parametrize("rows,ncols,panel", [(64, 128, 0), (300, 256, 1), (1024, 512, 0), (2048, 1024, 5), (520, 2048, 0)])
def test_qr_trailing_update_kernels_match_numpy(rows, ncols, panel):
    """The projection-coefficient kernel C = P^H T and the rank-64 trailing update T -= P C of the QR phase (csrc/tn_jacobi.cu) on their
    own: split-K with atomics and without, one to 31 trailing tiles, row counts that are / are not multiples of the tile sizes."""
    import tnb200
    from tnb200.api import _ptr, _f, check
    rng = np.random.default_rng(rows + ncols + panel)
    ctx = tnb200.Context.default()
    Q = _f(crandn(rng, rows, ncols))
    nt = ncols - 64 * (panel + 1)
    Cout = np.zeros(64 * nt, dtype=np.complex128)
    Qout = np.zeros((rows, ncols), dtype=np.complex128, order='F')
    check(ctx.lib.tn_jacobi_qr_update_pass(ctx.h, _ptr(Q), rows, ncols, panel, _ptr(Cout), _ptr(Qout)))
    P, T = Q[:, 64 * panel:64 * (panel + 1)], Q[:, 64 * (panel + 1):]
    Cref = P.conj().T @ T
    Cg = np.reshape(Cout, (64, nt), order='F')
    assert np.linalg.norm(Cg - Cref) <= 1e-13 * np.linalg.norm(Cref)
    want = Q.copy()
    want[:, 64 * (panel + 1):] = T - P @ Cg
    assert np.linalg.norm(Qout - want) <= 1e-13 * np.linalg.norm(want)
    assert np.array_equal(Qout[:, :64 * (panel + 1)], Q[:, :64 * (panel + 1)])
