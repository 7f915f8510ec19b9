"""GPU parity: strided complex-FP64 contraction kernel (K1) vs NumPy (tolerance: relative
Frobenius <= 1e-13, SURVEY.md section 4 tier 4)."""
import numpy as np
import pytest

from gpu_util import crandn, relerr

pytestmark = pytest.mark.gpu
BIG = 1 << 40


def test_plain_nn_and_edges():
    import tnb200
    rng = np.random.default_rng(0)
    for (M, N, K) in [(1, 1, 1), (7, 5, 3), (64, 64, 64), (129, 65, 33), (300, 200, 150), (256, 20, 20), (40, 300, 17)]:
        A, B = crandn(rng, M, K), crandn(rng, K, N)
        Af, Bf = np.asfortranarray(A), np.asfortranarray(B)
        C = tnb200.contract_strided(M, N, K, Af.reshape(-1, order='F'), (BIG, 1, 0), (BIG, M, 0), 0,
                                    Bf.reshape(-1, order='F'), (BIG, 1, 0), (BIG, K, 0), 0, M * N, (BIG, 1, 0), (BIG, M, 0))
        assert relerr(C.reshape(M, N, order='F'), A @ B) < 1e-13


def test_conj_transpose_and_alpha():
    import tnb200
    rng = np.random.default_rng(1)
    M, N, K = 70, 90, 130
    A, B = crandn(rng, K, M), crandn(rng, N, K)     # C = alpha * A^H B^T(conj)
    C = tnb200.contract_strided(M, N, K, np.asfortranarray(A).reshape(-1, order='F'), (BIG, K, 0), (BIG, 1, 0), 1,
                                np.asfortranarray(B).reshape(-1, order='F'), (BIG, N, 0), (BIG, 1, 0), 1,
                                M * N, (BIG, 1, 0), (BIG, M, 0), alpha=0.5 - 2j)
    assert relerr(C.reshape(M, N, order='F'), (0.5 - 2j) * (A.conj().T @ B.conj().T)) < 1e-13


def test_fused_two_level_indices():
    """T2(a,s,w2,b') = sum_{w,s'} T1(a,w,s',b') M(w,s,s',w2): rows (a,b') and the MPO operand use
    two-level strides -- the permute-into-tile path (no transposed copies)."""
    import tnb200
    rng = np.random.default_rng(2)
    ca, w, d, cb, w2 = 37, 5, 2, 29, 4
    T1 = crandn(rng, ca, w, d, cb)
    Mo = crandn(rng, w, d, d, w2)
    want = np.einsum('awtb,wstx->asxb', T1, Mo)
    C = tnb200.contract_strided(ca * cb, d * w2, w * d,
                                np.asfortranarray(T1).reshape(-1, order='F'), (ca, 1, ca * w * d), (BIG, ca, 0), 0,
                                np.asfortranarray(Mo).reshape(-1, order='F'), (w, 1, w * d), (d, w, w * d * d), 0,
                                ca * d * w2 * cb, (ca, 1, ca * d * w2), (BIG, ca, 0))
    assert relerr(C.reshape(ca, d, w2, cb, order='F'), want) < 1e-13


def test_streamk_partial_waves_and_small_grids():
    """Shapes that take the persistent stream-K path (partial last wave / less than one wave of 128x32 tiles on 296 CTA
    slots, ragged edges, K not a multiple of 8): partial tiles are summed with red.global.add.f64 into a zeroed C."""
    import tnb200
    rng = np.random.default_rng(3)
    for (M, N, K) in [(512, 128, 640), (1000, 250, 333), (2048, 512, 300), (1283, 997, 129), (4096, 1024, 136), (130, 4100, 200)]:
        A, B = crandn(rng, M, K), crandn(rng, K, N)
        C = tnb200.contract_strided(M, N, K, np.asfortranarray(A).reshape(-1, order='F'), (BIG, 1, 0), (BIG, M, 0), 0,
                                    np.asfortranarray(B).reshape(-1, order='F'), (BIG, 1, 0), (BIG, K, 0), 0, M * N, (BIG, 1, 0), (BIG, M, 0))
        assert relerr(C.reshape(M, N, order='F'), A @ B) < 1e-13, (M, N, K)


def test_streamk_conj_kfast_operands():
    """Stream-K with both operands k-contiguous and conjugated (the environment-update form conj(A)^T X)."""
    import tnb200
    rng = np.random.default_rng(4)
    M, N, K = 700, 300, 520
    A, B = crandn(rng, K, M), crandn(rng, N, K)
    C = tnb200.contract_strided(M, N, K, np.asfortranarray(A).reshape(-1, order='F'), (BIG, K, 0), (BIG, 1, 0), 1,
                                np.asfortranarray(B).reshape(-1, order='F'), (BIG, N, 0), (BIG, 1, 0), 1,
                                M * N, (BIG, 1, 0), (BIG, M, 0), alpha=1.5 + 0.25j)
    assert relerr(C.reshape(M, N, order='F'), (1.5 + 0.25j) * (A.conj().T @ B.conj().T)) < 1e-13


def test_two_level_k_blocks_aligned_to_tile():
    """k = (k0, k1) with the level-1 block a multiple of the 8-deep k-tile (pointer-walking load path, KMODE 2):
    C[m,n] = sum_{k0,k1} A(k0,m,k1) B(n,k0,k1)."""
    import tnb200
    rng = np.random.default_rng(7)
    for (k0, k1, M, N) in [(16, 6, 50, 40), (8, 9, 130, 70), (32, 2, 300, 64)]:
        A, B = crandn(rng, k0, M, k1), crandn(rng, N, k0, k1)
        want = np.einsum('amb,nab->mn', A, B)
        C = tnb200.contract_strided(M, N, k0 * k1, np.asfortranarray(A).reshape(-1, order='F'), (BIG, k0, 0), (k0, 1, k0 * M), 0,
                                    np.asfortranarray(B).reshape(-1, order='F'), (k0, N, N * k0), (BIG, 1, 0), 0, M * N, (BIG, 1, 0), (BIG, M, 0))
        assert relerr(C.reshape(M, N, order='F'), want) < 1e-13, (k0, k1, M, N)


@pytest.mark.parametrize("variant", [1, 2])
def test_bulk_copy_kernel_variants_match_numpy(variant):
    """The bulk-copy (TMA engine: cp.async.bulk + mbarrier) variants of the 128 x 32 kernel -- an experiment that is off by default
    (TN_GEMM_BULK=1: 8-deep k-tiles x 4 stages, 2: 16-deep x 2 stages) -- on shapes that take them: full tiles, exactly one
    wave of CTAs, B k-fast and B n-fast, with conjugation and alpha.  The switch is read once per process, hence the subprocess."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent('''
        import sys, numpy as np
        sys.path[:0] = [%r, %r]
        import tnb200
        from gpu_util import crandn, relerr
        BIG = 1 << 40
        rng = np.random.default_rng(11)
        M, N, K = 128 * 37, 32 * 8, 96
        A = crandn(rng, M, K)
        Bk = crandn(rng, K, N)                   # k-fast: column-major K x N
        C = tnb200.contract_strided(M, N, K, np.asfortranarray(A).reshape(-1, order='F'), (BIG, 1, 0), (BIG, M, 0), 0,
                                    np.asfortranarray(Bk).reshape(-1, order='F'), (BIG, 1, 0), (BIG, K, 0), 1, M * N, (BIG, 1, 0), (BIG, M, 0), alpha=0.5 - 2j)
        assert relerr(C.reshape(M, N, order='F'), (0.5 - 2j) * (A @ Bk.conj())) < 1e-13
        Bn = crandn(rng, N, K)                   # n-fast: column-major N x K, used as its transpose
        C = tnb200.contract_strided(M, N, K, np.asfortranarray(A).reshape(-1, order='F'), (BIG, 1, 0), (BIG, M, 0), 0,
                                    np.asfortranarray(Bn).reshape(-1, order='F'), (BIG, N, 0), (BIG, 1, 0), 0, M * N, (BIG, 1, 0), (BIG, M, 0))
        assert relerr(C.reshape(M, N, order='F'), A @ Bn.T) < 1e-13
        print("ok", tnb200.Context.default().counters()["launches"])
    ''') % (os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tensornetworks.jl_b200"), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TN_GEMM_BULK=str(variant))
    env["PYTHONPATH"] = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) + os.pathsep + env.get("PYTHONPATH", "")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
