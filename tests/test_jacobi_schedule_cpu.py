"""CPU checks of two pieces of host-visible logic behind the GPU truncated SVD (no GPU needed):
  * the split pair schedule (csrc/tn_svd.cu "Split schedule", restated in tools/jacobi_sched_emul.py) converges like the circle method when the
    blocked one-sided Jacobi iteration is emulated in NumPy on a QR-preconditioned matrix, and gives LAPACK's singular values;
  * the tile numbering of jacobi_gram64_kernel (csrc/tn_jacobi.cu): the nine 8x8 tiles per warp cover every unordered pair of the eight
    8-column groups exactly once, so writing each off-diagonal tile and its mirror image yields the full Hermitian 64 x 64 block."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def _sweeps(X, steps, b, tol):
    import jacobi_emul as je
    W = X.copy()
    for sweep in range(30):
        offmax = 0.0
        for st in steps:
            for p, q in st:
                cols = np.r_[p * b:(p + 1) * b, q * b:(q + 1) * b]
                P = W[:, cols]
                G = P.conj().T @ P
                G = (G + G.conj().T) / 2
                dg = np.sqrt(np.abs(np.diag(G).real))
                R = np.abs(G) / np.maximum(np.outer(dg, dg), 1e-300)
                np.fill_diagonal(R, 0)
                offmax = max(offmax, R.max())
                if R.max() <= tol:
                    continue
                J, _ = je.evd_jacobi(G, tol, 1)
                W[:, cols] = P @ J
        if offmax <= tol or offmax <= 1e-9:
            break
    return sweep + 1, np.sort(np.linalg.norm(W, axis=0))[::-1]


def test_split_schedule_converges_like_the_circle_method():
    import jacobi_sched_emul as js
    n, b = 128, 8                          # 16 column blocks, as a 512 x 512 problem has with 32-column blocks
    rng = np.random.default_rng(3)
    u, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    v, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    A = (u * np.exp(-np.arange(n) * 20.0 / n)) @ v.conj().T
    R1 = np.linalg.qr(A)[1]
    X = np.linalg.qr(R1.conj().T)[1].conj().T        # the product's preconditioner: Jacobi runs on R2^H
    tol = 3 * np.sqrt(n) * 2.2e-16
    ref = np.linalg.svd(A, compute_uv=False)
    counts = {}
    for groups in (1, 2, 4):
        assert js.check(n // b, groups)[1] == n // b - 1
        counts[groups], s = _sweeps(X, js.flat_steps(n // b, groups), b, tol)
        assert np.max(np.abs(s - ref)) <= 1e-12 * ref[0]
    assert counts[2] <= counts[1] + 1 and counts[4] <= counts[1] + 1, counts


def test_gram_kernel_tile_numbering_covers_every_group_pair_once():
    seen = {}
    for warp in range(4):
        tiles = [(2 * warp + ii, (2 * warp + ii + d) & 7) for ii in range(2) for d in range(4)] + [(warp, warp + 4)]
        assert len(tiles) == 9
        for bi, bj in tiles:
            key = (min(bi, bj), max(bi, bj))
            assert key not in seen, (key, warp, seen[key])
            seen[key] = warp
        # fragments a warp loads per k-step: groups (2w + d) & 7 for d = 0..4, plus w and w + 4
        frag = {(2 * warp + d) & 7 for d in range(5)} | {warp, warp + 4}
        assert all(bi in frag and bj in frag for bi, bj in tiles)
    assert len(seen) == 36 and set(seen) == {(i, j) for i in range(8) for j in range(i, 8)}
