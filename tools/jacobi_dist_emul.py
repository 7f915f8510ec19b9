"""CPU emulation of the DISTRIBUTED ordering planned for the multi-GPU one-sided block Jacobi: every rank
holds two column super-blocks; a step rotates all cross pairs of the rank's two super-blocks, then super-blocks move round-robin
(2G-1 steps per sweep); pairs inside a super-block are rotated once per sweep.  Compares the number of sweeps with the single-GPU
round-robin order on the same matrix.  Design aid, not product code.
Usage: python tools/jacobi_dist_emul.py n b G [kind]"""
import sys
import numpy as np
from jacobi_emul import rr_pairs, evd_jacobi, block_jacobi


def rotate_pair(W, V, b, p, q, tol, inner):
    cols = np.r_[p * b:(p + 1) * b, q * b:(q + 1) * b]
    P = W[:, cols]
    G = P.conj().T @ P
    G = (G + G.conj().T) / 2
    dg = np.sqrt(np.abs(np.diag(G).real))
    R = np.abs(G) / np.maximum(np.outer(dg, dg), 1e-300)
    np.fill_diagonal(R, 0)
    off = R.max()
    if off <= tol:
        return off
    J, _ = evd_jacobi(G, tol, inner)
    W[:, cols] = P @ J
    V[:, cols] = V[:, cols] @ J
    return off


def dist_block_jacobi(A, b, G, inner=1, verbose=True):
    m, n = A.shape
    W = A.astype(complex).copy()
    V = np.eye(n, dtype=complex)
    nb = n // b
    nsb = 2 * G                      # super-blocks
    assert nb % nsb == 0
    k = nb // nsb                    # column blocks per super-block
    tol = np.sqrt(m) * 2.2e-16
    for sweep in range(60):
        offmax = 0.0
        # pairs inside every super-block (all ranks in parallel: 2 super-blocks each)
        if k > 1:
            for st in range(k - 1 if k % 2 == 0 else k):
                ps, qs = rr_pairs(k if k % 2 == 0 else k + 1, st)
                for sb in range(nsb):
                    for p, q in zip(ps, qs):
                        if p < k and q < k:
                            offmax = max(offmax, rotate_pair(W, V, b, sb * k + p, sb * k + q, tol, inner))
        # cross pairs: round-robin over the super-blocks, k matchings per meeting
        for st in range(nsb - 1):
            Ps, Qs = rr_pairs(nsb, st)
            for shift in range(k):
                for P, Q in zip(Ps, Qs):         # the G ranks work in parallel on their (P, Q)
                    for i in range(k):
                        offmax = max(offmax, rotate_pair(W, V, b, P * k + i, Q * k + (i + shift) % k, tol, inner))
        if verbose:
            print("dist sweep", sweep + 1, "offmax %.3e" % offmax, flush=True)
        if offmax <= tol:
            break
    return W, V, np.linalg.norm(W, axis=0), sweep + 1


if __name__ == "__main__":
    n = int(sys.argv[1]); b = int(sys.argv[2]); G = int(sys.argv[3]); kind = sys.argv[4] if len(sys.argv) > 4 else "randn"
    rng = np.random.default_rng(0)
    if kind == "randn":
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    else:
        u, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        v, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        A = (u * np.exp(-np.arange(n) * 30.0 / n)) @ v.conj().T
        # what the device feeds the Jacobi after its two QR steps: X = R2^H (columns nearly orthogonal for graded spectra)
        q1, r1 = np.linalg.qr(A); q2, r2 = np.linalg.qr(r1.conj().T); A = r2.conj().T
    so = np.linalg.svd(A, compute_uv=False)
    _, _, s1, sw1 = block_jacobi(A, b, 1, verbose=False)
    W, V, s2, sw2 = dist_block_jacobi(A, b, G, 1, verbose=False)
    print({"n": n, "b": b, "G": G, "kind": kind, "sweeps_round_robin": sw1, "sweeps_distributed": sw2,
           "sv_err_distributed": float(np.max(np.abs(np.sort(s2)[::-1] - so)) / so[0]),
           "orthV": float(np.linalg.norm(V.conj().T @ V - np.eye(n)))})
