"""CPU emulation of the planned GPU QR preconditioner: right-looking block Gram-Schmidt with shifted-CholeskyQR3
panels (all GEMM-shaped), applied twice ("qr2": A = Q1 R1, R1^H = Q2 R2, Jacobi on X = R2^H).  Design aid only."""
import sys
import numpy as np
sys.path.insert(0, 'tools')
from jacobi_emul import block_jacobi

EPS = 2.220446049250313e-16

def chol_shift(G, shift):
    n = G.shape[0]
    G = G + shift * np.eye(n)
    R = np.zeros_like(G)
    for j in range(n):      # upper Cholesky G = R^H R with pivot guard
        d = G[j, j].real - np.sum(np.abs(R[:j, j]) ** 2)
        if d <= 0 or not np.isfinite(d):
            R[j, j] = 1.0; R[j, j + 1:] = 0   # zero/noise column: leave it
            continue
        R[j, j] = np.sqrt(d)
        R[j, j + 1:] = (G[j, j + 1:] - R[:j, j].conj() @ R[:j, j + 1:]) / R[j, j]
    return R

def panel_scholqr3(P, passes=3):
    m, b = P.shape
    Rtot = np.eye(b, dtype=complex)
    Q = P.copy()
    for it in range(passes):
        G = Q.conj().T @ Q
        G = (G + G.conj().T) / 2
        shift = 11 * (m * b + b * (b + 1)) * EPS * np.linalg.norm(G, 2) if it == 0 else 0.0
        R = chol_shift(G, shift)
        Q = Q @ np.linalg.inv(R)
        Rtot = R @ Rtot
    return Q, Rtot

def bgs_qr(A, b=64, twice=True):
    m, n = A.shape
    def one_pass(A):
        Q = A.astype(complex).copy(); R = np.zeros((n, n), dtype=complex)
        for k in range(0, n, b):
            e = min(k + b, n)
            Qk, Rkk = panel_scholqr3(Q[:, k:e])
            Q[:, k:e] = Qk; R[k:e, k:e] = Rkk
            if e < n:
                C = Qk.conj().T @ Q[:, e:]
                Q[:, e:] -= Qk @ C
                R[k:e, e:] = C
        return Q, R
    Q, R = one_pass(A)
    if twice:
        Q, R2 = one_pass(Q); R = R2 @ R
    return Q, R

def svd_precond(A, b=32, inner=1, nqr=2):
    m, n = A.shape
    Q1, R1 = bgs_qr(A)
    if nqr == 2:
        Q2, R2 = bgs_qr(R1.conj().T)
        X = R2.conj().T
    else:
        Q2 = None; X = R1.conj().T       # Jacobi on R1^H:  R1^H V = U' S  ->  A = Q1 R1 = Q1 V S U'^H
    W, V, s, sw = block_jacobi(X, b, inner, verbose=False)
    order = np.argsort(-s); s = s[order]; W = W[:, order]; V = V[:, order]
    Up = W / np.where(s > 0, s, 1)
    if nqr == 2:
        U = Q1 @ Up; Vf = Q2 @ V        # A = Q1 R1, R1 = X Q2^H  (R1^H = Q2 R2 = Q2 X^H)
    else:
        U = Q1 @ V; Vf = Up
    return U, s, Vf, sw

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rng = np.random.default_rng(0)
    cr = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    def hidden():
        u, _ = np.linalg.qr(cr(n, n)); v, _ = np.linalg.qr(cr(n, n)); return (u * np.exp(-np.arange(n) * 30.0 / n)) @ v.conj().T
    def visible():
        lam = np.exp(-np.arange(n // 2) * 28.0 / (n // 2)); return (np.tile(lam, 2)[:, None] * cr(n, n)) * np.tile(lam, 2)[None, :]
    def rankdef():
        return cr(n, n // 2) @ cr(n // 2, n)
    def zerocols():
        A = cr(n, n); A[:, 5] = 0; A[:, n - 3:] = 0; return A
    def tall():
        return cr(2 * n, n) * np.exp(-np.arange(n) * 20.0 / n)[None, :]
    for name, A in (("randn", cr(n, n)), ("hidden", hidden()), ("visible", visible()), ("rankdef", rankdef()), ("zerocols", zerocols()), ("tall", tall())):
        so = np.linalg.svd(A, compute_uv=False)
        for nqr in (1, 2):
            U, s, V, sw = svd_precond(A, nqr=nqr)
            k = int(np.sum(so > 1e-13 * so[0]))
            rec = np.linalg.norm((U * s) @ V.conj().T - A) / np.linalg.norm(A)
            oU = np.linalg.norm(U[:, :k].conj().T @ U[:, :k] - np.eye(k)); oV = np.linalg.norm(V[:, :k].conj().T @ V[:, :k] - np.eye(k))
            print("%-9s nqr=%d sweeps %2d  sv_abs %.1e  recon %.1e  orthU(k=%d) %.1e orthV %.1e" % (name, nqr, sw, np.max(np.abs(s - so)) / so[0], rec, k, oU, oV), flush=True)
