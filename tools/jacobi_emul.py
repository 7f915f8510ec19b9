"""CPU emulation of the blocked one-sided Jacobi SVD (algorithm design aid, not product code).
Usage: python tools/jacobi_emul.py n b inner_sweeps [kind]"""
import sys
import numpy as np

def rr_pairs(n, step):
    ps, qs = [], []
    for k in range(n // 2):
        if n == 2: a, b = 0, 1
        elif k == 0: a, b = n - 1, step
        else: a, b = (step + k) % (n - 1), (step - k + n - 1) % (n - 1)
        ps.append(min(a, b)); qs.append(max(a, b))
    return np.array(ps), np.array(qs)

def evd_jacobi(G, tol, max_sweeps):
    n = G.shape[0]
    G = G.copy(); J = np.eye(n, dtype=complex)
    nrot = 0
    for sweep in range(max_sweeps):
        rotated = False
        for step in range(n - 1):
            ps, qs = rr_pairs(n, step)
            a, b, c = G[ps, ps].real, G[qs, qs].real, G[ps, qs]
            absc = np.abs(c)
            act = (absc > tol * np.sqrt(np.abs(a * b))) & (absc > 0)
            if not act.any(): continue
            rotated = True
            zeta = np.where(act, (b - a) / (2 * np.where(act, absc, 1)), 0)
            t = np.where(zeta >= 0, 1.0, -1.0) / (np.abs(zeta) + np.sqrt(1 + zeta ** 2))
            cs = np.where(act, 1 / np.sqrt(1 + t * t), 1.0)
            s = np.where(act, cs * t * c / np.where(act, absc, 1), 0)
            nrot += act.sum()
            for M in (G, J):
                x, y = M[:, ps].copy(), M[:, qs].copy()
                M[:, ps] = cs * x - y * np.conj(s); M[:, qs] = x * s + cs * y
            x, y = G[ps, :].copy(), G[qs, :].copy()
            G[ps, :] = cs[:, None] * x - s[:, None] * y; G[qs, :] = np.conj(s)[:, None] * x + cs[:, None] * y
            G[ps, qs] = 0; G[qs, ps] = 0
        if not rotated: break
    return J, sweep + 1

def block_jacobi(A, b, inner, verbose=True, sort_eig=False):
    m, n = A.shape
    W = A.astype(complex).copy(); V = np.eye(n, dtype=complex)
    nb = n // b
    tol = np.sqrt(m) * 2.2e-16
    hist = []
    for sweep in range(40):
        offmax = 0.0; inner_tot = 0
        for step in range(nb - 1):
            ps, qs = rr_pairs(nb, step)
            for p, q in zip(ps, qs):
                cols = np.r_[p * b:(p + 1) * b, q * b:(q + 1) * b]
                P = W[:, cols]
                G = P.conj().T @ P
                G = (G + G.conj().T) / 2
                dg = np.sqrt(np.abs(np.diag(G).real))
                R = np.abs(G) / np.maximum(np.outer(dg, dg), 1e-300)
                np.fill_diagonal(R, 0)
                off = R.max(); offmax = max(offmax, off)
                if off <= tol: continue
                if inner <= 0:
                    w, J = np.linalg.eigh(G); 
                    if sort_eig: J = J[:, ::-1]
                    isw = 0
                else:
                    J, isw = evd_jacobi(G, tol, inner)
                inner_tot += isw
                W[:, cols] = P @ J; V[:, cols] = V[:, cols] @ J
        hist.append(offmax)
        if verbose: print("sweep", sweep + 1, "offmax %.3e" % offmax, "inner sweeps", inner_tot, flush=True)
        if offmax <= tol: break
    s = np.linalg.norm(W, axis=0)
    return W, V, s, sweep + 1

if __name__ == "__main__":
    n = int(sys.argv[1]); b = int(sys.argv[2]); inner = int(sys.argv[3]); kind = sys.argv[4] if len(sys.argv) > 4 else "randn"
    rng = np.random.default_rng(0)
    if kind == "randn":
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    else:   # decaying spectrum like an MPS bond
        u, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        v, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        A = (u * np.exp(-np.arange(n) * 30.0 / n)) @ v.conj().T
    W, V, s, sw = block_jacobi(A, b, inner)
    so = np.linalg.svd(A, compute_uv=False)
    print("sweeps", sw, "sv err", np.max(np.abs(np.sort(s)[::-1] - so)) / so[0], "orthV", np.linalg.norm(V.conj().T @ V - np.eye(n)))
