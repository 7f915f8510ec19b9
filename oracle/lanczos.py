"""Lowest-eigenpair Lanczos with thick restart -- restates the published
algorithm of KrylovKit.jl ``eigsolve(f, x0, 1, :SR, Lanczos(krylovdim, maxiter,
tol))`` as the reference calls it at /root/reference/src/algorithms/mps/dmrg.jl:51-53
(krylovdim=3, maxiter=2, tol=1e-14, ishermitian=true).

KrylovKit is NOT vendored and NOT version-pinned by the reference, so only the
schedule is restated: build a ``krylovdim``-dimensional Krylov factorisation
(one operator application per vector), Rayleigh-Ritz, and if not converged keep
``div(3*krylovdim + 2*converged, 5)`` Ritz vectors (= 1 here) plus the residual
and expand again; at most ``maxiter`` such rounds, i.e. <= 5 applications for
the reference's settings.  Every new vector is orthogonalised against the whole
basis (two Gram-Schmidt passes); alpha is the real part of the diagonal
coefficient; the projected matrix is real symmetric.  The CUDA path
(csrc/tn_lanczos.cu) uses exactly this schedule."""
import numpy as np


def eigsolve_lowest(matvec, x0, krylovdim=3, maxiter=2, tol=1e-14):
    """Returns (theta, x, info): lowest Ritz value, unit-norm Ritz vector shaped
    like ``x0`` and ``info = dict(numops, numiter, normres, converged)``."""
    shape = x0.shape
    v = x0.reshape(-1).astype(np.complex128)
    v = v / np.linalg.norm(v)
    V = [v]
    T = np.zeros((krylovdim, krylovdim))
    numops, numiter = 0, 1

    def expand(K):
        """Apply the operator to V[K-1]; fill column K-1 of T; return residual, beta."""
        w = matvec(V[K - 1].reshape(shape)).reshape(-1)
        alpha = np.real(np.vdot(V[K - 1], w))
        for _ in range(2):
            for q in V:
                w = w - np.vdot(q, w) * q
        T[K - 1, K - 1] = alpha
        return w, float(np.linalg.norm(w))

    r, beta = expand(1)
    numops += 1
    converged = 0
    while True:
        K = len(V)
        if K == krylovdim or beta <= tol:
            D, U = np.linalg.eigh(T[:K, :K])
            f = U[K - 1, :] * beta
            converged = 0
            while converged < K and abs(f[converged]) <= tol:
                converged += 1
            if converged >= 1 or beta <= tol:
                break
        if K < krylovdim:
            V.append(r / beta)
            T[K - 1, K] = T[K, K - 1] = beta
            r, beta = expand(K + 1)
            numops += 1
        else:
            if numiter == maxiter:
                break
            keep = (3 * krylovdim + 2 * converged) // 5
            Vm = np.stack(V, axis=1)
            newV = [Vm @ U[:, j] for j in range(keep)]
            T[:, :] = 0.0
            for j in range(keep):
                T[j, j] = D[j]
                T[j, keep] = T[keep, j] = f[j]
            newV.append(r / beta)
            V = newV
            r, beta = expand(keep + 1)
            numops += 1
            numiter += 1
    K = len(V)
    D, U = np.linalg.eigh(T[:K, :K])
    x = np.stack(V, axis=1) @ U[:, 0]
    x = x / np.linalg.norm(x)
    return float(D[0]), x.reshape(shape), dict(numops=numops, numiter=numiter, normres=abs(U[K - 1, 0]) * beta, converged=converged)
