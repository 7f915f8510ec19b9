import numpy as np
import oracle
from oracle.gmps import GMPS as OGMPS


def crandn(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def random_complex_mps(rng, N, d, chi, center=1):
    psi = oracle.randomMPS(d, N, chi, rng)
    psi = OGMPS(1, d, [t + 0.4j * rng.standard_normal(t.shape) for t in psi.tensors], 0)
    psi.movecenter(center)
    psi.normalize()
    return psi


def random_mpo(rng, N, d, w):
    W = [crandn(rng, 1 if i == 0 else w, d, d, 1 if i == N - 1 else w) for i in range(N)]
    return OGMPS(2, d, W, 0)


def relerr(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)
