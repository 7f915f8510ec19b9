"""NumPy emulation of the strided contraction kernel's ADDRESSING (tn_common.cuh: Idx2 / GemmDesc), used on the CPU-only
build machine to check the index wiring of a new GEMM call sequence before it is transcribed into tn_mps.cu.
Development tooling only: not imported by the product, the tests or the bench."""
import numpy as np

BIG = 0x7fffffff


def idx1(stride):
    return (BIG, int(stride), 0)


def idx2(n0, s0, s1):
    return (int(n0), int(s0), int(s1))


def _offs(n, ix):
    n0, s0, s1 = ix
    i = np.arange(n)
    return (i % n0) * s0 + (i // n0) * s1


def zgemm(M, N, K, A, am, ak, conjA, B, bk, bn, conjB, C, cm, cn, alpha=1.0, beta=0.0):
    """C[m,n] = alpha * sum_k op(A)[m,k] op(B)[k,n] + beta * C[m,n] on FLAT buffers (element offsets as in the kernel)."""
    A = np.asarray(A).reshape(-1)
    B = np.asarray(B).reshape(-1)
    a = A[_offs(M, am)[:, None] + _offs(K, ak)[None, :]]
    b = B[_offs(K, bk)[:, None] + _offs(N, bn)[None, :]]
    if conjA:
        a = a.conj()
    if conjB:
        b = b.conj()
    co = _offs(M, cm)[:, None] + _offs(N, cn)[None, :]
    assert len(np.unique(co)) == M * N, "output addressing is not injective"
    old = C[co] if beta != 0 else 0
    C[co] = alpha * (a @ b) + beta * old
    return C


def flat(x):
    """column-major flat copy (the device layout)."""
    return np.asfortranarray(x).reshape(-1, order='F').copy()
