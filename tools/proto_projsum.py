"""CPU check of the GEMM index wiring used by tn_projsum.cu (project / squared product / one-site product), run with
tools/gemm_emul.py against the oracle.  Development tooling only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import oracle
from gemm_emul import zgemm, idx1, idx2, flat
from gpu_util import random_complex_mps, random_mpo, crandn

rng = np.random.default_rng(0)
N, d = 6, 2
Z = lambda n: np.zeros(n, dtype=np.complex128)


def F(x, shape):
    return np.reshape(x, shape, order='F')


def blk3(x):
    return x.reshape(x.shape[0], 1, x.shape[-1]) if x.ndim == 2 else x


def phi2_nompo(L, V1, V2, R):
    """phi = conj(project) for an overlap env: blocks (a,1,b)."""
    L, R = blk3(L), blk3(R)
    ca, _, cb = L.shape; ca1 = V1.shape[2]; ca2 = V2.shape[2]; cb2 = R.shape[2]
    X1 = Z(cb * d * ca1)
    zgemm(cb, d * ca1, ca, flat(L), idx1(ca), idx1(1), 1, flat(V1), idx1(1), idx1(ca), 0, X1, idx1(1), idx1(cb))
    X2 = Z(cb * d * d * ca2)
    zgemm(cb * d, d * ca2, ca1, X1, idx1(1), idx1(cb * d), 0, flat(V2), idx1(1), idx1(ca1), 0, X2, idx1(1), idx1(cb * d))
    ph = Z(cb * d * d * cb2)
    zgemm(cb * d * d, cb2, ca2, X2, idx1(1), idx1(cb * d * d), 0, flat(R), idx1(1), idx1(ca2), 1, ph, idx1(1), idx1(cb * d * d))
    return F(ph, (cb, d, d, cb2))


def phi1_nompo(L, V, R):
    L, R = blk3(L), blk3(R)
    ca, _, cb = L.shape; ca1 = V.shape[2]; cb2 = R.shape[2]
    X1 = Z(cb * d * ca1)
    zgemm(cb, d * ca1, ca, flat(L), idx1(ca), idx1(1), 1, flat(V), idx1(1), idx1(ca), 0, X1, idx1(1), idx1(cb))
    ph = Z(cb * d * cb2)
    zgemm(cb * d, cb2, ca1, X1, idx1(1), idx1(cb * d), 0, flat(R), idx1(1), idx1(ca1), 1, ph, idx1(1), idx1(cb * d))
    return F(ph, (cb, d, cb2))


def build_w2(M1, M2):
    w, _, _, w1 = M1.shape; w2 = M2.shape[3]
    W = np.einsum('wsaj,jtbx->wabstx', M1, M2)      # (w, s1', s2', s1, s2, w2)
    return flat(W)


def phi2_mpo(L, V1, V2, M1, M2, R):
    ca, w, cb = L.shape; ca1 = V1.shape[2]; ca2 = V2.shape[2]; _, w2, cb2 = R.shape
    d2 = d * d
    W = build_w2(M1, M2)
    th = Z(ca * d2 * ca2)
    zgemm(ca * d, d * ca2, ca1, flat(V1), idx1(1), idx1(ca * d), 0, flat(V2), idx1(1), idx1(ca1), 0, th, idx1(1), idx1(ca * d))
    # T1(b,w,s1,s2,a'') = sum_a conj(L(a,w,b)) th(a,s1,s2,a''); m = (w,b) in L's order, written transposed
    T1 = Z(cb * w * d2 * ca2)
    zgemm(w * cb, d2 * ca2, ca, flat(L), idx1(ca), idx1(1), 1, th, idx1(1), idx1(ca), 0, T1, idx2(w, cb, 1), idx1(w * cb))
    # T2(b,s1',s2',w2,a'') = sum_{(w,s1,s2)} T1(b,(w,s1,s2),a'') conj(W[(w,s1',s2'),(s1,s2,w2)])
    T2 = Z(cb * d2 * w2 * ca2)
    zgemm(cb * ca2, d2 * w2, w * d2, T1, idx2(cb, 1, cb * w * d2), idx1(cb), 0,
          W, idx2(w, 1, w * d2), idx2(d2, w, w * d2 * d2), 1, T2, idx2(cb, 1, cb * d2 * w2), idx1(cb))
    # phi[(b,s1',s2'),b'] = sum_{(w2,a'')} T2[(b,s1',s2'),(w2,a'')] conj(R(a'',w2,b'))
    ph = Z(cb * d2 * cb2)
    zgemm(cb * d2, cb2, w2 * ca2, T2, idx1(1), idx1(cb * d2), 0, flat(R), idx2(w2, ca2, 1), idx1(ca2 * w2), 1, ph, idx1(1), idx1(cb * d2))
    return F(ph, (cb, d, d, cb2))


def phi1_mpo(L, V, M, R):
    ca, w, cb = L.shape; ca1 = V.shape[2]; _, w2, cb2 = R.shape
    T1 = Z(cb * w * d * ca1)
    zgemm(w * cb, d * ca1, ca, flat(L), idx1(ca), idx1(1), 1, flat(V), idx1(1), idx1(ca), 0, T1, idx2(w, cb, 1), idx1(w * cb))
    # T2(b,s',w2,a') = sum_{(w,s)} T1(b,w,s,a') conj(M(w,s,s',w2))
    T2 = Z(cb * d * w2 * ca1)
    zgemm(cb * ca1, d * w2, w * d, T1, idx2(cb, 1, cb * w * d), idx1(cb), 0,
          flat(M), idx1(1), idx1(w * d), 1, T2, idx2(cb, 1, cb * d * w2), idx1(cb))
    ph = Z(cb * d * cb2)
    zgemm(cb * d, cb2, w2 * ca1, T2, idx1(1), idx1(cb * d), 0, flat(R), idx2(w2, ca1, 1), idx1(ca1 * w2), 1, ph, idx1(1), idx1(cb * d))
    return F(ph, (cb, d, cb2))


def product1(L, M, A, R, coeff):
    """one-site H_eff: out(a,s,a') = sum L(a,w,b) M(w,s,s',w') A(b,s',b') R(a',w',b')."""
    ca, w, cb = L.shape; cb2 = A.shape[2]; ca2, w2, _ = R.shape
    T1 = Z(ca * w * d * cb2)
    zgemm(ca * w, d * cb2, cb, flat(L), idx1(1), idx1(ca * w), 0, flat(A), idx1(1), idx1(cb), 0, T1, idx1(1), idx1(ca * w))
    T2 = Z(ca * d * w2 * cb2)
    zgemm(ca * cb2, d * w2, w * d, T1, idx2(ca, 1, ca * w * d), idx1(ca), 0,
          flat(M), idx2(w, 1, w * d), idx2(d, w, w * d * d), 0, T2, idx2(ca, 1, ca * d * w2), idx1(ca))
    out = Z(ca * d * ca2)
    zgemm(ca * d, ca2, w2 * cb2, T2, idx1(1), idx1(ca * d), 0, flat(R), idx1(ca2), idx1(1), 0, out, idx1(1), idx1(ca * d), alpha=coeff)
    return F(out, (ca, d, ca2))


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


V = random_complex_mps(rng, N, d, 5, center=1)
psi = random_complex_mps(rng, N, d, 6, center=3)
H = random_mpo(rng, N, d, 3)
for c in (2, 3, 5):
    P = oracle.ProjMPS([V, psi], rank=1, center=c)
    for direction in (False, True):
        site = c - 1 if direction else c
        if site < 1 or site + 1 > N:
            continue
        want = np.conj(P.project(None, direction, 2))
        got = phi2_nompo(P.block(site - 1), V[site], V[site + 1], P.block(site + 2))
        print("phi2 nompo", c, direction, rel(got, want))
    want = np.conj(P.project(None, False, 1))
    got = phi1_nompo(P.block(c - 1), V[c], P.block(c + 1))
    print("phi1 nompo", c, rel(got, want))
    Q = oracle.ProjMPS([V, H, psi], rank=1, center=c)
    site = c
    want = np.conj(Q.project(None, False, 2))
    got = phi2_mpo(Q.block(site - 1), V[site], V[site + 1], H[site], H[site + 1], Q.block(site + 2))
    print("phi2 mpo", c, rel(got, want))
    want = np.conj(Q.project(None, False, 1))
    got = phi1_mpo(Q.block(c - 1), V[c], H[c], Q.block(c + 1))
    print("phi1 mpo", c, rel(got, want))
    E = oracle.ProjMPS([psi, H, psi], rank=2, center=c, coeff=0.7 - 0.2j)
    A = crandn(rng, *psi[c].shape)
    want = E.product(A, False, 1)
    got = product1(E.block(c - 1), H[c], A, E.block(c + 1), 0.7 - 0.2j)
    print("product1", c, rel(got, want))
