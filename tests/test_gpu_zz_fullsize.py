"""GPU parity at BASELINE.json's full single-GPU sizes (C2: chi = 1024, w = 5, d = 2; the 2048 x 2048 truncated SVD of a C2 bond),
where the oracle cannot produce the whole answer in seconds, through size-independent properties:
  H_eff matvec -- linearity, the adjoint identity <y, H x> = <H^dag y, x> with the conjugate-transposed blocks, and exact oracle
                  values (NumPy einsum) on a random subset of output rows;
  truncated SVD -- orthonormal factors, reconstruction of the truncated matrix, sorted singular values, the reference's rank rule.
Tolerances: 1e-12 relative (contractions; 1e-11 for the adjoint identity, whose reference value is an inner product ~ 1/sqrt(n) of the
norms), 5e-12 sigma_max (singular values) and 1e-11 (orthonormality / reconstruction) at n = 2048.  (Sorts last among the GPU suites.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CHI, D, W = 1024, 2, 5


def heff_properties(make_product, chi, rng, nrows=3):
    """``make_product(L, R, M1, M2)`` returns a function theta -> H_eff theta.  Shared with the CPU self-check below."""
    def crandn(*s):
        return rng.standard_normal(s) + 1j * rng.standard_normal(s)
    L, R = crandn(chi, W, chi) / np.sqrt(chi), crandn(chi, W, chi) / np.sqrt(chi)
    M1, M2 = crandn(W, D, D, W), crandn(W, D, D, W)
    H = make_product(L, R, M1, M2)
    x, y = crandn(chi, D, D, chi), crandn(chi, D, D, chi)
    Hx, Hy = H(x), H(y)
    al, be = 0.7 - 0.3j, -1.1 + 0.4j
    lin = np.linalg.norm(H(al * x + be * y) - (al * Hx + be * Hy)) / np.linalg.norm(Hx)
    # adjoint: blocks conjugate-transposed in the bond indices, MPO tensors conjugated with physical in/out swapped
    Ld, Rd = np.conj(np.transpose(L, (2, 1, 0))), np.conj(np.transpose(R, (2, 1, 0)))
    M1d, M2d = np.conj(np.transpose(M1, (0, 2, 1, 3))), np.conj(np.transpose(M2, (0, 2, 1, 3)))
    Hd = make_product(Ld, Rd, M1d, M2d)
    adj = abs(np.vdot(y, Hx) - np.vdot(Hd(y), x)) / abs(np.vdot(y, Hx))
    # exact values on a few output rows a: out(a,s1,s2,a') = sum L(a,w,b) M1(w,s1,t1,x) M2(x,s2,t2,y) theta(b,t1,t2,c) R(a',y,c)
    rows = rng.choice(chi, size=nrows, replace=False)
    T = np.einsum('awb,btuc->awtuc', L[rows], x)
    T = np.einsum('awtuc,wstx,xvuy->asvyc', T, M1, M2)
    want = np.einsum('asvyc,eyc->asve', T, R)
    sub = np.linalg.norm(Hx[rows] - want) / np.linalg.norm(want)
    return lin, adj, sub


def svd_properties(svd_fn, n, rng, **kw):
    """``svd_fn(x, **kw)`` returns (U, s, Vh) of the truncated SVD."""
    u, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    v, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    s_true = np.exp(-np.arange(n) * (20.0 / n))                # graded like a DMRG bond; known exactly
    x = (u * s_true) @ v.conj().T
    U, s, Vh = svd_fn(x, **kw)
    k = len(s)
    # the reference's rule on the known spectrum (tensors.jl:201-215)
    kk = n
    if kw.get("maxdim", 0) not in (0,) and kw["maxdim"] < n:
        kk = kw["maxdim"]
    if kw.get("cutoff", 0.0) != 0.0:
        s2 = s_true ** 2
        c = np.cumsum(s2[::-1])[::-1] / np.sum(s2)
        keep = int(np.nonzero(c > kw["cutoff"])[0][-1]) + 1
        kk = min(kk, keep)
    return dict(rank_ok=(k == kk), sorted_ok=bool(np.all(np.diff(s) <= 1e-15)), sv_err=float(np.max(np.abs(s - s_true[:k])) / s_true[0]),
                orthU=float(np.linalg.norm(U.conj().T @ U - np.eye(k)) / np.sqrt(k)), orthV=float(np.linalg.norm(Vh @ Vh.conj().T - np.eye(k)) / np.sqrt(k)),
                recon=float(np.linalg.norm((U * s) @ Vh - (u[:, :k] * s_true[:k]) @ v[:, :k].conj().T) / s_true[0]))


def _gpu_product(L, R, M1, M2):
    import tnb200
    chi = L.shape[0]
    dims = [1, chi, chi, chi, 1]
    sites = [np.zeros((dims[i], D, dims[i + 1]), dtype=np.complex128) for i in range(4)]      # placeholders: only the blocks matter
    wd = [1, W, W, W, 1]
    mpo = [np.zeros((wd[0], D, D, wd[1]), dtype=np.complex128), M1, M2, np.zeros((wd[3], D, D, wd[4]), dtype=np.complex128)]
    psi = tnb200.GMPS(1, D, sites, 0)
    psi.center = 2
    H = tnb200.GMPS(2, D, mpo)
    env = tnb200.ProjMPS(psi, H, psi, center=2)
    env.setblock(1, L)
    env.setblock(4, R)
    keep = (psi, H, env)
    return lambda th: (keep, env.product(th, False))[1]


def test_heff_matvec_properties_at_c2_size():
    lin, adj, sub = heff_properties(_gpu_product, CHI, np.random.default_rng(0))
    assert lin < 1e-12 and adj < 1e-11 and sub < 1e-12, (lin, adj, sub)


def test_truncated_svd_properties_at_c2_size():
    import tnb200

    def svd_fn(x, **kw):
        U, S, Vh = tnb200.svd(x, 2, **kw)
        return U, np.real(np.diag(S)), Vh
    for kw in (dict(maxdim=1024), dict(cutoff=1e-12)):
        r = svd_properties(svd_fn, 2048, np.random.default_rng(1), **kw)
        assert r["rank_ok"] and r["sorted_ok"], r
        assert r["sv_err"] < 5e-12 and r["orthU"] < 1e-11 and r["orthV"] < 1e-11 and r["recon"] < 1e-11, r


def test_dmrg_chi4096_on_the_4x6_j1j2_cylinder_reproduces_exact_diagonalisation():
    """BASELINE config 5 / the north-star target at the size where the answer is known.  Two-site DMRG (cutoff = 0) of the J1-J2 model on
    a 4 x 6 cylinder (N = 24, MPO bond 20) with maxdim ramped 64 -> 4096: the central bond of a 24-site chain is exact at 2^12, so the
    energy has to reproduce exact diagonalisation in the S^z = 0 sector (tools/ed_j1j2.py: bit manipulation + ARPACK, no MPS code;
    -47.290880085317, itself checked against the 12-site full-space values of tests/models.py).  The central bond of the last passes is
    bench.py's headline workload: Theta (2048, 2, 2, 2048), w = 20, followed by a 4096 x 4096 factorisation kept at rank 4096.
    Tolerance 1e-10 relative (north_star); measured 7e-13 (profiles/r02b_c5_exact_4x6_chi4096.jsonl)."""
    import ctypes as C
    import tnb200
    from tnb200.mpo import MPO
    from tnb200._lib import check, tn_lanczos_t
    e_ed = -47.290880085317
    ctx = tnb200.Context.default()
    N = 24
    gH = MPO(N, 2, tnb200.models.j1j2_cylinder_terms(4, 6), ctx=ctx)
    g = tnb200.GMPS(1, 2, tnb200.models.random_canonical_mps(N, 2, 16, seed=7), 1)
    g.movecenter(1)
    Hs = tnb200.ProjMPS(g, gH, g, center=1)
    direction, energy, maxbond = False, None, 0
    for chi, passes in ((64, 4), (256, 2), (1024, 2), (2048, 2), (4096, 1)):
        for _ in range(passes):
            e, mb = C.c_double(), C.c_int64()
            check(g.lib.tn_dmrg_sweep(g.h, Hs.h, int(direction), tn_lanczos_t(3, 2, 1e-14), tnb200.Trunc(0.0, chi, 1), C.byref(e), C.byref(mb)))
            direction = not direction
            energy, maxbond = e.value, mb.value
    assert maxbond == 4096
    assert abs(energy - e_ed) <= 1e-10 * abs(e_ed), (energy, e_ed)
