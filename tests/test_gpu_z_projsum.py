"""GPU parity for the projector branch (SURVEY 8(a) A4/A5 fan-out over ProjMPSSum, 8(f) rank 3): project (with and
without an MPO layer, one and two sites), the squared rank-1 penalty, products of sums, the one-site rank-2 product,
excited-state / multi-MPO / one-site DMRG and vmps, all through the C ABI against the oracle.
Tolerances: contractions 1e-13 relative Frobenius, energies 1e-10 relative, overlaps / costs 1e-8.
(The file sorts after the other GPU suites on purpose: it was written in a session without GPU time left.)"""
import numpy as np
import pytest

import oracle
from gpu_util import crandn, random_complex_mps, random_mpo, relerr
from models import tfim, dense_hamiltonian, mps_to_dense

pytestmark = pytest.mark.gpu


def _dense(g):
    ts = g.tensors
    v = ts[0]
    for t in ts[1:]:
        v = np.tensordot(v, t, axes=([v.ndim - 1], [0]))
    return v.reshape(-1)


@pytest.mark.parametrize("N,chiv,chip,w", [(6, 5, 6, 3), (7, 12, 9, 4)])
def test_project_and_one_site_product(N, chiv, chip, w):
    import tnb200
    rng = np.random.default_rng(N + chiv)
    V = random_complex_mps(rng, N, 2, chiv, center=1)
    psi = random_complex_mps(rng, N, 2, chip, center=3)
    H = random_mpo(rng, N, 2, w)
    gV, gpsi, gH = tnb200.GMPS.from_host(V), tnb200.GMPS.from_host(psi), tnb200.GMPS.from_host(H)
    for c in (2, 3, N - 1):
        P = oracle.ProjMPS([V, psi], rank=1, center=c)
        Q = oracle.ProjMPS([V, H, psi], rank=1, center=c)
        E = oracle.ProjMPS([psi, H, psi], rank=2, center=c, coeff=0.7 - 0.2j)
        G = tnb200.ProjMPS(gV, None, gpsi, center=c)
        GQ = tnb200.ProjMPS(gV, gH, gpsi, center=c)
        GE = tnb200.ProjMPS(gpsi, gH, gpsi, coeff=0.7 - 0.2j, center=c)
        for direction in (False, True):
            assert relerr(G.project(None, direction, 2), P.project(None, direction, 2)) < 1e-13
            assert relerr(GQ.project(None, direction, 2), Q.project(None, direction, 2)) < 1e-13
        assert relerr(G.project(None, False, 1), P.project(None, False, 1)) < 1e-13
        assert relerr(GQ.project(None, False, 1), Q.project(None, False, 1)) < 1e-13
        A = crandn(rng, *psi[c].shape)
        assert relerr(GE.product(A, False, 1), E.product(A, False, 1)) < 1e-13


def test_squared_projector_and_sum_product():
    import tnb200
    rng = np.random.default_rng(21)
    N = 7
    V = random_complex_mps(rng, N, 2, 6, center=1)
    V2 = random_complex_mps(rng, N, 2, 3, center=1)
    psi = random_complex_mps(rng, N, 2, 8, center=4)
    Ha, Hb = random_mpo(rng, N, 2, 3), random_mpo(rng, N, 2, 5)
    gV, gV2, gpsi = tnb200.GMPS.from_host(V), tnb200.GMPS.from_host(V2), tnb200.GMPS.from_host(psi)
    gHa, gHb = tnb200.GMPS.from_host(Ha), tnb200.GMPS.from_host(Hb)
    c = 4
    S = oracle.ProjMPS([V, psi], rank=2, squared=True, coeff=2.5, center=c)
    GS = tnb200.ProjMPS(gV, None, gpsi, coeff=2.5, center=c, squared=True)
    for nsites, direction in ((2, False), (2, True), (1, False)):
        site = c - nsites + 1 if direction else c
        shape = (psi[site].shape[0],) + (2,) * nsites + (psi[site + nsites - 1].shape[2],)
        A = crandn(rng, *shape)
        assert relerr(GS.product(A, direction, nsites), S.product(A, direction, nsites)) < 1e-12
    assert abs(GS.calculate() - S.calculate()) < 1e-12
    # sum of two MPO terms and two penalties
    osum = oracle.ProjMPSSum([oracle.ProjMPS([psi, Ha, psi], rank=2, center=c, coeff=0.3),
                              oracle.ProjMPS([psi, Hb, psi], rank=2, center=c, coeff=-1.1),
                              S, oracle.ProjMPS([V2, psi], rank=2, squared=True, coeff=0.4, center=c)], center=c)
    gsum = tnb200.ProjMPSSum([tnb200.ProjMPS(gpsi, gHa, gpsi, coeff=0.3, center=c), tnb200.ProjMPS(gpsi, gHb, gpsi, coeff=-1.1, center=c),
                              GS, tnb200.ProjMPS(gV2, None, gpsi, coeff=0.4, center=c, squared=True)], center=c)
    for nsites in (2, 1):
        shape = (psi[c].shape[0],) + (2,) * nsites + (psi[c + nsites - 1].shape[2],)
        A = crandn(rng, *shape)
        assert relerr(gsum.product(A, False, nsites), osum.product(A, False, nsites)) < 1e-12
    assert abs(gsum.calculate() - osum.calculate()) < 1e-11 * abs(osum.calculate())
    gsum.movecenter(2)
    osum.movecenter(2)
    A = crandn(rng, psi[2].shape[0], 2, 2, psi[3].shape[2])
    assert relerr(gsum.product(A, False, 2), osum.product(A, False, 2)) < 1e-12


def test_excited_state_dmrg_matches_oracle_and_ed():
    import tnb200
    sh = oracle.spinhalf()
    N = 8
    H = tfim(N)
    ev = np.linalg.eigvalsh(dense_hamiltonian(sh, H).toarray())
    M = oracle.MPO(sh, H)
    p0 = oracle.randomMPS(2, N, 4, np.random.default_rng(1))
    p1 = oracle.randomMPS(2, N, 4, np.random.default_rng(2))
    gM = tnb200.GMPS.from_host(M)
    g0, E0 = tnb200.dmrg(tnb200.GMPS.from_host(p0), gM, maxdim=32, cutoff=1e-14, maxsweeps=20)
    assert abs(E0 - ev[0]) < 1e-10 * abs(ev[0])
    ho, hg = [], []
    o0, _ = oracle.dmrg(p0.copy(), M, maxdim=32, cutoff=1e-14, maxsweeps=20)
    oracle.dmrg(p1.copy(), M, o0, coeffs=[1.0, 20.0], maxdim=32, cutoff=1e-14, maxsweeps=30, history=ho)
    g1, E1 = tnb200.dmrg(tnb200.GMPS.from_host(p1), gM, g0, coeffs=[1.0, 20.0], maxdim=32, cutoff=1e-14, maxsweeps=30, history=hg)
    assert abs(E1 - ev[1]) < 1e-9 * abs(ev[1])
    assert abs(E1 - ho[-1][1]) < 1e-10 * abs(ho[-1][1])
    assert abs(np.vdot(_dense(g0), _dense(g1))) < 1e-7


def test_two_mpo_terms_and_one_site_dmrg():
    import tnb200
    sh = oracle.spinhalf()
    N = 8
    Ha, Hb = oracle.OpList(N), oracle.OpList(N)
    for i in range(1, N + 1):
        Ha.add("x", i, 1.0)
        Ha.add("z", i, 0.05)
    for i in range(1, N):
        Hb.add(["z", "z"], [i, i + 1], 1.2)
    E_ref = np.linalg.eigvalsh(dense_hamiltonian(sh, tfim(N)).toarray())[0]
    Ma, Mb = oracle.MPO(sh, Ha), oracle.MPO(sh, Hb)
    p = oracle.randomMPS(2, N, 4, np.random.default_rng(7))
    g, E = tnb200.dmrg(tnb200.GMPS.from_host(p), tnb200.GMPS.from_host(Ma), tnb200.GMPS.from_host(Mb), maxdim=32, cutoff=1e-14, maxsweeps=20)
    assert abs(E - E_ref) < 1e-10 * abs(E_ref)
    # one-site DMRG keeps the bond dimension: same sweeps as the oracle from the same chi = 8 start
    q = oracle.randomMPS(2, N, 8, np.random.default_rng(8))
    ho, hg = [], []
    oracle.dmrg(q.copy(), oracle.MPO(sh, tfim(N)), nsites=1, maxsweeps=6, history=ho)
    tnb200.dmrg(tnb200.GMPS.from_host(q), tnb200.GMPS.from_host(oracle.MPO(sh, tfim(N))), nsites=1, maxsweeps=6, history=hg)
    assert len(ho) == len(hg)
    for a, b in zip(ho, hg):
        assert a[2] == b[2] and abs(a[1] - b[1]) < 1e-7 * abs(a[1])
    assert abs(ho[-1][1] - hg[-1][1]) < 1e-10 * abs(ho[-1][1])


def test_vmps_matches_oracle():
    import tnb200
    rng = np.random.default_rng(12)
    N = 8
    a = random_complex_mps(rng, N, 2, 6, center=1)
    b = random_complex_mps(rng, N, 2, 6, center=1)
    target = mps_to_dense(a) + mps_to_dense(b)
    for nsites, maxdim in ((2, 16), (2, 4), (1, 16)):
        ho, hg = [], []
        po = oracle.vmps(a, b, nsites=nsites, maxdim=maxdim, cutoff=0.0 if maxdim == 4 else 1e-14, maxsweeps=6, history=ho)
        pg = tnb200.vmps(tnb200.GMPS.from_host(a), tnb200.GMPS.from_host(b), nsites=nsites, maxdim=maxdim,
                         cutoff=0.0 if maxdim == 4 else 1e-14, maxsweeps=6, history=hg)
        assert len(ho) == len(hg)
        assert abs(hg[-1][1] - ho[-1][1]) < 1e-8 * abs(ho[-1][1])
        assert hg[-1][2] == ho[-1][2]
        vo, vg = mps_to_dense(po), _dense(pg)
        assert abs(abs(np.vdot(vo, vg)) - np.linalg.norm(vo) * np.linalg.norm(vg)) < 1e-8 * np.linalg.norm(vo) ** 2
        if maxdim == 16 and nsites == 2:
            assert np.linalg.norm(vg - target) < 1e-9 * np.linalg.norm(target)


def test_eigsolve_with_caller_map_matches_fused_eigsolve():
    """tn_eigsolve_fn (caller-supplied linear map on device vectors) with the map = tn_env_product_dev reproduces tn_eigsolve."""
    import ctypes as C
    import torch
    import tnb200
    from tnb200 import _lib
    sh = oracle.spinhalf()
    H = oracle.MPO(sh, tfim(8))
    psi = random_complex_mps(np.random.default_rng(3), 8, 2, 8, center=4)
    gpsi, gH = tnb200.GMPS.from_host(psi), tnb200.GMPS.from_host(H)
    G = tnb200.ProjMPS(gpsi, gH, gpsi, center=4)
    A0 = np.tensordot(psi[4], psi[5], axes=([2], [0]))
    e_ref, v_ref, nops_ref = G.eigsolve(A0, False)
    ctx = tnb200.Context.default()
    lib = ctx.lib
    th0 = torch.from_numpy(np.reshape(A0, -1, order='F').copy()).cuda()
    th1 = torch.zeros_like(th0)
    torch.cuda.synchronize()
    calls = []

    def cb(_u, pin, pout):
        calls.append(1)
        return lib.tn_env_product_dev(G.h, C.c_void_p(pin), 0, C.c_void_p(pout), 1)
    e, nops = C.c_double(), C.c_int32()
    _lib.check(lib.tn_eigsolve_fn(ctx.h, th0.numel(), C.c_void_p(th0.data_ptr()), C.c_void_p(th1.data_ptr()),
                                  _lib.tn_lanczos_t(3, 2, 1e-14), _lib.APPLY_FN(cb), None, C.byref(e), C.byref(nops)))
    assert nops.value == nops_ref == len(calls)
    assert abs(e.value - e_ref) < 1e-12 * abs(e_ref)
    v = th1.cpu().numpy().reshape(A0.shape, order='F')
    assert abs(abs(np.vdot(v, v_ref)) - 1.0) < 1e-10


def test_sharded_dmrg_world1_matches_unsharded():
    """The sharded sweep's GPU backend (tn_contract_strided_dev, tn_eigsolve_fn, tn_mps_site_ptr, tn_mps_replacesites_dev) at
    world size 1 against the fused single-GPU sweep and the oracle."""
    import tnb200
    from tnb200.sharded import GpuBackend, sharded_dmrg
    from models import xxz
    sh = oracle.spinhalf()
    N = 10
    M = oracle.MPO(sh, xxz(N, 1.0))
    p0 = oracle.randomMPS(2, N, 4, np.random.default_rng(1))
    ho, hg, hs = [], [], []
    oracle.dmrg(p0.copy(), M, maxdim=24, maxsweeps=4, history=ho)
    tnb200.dmrg(tnb200.GMPS.from_host(p0), tnb200.GMPS.from_host(M), maxdim=24, maxsweeps=4, history=hg)
    ctx = tnb200.Context.default()
    sharded_dmrg(tnb200.GMPS.from_host(p0), [M[i] for i in range(1, N + 1)], GpuBackend(ctx, "cuda"), maxdim=24, maxsweeps=4, history=hs)
    assert len(ho) == len(hs) == len(hg)
    for a, b, c in zip(ho, hs, hg):
        assert a[2] == b[2] == c[2]
        assert abs(a[1] - b[1]) < 1e-10 * abs(a[1]) and abs(c[1] - b[1]) < 1e-10 * abs(a[1])


@pytest.mark.parametrize("name", ["tfim", "xxz", "j1j2"])
def test_mpo_builder_with_device_compression(name):
    """MPO(st, H) (mpo.jl:323-459): host assembly + tn_mpo_compress; same operator and bond dimensions as the oracle's
    builder, and DMRG on it reaches the exact-diagonalisation energy."""
    import tnb200
    from tnb200.mpo import MPO
    from models import xxz, j1j2_cylinder, KAT
    sh = oracle.spinhalf()
    H, key = {"tfim": (tfim(8), ("tfim", 8)), "xxz": (xxz(10, 0.5), ("xxz0.5", 10)), "j1j2": (j1j2_cylinder(4, 3), ("j1j2_4x3", 12))}[name]
    N = len(H)
    terms = [([sh.op(o) for o in ops], sites, c) for ops, sites, c in zip(H.ops, H.sites, H.coeffs)]
    g = MPO(N, 2, terms)
    ref = oracle.MPO(sh, H)
    assert [int(x[3]) for x in g.dims()] == [ref[i].shape[3] for i in range(1, N + 1)]
    if N <= 10:
        ts = g.tensors
        t = ts[0]
        for x in ts[1:]:
            t = np.tensordot(t, x, axes=([t.ndim - 1], [0]))
        t = np.transpose(t[0, ..., 0], list(range(0, 2 * N, 2)) + list(range(1, 2 * N, 2))).reshape(2 ** N, 2 ** N)
        assert np.abs(t - dense_hamiltonian(sh, H).toarray()).max() < 1e-11
    p0 = oracle.randomMPS(2, N, 4, np.random.default_rng(1234))
    _, E = tnb200.dmrg(tnb200.GMPS.from_host(p0), g, maxdim=64, cutoff=1e-14, maxsweeps=30)
    assert abs((E - KAT[key]) / KAT[key]) < 1e-10


def test_tebd_front_end_matches_oracle():
    """tnb200.evolve.tebd (tebd.jl:3-100: trotterize on the host, gates / norm / energy on the device) against oracle.tebd."""
    import tnb200
    from tnb200.evolve import tebd
    sh = oracle.spinhalf()
    N = 8
    H = -1 * tfim(N, 1.0, 0.0, 1.0)
    terms = [([sh.op(o) for o in ops], sites, c) for ops, sites, c in zip(H.ops, H.sites, H.coeffs)]
    p0 = oracle.randomMPS(2, N, 4, np.random.default_rng(5))

    class Obs:
        def __init__(self):
            self.rec = []

        def measure(self, time, psi, norm, energy):
            self.rec.append((time, norm, energy, psi.maxbonddim()))

        def checkdone(self):
            return False
    oo, og = Obs(), Obs()
    _, Eo = oracle.tebd(sh, p0.copy(), H, 0.02, 0.4, 0.1, observers=[oo], cutoff=1e-12, maxdim=16)
    _, Eg = tebd(tnb200.GMPS.from_host(p0), terms, 0.02, 0.4, 0.1, observers=[og], cutoff=1e-12, maxdim=16)
    assert len(oo.rec) == len(og.rec) == 5
    for a, b in zip(oo.rec, og.rec):
        assert a[0] == b[0] and abs(a[3] - b[3]) <= 1      # a singular value at the 1e-12 cutoff edge may fall either side
        assert abs(a[1] - b[1]) < 1e-9 * max(1.0, abs(a[1])) and abs(a[2] - b[2]) < 1e-9 * abs(a[2])
    assert abs(Eo - Eg) < 1e-9 * abs(Eo)


def test_qjmc_norm_based_branch_matches_oracle():
    """tn_qjmc_run with classical = 0 (qjmc.jl:65-87) against the oracle with the same per-step uniforms."""
    import tnb200
    sh = oracle.spinhalf()
    N, dt, steps = 7, 0.02, 80
    H = tfim(N, 1.0, 2.0, 1.0)
    J = oracle.OpList(N)
    for i in range(1, N + 1):
        J.add("s-", i, np.sqrt(0.9))
    u = np.random.default_rng(12).random(3 * steps)

    class Feeder:                                     # the oracle draws sequentially; the C ABI indexes 3 per step
        def __init__(self):
            self.step, self.slot = 0, 0

        def __call__(self):
            v = u[3 * self.step + self.slot]
            self.slot += 1
            return v
    f = Feeder()
    psi = oracle.productMPS(sh, ["up" if i % 2 else "dn" for i in range(1, N + 1)])
    psi.movecenter(1)
    zs = oracle.OpList(N)
    for i in range(1, N + 1):
        zs.add("z", i)

    class Obs:
        def __init__(self):
            self.m = []

        def measure(self, time, p, jumps, jt):
            self.m.append(np.real(oracle.inner(sh, p, zs, p)))
            f.step, f.slot = len(self.m) - 1, 0
    ob = Obs()
    kw = dict(cutoff=1e-10, maxdim=16)
    jumps, times = oracle.qjmc_simulation(sh, psi, H, J, steps * dt, dt, [ob], uniforms=f, classical=False, **kw)
    _, gl = oracle.qjmc_gates(sh, H, J, dt)
    g = tnb200.GMPS.from_host(oracle.productMPS(sh, ["up" if i % 2 else "dn" for i in range(1, N + 1)]))
    g.movecenter(1)
    gg = tnb200.GateList.from_host(2, gl)
    gj, gt, obs = tnb200.qjmc_simulation(g, gg, list(range(1, N + 1)), [sh.op("s-")] * N, [np.sqrt(0.9)] * N, steps, dt,
                                         uniforms=u, obs_op=sh.op("z"), save_every=1, classical=False, **kw)
    assert gj == jumps and len(jumps) > 0
    assert np.allclose(gt, times)
    assert np.max(np.abs(np.real(obs) - np.array(ob.m[1:]))) < 1e-8


@pytest.mark.parametrize("B,m,n,kw", [(5, 48, 48, dict()), (4, 96, 64, dict(cutoff=1e-10)), (3, 200, 130, dict(maxdim=70)),
                                      (6, 130, 260, dict(cutoff=1e-12, maxdim=100)), (8, 256, 256, dict(maxdim=128))])
def test_batched_svd_matches_lapack_and_single(B, m, n, kw):
    """tn_svd_trunc_batched: B same-shape problems in one batched factorisation (stacked workspace, one launch per pipeline
    stage) against LAPACK singular values, the reference's truncation rule and reconstruction of the truncated matrix."""
    import tnb200
    rng = np.random.default_rng(B * 1000 + m + n)
    mats = []
    for b in range(B):
        x = crandn(rng, m, n)
        if b % 2 == 1:                                   # graded spectrum: the truncation rule has something to cut
            u, s, vh = np.linalg.svd(x, full_matrices=False)
            x = (u * (s[0] * np.exp(-np.arange(len(s)) * (24.0 / len(s))))) @ vh
        mats.append(x)
    res = tnb200.svd_batched(np.stack(mats), **kw)
    assert len(res) == B
    for x, (U, S, Vh) in zip(mats, res):
        _, So, _ = oracle.svd(x, 2, **kw)
        so = np.real(np.diag(So))
        assert S.shape == so.shape, (S.shape, so.shape)
        assert np.max(np.abs(S - so)) < 1e-12 * so[0]
        k = len(S)
        assert np.linalg.norm(U.conj().T @ U - np.eye(k)) < 1e-11 * np.sqrt(k)
        assert np.linalg.norm(Vh @ Vh.conj().T - np.eye(k)) < 1e-11 * np.sqrt(k)
        u, s, vh = np.linalg.svd(x, full_matrices=False)
        want = (u[:, :k] * s[:k]) @ vh[:k]
        assert np.linalg.norm((U * S) @ Vh - want) < 1e-11 * s[0] * np.sqrt(k)


def test_qjmc_ensemble_batching_rounds_equal_plain_ensemble():
    """TN_QJMC_BATCH=1: the ensemble workers' SVDs go through SvdBatcher rounds (svd_batched_factor per shape group); jump records
    and observables must equal the plain multi-stream ensemble.  Runs in a subprocess with a time limit."""
    import json
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "run_batched_ensemble.py")], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    a, b = out["0"], out["1"]
    assert a["nj"] == b["nj"] and a["jumps"] == b["jumps"] and sum(a["nj"]) > 0
    assert np.allclose(np.array(a["obs"]), np.array(b["obs"]), atol=1e-9)


def test_sharded_heff_pipelined_world1_matches_plain():
    import torch
    import tnb200
    from tnb200.sharded import ShardedHeff, GpuContractor
    rng = np.random.default_rng(0)
    chi, d, w, w1, w2 = 48, 2, 7, 6, 5
    L, R = crandn(rng, chi, w, chi), crandn(rng, chi, w2, chi)
    M1, M2 = crandn(rng, w, d, d, w1), crandn(rng, w1, d, d, w2)
    theta = crandn(rng, chi, d, d, chi)
    want = np.einsum('awb,wstx,xuvy,btvc,eyc->asue', L, M1, M2, theta, R)
    ctx = tnb200.Context.default()
    sh = ShardedHeff(L, R, M1, M2, 0, 1, GpuContractor(ctx), "cuda")
    th = torch.from_numpy(np.reshape(theta, -1, order='F').copy()).cuda()
    torch.cuda.synchronize()
    for ns in (1, 3, 4):
        out = sh.apply_pipelined(th, ns, ctx.stream()).cpu().numpy().reshape(chi, d, d, chi, order='F')
        assert relerr(out, want) < 1e-13


def test_qjmc_front_end_with_observers_equals_single_call():
    import tnb200
    from tnb200 import models
    from tnb200.evolve import qjmc, QJMCOperators, QJMCEntropy, QJMCActivity
    N, d, dt, steps, chi = 8, 2, 0.02, 40, 8
    gamma = 0.9
    onsite = -1j * (1.0 * models.X + 2.0 * models.Z) - 0.5 * gamma * (models.SM.conj().T @ models.SM)
    bond = -1j * 1.0 * np.kron(models.Z, models.Z)
    ss, gg = models.trotter_gates(N, onsite, bond, dt, evol="imag", order=2)
    tens = models.random_canonical_mps(N, d, chi, seed=5)
    u = np.random.default_rng(3).random(3 * steps)
    u[[3 * 5 + 1, 3 * 17 + 1, 3 * 31 + 1]] = 0.99999      # the jump test of these steps always fires (u1 > exp(-sum(rates) dt))
    gl = tnb200.GateList(d, ss, gg)
    args = (list(range(1, N + 1)), [models.SM] * N, [np.sqrt(gamma)] * N)
    a = tnb200.GMPS(1, d, tens, 1)
    j1, t1, o1 = tnb200.qjmc_simulation(a, gl, *args, steps, dt, uniforms=u, obs_op=models.Z, save_every=10, cutoff=1e-12, maxdim=chi)
    b = tnb200.GMPS(1, d, tens, 1)
    zterms = [([models.Z], [i], 1.0) for i in range(1, N + 1)]
    obs, ent, act = QJMCOperators(zterms), QJMCEntropy(), QJMCActivity()
    j2, t2 = qjmc(b, gl, *args, steps * dt, dt, observers=[obs, act], save=10 * dt, uniforms=u, cutoff=1e-12, maxdim=chi)
    assert j1 == j2 and np.allclose(t1, t2) and len(j1) > 0
    assert len(obs.times) == 5 and act.jumps == len(j2)
    assert np.max(np.abs(np.real(np.array(obs.measurements[1:])) - np.real(o1))) < 1e-9
    # the entropy observer moves the orthogonality centre (entropy() does, gmps.jl:184-189), so it gets its own short run
    c = tnb200.GMPS(1, d, tens, 1)
    qjmc(c, gl, *args, 10 * dt, dt, observers=[ent], save=5 * dt, uniforms=u, cutoff=1e-12, maxdim=chi)
    assert len(ent.times) == 3 and all(len(e) == N - 1 and min(e) > -1e-12 for e in ent.measurements)


def test_applygates_fidelity_matches_oracle():
    import tnb200
    from models import xxz
    sh = oracle.spinhalf()
    N = 8
    gl = oracle.trotterize(sh, -1 * xxz(N, 0.7), 0.05, order=2)        # no on-site terms: two-site gates only
    rng = np.random.default_rng(9)
    psi = random_complex_mps(rng, N, 2, 16, center=1)
    g = tnb200.GMPS.from_host(psi)
    gg = tnb200.GateList.from_host(2, gl)
    for kw in (dict(cutoff=1e-12, maxdim=0), dict(cutoff=0.0, maxdim=6)):
        eo = oracle.applygates(psi, gl, error=True, **kw)
        eg = tnb200.applygates(g, gg, error=True, **kw)
        assert abs(eo - eg) < 1e-9 * max(abs(eo), 1e-30), (eo, eg)
        assert psi.maxbonddim() == g.maxbonddim()


def test_itebd_step_matches_oracle():
    """tn_itebd_apply_gate (itebd.jl:71-119, two-site cell) against the oracle: Schmidt values of both bonds, accumulated log-norms and
    bond energy after the same number of steps; and the exact TFIM energy per site."""
    from scipy.integrate import quad
    from oracle.itebd import iMPS, itebd_gate, itebd_apply_gates_mps, bond_energy
    from tnb200.evolve import IGMPS
    sh = oracle.spinhalf()
    g = 2.0
    H = oracle.OpList(2)
    H.add(["z", "z"], [1, 2], -1.0)
    H.add("x", 1, -g)
    h2 = H.sitetensor(sh, 1)
    po = iMPS(2, np.array([1.0, 0.3]))
    pg = IGMPS.product(2, np.array([1.0, 0.3]))
    for dt, n in ((0.05, 60), (0.01, 60)):
        gate = itebd_gate(sh, -1 * H, dt)
        for _ in range(n):
            itebd_apply_gates_mps(po, gate, maxdim=8, cutoff=1e-12)
        pg.apply_gate(gate, n, cutoff=1e-12, maxdim=8)
        for i in (1, 2):
            t, s, nrm = pg.site(i)
            assert s.shape == po.singulars[i - 1].shape
            assert np.max(np.abs(s - po.singulars[i - 1])) < 1e-8
            assert abs(nrm - po.norms[i - 1]) < 1e-8 * max(1.0, abs(po.norms[i - 1]))
        assert abs(pg.bond_energy(h2) - bond_energy(po, h2)) < 1e-8
    exact = -quad(lambda k: np.sqrt(1 + g * g - 2 * g * np.cos(k)), -np.pi, np.pi)[0] / (2 * np.pi)
    pg.apply_gate(itebd_gate(sh, -1 * H, 0.01), 300, cutoff=1e-12, maxdim=8)
    assert abs(pg.bond_energy(h2).real - exact) < 1e-4 * abs(exact)


# ---- committed golden vectors of the projector branch (tests/golden/projector_golden.npz; no oracle call on the checked values) ----
def _golden2():
    import os
    P2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "projector_golden.npz"))

    def lst(prefix):
        out, i = [], 0
        while f"{prefix}{i}" in P2:
            out.append(P2[f"{prefix}{i}"])
            i += 1
        return out
    return P2, lst


def test_projector_branch_against_golden():
    import tnb200
    P2, lst = _golden2()
    V, psi, H = tnb200.GMPS(1, 2, lst("pj_V"), 1), tnb200.GMPS(1, 2, lst("pj_psi"), 3), tnb200.GMPS(2, 2, lst("pj_H"), 0)
    assert relerr(tnb200.ProjMPS(V, None, psi, center=3).project(None, False, 2), P2["pj_project2"]) < 1e-13
    assert relerr(tnb200.ProjMPS(V, H, psi, center=3).project(None, False, 2), P2["pj_project2_mpo"]) < 1e-13
    assert relerr(tnb200.ProjMPS(V, H, psi, center=3).project(None, False, 1), P2["pj_project1_mpo"]) < 1e-13
    sq = tnb200.ProjMPS(V, None, psi, coeff=2.5, center=3, squared=True).product(P2["pj_theta"], False, 2)
    assert relerr(sq, P2["pj_squared_out"]) < 1e-12
    p1 = tnb200.ProjMPS(psi, H, psi, coeff=0.7 - 0.2j, center=3).product(P2["pj_A1"], False, 1)
    assert relerr(p1, P2["pj_product1"]) < 1e-13


def test_excited_dmrg_and_vmps_against_golden():
    import tnb200
    P2, lst = _golden2()
    M = tnb200.GMPS(2, 2, lst("ex_mpo"), 0)
    g0 = tnb200.GMPS(1, 2, lst("ex_gs"), int(P2["ex_gs_center"]))
    p1 = tnb200.GMPS(1, 2, lst("ex_start"), 1)
    hist = []
    tnb200.dmrg(p1, M, g0, coeffs=[1.0, 20.0], maxdim=32, cutoff=1e-14, maxsweeps=30, history=hist)
    e, want = np.array([h[1] for h in hist]), P2["ex_energy"]
    assert e.shape == want.shape
    assert np.max(np.abs((e - want) / want)) < 1e-7 and abs((e[-1] - want[-1]) / want[-1]) < 1e-10
    a, b = tnb200.GMPS(1, 2, lst("vm_a"), 1), tnb200.GMPS(1, 2, lst("vm_b"), 1)
    hist = []
    tnb200.vmps(a, b, maxdim=4, cutoff=0.0, maxsweeps=6, history=hist)
    c = np.array([h[1] for h in hist])
    assert np.max(np.abs(c - P2["vm_cost"])) < 1e-8 * np.max(np.abs(P2["vm_cost"]))
    assert [h[2] for h in hist] == list(P2["vm_maxbond"])


def test_itebd_against_golden():
    from tnb200.evolve import IGMPS
    P2, _ = _golden2()
    pg = IGMPS.product(2, np.array([1.0, 0.3]))
    pg.apply_gate(P2["it_gate"], 40, cutoff=1e-12, maxdim=8)
    for i, key in ((1, "it_sing1"), (2, "it_sing2")):
        _, s, nrm = pg.site(i)
        assert s.shape == P2[key].shape and np.max(np.abs(s - P2[key])) < 1e-8
        assert abs(nrm - P2["it_norms"][i - 1]) < 1e-8 * max(1.0, abs(P2["it_norms"][i - 1]))


def test_thermal_example_energy_matches_oracle():
    """examples/thermal.jl at N = 8: evolve the identity MPO in imaginary time on the device, then
    trace(H, adjoint(U), U) / trace(adjoint(U), U) through tnb200.evolve.thermal_energy against the oracle's trace()."""
    import tnb200
    from tnb200.evolve import thermal_energy
    sh = oracle.spinhalf()
    N, dt, steps = 8, 0.02, 15
    Hl = tfim(N, 1.0, 0.0, 1.0)
    gl = oracle.trotterize(sh, -1 * Hl, dt)
    U = oracle.productMPO(sh, ["id"] * N)
    g = tnb200.GMPS.from_host(U)
    gg = tnb200.GateList.from_host(2, gl)
    for _ in range(steps):
        oracle.applygates(U, gl, cutoff=1e-10, maxdim=16)
        tnb200.applygates(g, gg, cutoff=1e-10, maxdim=16)
    M = oracle.MPO(sh, Hl)
    want = oracle.trace(M, oracle.adjoint(U), U) / oracle.trace(oracle.adjoint(U), U)
    got = thermal_energy(g, M.tensors)
    assert abs(got - want) < 1e-8 * abs(want)


def test_applympo_matches_oracle():
    import tnb200
    rng = np.random.default_rng(0)
    N = 7
    O, psi = random_mpo(rng, N, 2, 3), random_complex_mps(rng, N, 2, 5, center=2)
    gO, gpsi = tnb200.GMPS.from_host(O), tnb200.GMPS.from_host(psi)
    for kw in (dict(), dict(cutoff=1e-10), dict(maxdim=6)):
        want = oracle.applyMPO(O, psi, **kw)
        got = tnb200.applyMPO(gO, gpsi, **kw)
        assert got.center == 1 and [int(x[0]) for x in got.dims()] == [t.shape[0] for t in want.tensors]
        vo, vg = mps_to_dense(want), _dense(got)
        assert abs(np.vdot(vo, vg) - np.vdot(vo, vo)) < 1e-9 * abs(np.vdot(vo, vo))          # same vector (gauge-independent)
        assert abs(np.linalg.norm(vg) - np.linalg.norm(vo)) < 1e-9 * np.linalg.norm(vo)


def test_tebd_with_projector_matches_oracle():
    """tebd.jl:22-41,67-73 on the device (MPSProjector overlap + scaled copy, tn_vmps_sweep) against the oracle and exact diagonalisation."""
    import tnb200
    from tnb200.evolve import tebd as gtebd
    from oracle.tebd import tebd as otebd
    sh = oracle.spinhalf()
    N = 8
    Hl = tfim(N)
    ev = np.linalg.eigvalsh(dense_hamiltonian(sh, Hl).toarray())
    g0, _ = oracle.dmrg(oracle.randomMPS(2, N, 4, np.random.default_rng(1)), oracle.MPO(sh, Hl), maxdim=32, cutoff=1e-14, maxsweeps=20)
    p = oracle.randomMPS(2, N, 4, np.random.default_rng(3))
    Hm = -1 * Hl
    terms = [([sh.op(o) for o in ops], sites, c) for ops, sites, c in zip(Hm.ops, Hm.sites, Hm.coeffs)]
    _, Eo = otebd(sh, p.copy(), Hm, 0.02, 2.0, 1.0, projectors=[g0], cutoff=1e-12, maxdim=16, projection_every=5)
    psi, Eg = gtebd(tnb200.GMPS.from_host(p), terms, 0.02, 2.0, 1.0, projectors=[tnb200.GMPS.from_host(g0)], cutoff=1e-12, maxdim=16,
                    projection_every=5)
    assert abs(Eo - Eg) < 1e-8 * abs(Eo)
    assert abs(tnb200.GMPS.from_host(g0).overlap(psi)) < 1e-9
    psi, Eg = gtebd(psi, terms, 0.02, 4.0, 1.0, projectors=[tnb200.GMPS.from_host(g0)], cutoff=1e-12, maxdim=16, projection_every=5)
    assert abs(-Eg - ev[1]) < 1e-4 * abs(ev[1])


def test_three_call_svd_with_external_schedule_matches_fused_svd():
    """tn_svd_dist_begin / _step / _finish driven by dist_jacobi_sweeps at world 1 (round-robin schedule from Python) against the fused
    tn_svd_trunc and LAPACK; then the sharded DMRG sweep with the SVD engine against the fused sweep."""
    import torch
    import tnb200
    from tnb200.sharded import GpuSvdEngine, GpuBackend, dist_jacobi_sweeps, sharded_dmrg
    from models import xxz
    ctx = tnb200.Context.default()
    eng = GpuSvdEngine(ctx, "cuda")
    rng = np.random.default_rng(5)
    for (m, n, kw) in ((200, 200, dict()), (300, 130, dict(cutoff=1e-10)), (96, 260, dict(maxdim=50))):
        x = crandn(rng, m, n)
        u, s, vh = np.linalg.svd(x, full_matrices=False)
        x = (u * (s[0] * np.exp(-np.arange(len(s)) * (16.0 / len(s))))) @ vh
        xd = torch.from_numpy(np.reshape(x, -1, order='F').copy()).cuda()
        torch.cuda.synchronize()
        nb, tol = eng.begin(xd, m, n)
        sweeps = dist_jacobi_sweeps(eng, nb, tol, 0, 1, None)
        k = eng.finish(kw.get("cutoff", 0.0), kw.get("maxdim", 0), 1, sweeps)
        _, So, _ = oracle.svd(x, 2, **kw)
        so = np.real(np.diag(So))
        assert k == len(so) and sweeps < 30
        U = torch.zeros(m * k, dtype=torch.complex128, device="cuda")
        S = torch.zeros(k, dtype=torch.float64, device="cuda")
        Vh = torch.zeros(k * n, dtype=torch.complex128, device="cuda")
        torch.cuda.synchronize()
        import ctypes as C
        tnb200._lib.check(ctx.lib.tn_svd_dist_factors(ctx.h, C.c_void_p(U.data_ptr()), C.c_void_p(S.data_ptr()), C.c_void_p(Vh.data_ptr())))
        Uh, Sh, Vhh = U.cpu().numpy().reshape(m, k, order='F'), S.cpu().numpy(), Vh.cpu().numpy().reshape(k, n, order='F')
        assert np.max(np.abs(Sh - so)) < 1e-12 * so[0]
        uu, ss, vv = np.linalg.svd(x, full_matrices=False)
        assert np.linalg.norm((Uh * Sh) @ Vhh - (uu[:, :k] * ss[:k]) @ vv[:k]) < 1e-11 * ss[0] * np.sqrt(k)
    sh = oracle.spinhalf()
    N = 10
    M = oracle.MPO(sh, xxz(N, 1.0))
    p0 = oracle.randomMPS(2, N, 4, np.random.default_rng(1))
    hg, hs = [], []
    tnb200.dmrg(tnb200.GMPS.from_host(p0), tnb200.GMPS.from_host(M), maxdim=24, maxsweeps=3, history=hg)
    sharded_dmrg(tnb200.GMPS.from_host(p0), [M[i] for i in range(1, N + 1)], GpuBackend(ctx, "cuda"), maxdim=24, maxsweeps=3, history=hs,
                 svd_engine=eng)
    for a, b in zip(hg, hs):
        assert a[2] == b[2] and abs(a[1] - b[1]) < 1e-10 * abs(a[1])


def test_balanced_sharded_heff_world1_matches_einsum():
    import torch
    import tnb200
    from tnb200.sharded import BalancedShardedHeff, GpuContractor
    rng = np.random.default_rng(0)
    chi, d, w, w1, w2 = 48, 2, 7, 6, 5
    L, R = crandn(rng, chi, w, chi), crandn(rng, chi, w2, chi)
    M1, M2 = crandn(rng, w, d, d, w1), crandn(rng, w1, d, d, w2)
    theta = crandn(rng, chi, d, d, chi)
    want = np.einsum('awb,wstx,xuvy,btvc,eyc->asue', L, M1, M2, theta, R)
    ctx = tnb200.Context.default()
    sh = BalancedShardedHeff(L, R, M1, M2, 0, 1, GpuContractor(ctx), "cuda")
    th = torch.from_numpy(np.reshape(theta, -1, order='F').copy()).cuda()
    torch.cuda.synchronize()
    out = sh.apply(th).cpu().numpy().reshape(chi, d, d, chi, order='F')
    assert relerr(out, want) < 1e-13
