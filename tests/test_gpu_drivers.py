"""GPU end-to-end parity: DMRG energies (1e-10 relative, exact-diagonalisation KATs), TEBD gate
application (spectra / observables 1e-8), thermal-MPO gates (rank 2) and a QJMC trajectory with
host-supplied uniforms -- the north_star tolerances."""
import numpy as np
import pytest

import oracle
from gpu_util import random_complex_mps, relerr
from models import tfim, xxz, j1j2_cylinder, KAT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key,H", [(("tfim", 10), tfim(10)), (("heis", 10), xxz(10, 1.0)), (("j1j2_4x3", 12), j1j2_cylinder(4, 3))])
def test_dmrg_energy_matches_ed_and_oracle(key, H):
    import tnb200
    sh = oracle.spinhalf()
    M = oracle.MPO(sh, H)
    psi0 = oracle.randomMPS(2, len(H), 4, np.random.default_rng(1234))
    ho, hg = [], []
    po = psi0.copy()
    oracle.dmrg(po, M, maxdim=64, cutoff=1e-14, maxsweeps=12, history=ho)
    g = tnb200.GMPS.from_host(psi0)
    gM = tnb200.GMPS.from_host(M)
    g, E = tnb200.dmrg(g, gM, maxdim=64, cutoff=1e-14, maxsweeps=12, history=hg)
    assert abs((E - KAT[key]) / KAT[key]) < 1e-10
    assert abs((hg[-1][1] - ho[-1][1]) / ho[-1][1]) < 1e-10
    # truncated spectrum at the middle bond, 1e-8
    mid = len(H) // 2
    so, sg = po.spectrum(mid), g.spectrum(mid)
    k = min(len(so), len(sg))
    assert np.max(np.abs(so[:k] - sg[:k])) < 1e-8


def test_example_dmrg_config_small_maxdim():
    """examples/dmrg.jl settings (cutoff=1e-12, maxdim=32) at N=14 against the ED table."""
    import tnb200
    sh = oracle.spinhalf()
    H = tfim(14)
    g = tnb200.GMPS.from_host(oracle.randomMPS(2, 14, 1, np.random.default_rng(1234)))
    gM = tnb200.GMPS(2, 2, tnb200.models.tfim_mpo(14))
    g, E = tnb200.dmrg(g, gM, nsites=2, cutoff=1e-12, maxdim=32, maxsweeps=30)
    assert abs((E - KAT[("tfim", 14)]) / KAT[("tfim", 14)]) < 1e-10


def test_model_builders_match_reference_mpo():
    """tnb200.models hand-written MPOs / gates == the reference algorithm's (restated in the oracle)."""
    import tnb200
    sh = oracle.spinhalf()
    from test_oracle_kat import mpo_to_dense
    from oracle.gmps import GMPS as OG
    for ten, H in ((tnb200.models.tfim_mpo(6), tfim(6)), (tnb200.models.xxz_mpo(6, 0.5), xxz(6, 0.5))):
        assert np.allclose(mpo_to_dense(OG(2, 2, ten, 0)), mpo_to_dense(oracle.MPO(sh, H)), atol=1e-12)
    gl = oracle.trotterize(sh, -1 * tfim(7, 1.0, 0.3, 0.8), 0.01)
    X, Z, I2 = tnb200.models.X, tnb200.models.Z, tnb200.models.I2
    ss, gg = tnb200.models.trotter_gates(7, -(1.0 * X + 0.3 * Z), -0.8 * np.kron(Z, Z), 0.01)
    assert ss == gl.sites
    for ra, rb in zip(gg, gl.gates):
        for a, b in zip(ra, rb):
            assert np.allclose(a, b, atol=1e-14)


def _z_expect(g, N):
    import tnb200
    return np.real(g.expect([tnb200.models.Z] * N, list(range(1, N + 1))))


def test_tebd_gates_match_oracle():
    import tnb200
    sh = oracle.spinhalf()
    N = 10
    H = tfim(N, 1.0, 0.2, 0.9)
    gl = oracle.trotterize(sh, -1 * H, 0.02, evol="imag", order=2)
    psi = random_complex_mps(np.random.default_rng(7), N, 2, 6, center=1)
    g = tnb200.GMPS.from_host(psi)
    gg = tnb200.GateList.from_host(2, gl)
    zs = oracle.OpList(N)
    for i in range(1, N + 1):
        zs.add("z", i)
    kw = dict(cutoff=1e-12, maxdim=12, mindim=1)
    lognorm_o = lognorm_g = 0.0
    for step in range(6):
        oracle.applygates(psi, gl, **kw)
        lognorm_o += np.log(np.real(psi.norm()))
        psi.normalize()
        tnb200.applygates(g, gg, **kw)
        lognorm_g += np.log(np.real(g.norm()))
        g.normalize()
        assert g.center == psi.center
        assert [g.bonddim(i) for i in range(1, N)] == [psi.bonddim(i) for i in range(1, N)]
    assert abs(lognorm_g - lognorm_o) < 1e-10
    assert np.max(np.abs(_z_expect(g, N) - np.real(oracle.inner(sh, psi, zs, psi)))) < 1e-8
    assert np.max(np.abs(g.spectrum(5) - psi.spectrum(5))) < 1e-8


def test_thermal_mpo_gates_rank2():
    """examples/thermal.jl: the same applygates! on a rank-2 GMPS (identity MPO evolved in imaginary time)."""
    import tnb200
    sh = oracle.spinhalf()
    N = 6
    gl = oracle.trotterize(sh, -1 * tfim(N, 1.0, 0.0, 1.0), 0.05)
    U = oracle.productMPO(sh, ["id"] * N)
    g = tnb200.GMPS.from_host(U)
    gg = tnb200.GateList.from_host(2, gl)
    kw = dict(cutoff=1e-10, maxdim=16, mindim=1)
    U.movecenter(1)
    g.movecenter(1)
    for _ in range(4):
        oracle.applygates(U, gl, **kw)
        tnb200.applygates(g, gg, **kw)
    assert [g.bonddim(i) for i in range(1, N)] == [U.bonddim(i) for i in range(1, N)]
    def dense(p):
        v = p[1]
        for i in range(2, N + 1):
            v = np.tensordot(v, p[i], axes=([v.ndim - 1], [0]))
        return v.reshape(-1)
    assert relerr(dense(g), dense(U)) < 1e-8


def test_qjmc_trajectory_matches_oracle():
    import tnb200
    sh = oracle.spinhalf()
    N, dt, steps = 7, 0.02, 80
    H = tfim(N, 1.0, 2.0, 1.0)
    J = oracle.OpList(N)
    for i in range(1, N + 1):
        J.add("s-", i, np.sqrt(0.9))
    u = np.random.default_rng(11).random(3 * steps)
    # the oracle consumes uniforms sequentially: u0 every step, u1 (jump test) every step, u2 only on a jump,
    # whereas the C ABI indexes 3 per step; feed the oracle the same per-step triples.
    class Feeder:
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
            f.step, f.slot = len(self.m) - 1, 0       # next step's triple
    ob = Obs()
    kw = dict(cutoff=1e-10, maxdim=16)
    jumps, times = oracle.qjmc_simulation(sh, psi, H, J, steps * dt, dt, [ob], uniforms=f, **kw)
    _, gl = oracle.qjmc_gates(sh, H, J, dt)
    g = tnb200.GMPS.from_host(oracle.productMPS(sh, ["up" if i % 2 else "dn" for i in range(1, N + 1)]))
    g.movecenter(1)
    gg = tnb200.GateList.from_host(2, gl)
    gj, gt, obs = tnb200.qjmc_simulation(g, gg, list(range(1, N + 1)), [sh.op("s-")] * N, [np.sqrt(0.9)] * N, steps, dt,
                                         uniforms=u, obs_op=sh.op("z"), save_every=1, **kw)
    assert gj == jumps and len(jumps) > 0
    assert np.allclose(gt, times)
    assert np.max(np.abs(np.real(obs) - np.array(ob.m[1:]))) < 1e-8


def test_qjmc_ensemble_equals_single_trajectories():
    """tn_qjmc_ensemble (worker threads / streams inside the library, dynamic trajectory hand-out) reproduces
    tn_qjmc_run for the same (seed, trajectory id) keys, independently of the worker count."""
    import tnb200
    from tnb200 import models
    N, d, dt, steps, chi = 8, 2, 0.02, 30, 8
    gamma = 0.9
    onsite = -1j * (1.0 * models.X + 2.0 * models.Z) - 0.5 * gamma * (models.SM.conj().T @ models.SM)
    bond = -1j * 1.0 * np.kron(models.Z, models.Z)
    ss, gg = models.trotter_gates(N, onsite, bond, dt, evol="imag", order=2)
    tens = models.random_canonical_mps(N, d, chi, seed=5)
    ids = [3, 11, 7, 0, 42]
    kw = dict(cutoff=1e-12, maxdim=chi)
    nj, jumps, times, obs = tnb200.qjmc_ensemble(tens, 1, ss, gg, list(range(1, N + 1)), [models.SM] * N, [np.sqrt(gamma)] * N, steps, dt,
                                                 ids, workers=3, seed=9, obs_op=models.Z, save_every=5, **kw)
    assert obs.shape == (len(ids), steps // 5, N)
    gl = tnb200.GateList(d, ss, gg)
    for k, t in enumerate(ids):
        psi = tnb200.GMPS(1, d, tens, 1)
        j1, t1, o1 = tnb200.qjmc_simulation(psi, gl, list(range(1, N + 1)), [models.SM] * N, [np.sqrt(gamma)] * N, steps, dt,
                                            seed=9, trajectory=t, obs_op=models.Z, save_every=5, **kw)
        assert list(jumps[k, :nj[k]]) == j1 and np.allclose(times[k, :nj[k]], t1)
        assert np.max(np.abs(obs[k] - o1)) < 1e-9
    assert nj.sum() > 0


def test_inner_oplist_matches_oracle():
    """tn_inner_oplist (mps.jl:87-134): operator strings with gaps, one- to three-site terms, bra != ket, complex
    coefficients; and the TEBD energy measurement inner(st, psi, H, psi) of tebd.jl:51."""
    import tnb200
    sh = oracle.spinhalf()
    rng = np.random.default_rng(21)
    N = 9
    psi = random_complex_mps(rng, N, 2, 12, center=4)
    phi = random_complex_mps(rng, N, 2, 7, center=7)
    ol = oracle.OpList(N)
    ol.add("z", 1, 0.5)
    ol.add(["x", "x"], [3, 4], 1.2)
    ol.add(["s+", "z", "s-"], [2, 5, 9], 0.3 - 0.4j)       # gaps between the operators
    ol.add(["y", "n"], [8, 9], -2.0)
    ol.add(["z", "z"], [6, 7])
    ol.add("x", N, 1.5j)
    g, h = tnb200.GMPS.from_host(psi), tnb200.GMPS.from_host(phi)
    terms = [([sh.op(o) for o in ops], sites, co) for ops, sites, co in zip(ol.ops, ol.sites, ol.coeffs)]
    want = oracle.inner(sh, psi, ol, phi)
    got = tnb200.inner(g, terms, h)
    assert np.max(np.abs(got - want)) < 1e-12 * max(1.0, np.max(np.abs(want)))
    H = tfim(N, 1.0, 0.3, 1.1)
    hterms = [([sh.op(o) for o in ops], sites, co) for ops, sites, co in zip(H.ops, H.sites, H.coeffs)]
    e_want = np.sum(oracle.inner(sh, psi, H, psi))
    e_got = np.sum(tnb200.inner(g, hterms))
    assert abs(e_got - e_want) < 1e-11 * abs(e_want)
    assert abs(e_got.imag) < 1e-11 * abs(e_want)
