"""Tier-2 oracle known-answer tests: the restated algorithms reproduce exact-
diagonalisation energies to the north_star
tolerance (1e-10 relative), and the MPO builder reproduces the dense operator."""
import numpy as np
import pytest

import oracle
from models import tfim, xxz, j1j2_cylinder, dense_hamiltonian, ed_ground_energy, mps_to_dense, KAT


def mpo_to_dense(M):
    t = M[1]
    for i in range(2, len(M) + 1):
        t = np.tensordot(t, M[i], axes=([t.ndim - 1], [0]))
    t = t[0, ..., 0]
    N = len(M)
    outs = list(range(0, 2 * N, 2))
    ins = list(range(1, 2 * N, 2))
    t = np.transpose(t, outs + ins)
    return t.reshape(2 ** N, 2 ** N)


@pytest.mark.parametrize("name,H", [("tfim", tfim(6)), ("xxz", xxz(6, 0.5)), ("j1j2", j1j2_cylinder(2, 3))])
def test_mpo_builder_matches_dense(name, H):
    sh = oracle.spinhalf()
    M = oracle.MPO(sh, H)
    assert np.allclose(mpo_to_dense(M), dense_hamiltonian(sh, H).toarray(), atol=1e-12)


def test_ed_table_is_reproduced_by_independent_ed():
    sh = oracle.spinhalf()
    assert np.isclose(ed_ground_energy(sh, tfim(8)), KAT[("tfim", 8)], rtol=0, atol=1e-10)
    assert np.isclose(ed_ground_energy(sh, xxz(8, 1.0)), KAT[("heis", 8)], rtol=0, atol=1e-10)
    assert np.isclose(ed_ground_energy(sh, xxz(10, 0.5)), KAT[("xxz0.5", 10)], rtol=0, atol=1e-10)


@pytest.mark.parametrize("key,H", [
    (("tfim", 8), tfim(8)), (("tfim", 10), tfim(10)),
    (("heis", 8), xxz(8, 1.0)), (("heis", 10), xxz(10, 1.0)),
    (("xxz0.5", 10), xxz(10, 0.5)),
    (("j1j2_4x3", 12), j1j2_cylinder(4, 3)),
])
def test_dmrg_reaches_ed_energy(key, H):
    sh = oracle.spinhalf()
    M = oracle.MPO(sh, H)
    psi = oracle.randomMPS(2, len(H), 4, np.random.default_rng(1234))
    psi, E = oracle.dmrg(psi, M, maxdim=64, cutoff=1e-14, maxsweeps=30)
    assert abs((np.real(E) - KAT[key]) / KAT[key]) < 1e-10
    # energy returned by the eigensolver == <psi|H|psi>
    v = mps_to_dense(psi)
    assert np.isclose(np.real(np.vdot(v, dense_hamiltonian(sh, H) @ v)), KAT[key], rtol=1e-10)


def test_dmrg_optimal_order_same_energy():
    sh = oracle.spinhalf()
    H = tfim(10)
    M = oracle.MPO(sh, H)
    hist_a, hist_b = [], []
    pa = oracle.randomMPS(2, 10, 2, np.random.default_rng(5))
    pb = pa.copy()
    oracle.dmrg(pa, M, maxdim=16, maxsweeps=4, history=hist_a)
    oracle.dmrg(pb, M, maxdim=16, maxsweeps=4, history=hist_b, optimal_order=True)
    for a, b in zip(hist_a, hist_b):
        assert abs(a[1] - b[1]) < 1e-10 * abs(a[1]) and a[2] == b[2]


def test_tebd_imaginary_time_converges_to_ground_state():
    sh = oracle.spinhalf()
    N = 8
    H = tfim(N, 1.0, 0.0, 0.5)          # paramagnetic: gap ~ 1, converges within T = 8
    E0 = ed_ground_energy(sh, H)
    psi = oracle.randomMPS(2, N, 8, np.random.default_rng(11))
    dt = 0.01
    psi, E = oracle.tebd(sh, psi, -1 * H, dt, 8.0, 1.0, cutoff=1e-12, maxdim=16)
    # tebd evolves with exp(+dt*(-H)); the energy it reports is <-H> (tebd.jl:90)
    assert abs(-E - E0) < 2e-4           # second-order Trotter error at dt = 0.01


def test_qjmc_matches_dense_state_trajectory():
    """Same Trotter gates + same uniforms applied to a dense state vector give the
    same jump record and <z_i>: checks gate application, truncation-free SVD
    splitting, emission rates and the jump update in one go (N small, no truncation)."""
    sh = oracle.spinhalf()
    N, dt, steps = 5, 0.05, 60
    H = tfim(N, 1.0, 0.3, 0.7)
    J = oracle.OpList(N)
    for i in range(1, N + 1):
        J.add("s-", i, np.sqrt(0.8))
    u = np.random.default_rng(3).random(3 * steps + 8)
    it = iter(u)
    psi = oracle.productMPS(sh, ["up" if i % 2 else "dn" for i in range(1, N + 1)])
    psi.movecenter(1)
    zs = oracle.OpList(N)
    for i in range(1, N + 1):
        zs.add("z", i)
    ob = oracle.qjmc.QJMCOperators(zs, sh)
    jumps, times = oracle.qjmc_simulation(sh, psi, H, J, steps * dt, dt, [ob], uniforms=lambda: next(it), cutoff=0, maxdim=0)

    # dense replay
    _, gates = oracle.qjmc_gates(sh, H, J, dt)
    it = iter(u)
    v = mps_to_dense(oracle.productMPS(sh, ["up" if i % 2 else "dn" for i in range(1, N + 1)])).reshape((2,) * N)

    def apply(v, g, site):
        n = g.ndim // 2
        gm = np.transpose(g, [2 * k for k in range(n)] + [2 * k + 1 for k in range(n)])  # (outs..., ins...)
        v = np.tensordot(gm, v, axes=(list(range(n, 2 * n)), list(range(site - 1, site - 1 + n))))
        return np.moveaxis(v, list(range(n)), list(range(site - 1, site - 1 + n)))
    sm = sh.op("s-")
    dj, dz = [], []
    for step in range(steps):
        for rs, rg in zip(gates.sites, gates.gates):
            for s, g in zip(rs, rg):
                v = apply(v, g, s)
        next(it)
        v = v / np.linalg.norm(v)
        rates = np.array([0.8 * np.linalg.norm(apply(v, sm, s)) ** 2 for s in range(1, N + 1)])
        if next(it) > np.exp(-rates.sum() * dt):
            r = next(it)
            k = int(np.nonzero(r < np.cumsum(rates) / rates.sum())[0][0])
            v = apply(v, sm, k + 1)
            v = v / np.linalg.norm(v)
            dj.append(k + 1)
        dz.append([np.real(np.vdot(v, apply(v, sh.op("z"), s))) for s in range(1, N + 1)])
    assert jumps == dj and len(jumps) > 0
    got = np.real(np.array(ob.measurements[1:]))
    assert np.allclose(got, np.array(dz), atol=1e-8)


def test_qjmc_norm_based_branch_matches_dense_state_trajectory():
    """classical=false (qjmc.jl:65-87): the jump fires when u0 exceeds the norm^2 left after the non-unitary gates; the
    channel comes from the next uniform.  Dense replay with the same gates and uniforms."""
    sh = oracle.spinhalf()
    N, dt, steps = 5, 0.05, 60
    H = tfim(N, 1.0, 0.3, 0.7)
    J = oracle.OpList(N)
    for i in range(1, N + 1):
        J.add("s-", i, np.sqrt(0.8))
    u = np.random.default_rng(4).random(2 * steps + 8)
    it = iter(u)
    psi = oracle.productMPS(sh, ["up" if i % 2 else "dn" for i in range(1, N + 1)])
    psi.movecenter(1)
    zs = oracle.OpList(N)
    for i in range(1, N + 1):
        zs.add("z", i)
    ob = oracle.qjmc.QJMCOperators(zs, sh)
    jumps, times = oracle.qjmc_simulation(sh, psi, H, J, steps * dt, dt, [ob], uniforms=lambda: next(it), cutoff=0, maxdim=0,
                                          classical=False)
    _, gates = oracle.qjmc_gates(sh, H, J, dt)
    it = iter(u)
    v = mps_to_dense(oracle.productMPS(sh, ["up" if i % 2 else "dn" for i in range(1, N + 1)])).reshape((2,) * N)

    def apply(v, g, site):
        n = g.ndim // 2
        gm = np.transpose(g, [2 * k for k in range(n)] + [2 * k + 1 for k in range(n)])
        v = np.tensordot(gm, v, axes=(list(range(n, 2 * n)), list(range(site - 1, site - 1 + n))))
        return np.moveaxis(v, list(range(n)), list(range(site - 1, site - 1 + n)))
    sm = sh.op("s-")
    dj, dz = [], []
    for step in range(steps):
        for rs, rg in zip(gates.sites, gates.gates):
            for s, g in zip(rs, rg):
                v = apply(v, g, s)
        r0 = next(it)
        prob = np.linalg.norm(v) ** 2
        v = v / np.linalg.norm(v)
        if r0 > prob:
            rates = np.array([0.8 * np.linalg.norm(apply(v, sm, s)) ** 2 for s in range(1, N + 1)])
            r = next(it)
            k = int(np.nonzero(r < np.cumsum(rates) / rates.sum())[0][0])
            v = apply(v, sm, k + 1)
            v = v / np.linalg.norm(v)
            dj.append(k + 1)
        dz.append([np.real(np.vdot(v, apply(v, sh.op("z"), s))) for s in range(1, N + 1)])
    assert jumps == dj and len(jumps) > 0
    assert np.allclose(np.real(np.array(ob.measurements[1:])), np.array(dz), atol=1e-8)


@pytest.mark.parametrize("classical", [True, False])
def test_qjmc_ensemble_average_follows_the_lindblad_equation(classical):
    """SURVEY 8(c) pin: the average of <z_i> over quantum-jump trajectories (both jump rules, qjmc.jl:65-112) equals the solution
    of the Lindblad master equation d rho/dt = -i[H, rho] + sum_k (L_k rho L_k^dag - {L_k^dag L_k, rho}/2), integrated exactly for
    N = 3 as a 64 x 64 matrix exponential.  300 trajectories: statistical error 0.06 per observable, tolerance 0.15."""
    import scipy.linalg as sla
    sh = oracle.spinhalf()
    N, gamma, dt, steps = 3, 0.8, 0.02, 25
    H = tfim(N, 1.0, 0.3, 0.7)
    J = oracle.OpList(N)
    for i in range(1, N + 1):
        J.add("s-", i, np.sqrt(gamma))
    Hd = dense_hamiltonian(sh, H).toarray()

    def op_at(o, i):
        mats = [np.eye(2)] * N
        mats = list(mats)
        mats[i] = o
        out = mats[0]
        for x in mats[1:]:
            out = np.kron(out, x)
        return out
    D = 2 ** N
    I = np.eye(D)
    Lv = -1j * (np.kron(Hd, I) - np.kron(I, Hd.T))                      # row-major vectorisation of rho
    for i in range(N):
        L = np.sqrt(gamma) * op_at(sh.op("s-"), i)
        LdL = L.conj().T @ L
        Lv += np.kron(L, L.conj()) - 0.5 * np.kron(LdL, I) - 0.5 * np.kron(I, LdL.T)
    names = ["up", "dn", "up"]
    v0 = mps_to_dense(oracle.productMPS(sh, names))
    rhoT = (sla.expm(Lv * dt * steps) @ np.outer(v0, v0.conj()).reshape(-1)).reshape(D, D)
    exact = np.array([np.real(np.trace(rhoT @ op_at(sh.op("z"), i))) for i in range(N)])
    zs = oracle.OpList(N)
    for i in range(1, N + 1):
        zs.add("z", i)
    rng = np.random.default_rng(0)
    M, acc = 300, np.zeros(N)
    for _ in range(M):
        p = oracle.productMPS(sh, names)
        p.movecenter(1)
        oracle.qjmc_simulation(sh, p, H, J, steps * dt, dt, uniforms=rng.random, classical=classical, cutoff=0, maxdim=0)
        acc += np.real(oracle.inner(sh, p, zs, p))
    assert np.max(np.abs(acc / M - exact)) < 0.15, (acc / M, exact)


def test_thermal_mpo_energy_matches_exact_canonical_ensemble():
    """examples/thermal.jl: the identity MPO evolved to beta/2 with the Trotter gates (rank-2 applygates!), energy =
    trace(H, adjoint(U), U) / trace(adjoint(U), U) (mpo.jl:229-252), against tr(H exp(-beta H)) / tr(exp(-beta H)) from exact
    diagonalisation (N = 6, beta = 1)."""
    sh = oracle.spinhalf()
    N, beta, dt = 6, 1.0, 0.01
    Hl = tfim(N, 1.0, 0.0, 1.0)
    gl = oracle.trotterize(sh, -1 * Hl, dt)
    U = oracle.productMPO(sh, ["id"] * N)
    for _ in range(int(round(beta / 2 / dt))):
        oracle.applygates(U, gl, cutoff=1e-12, maxdim=64)
    M = oracle.MPO(sh, Hl)
    E = oracle.trace(M, oracle.adjoint(U), U) / oracle.trace(oracle.adjoint(U), U)
    ev = np.linalg.eigvalsh(dense_hamiltonian(sh, Hl).toarray())
    w = np.exp(-beta * (ev - ev[0]))
    exact = float(np.sum(ev * w) / np.sum(w))
    assert abs(E.imag) < 1e-10 and abs(E.real - exact) < 2e-4 * abs(exact)        # second-order Trotter error at dt = 0.01
