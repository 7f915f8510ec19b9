"""Oracle known answers for the projector branch of the path (SURVEY 8(a) A4/A5 fan-out over
ProjMPSSum, 8(f) rank 3): excited-state DMRG with a squared rank-1 projector penalty
(dmrg.jl:128-154, projmps.jl:135-143) against exact diagonalisation, and vmps (vmps.jl:1-105)
against dense-vector arithmetic."""
import numpy as np

import oracle
from gpu_util import random_complex_mps
from models import tfim, xxz, dense_hamiltonian, mps_to_dense


def test_squared_projector_product_is_rank_one_penalty():
    rng = np.random.default_rng(3)
    N = 6
    V = random_complex_mps(rng, N, 2, 4, center=1)
    psi = random_complex_mps(rng, N, 2, 5, center=3)
    P = oracle.ProjMPS([V, psi], rank=2, squared=True, coeff=2.5, center=3)
    A = psi[3]
    theta = np.tensordot(A, psi[4], axes=([2], [0]))
    out = P.product(theta, False, 2)
    # dense check: the effective operator is coeff * |v><v| with v the projection of V into psi's basis
    v = np.conj(P.project(theta, False, 2))
    assert np.allclose(out, 2.5 * v * np.vdot(v, theta), atol=1e-13)
    # <theta|P theta> = coeff |<V|psi>|^2 when theta is psi's own two-site tensor
    ov = np.vdot(mps_to_dense(V), mps_to_dense(psi))
    assert np.isclose(np.vdot(theta, out), 2.5 * abs(ov) ** 2, atol=1e-12)
    # calculate() ignores ``squared``: coeff * <V|psi>
    assert np.isclose(P.calculate(), 2.5 * ov, atol=1e-12)


def test_excited_state_dmrg_matches_ed():
    sh = oracle.spinhalf()
    N = 8
    H = tfim(N)
    ev = np.linalg.eigvalsh(dense_hamiltonian(sh, H).toarray())
    M = oracle.MPO(sh, H)
    psi0 = oracle.randomMPS(2, N, 4, np.random.default_rng(1))
    psi0, E0 = oracle.dmrg(psi0, M, maxdim=32, cutoff=1e-14, maxsweeps=20)
    assert abs(E0 - ev[0]) < 1e-10 * abs(ev[0])
    psi1 = oracle.randomMPS(2, N, 4, np.random.default_rng(2))
    psi1, E1 = oracle.dmrg(psi1, M, psi0, coeffs=[1.0, 20.0], maxdim=32, cutoff=1e-14, maxsweeps=30)
    assert abs(E1 - ev[1]) < 1e-9 * abs(ev[1])
    assert abs(np.vdot(mps_to_dense(psi0), mps_to_dense(psi1))) < 1e-7


def test_two_mpo_sum_equals_single_mpo():
    """dmrg(psi, H1, H2) optimises H1 + H2 (projmpssum.jl:63-73)."""
    sh = oracle.spinhalf()
    N = 8
    Ha, Hb = oracle.OpList(N), oracle.OpList(N)
    for i in range(1, N + 1):
        Ha.add("x", i, 1.0)
        Ha.add("z", i, 0.05)
    for i in range(1, N):
        Hb.add(["z", "z"], [i, i + 1], 1.2)
    E_ref = np.linalg.eigvalsh(dense_hamiltonian(sh, tfim(N)).toarray())[0]
    psi = oracle.randomMPS(2, N, 4, np.random.default_rng(7))
    psi, E = oracle.dmrg(psi, oracle.MPO(sh, Ha), oracle.MPO(sh, Hb), maxdim=32, cutoff=1e-14, maxsweeps=20)
    assert abs(E - E_ref) < 1e-10 * abs(E_ref)


def test_vmps_recovers_sum_of_two_mps():
    rng = np.random.default_rng(11)
    N = 7
    a = random_complex_mps(rng, N, 2, 3, center=1)
    b = random_complex_mps(rng, N, 2, 2, center=1)
    target = mps_to_dense(a) + mps_to_dense(b)
    hist = []
    psi = oracle.vmps(a, b, maxdim=16, cutoff=1e-14, maxsweeps=6, history=hist)
    assert np.linalg.norm(mps_to_dense(psi) - target) < 1e-10 * np.linalg.norm(target)
    # at the optimum cost = |S|^2 - 2 <S|S> = -|S|^2
    assert np.isclose(hist[-1][0 + 1], -np.vdot(target, target), rtol=1e-10)


def test_vmps_truncated_is_best_sweep_approximation():
    rng = np.random.default_rng(12)
    N = 8
    a = random_complex_mps(rng, N, 2, 6, center=1)
    b = random_complex_mps(rng, N, 2, 6, center=1)
    target = mps_to_dense(a) + mps_to_dense(b)
    hist = []
    psi = oracle.vmps(a, b, maxdim=4, cutoff=0.0, maxsweeps=8, history=hist)
    assert psi.maxbonddim() <= 4
    v = mps_to_dense(psi)
    # variational optimum: the residual is orthogonal to psi (<psi|target> = <psi|psi>), costs decrease monotonically
    assert np.isclose(np.vdot(v, target), np.vdot(v, v), rtol=1e-6)
    costs = [np.real(h[1]) for h in hist]
    assert all(costs[i + 1] <= costs[i] + 1e-9 for i in range(len(costs) - 1))


def test_tebd_with_projector_reaches_first_excited_state():
    """tebd.jl:22-41,67-73: imaginary-time TEBD with the ground state projected out (vmps(psi, -P psi) every few steps) converges
    to the first excited state (exact diagonalisation)."""
    from oracle.tebd import tebd, overlap
    sh = oracle.spinhalf()
    N = 8
    Hl = tfim(N)
    ev = np.linalg.eigvalsh(dense_hamiltonian(sh, Hl).toarray())
    g0, _ = oracle.dmrg(oracle.randomMPS(2, N, 4, np.random.default_rng(1)), oracle.MPO(sh, Hl), maxdim=32, cutoff=1e-14, maxsweeps=20)
    p = oracle.randomMPS(2, N, 4, np.random.default_rng(3))
    psi, E = tebd(sh, p, -1 * Hl, 0.02, 6.0, 1.0, projectors=[g0], cutoff=1e-12, maxdim=16, projection_every=5)
    assert abs(-E - ev[1]) < 1e-4 * abs(ev[1])          # second-order Trotter error at dt = 0.02
    assert abs(overlap(g0, psi)) < 1e-10
