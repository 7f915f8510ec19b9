"""Host half of the product's MPO(st, H) (tnb200/mpo.py): the finite-state-machine assembly reproduces the dense Hamiltonian
for on-site, nearest-neighbour and long-range (cylinder) operator lists; and the reference's compression sweeps
(mpo.jl:443-457, emulated here with the oracle's svd) bring it to the same bond dimensions as the oracle's MPO(st, H)."""
import numpy as np
import pytest

import oracle
from models import tfim, xxz, j1j2_cylinder, dense_hamiltonian


def _terms(sh, H):
    return [([sh.op(o) for o in ops], sites, c) for ops, sites, c in zip(H.ops, H.sites, H.coeffs)]


def _dense(Ts):
    t = Ts[0]
    for x in Ts[1:]:
        t = np.tensordot(t, x, axes=([t.ndim - 1], [0]))
    t = t[0, ..., 0]
    N = len(Ts)
    return np.transpose(t, list(range(0, 2 * N, 2)) + list(range(1, 2 * N, 2))).reshape(2 ** N, 2 ** N)


@pytest.mark.parametrize("name,H", [("tfim", tfim(6)), ("tfim2", tfim(2)), ("xxz", xxz(6, 0.5)), ("j1j2", j1j2_cylinder(2, 3)),
                                    ("j1j2_3x3", j1j2_cylinder(3, 3))])
def test_fsm_assembly_matches_dense(name, H):
    from tnb200.mpo import fsm_tensors
    sh = oracle.spinhalf()
    Ts = fsm_tensors(len(H), 2, _terms(sh, H))
    assert Ts[0].shape[0] == 1 and Ts[-1].shape[3] == 1
    assert np.allclose(_dense(Ts), dense_hamiltonian(sh, H).toarray(), atol=1e-12)


def test_unsorted_sites_and_complex_coefficients():
    from tnb200.mpo import fsm_tensors
    sh = oracle.spinhalf()
    X, Y, Z = sh.op("x"), sh.op("y"), sh.op("z")
    terms = [([Z, X], [4, 1], 0.3 - 0.2j), ([Y], [2], 1.5), ([X, Y, Z], [2, 3, 5], -0.7j)]
    Ts = fsm_tensors(5, 2, terms)
    I = np.eye(2)

    def kron(*ms):
        out = np.eye(1)
        for m in ms:
            out = np.kron(out, m)
        return out
    want = (0.3 - 0.2j) * kron(X, I, I, Z, I) + 1.5 * kron(I, Y, I, I, I) - 0.7j * kron(I, X, Y, I, Z)
    assert np.allclose(_dense(Ts), want, atol=1e-13)


def test_compression_sweeps_reach_the_oracle_bond_dimensions():
    from tnb200.mpo import fsm_tensors
    sh = oracle.spinhalf()
    H = j1j2_cylinder(4, 3)
    O = [t.copy() for t in fsm_tensors(len(H), 2, _terms(sh, H))]
    N = len(O)
    for i in range(N - 1):                      # mpo.jl:443-449
        U, S, V = oracle.svd(O[i], 4, cutoff=1e-15)
        O[i] = np.tensordot(U, S, axes=([3], [0]))
        O[i + 1] = np.tensordot(V, O[i + 1], axes=([1], [0]))
    for i in range(N - 1, 0, -1):               # mpo.jl:451-457
        U, S, V = oracle.svd(O[i], 1, cutoff=1e-15)
        O[i] = np.tensordot(S, U, axes=([1], [0]))
        O[i - 1] = np.tensordot(O[i - 1], V, axes=([3], [1]))
    ref = oracle.MPO(sh, H)
    assert [t.shape[3] for t in O] == [ref[i].shape[3] for i in range(1, N + 1)]
    assert np.allclose(_dense(O), dense_hamiltonian(sh, H).toarray(), atol=1e-10)


def test_product_trotterize_equals_oracle_trotterize():
    """tnb200.evolve.trotterize (host half of the tebd front-end) against the oracle's gatelist.jl:75-121 restatement."""
    from tnb200.evolve import trotterize
    sh = oracle.spinhalf()
    site_dep = oracle.OpList(6)
    for i in range(1, 7):
        site_dep.add("x", i, 0.3 * i)
    for i in range(1, 6):
        site_dep.add(["z", "z"], [i, i + 1], 1.0 + 0.1 * i)
        site_dep.add(["x", "y"], [i, i + 1], 0.2j)
    onsite_only = oracle.OpList(5)
    for i in range(1, 6):
        onsite_only.add("x", i, 0.3 * i)
    for H in (-1 * tfim(7), -1 * xxz(6, 0.5), site_dep, onsite_only):
        for evol in ("imag", "real"):
            for order in (1, 2):
                gl = oracle.trotterize(sh, H, 0.01, order=order, evol=evol)
                rs, rg = trotterize(len(H), 2, _terms(sh, H), 0.01, order=order, evol=evol)
                assert rs == gl.sites
                for ra, rb in zip(rg, gl.gates):
                    for a, b in zip(ra, rb):
                        assert a.shape == b.shape and np.abs(a - b).max() < 1e-14


def test_product_cylinder_terms_equal_the_oracle_oplist():
    from tnb200 import models as pm
    from tnb200.mpo import fsm_tensors
    sh = oracle.spinhalf()
    for Lx, Ly in ((3, 3), (2, 4)):
        Ts = fsm_tensors(Lx * Ly, 2, pm.j1j2_cylinder_terms(Lx, Ly))
        assert np.allclose(_dense(Ts), dense_hamiltonian(sh, j1j2_cylinder(Lx, Ly)).toarray(), atol=1e-12)


def test_doubled_site_identity_for_the_thermal_energy():
    """tr(H U^dag U) / tr(U^dag U) (mpo.jl:229-252, the measurement of examples/thermal.jl) equals <<U|1 (x) H^T|U>> / <<U|U>> on
    doubled sites -- the form tnb200.evolve.thermal_energy evaluates with the environment kernels."""
    from gpu_util import random_mpo
    from oracle.gmps import GMPS
    from tnb200.evolve import doubled_state_tensors, doubled_operator_tensors
    sh = oracle.spinhalf()
    rng = np.random.default_rng(0)
    N = 5
    H = oracle.MPO(sh, tfim(N, 1.0, 0.3, 0.7))
    Hc = random_mpo(rng, N, 2, 3)                        # a non-symmetric complex operator: the transpose matters
    U = random_mpo(rng, N, 2, 4)
    for Hx in (H, Hc):
        want = oracle.trace(Hx, oracle.adjoint(U), U) / oracle.trace(oracle.adjoint(U), U)
        Ud = GMPS(1, 4, doubled_state_tensors(U.tensors), 0)
        Hd = GMPS(2, 4, doubled_operator_tensors(Hx.tensors), 0)
        got = oracle.ProjMPS([Ud, Hd, Ud], rank=2, center=1).calculate() / oracle.ProjMPS([Ud, Ud], rank=1, center=1).calculate()
        assert abs(got - want) < 1e-12 * abs(want)


def test_oracle_trace_matches_dense():
    from gpu_util import random_mpo
    rng = np.random.default_rng(1)
    A, B, C = random_mpo(rng, 4, 2, 3), random_mpo(rng, 4, 2, 2), random_mpo(rng, 4, 2, 4)

    def dense(M):
        return _dense([M[i] for i in range(1, len(M) + 1)])
    for args in ((A,), (A, B), (A, B, C)):
        want = np.trace(np.linalg.multi_dot([dense(x) for x in args]) if len(args) > 1 else dense(args[0]))
        assert abs(oracle.trace(*args) - want) < 1e-12 * abs(want)
    assert np.allclose(dense(oracle.adjoint(A)), dense(A).conj().T)


def test_oracle_applympo_matches_dense():
    """applyMPO (mpo.jl:105-143 restated): O|psi> as a dense vector, canonical form with the centre at site 1, truncation kwargs."""
    from gpu_util import random_mpo, random_complex_mps
    from models import mps_to_dense
    rng = np.random.default_rng(0)
    N = 6
    O, psi = random_mpo(rng, N, 2, 3), random_complex_mps(rng, N, 2, 4, center=2)
    phi = oracle.applyMPO(O, psi)
    want = _dense([O[i] for i in range(1, N + 1)]) @ mps_to_dense(psi)
    assert np.linalg.norm(mps_to_dense(phi) - want) < 1e-12 * np.linalg.norm(want)
    assert phi.center == 1
    for i in range(2, N + 1):                      # right-orthonormal sites
        t = phi[i].reshape(phi[i].shape[0], -1)
        assert np.allclose(t @ t.conj().T, np.eye(t.shape[0]), atol=1e-12)
    small = oracle.applyMPO(O, psi, maxdim=3)
    assert small.maxbonddim() <= 3
