"""Oracle known answer for the infinite-TEBD gate step (itebd.jl:71-119 restated in oracle/itebd.py): imaginary-time evolution of
the infinite transverse-field Ising chain reaches the exact ground-state energy per site (free-fermion integral)."""
import numpy as np
from scipy.integrate import quad

import oracle
from oracle.itebd import iMPS, itebd_gate, itebd_apply_gates_mps, bond_energy


def test_itebd_tfim_ground_energy_per_site():
    sh = oracle.spinhalf()
    g = 2.0
    exact = -quad(lambda k: np.sqrt(1 + g * g - 2 * g * np.cos(k)), -np.pi, np.pi)[0] / (2 * np.pi)
    H = oracle.OpList(2)
    H.add(["z", "z"], [1, 2], -1.0)
    H.add("x", 1, -g)
    h2 = H.sitetensor(sh, 1)
    psi = iMPS(2, np.array([1.0, 0.3]))
    for dt, n in ((0.05, 200), (0.01, 300)):
        gate = itebd_gate(sh, -1 * H, dt)
        for _ in range(n):
            itebd_apply_gates_mps(psi, gate, maxdim=8, cutoff=1e-12)
    assert abs(bond_energy(psi, h2).real - exact) < 1e-4 * abs(exact)
    # Vidal form: normalised singular values on both bonds, accumulated log-norms negative for this (decaying) evolution
    for s in psi.singulars:
        assert abs(np.sum(s ** 2) - 1.0) < 1e-12
    assert psi.maxbonddim() <= 8


def test_itebd_three_site_cell_runs_and_keeps_normalisation():
    """the restatement follows the reference for any cell length (the device path is restricted to two sites)"""
    sh = oracle.spinhalf()
    H = oracle.OpList(3)
    H.add(["z", "z"], [1, 2], -1.0)
    H.add("x", 1, -1.5)
    H.add(["z", "x", "z"], [1, 2, 3], -0.3)          # the cell gate spans the interaction range, which must equal the cell length
    psi = iMPS(3, np.array([1.0, 0.2]))
    gate = itebd_gate(sh, -1 * H, 0.02)
    assert gate.ndim == 6
    for _ in range(30):
        itebd_apply_gates_mps(psi, gate, maxdim=6, cutoff=1e-12)
    for s in psi.singulars:
        assert abs(np.sum(s ** 2) - 1.0) < 1e-12
