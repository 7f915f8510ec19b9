"""Golden vectors of the projector branch, vmps, the norm-based QJMC branch and the iTEBD gate step, frozen from the CPU oracle
(python tests/golden/make_golden_projector.py -> projector_golden.npz).  Same status as hotpath_golden.npz: the reference (Julia)
cannot run in the build image; the oracle is pinned by exact diagonalisation / exact free-fermion energies (tests/test_oracle_*.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from oracle.itebd import iMPS, itebd_gate, itebd_apply_gates_mps  # noqa: E402
from gpu_util import crandn, random_complex_mps, random_mpo  # noqa: E402
from models import tfim  # noqa: E402


def main():
    rng = np.random.default_rng(4242)
    out = {}
    sh = oracle.spinhalf()
    # project() with and without an MPO layer, squared product, at centre 3 of a 6-site chain
    N = 6
    V = random_complex_mps(rng, N, 2, 5, center=1)
    psi = random_complex_mps(rng, N, 2, 6, center=3)
    H = random_mpo(rng, N, 2, 3)
    for name, obj in (("V", V), ("psi", psi), ("H", H)):
        for i, t in enumerate(obj.tensors):
            out[f"pj_{name}{i}"] = t
    out["pj_project2"] = oracle.ProjMPS([V, psi], rank=1, center=3).project(None, False, 2)
    out["pj_project2_mpo"] = oracle.ProjMPS([V, H, psi], rank=1, center=3).project(None, False, 2)
    out["pj_project1_mpo"] = oracle.ProjMPS([V, H, psi], rank=1, center=3).project(None, False, 1)
    theta = crandn(rng, psi[3].shape[0], 2, 2, psi[4].shape[2])
    out["pj_theta"] = theta
    out["pj_squared_out"] = oracle.ProjMPS([V, psi], rank=2, squared=True, coeff=2.5, center=3).product(theta, False, 2)
    A1 = crandn(rng, *psi[3].shape)
    out["pj_A1"] = A1
    out["pj_product1"] = oracle.ProjMPS([psi, H, psi], rank=2, center=3, coeff=0.7 - 0.2j).product(A1, False, 1)
    # excited-state DMRG (TFIM N=8): ground state, then first excited state with a penalty of 20 on the ground state
    M = oracle.MPO(sh, tfim(8))
    p0 = oracle.randomMPS(2, 8, 4, np.random.default_rng(1))
    p1 = oracle.randomMPS(2, 8, 4, np.random.default_rng(2))
    g0, _ = oracle.dmrg(p0.copy(), M, maxdim=32, cutoff=1e-14, maxsweeps=20)
    hist = []
    oracle.dmrg(p1.copy(), M, g0, coeffs=[1.0, 20.0], maxdim=32, cutoff=1e-14, maxsweeps=30, history=hist)
    for name, obj in (("mpo", M), ("gs", g0), ("start", p1)):
        for i, t in enumerate(obj.tensors):
            out[f"ex_{name}{i}"] = t
    out["ex_gs_center"] = np.array(g0.center)
    out["ex_energy"] = np.array([h[1] for h in hist])
    out["ex_maxbond"] = np.array([h[2] for h in hist])
    # vmps: sum of two MPS compressed to bond dimension 4
    a = random_complex_mps(rng, 8, 2, 6, center=1)
    b = random_complex_mps(rng, 8, 2, 6, center=1)
    for name, obj in (("a", a), ("b", b)):
        for i, t in enumerate(obj.tensors):
            out[f"vm_{name}{i}"] = t
    hist = []
    oracle.vmps(a, b, maxdim=4, cutoff=0.0, maxsweeps=6, history=hist)
    out["vm_cost"] = np.array([h[1] for h in hist])
    out["vm_maxbond"] = np.array([h[2] for h in hist])
    # iTEBD: Schmidt values and log-norms after 40 steps of imaginary time (TFIM g = 2, dt = 0.05, maxdim 8)
    Hc = oracle.OpList(2)
    Hc.add(["z", "z"], [1, 2], -1.0)
    Hc.add("x", 1, -2.0)
    gate = itebd_gate(sh, -1 * Hc, 0.05)
    ip = iMPS(2, np.array([1.0, 0.3]))
    for _ in range(40):
        itebd_apply_gates_mps(ip, gate, maxdim=8, cutoff=1e-12)
    out["it_gate"] = gate
    out["it_sing1"], out["it_sing2"] = ip.singulars[0], ip.singulars[1]
    out["it_norms"] = np.array(ip.norms)
    np.savez_compressed(os.path.join(HERE, "projector_golden.npz"), **out)
    print("wrote projector_golden.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
