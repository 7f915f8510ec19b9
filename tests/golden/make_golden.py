"""Generates the committed golden vectors from the CPU oracle (python tests/golden/make_golden.py).
The reference itself (Julia) cannot run in the build image, so these freeze the ORACLE's outputs on seeded
inputs; the oracle in turn is pinned by exact-diagonalisation known answers (tests/test_oracle_kat.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from gpu_util import crandn, random_complex_mps, random_mpo  # noqa: E402
from models import tfim, xxz  # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    out = {}
    # H_eff matvec + environment blocks on a random complex MPS / MPO
    N, chi, w = 6, 12, 4
    psi = random_complex_mps(rng, N, 2, chi, center=3)
    H = random_mpo(rng, N, 2, w)
    P = oracle.ProjMPS([psi, H, psi], rank=2, center=3, coeff=1.0)
    theta = crandn(rng, psi[3].shape[0], 2, 2, psi[4].shape[2])
    for i, t in enumerate(psi.tensors):
        out[f"mv_psi{i}"] = t
    for i, t in enumerate(H.tensors):
        out[f"mv_mpo{i}"] = t
    out["mv_theta"] = theta
    out["mv_L"] = P.block(2)
    out["mv_R"] = P.block(5)
    out["mv_out"] = P.product(theta, False, 2)
    out["mv_calculate"] = np.array(P.calculate())
    # truncated SVD: singular values + kept rank under the reference's rule
    x = crandn(rng, 40, 56) * np.exp(-0.35 * np.arange(56))[None, :]
    out["svd_x"] = x
    for name, kw in (("full", {}), ("cut", dict(cutoff=1e-10)), ("max", dict(maxdim=9)), ("min", dict(cutoff=1e-2, mindim=7))):
        _, S, _ = oracle.svd(x, 2, **kw)
        out[f"svd_S_{name}"] = np.real(np.diag(S))
    # DMRG energy per sweep (TFIM N=12 from a seeded random chi=2 state; XXZ delta=0.5 N=10)
    sh = oracle.spinhalf()
    for name, Hl, n in (("tfim12", tfim(12), 12), ("xxz10", xxz(10, 0.5), 10)):
        M = oracle.MPO(sh, Hl)
        p0 = oracle.randomMPS(2, n, 2, np.random.default_rng(77))
        for i, t in enumerate(p0.tensors):
            out[f"dmrg_{name}_psi{i}"] = t
        for i, t in enumerate(M.tensors):
            out[f"dmrg_{name}_mpo{i}"] = t
        hist = []
        oracle.dmrg(p0.copy(), M, maxdim=24, cutoff=1e-13, maxsweeps=6, history=hist)
        out[f"dmrg_{name}_energy"] = np.array([h[1] for h in hist])
        out[f"dmrg_{name}_maxbond"] = np.array([h[2] for h in hist])
    # one TEBD step (3 gate rows) on a random MPS: bond dimensions, log-norm and <z_i>
    Nt = 8
    gl = oracle.trotterize(sh, -1 * tfim(Nt, 1.0, 0.1, 0.8), 0.05)
    p = random_complex_mps(rng, Nt, 2, 5, center=1)
    for i, t in enumerate(p.tensors):
        out[f"tebd_psi{i}"] = t
    out["tebd_nrows"] = np.array(len(gl.sites))
    for r, (rs, rg) in enumerate(zip(gl.sites, gl.gates)):
        out[f"tebd_r{r}_sites"] = np.array(rs)
        for i, g_ in enumerate(rg):
            out[f"tebd_r{r}_g{i}"] = g_
    oracle.applygates(p, gl, cutoff=1e-12, maxdim=8)
    out["tebd_lognorm"] = np.array(np.log(np.real(p.norm())))
    p.normalize()
    zs = oracle.OpList(Nt)
    for i in range(1, Nt + 1):
        zs.add("z", i)
    out["tebd_z"] = np.real(oracle.inner(sh, p, zs, p))
    out["tebd_bonds"] = np.array([p.bonddim(i) for i in range(1, Nt)])
    out["tebd_center"] = np.array(p.center)
    np.savez_compressed(os.path.join(HERE, "hotpath_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "hotpath_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
