"""GPU parity against the committed golden vectors (no oracle call on the checked values): matvec / blocks 1e-13,
singular values 1e-12 sigma_max with the identical kept rank, DMRG energies per sweep 1e-10, TEBD observables 1e-8."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_golden.npz"))


def _list(prefix):
    out, i = [], 0
    while f"{prefix}{i}" in G:
        out.append(G[f"{prefix}{i}"])
        i += 1
    return out


def test_matvec_blocks_calculate():
    import tnb200
    psi = tnb200.GMPS(1, 2, _list("mv_psi"), 3)
    H = tnb200.GMPS(2, 2, _list("mv_mpo"), 0)
    P = tnb200.ProjMPS(psi, H, psi, center=3)
    for idx, key in ((2, "mv_L"), (5, "mv_R")):
        assert np.linalg.norm(P.block(idx) - G[key]) <= 1e-13 * np.linalg.norm(G[key])
    out = P.product(G["mv_theta"], False)
    assert np.linalg.norm(out - G["mv_out"]) <= 1e-13 * np.linalg.norm(G["mv_out"])
    assert abs(P.calculate() - complex(G["mv_calculate"])) <= 1e-12 * abs(complex(G["mv_calculate"]))


def test_svd_truncation():
    import tnb200
    for name, kw in (("full", {}), ("cut", dict(cutoff=1e-10)), ("max", dict(maxdim=9)), ("min", dict(cutoff=1e-2, mindim=7))):
        _, S, _ = tnb200.svd(G["svd_x"], 2, **kw)
        s, want = np.real(np.diag(S)), G[f"svd_S_{name}"]
        assert s.shape == want.shape, name
        assert np.max(np.abs(s - want)) <= 1e-12 * want[0]


def test_dmrg_histories():
    import tnb200
    for name in ("tfim12", "xxz10"):
        g = tnb200.GMPS(1, 2, _list(f"dmrg_{name}_psi"), 1)
        M = tnb200.GMPS(2, 2, _list(f"dmrg_{name}_mpo"), 0)
        hist = []
        tnb200.dmrg(g, M, maxdim=24, cutoff=1e-13, maxsweeps=6, history=hist)
        e = np.array([h[1] for h in hist])
        assert np.max(np.abs((e - G[f"dmrg_{name}_energy"]) / G[f"dmrg_{name}_energy"])) < 1e-10
        assert [h[2] for h in hist] == list(G[f"dmrg_{name}_maxbond"])


def test_tebd_step():
    import tnb200
    g = tnb200.GMPS(1, 2, _list("tebd_psi"), 1)
    nrows = int(G["tebd_nrows"])
    sites = [list(G[f"tebd_r{r}_sites"]) for r in range(nrows)]
    gates = [_list(f"tebd_r{r}_g") for r in range(nrows)]
    gl = tnb200.GateList(2, sites, gates)
    tnb200.applygates(g, gl, cutoff=1e-12, maxdim=8)
    assert abs(np.log(np.real(g.norm())) - float(G["tebd_lognorm"])) < 1e-10
    g.normalize()
    N = len(g)
    assert g.center == int(G["tebd_center"])
    assert [g.bonddim(i) for i in range(1, N)] == list(G["tebd_bonds"])
    z = np.real(g.expect([tnb200.models.Z] * N, list(range(1, N + 1))))
    assert np.max(np.abs(z - G["tebd_z"])) < 1e-8
