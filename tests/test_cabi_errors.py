"""GPU: error behaviour at the C-ABI boundary mirrors the reference's error("...") convention: bad arguments
return a non-zero status with a message instead of crashing."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_errors_are_reported_not_fatal():
    import tnb200
    rng = np.random.default_rng(0)
    t = [rng.standard_normal((1, 2, 3)) + 0j, rng.standard_normal((3, 2, 1)) + 0j]
    psi = tnb200.GMPS(1, 2, t, 1)
    with pytest.raises(tnb200.TNError, match="out of range"):
        psi.movecenter(5)                                   # gmps.jl:91 error("The index is out of range.")
    with pytest.raises(tnb200.TNError):
        tnb200.GMPS(1, 2, [t[0], rng.standard_normal((4, 2, 1)) + 0j], 1)     # mismatching bond dimensions
    with pytest.raises(tnb200.TNError):
        tnb200.GMPS(3, 2, t, 1)                             # rank must be 1 or 2
    H = tnb200.GMPS(2, 2, tnb200.models.tfim_mpo(3))
    with pytest.raises(tnb200.TNError):
        tnb200.ProjMPS(psi, H, psi)                         # projmps.jl:30 "GMPS must share the same length."
    with pytest.raises(tnb200.TNError):
        tnb200.dmrg(H, H)                                   # dmrg.jl:132 "Psi must be a GMPS of rank 1 (vector)."
    # gate / jump / operator sites beyond the chain are rejected before psi is touched (gatelist.jl:209-217 would throw a BoundsError)
    g2 = rng.standard_normal((2, 2, 2, 2)) + 0j
    for bad in (2, 0, 7):
        with pytest.raises(tnb200.TNError, match="out of range"):
            tnb200.applygates(psi, tnb200.GateList(2, [[bad]], [[g2]]), cutoff=0.0)
    with pytest.raises(tnb200.TNError, match="out of range"):
        tnb200.applygates(psi, tnb200.GateList(2, [[3]], [[np.eye(2) + 0j]]), cutoff=0.0)
    with pytest.raises(tnb200.TNError, match="out of range"):
        psi.expect([np.eye(2) + 0j], [3])
    gl = tnb200.GateList(2, [[1]], [[g2]])
    with pytest.raises(tnb200.TNError, match="out of range"):
        tnb200.qjmc_simulation(psi, gl, [4], [np.eye(2) + 0j], [1.0], 1, 0.01, uniforms=np.zeros(3))
    # the library is still usable afterwards
    psi.movecenter(2)
    assert abs(abs(psi.norm()) - np.linalg.norm(np.tensordot(t[0], t[1], axes=([2], [0])))) < 1e-12


def test_two_site_chain_and_chi1_edge_cases():
    """N = 2 (both environment blocks are edge blocks) and a chi = 1 product state."""
    import oracle
    import tnb200
    sh = oracle.spinhalf()
    from models import tfim, ed_ground_energy
    for N in (2, 3):
        Hl = tfim(N, 1.0, 0.3, 0.7)
        M = tnb200.GMPS(2, 2, tnb200.models.tfim_mpo(N, 1.0, 0.3, 0.7))
        g = tnb200.GMPS.from_host(oracle.productMPS(sh, ["up"] * N))
        g, E = tnb200.dmrg(g, M, maxdim=8, cutoff=1e-14, maxsweeps=4)
        assert abs(E - ed_ground_energy(sh, Hl)) < 1e-10 * abs(E)
