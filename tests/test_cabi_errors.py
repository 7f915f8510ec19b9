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
