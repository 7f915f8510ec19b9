"""Model builders shared by the tests and bench (host-side inputs only).
Built with the oracle's restatement of OpList / spinhalf, exactly as the
reference's examples build them (examples/dmrg.jl:9-24, examples/qjmc.jl:12-40)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle


def tfim(N, h=1.0, g=0.05, J=1.2):
    H = oracle.OpList(N)
    for i in range(1, N + 1):
        H.add("x", i, h)
        H.add("z", i, g)
    for i in range(1, N):
        H.add(["z", "z"], [i, i + 1], J)
    return H


def xxz(N, delta=1.0):
    H = oracle.OpList(N)
    for i in range(1, N):
        H.add(["x", "x"], [i, i + 1], 1.0)
        H.add(["y", "y"], [i, i + 1], 1.0)
        H.add(["z", "z"], [i, i + 1], delta)
    return H


def j1j2_cylinder(Lx, Ly, J1=1.0, J2=0.5):
    """site = x*Ly + y (+1), periodic in y, open in x."""
    N = Lx * Ly
    H = oracle.OpList(N)
    bonds = set()

    def idx(x, y):
        return x * Ly + (y % Ly) + 1
    for x in range(Lx):
        for y in range(Ly):
            i = idx(x, y)
            for (dx, dy, J) in ((0, 1, J1), (1, 0, J1), (1, 1, J2), (1, -1, J2)):
                if x + dx >= Lx:
                    continue
                j = idx(x + dx, y + dy)
                if i == j:
                    continue
                key = (min(i, j), max(i, j), J)
                if key in bonds:
                    continue
                bonds.add(key)
    for (i, j, J) in sorted(bonds):
        for o in ("x", "y", "z"):
            H.add([o, o], [i, j], J)
    return H


def dense_hamiltonian(st, H):
    """Sparse 2^N x 2^N matrix of an OpList; site 1 is the most significant
    factor (independent of any MPS code under test)."""
    N, d = len(H), st.dim
    tot = sp.csr_matrix((d ** N, d ** N), dtype=np.complex128)
    for ops, sites, c in zip(H.ops, H.sites, H.coeffs):
        mats = [sp.identity(d, dtype=np.complex128, format="csr")] * N
        mats = list(mats)
        for o, s in zip(ops, sites):
            mats[s - 1] = sp.csr_matrix(st.op(o))
        m = mats[0]
        for k in range(1, N):
            m = sp.kron(m, mats[k], format="csr")
        tot = tot + c * m
    return tot


def ed_ground_energy(st, H):
    M = dense_hamiltonian(st, H)
    if M.shape[0] <= 512:
        return float(np.linalg.eigvalsh(M.toarray())[0])
    return float(spla.eigsh(M, k=1, which="SA", tol=1e-13)[0][0])


def mps_to_dense(psi):
    """Contract an MPS to a dense vector, site 1 most significant."""
    v = psi[1]
    for i in range(2, len(psi) + 1):
        v = np.tensordot(v, psi[i], axes=([v.ndim - 1], [0]))
    return v.reshape(-1)  # C-order: first site slowest == most significant


# exact-diagonalisation known answers (properties of the Hamiltonians)
KAT = {
    ("tfim", 8): -10.697775115120, ("tfim", 10): -13.517133134389,
    ("tfim", 12): -16.342203690567, ("tfim", 14): -19.171273562215,
    ("heis", 8): -13.499730394752, ("heis", 10): -17.032140829132,
    ("heis", 12): -20.568362531362, ("heis", 14): -24.106898647449,
    ("xxz0.5", 10): -14.361002811946, ("xxz0.5", 12): -17.352593323708,
    ("j1j2_4x3", 12): -21.967151676841, ("j1j2_3x4", 12): -24.688115833628,
}
