"""Host-side input builders for benches and examples (tiny 2x2 / 4x4 algebra, NumPy only):
hand-written finite-state-machine MPOs and second-order Trotter gates for the spin-1/2
models of BASELINE.json's configs.  Pauli convention of the reference's spinhalf()
(/root/reference/src/lattices/spinhalf.jl:12-14).  These do not go through the reference's
generic MPO(st, H) algorithm (mpo.jl:323-459, a "next" row); tests check them against it."""
import numpy as np
import scipy.linalg as sla

X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
I2 = np.eye(2, dtype=np.complex128)
SM = np.array([[0, 0], [1, 0]], dtype=np.complex128)   # "s-" (spinhalf.jl:20)


def _fsm_mpo(N, onsite, couplings):
    """MPO (w_l, out, in, w_r) for sum_i onsite + sum_i sum_k c_k A_k(i) B_k(i+1)."""
    w = 2 + len(couplings)
    W = np.zeros((w, 2, 2, w), dtype=np.complex128)
    W[0, :, :, 0] = I2
    W[w - 1, :, :, w - 1] = I2
    W[0, :, :, w - 1] = onsite
    for k, (c, A, B) in enumerate(couplings):
        W[0, :, :, 1 + k] = c * A
        W[1 + k, :, :, w - 1] = B
    tensors = []
    for i in range(N):
        t = W
        if i == 0:
            t = t[0:1]
        if i == N - 1:
            t = t[:, :, :, w - 1:w]
        tensors.append(np.ascontiguousarray(t))
    return tensors


def xxz_mpo(N, delta=1.0):
    """H = sum_i (x x + y y + delta z z), w = 5."""
    return _fsm_mpo(N, np.zeros((2, 2)), [(1.0, X, X), (1.0, Y, Y), (delta, Z, Z)])


def tfim_mpo(N, h=1.0, g=0.05, J=1.2):
    """H = sum_i (h x + g z) + J sum_i z z, w = 3 (examples/dmrg.jl:9-24)."""
    return _fsm_mpo(N, h * X + g * Z, [(J, Z, Z)])


def two_site_terms(N, onsite, bond):
    """Per-bond dense operators the way sitetensor() assembles them (oplist.jl:134-183):
    the term starting at site i carries the on-site part of site i only; the last site's
    on-site part is a separate one-site term."""
    terms = []
    for i in range(1, N):
        terms.append((i, np.kron(onsite, I2) + bond))
    terms.append((N, onsite))
    return terms


def trotter_gates(N, onsite, bond, dt, evol="imag", order=2):
    """Second-order Trotter rows for a nearest-neighbour Hamiltonian, same schedule and gate
    layout (out1,in1,out2,in2) as trotterize(): gatelist.jl:75-121.  ``onsite``/``bond`` are the
    2x2 / 4x4 (kron(site_i, site_i+1)) operator matrices of H as passed to tebd (callers pass -H
    for imaginary time, -iH-... for QJMC)."""
    t = -1j * dt if evol == "real" else dt
    rows_sites, rows_gates = [], []
    terms = two_site_terms(N, onsite, bond)
    for r in (1, 2):
        time = t / 2 if (r < 2 and order == 2) else t
        ss, gg = [], []
        site = r
        while site <= N:
            s, h = terms[site - 1]
            if h.shape == (4, 4):
                u = sla.expm(time * h)                       # rows (o1,o2), cols (i1,i2), site i most significant
                g = u.reshape(2, 2, 2, 2).transpose(0, 2, 1, 3)   # (o1,o2,i1,i2) -> (o1,i1,o2,i2)
            else:
                g = sla.expm(time * h)
            if h.shape == (4, 4) or np.any(h != 0):   # no on-site term at the last site -> no gate (oplist.jl:171)
                ss.append(s)
                gg.append(np.ascontiguousarray(g))
            site += 2
        rows_sites.append(ss)
        rows_gates.append(gg)
    if order == 2:
        rows_sites.append(rows_sites[0])
        rows_gates.append(rows_gates[0])
    return rows_sites, rows_gates


def random_canonical_mps(N, d, chi, seed=0, dtype=np.complex128):
    """Seeded random right-canonical MPS with bond dimensions min(d^i, d^(N-i), chi), centre 1
    (throughput runs only; SURVEY 8(d) C3/C4 allow any seeded canonical state)."""
    rng = np.random.default_rng(seed)
    dims = [min(d ** i, d ** (N - i), chi) for i in range(N + 1)]
    tensors = []
    for i in range(N):
        l, r = dims[i], dims[i + 1]
        a = rng.standard_normal((l, d * r)) + 1j * rng.standard_normal((l, d * r))
        q, _ = np.linalg.qr(a.T.conj())          # (d*r, l) orthonormal columns
        tensors.append(np.ascontiguousarray(q.conj().T).reshape(l, d, r, order='F'))
    t0 = tensors[0]
    tensors[0] = t0 / np.linalg.norm(t0)
    return tensors


def j1j2_cylinder_terms(Lx, Ly, J1=1.0, J2=0.5):
    """Operator list [(ops, sites, coeff)] of H = J1 sum_<ij> sigma_i.sigma_j + J2 sum_<<ij>> sigma_i.sigma_j on an Lx x Ly
    cylinder (periodic in y, open in x), site = x*Ly + y + 1 -- the MPS ordering of BASELINE.json's config 5.
    Feed to tnb200.mpo.MPO (w = 20 after compression for Ly = 6)."""
    def idx(x, y):
        return x * Ly + (y % Ly) + 1
    bonds = set()
    for x in range(Lx):
        for y in range(Ly):
            i = idx(x, y)
            for dx, dy, J in ((0, 1, J1), (1, 0, J1), (1, 1, J2), (1, -1, J2)):
                if x + dx >= Lx:
                    continue
                j = idx(x + dx, y + dy)
                if i != j:
                    bonds.add((min(i, j), max(i, j), J))
    terms = []
    for i, j, J in sorted(bonds):
        for o in (X, Y, Z):
            terms.append(([o, o], [i, j], J))
    return terms
