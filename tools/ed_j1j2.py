"""Exact ground-state energy of H = J1 sum_<ij> sigma_i.sigma_j + J2 sum_<<ij>> sigma_i.sigma_j (Pauli matrices) on an Lx x Ly cylinder in the
S^z_total = 0 sector (independent of any MPS code: bit manipulation + scipy eigsh).  python tools/ed_j1j2.py Lx Ly
Checked against tests/models.py KAT for 4x3 / 3x4 (full-space ED); the 4 x 6 value pins the chi = 4096 DMRG run (tools/bench_c5_exact.py)."""
import sys, itertools, time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def bonds(Lx, Ly, J1=1.0, J2=0.5):
    def idx(x, y):
        return x * Ly + (y % Ly)
    out = set()
    for x in range(Lx):
        for y in range(Ly):
            i = idx(x, y)
            for dx, dy, J in ((0, 1, J1), (1, 0, J1), (1, 1, J2), (1, -1, J2)):
                if x + dx >= Lx:
                    continue
                j = idx(x + dx, y + dy)
                if i != j:
                    out.add((min(i, j), max(i, j), J))
    return sorted(out)


def ground_energy(Lx, Ly):
    N = Lx * Ly
    bs = bonds(Lx, Ly)
    # basis: all N-bit words with N/2 bits set, sorted
    states = np.fromiter((sum(1 << b for b in c) for c in itertools.combinations(range(N), N // 2)), dtype=np.int64, count=-1)
    states.sort()
    dim = len(states)
    diag = np.zeros(dim)
    rows, cols, vals = [], [], []
    ar = np.arange(dim)
    for i, j, J in bs:
        bi, bj = (states >> i) & 1, (states >> j) & 1
        diag += J * np.where(bi == bj, 1.0, -1.0)          # sigma^z sigma^z
        anti = bi != bj                                     # sigma^x sigma^x + sigma^y sigma^y = 2 (s+ s- + s- s+): flips antiparallel pairs
        flipped = states[anti] ^ ((1 << i) | (1 << j))
        rows.append(ar[anti]); cols.append(np.searchsorted(states, flipped)); vals.append(np.full(anti.sum(), 2.0 * J))
    H = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(dim, dim)) + sp.diags(diag)
    w = spla.eigsh(H, k=1, which="SA", tol=1e-14, ncv=40)[0]
    return float(w[0]), dim, len(bs)


if __name__ == "__main__":
    Lx, Ly = int(sys.argv[1]), int(sys.argv[2])
    t0 = time.time()
    e, dim, nb = ground_energy(Lx, Ly)
    print(f"J1-J2 (J2 = 0.5) {Lx} x {Ly} cylinder, N = {Lx * Ly}, {nb} bonds, Sz = 0 sector dim {dim}: E0 = {e:.12f}  ({time.time() - t0:.1f} s)")
