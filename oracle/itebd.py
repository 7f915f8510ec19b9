"""Infinite TEBD in Vidal form -- restates /root/reference/src/structures/mps/igmps.jl:8-60 (container) and
/root/reference/src/algorithms/mps/itebd.jl:3-119 (gate construction and ``_itebd_apply_gates_mps!``).

psi.singulars[i] sits to the LEFT of psi.tensors[i] (itebd.jl:78-83); sites are 1-based, the unit cell is periodic.
The reference measures energies through infinite environments (igmps.jl:155-340, out of scope); ``bond_energy`` below is the
oracle's own measurement for a two-site cell in (approximately) canonical form and is only used by tests."""
import numpy as np
from .tensors import contract, moveidx, combineidxs, svd, tensor_exp


class IGMPS:
    """igmps.jl:8-33 (the environment caches ``lefts``/``rights`` are not restated)."""

    def __init__(self, rank, dim, length):
        self.rank, self.dim = rank, dim
        self.tensors = [np.zeros((1,) + (dim,) * rank + (1,), dtype=np.complex128) for _ in range(length)]
        self.singulars = [np.ones(1) for _ in range(length)]
        self.norms = [0.0 for _ in range(length)]

    def __len__(self):
        return len(self.tensors)

    def maxbonddim(self):  # igmps.jl:39
        return max(t.shape[0] for t in self.tensors)

    def copy(self):
        o = IGMPS(self.rank, self.dim, len(self))
        o.tensors = [t.copy() for t in self.tensors]
        o.singulars = [s.copy() for s in self.singulars]
        o.norms = list(self.norms)
        return o


def iMPS(length, A):
    """igmps.jl:53-59: product state, the same local vector A on every site of the cell."""
    A = np.asarray(A, dtype=np.complex128)
    psi = IGMPS(1, A.shape[0], length)
    for i in range(length):
        psi.tensors[i] = A.reshape(1, A.shape[0], 1).copy()
    return psi


def itebd_gate(st, H, dt, evol="imag"):
    """itebd.jl:19-20: exp(+-dt * sitetensor(H, st, 1)) over the whole cell (callers pass -H for imaginary time)."""
    gate = H.sitetensor(st, 1)
    L = len(H)
    return tensor_exp((-1j if evol == "real" else 1) * dt * gate, [2 * i for i in range(1, L + 1)])


def itebd_apply_gates_mps(psi, gate, mindim=1, maxdim=0, cutoff=1e-12):
    """itebd.jl:71-119, any cell length."""
    L = len(psi)
    for i in range(1, L + 1):
        idxs = [(i - 1 + k) % L + 1 for k in range(L)]
        prod = np.diag(psi.singulars[idxs[0] - 1]).astype(np.complex128)
        for j in range(1, L + 1):
            prod = contract(prod, psi.tensors[idxs[j - 1] - 1], 1 + j, 1)
            j2 = 1 if j + 1 > L else j + 1
            prod = contract(prod, np.diag(psi.singulars[idxs[j2 - 1] - 1]).astype(np.complex128), 2 + j, 1)
        prod = contract(prod, gate, [1 + j for j in range(1, L + 1)], [2 * j for j in range(1, L + 1)])
        prod = moveidx(prod, 2, prod.ndim)
        tensors, singulars = [], []
        for _ in range(L - 1):
            prod, cmb = combineidxs(prod, list(range(3, prod.ndim + 1)))
            U, S, prod = svd(prod, 3, mindim=mindim, maxdim=maxdim, cutoff=cutoff)
            tensors.append(U)
            singulars.append(np.real(np.diag(S)).copy())
            prod = np.reshape(prod, (S.shape[0],) + tuple(cmb[1]), order='F')
        tensors.append(prod)
        s0 = psi.singulars[idxs[0] - 1]
        tensors[0] = contract(np.diag(1.0 / s0).astype(np.complex128), tensors[0], 2, 1)
        tensors[-1] = contract(tensors[-1], np.diag(1.0 / s0).astype(np.complex128), 3, 1)
        for j in range(2, L + 1):
            nrm = np.sqrt(np.sum(singulars[j - 2] ** 2))
            psi.norms[idxs[j - 1] - 1] += np.log(nrm)
            psi.singulars[idxs[j - 1] - 1] = singulars[j - 2] / nrm
        for j in range(1, L + 1):
            psi.tensors[idxs[j - 1] - 1] = tensors[j - 1]


def bond_energy(psi, h2):
    """<h> per bond of a two-site cell, h2 = (o1,i1,o2,i2): average over the two bonds of <Theta|h|Theta>/<Theta|Theta> with
    Theta = S_a G_a S_b G_b S_a (exact when the cell is in canonical form).  Test helper, not in the reference."""
    assert len(psi) == 2 and psi.rank == 1
    out = []
    for a, b in ((0, 1), (1, 0)):
        Sa, Sb = psi.singulars[a], psi.singulars[b]
        th = np.einsum('l,lsm,m,mtr,r->lstr', Sa, psi.tensors[a], Sb, psi.tensors[b], Sa)
        hth = np.einsum('sutv,lutr->lstr', h2, th)
        out.append(np.vdot(th, hth) / np.vdot(th, th))
    return complex(np.mean(out))
