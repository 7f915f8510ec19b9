"""GMPS container and gauge moves -- restates
/root/reference/src/structures/mps/gmps.jl, abstractmps.jl and mps.jl.

Site numbers are 1-based (``psi[i]`` is the reference's ``psi[i]``); ``center == 0``
means "no orthogonality centre set" (gmps.jl:20)."""
import numpy as np
from .tensors import contract, svd, moveidx


class GMPS:
    """gmps.jl:8-13.  rank 1 = MPS (chi_l, s, chi_r); rank 2 = MPO (w_l, out, in, w_r)."""

    def __init__(self, rank, dim, tensors, center=0):
        self.rank, self.dim = rank, dim
        self.tensors = [np.asarray(t, dtype=np.complex128) for t in tensors]
        self.center = center

    @classmethod
    def zeros(cls, rank, dim, length):  # gmps.jl:15-21
        return cls(rank, dim, [np.zeros((1,) + (dim,) * rank + (1,), np.complex128) for _ in range(length)], 0)

    def __len__(self):
        return len(self.tensors)

    def __getitem__(self, i):
        return self.tensors[i - 1]

    def __setitem__(self, i, x):
        self.tensors[i - 1] = np.asarray(x, dtype=np.complex128)

    def copy(self):
        return GMPS(self.rank, self.dim, [t.copy() for t in self.tensors], self.center)

    # --- abstractmps.jl:62-78
    def bonddim(self, site):
        if site < 1 or site > len(self):
            return None
        return self[site + 1].shape[0]

    def maxbonddim(self):
        D = 0
        for i in range(1, len(self)):
            D = max(D, self.bonddim(i))
        return D

    def scale(self, a):  # abstractmps.jl:99-109
        phi = self.copy()
        c = phi.center if phi.center != 0 else 1
        phi[c] = phi[c] * a
        return phi

    # --- gmps.jl:29-51
    def norm(self):
        """gmps.jl:29-37: sqrt(<A_c, A_c>) at the centre (complex scalar)."""
        if self.center == 0:
            self.movecenter(1)
        A = self[self.center]
        return np.sqrt(np.vdot(A, A) + 0j)

    def normalize(self):
        if self.center == 0:
            self.movecenter(1)
        self[self.center] = self[self.center] * self.norm() ** -1

    # --- gmps.jl:60-112
    def moveleft(self, idx, **kw):
        """gmps.jl:60-67: SVD of site ``idx`` w.r.t. its left bond; U (right-
        orthonormal) stays, S*V^H is absorbed into site idx-1."""
        if 1 < idx <= len(self):
            U, S, V = svd(self[idx], 1, **kw)
            V = contract(S, V, 2, 1)
            self[idx] = U
            self[idx - 1] = contract(self[idx - 1], V, 2 + self.rank, 2)

    def moveright(self, idx, **kw):
        """gmps.jl:75-82."""
        if 0 < idx < len(self):
            U, S, V = svd(self[idx], 2 + self.rank, **kw)
            V = contract(S, V, 2, 1)
            self[idx] = U
            self[idx + 1] = contract(V, self[idx + 1], 2, 1)

    def movecenter(self, idx, **kw):
        """gmps.jl:90-112."""
        N = len(self)
        if idx < 1 or idx > N:
            raise IndexError("The index is out of range.")
        if self.center == 0:
            for i in range(1, idx):
                self.moveright(i, **kw)
            for i in range(1, N - idx + 1):
                self.moveleft(N + 1 - i, **kw)
        elif idx > self.center:
            for i in range(self.center, idx):
                self.moveright(i, **kw)
        elif idx < self.center:
            for i in range(1, self.center - idx + 1):
                self.moveleft(self.center + 1 - i, **kw)
        self.center = idx

    def entropy(self, site):  # gmps.jl:184-189
        self.movecenter(site)
        _, S, _ = svd(self[site], -1)
        S2 = np.real(np.diag(S)) ** 2
        return float(-np.sum(S2 * np.log(S2)))

    def spectrum(self, site):
        """Singular values across bond (site, site+1); the gauge-invariant
        quantity parity tests compare (entropy() uses the same SVD)."""
        self.movecenter(site)
        _, S, _ = svd(self[site], -1)
        return np.real(np.diag(S)).copy()

    # --- gmps.jl:199-267
    def replacesites(self, A, site, direction=False, normalize=False, **kw):
        """gmps.jl:199-267.  Split an n-site tensor back into site tensors by
        repeated truncated SVD.  direction False = sweeping right (centre ends on
        the last site), True = sweeping left (centre ends on ``site``)."""
        r = self.rank
        nsites = (A.ndim - 2) // r
        N = len(self)
        if nsites == 1:  # gmps.jl:204-213
            self[site] = A
            nxt = site + 1 - 2 * int(direction)
            if 0 < nxt <= N:
                self.movecenter(nxt)
            if normalize:
                self.normalize()
            return None
        U = A
        for i in range(1, nsites):
            nd = U.ndim
            if direction:  # gmps.jl:218-235
                site1 = site + nsites - i
                lead = U.shape[:nd - r - 1]
                M = np.reshape(U, (int(np.prod(lead)), -1), order='F')
                Uu, S, V = svd(M, -1, **kw)
                k = S.shape[0]
                U = np.reshape(Uu @ S, lead + (k,), order='F')
                D = 1 if site1 == N else self[site1 + 1].shape[0]
                self[site1] = np.reshape(V, (k,) + (self.dim,) * r + (D,), order='F')
            else:  # gmps.jl:236-256
                site1 = site + i - 1
                trail = U.shape[1 + r:]
                M = np.reshape(U, (-1, int(np.prod(trail))), order='F')   # ((chi_l,s1..), rest)
                Uu, S, V = svd(M.T, -1, **kw)          # rows = rest, cols = (chi_l,s1..)
                k = S.shape[0]
                Unew = (Uu @ S).T                        # (k, rest)
                U = np.reshape(Unew, (k,) + trail, order='F')
                D = 1 if site1 == 1 else self[site1 - 1].shape[1 + r]
                Vt = np.reshape(V, (k, D) + (self.dim,) * r, order='F')
                self[site1] = moveidx(Vt, 1, -1)
        site1 = site if direction else site + nsites - 1
        self[site1] = U
        self.center = site1
        if normalize:
            self.normalize()
        return True


def randomGMPS(rank, dim, length, bonddim, rng=None):
    """gmps.jl:276-291.  Real Gaussian entries (Julia's ``randn``); the seeded
    NumPy generator replaces Julia's unseeded global RNG."""
    rng = np.random.default_rng(1234) if rng is None else rng
    tensors = []
    for i in range(1, length + 1):
        D1 = 1 if i == 1 else bonddim
        D2 = 1 if i == length else bonddim
        tensors.append(rng.standard_normal((D1,) + (dim,) * rank + (D2,)))
    psi = GMPS(rank, dim, tensors, 0)
    psi.movecenter(length)
    psi.movecenter(1)
    psi[1] = rng.standard_normal((1,) + (dim,) * rank + (min(dim ** rank, bonddim),))
    psi.normalize()
    return psi


def randomMPS(dim, length, bonddim, rng=None):  # mps.jl:32-34
    return randomGMPS(1, dim, length, bonddim, rng)


def productMPS(st, names):  # mps.jl:68-76
    return GMPS(1, st.dim, [np.reshape(st.state(n), (1, st.dim, 1)) for n in names], 0)


def productMPO(st, names):  # mpo.jl:86-94
    return GMPS(2, st.dim, [np.reshape(st.op(n), (1, st.dim, st.dim, 1)) for n in names], 0)


def inner(st, psi, oplist, phi):
    """mps.jl:87-134.  <psi| O_k |phi> for every term of ``oplist`` (times its
    coefficient), using left/right overlap blocks (bra bond, ket bond)."""
    from .projmps import ProjMPS
    projV = ProjMPS([psi, phi], rank=1, squared=False)
    out = np.zeros(len(oplist.sites), dtype=np.complex128)
    for site in range(1, len(psi) + 1):
        projV.movecenter(site)
        for idx in oplist.siteindexs(site):
            sites = oplist.sites[idx]
            rng = sites[-1] - sites[0] + 1
            left = projV.block(site - 1)
            right = projV.block(site + rng)
            for i in range(1, rng + 1):
                A = np.conj(psi[site - 1 + i])
                B = phi[site - 1 + i]
                if (site - 1 + i) in sites:
                    O = st.op(oplist.ops[idx][sites.index(site - 1 + i)])
                    B = np.einsum('st,ltr->lsr', O, B)
                left = np.tensordot(A, np.tensordot(left, B, axes=([1], [0])), axes=([0, 1], [0, 1]))      # sum_{a,b,s} left(a,b) A(a,s,c) B(b,s,d)
            out[idx] = oplist.coeffs[idx] * np.einsum('ab,ab->', left, right)
    return out


def applyop(st, psi, ops, sites):
    """mps.jl:141-152 (``applyop!``): local operators onto site tensors."""
    for o, s in zip(ops, sites):
        psi[s] = np.einsum('st,ltr->lsr', st.op(o), psi[s])
