"""Projected environments -- restates
/root/reference/src/structures/mps/projmps.jl, abstractprojmps.jl, projmpssum.jl.

Block layout: (bra bond, [MPO bond ...], ket bond); out-of-range blocks are
``ones(1,...,1)`` (abstractprojmps.jl:33-46).  Sites are 1-based."""
import numpy as np
from .tensors import contract, moveidx


class ProjMPS:
    """projmps.jl:1-42."""

    def __init__(self, objects, squared=False, center=1, coeff=1.0, rank=1):
        if squared and rank == 1:
            raise ValueError("Squared and rank one are incomptible.")
        if len(objects) < 2:
            raise ValueError("The projection must have a braket structure")
        if any(o.dim != objects[0].dim for o in objects):
            raise ValueError("GMPS must share the same physical dim.")
        if any(len(o) != len(objects[0]) for o in objects):
            raise ValueError("GMPS must share the same length.")
        if objects[0].rank != 1 or objects[-1].rank != 1:
            raise ValueError("The projection must have a braket structure")
        self.objects = list(objects)
        self.blocks = [self.edgeblock() for _ in range(len(objects[0]))]
        self.squared, self.rank, self.center, self.coeff = squared, rank, 0, coeff
        self.movecenter(center)

    def __len__(self):
        return len(self.objects[0])

    def edgeblock(self):  # abstractprojmps.jl:33-35
        return np.ones((1,) * len(self.objects), dtype=np.complex128)

    def block(self, idx):  # abstractprojmps.jl:43-46
        if idx < 1 or idx > len(self):
            return self.edgeblock()
        return self.blocks[idx - 1]

    def buildleft(self, idx):
        """projmps.jl:50-66: L'(a',w',b') = sum L(a,w,b) conj(A)(a,s,a') M(w,s,s',w') A(b,s',b')."""
        left = self.block(idx - 1)
        A1 = np.conj(self.objects[0][idx])
        A2 = self.objects[-1][idx]
        prod = contract(left, A1, 1, 1)
        for i in range(1, len(self.objects) - 1):
            M = self.objects[i][idx]
            prod = contract(prod, M, [1, prod.ndim - 1], [1, 2])
        prod = contract(prod, A2, [1, prod.ndim - 1], [1, 2])
        self.blocks[idx - 1] = prod

    def buildright(self, idx):
        """projmps.jl:74-95: mirror image; result (a_l, w_l, b_l)."""
        right = self.block(idx + 1)
        A1 = np.conj(self.objects[0][idx])
        A2 = self.objects[-1][idx]
        prod = contract(A2, right, 3, right.ndim)
        n = len(self.objects)
        for i in range(1, n - 1):
            M = self.objects[n - 1 - i][idx]
            prod = contract(M, prod, [3, 4], [2, prod.ndim])
        prod = contract(A1, prod, [2, 3], [2, prod.ndim])
        self.blocks[idx - 1] = prod

    def movecenter(self, idx):  # abstractprojmps.jl:60-81
        N = len(self)
        if self.center == 0:
            for i in range(1, idx):
                self.buildleft(i)
            for i in range(1, N - idx + 1):
                self.buildright(N + 1 - i)
        elif idx > self.center:
            for i in range(1, idx - self.center + 1):
                self.buildleft(self.center - 1 + i)
        elif idx < self.center:
            for i in range(1, self.center - idx + 1):
                self.buildright(self.center + 1 - i)
        self.center = idx

    def product(self, A, direction=False, nsites=2):
        """projmps.jl:103-145.  rank-2 branch (:107-134) is the H_eff matvec, in
        the REFERENCE's contraction order: (L.M1.M2) first, then Theta, then R."""
        if self.rank == 2 and not self.squared:
            site = self.center - nsites + 1 if direction else self.center
            left = self.block(site - 1)
            right = self.block(site + nsites)
            prod = moveidx(left, left.ndim, 2)
            nobj = len(self.objects)
            for i in range(1, nsites + 1):
                for j in range(1, nobj - 1):
                    M = self.objects[j][site - 1 + i]
                    if j == 1:
                        prod = contract(prod, M, 1 + 2 * i, 1)
                    else:
                        prod = contract(prod, M, [1 + 2 * i, prod.ndim - 1], [1, 2])
                prod = moveidx(prod, prod.ndim - 1, 2 * i + 2)
            prod = contract(prod, A, [2 * k for k in range(1, nsites + 2)], list(range(1, nsites + 2)))
            prod = contract(prod, right, list(range(2 + nsites, prod.ndim + 1)), list(range(2, right.ndim + 1)))
        else:
            prod = self.project(A, direction, nsites)
            prod2 = np.conj(prod) if self.squared else 1
            # NB projmps.jl:141 contracts ``prod`` with A without conjugation
            s = np.tensordot(prod, A, axes=(list(range(prod.ndim)), list(range(prod.ndim))))
            prod = s * prod2
        return prod * self.coeff

    def product_optimal(self, A, direction=False):
        """Same rank-2 two-site matvec in the flop-optimal order
        (L.Theta).W.R used by the CUDA path; equal to :meth:`product` up to
        rounding.  Single MPO layer, nsites = 2."""
        site = self.center - 1 if direction else self.center
        L = self.block(site - 1)
        R = self.block(site + 2)
        M1 = self.objects[1][site]
        M2 = self.objects[1][site + 1]
        T = np.tensordot(L, A, axes=([2], [0]))                 # (a,w,s1',s2',b')
        W = np.tensordot(M1, M2, axes=([3], [0]))               # (w,s1,s1',s2,s2',w2)
        T = np.tensordot(T, W, axes=([1, 2, 3], [0, 2, 4]))     # (a,b',s1,s2,w2)
        out = np.tensordot(T, R, axes=([4, 1], [1, 2]))         # (a,s1,s2,a')
        return out * self.coeff

    def project(self, A, direction=False, nsites=2):
        """projmps.jl:153-185."""
        site = self.center - nsites + 1 if direction else self.center
        left = self.block(site - 1)
        right = self.block(site + nsites)
        prod = moveidx(left, left.ndim, 1)
        for i in range(1, nsites + 1):
            prod = contract(prod, np.conj(self.objects[0][site - 1 + i]), 1 + i, 1)
            for j in range(1, len(self.objects) - 1):
                prod = contract(prod, self.objects[j][site - 1 + i], [1 + i, prod.ndim - 1], [1, 2])
            prod = moveidx(prod, prod.ndim - 1, 1 + i)
        prod = contract(prod, right, list(range(2 + nsites, prod.ndim + 1)), list(range(1, right.ndim)))
        return prod

    def calculate(self):
        """projmps.jl:192-216: the fully contracted <bra| ... |ket> at the centre."""
        site = self.center
        left = self.block(site - 1)
        right = self.block(site + 1)
        A1 = np.conj(self.objects[0][site])
        A2 = self.objects[-1][site]
        prod = contract(left, A1, 1, 1)
        for i in range(1, len(self.objects) - 1):
            prod = contract(prod, self.objects[i][site], [1, prod.ndim - 1], [1, 2])
        prod = contract(prod, A2, [1, prod.ndim - 1], [1, 2])
        return self.coeff * np.tensordot(prod, right, axes=(list(range(prod.ndim)), list(range(prod.ndim))))


class ProjMPSSum:
    """projmpssum.jl:1-108."""

    def __init__(self, projs, center=1):
        self.projs = list(projs)
        self.center = center
        self.movecenter(center)

    def __len__(self):
        return len(self.projs[0])

    def movecenter(self, idx):
        for p in self.projs:
            p.movecenter(idx)
        self.center = idx

    def product(self, A, direction=False, nsites=2):
        out = None
        for p in self.projs:
            t = p.product(A, direction, nsites)
            out = t if out is None else out + t
        return out

    def product_optimal(self, A, direction=False):
        out = None
        for p in self.projs:
            t = p.product_optimal(A, direction)
            out = t if out is None else out + t
        return out

    def project(self, A, direction=False, nsites=2):
        out = None
        for p in self.projs:
            t = p.project(A, direction, nsites)
            out = t if out is None else out + t
        return out

    def calculate(self):
        return sum(p.calculate() for p in self.projs)
