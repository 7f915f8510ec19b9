"""Operator lists -- restates /root/reference/src/structures/mps/oplist.jl.
Sites are 1-based as in the reference."""
import numpy as np


class OpList:
    """oplist.jl:1-15."""

    def __init__(self, length):
        self.length = length
        self.ops, self.sites, self.coeffs = [], [], []

    def __len__(self):
        return self.length

    def add(self, ops, sites, coeff=1.0):
        """oplist.jl:35-62 (``add!``): validates, sorts by site, appends."""
        if isinstance(ops, str):
            ops, sites = [ops], [sites]
        if len(ops) != len(sites):
            raise ValueError("The lists must be the same length.")
        perm = np.argsort(sites, kind="stable")
        sites = [int(sites[p]) for p in perm]
        ops = [ops[p] for p in perm]
        last = 0
        for s in sites:
            if s < 0 or s > self.length:
                raise ValueError(f"The sites must be between 1 and {self.length}.")
            if sites.count(s) > 1:
                raise ValueError("There are two or more operators on the same site.")
            if s <= last:
                raise ValueError("The site list must be ordered.")
            last = s
        self.ops.append(ops)
        self.sites.append(sites)
        self.coeffs.append(coeff)
        return self

    def copy(self):
        o = OpList(self.length)
        o.ops = [list(x) for x in self.ops]
        o.sites = [list(x) for x in self.sites]
        o.coeffs = list(self.coeffs)
        return o

    def __add__(self, other):  # oplist.jl:70-81
        o = OpList(max(self.length, other.length))
        o.ops = [list(x) for x in self.ops] + [list(x) for x in other.ops]
        o.sites = [list(x) for x in self.sites] + [list(x) for x in other.sites]
        o.coeffs = list(self.coeffs) + list(other.coeffs)
        return o

    def __rmul__(self, x):  # oplist.jl:84-98
        o = self.copy()
        o.coeffs = [c * x for c in o.coeffs]
        return o

    __mul__ = __rmul__

    def siterange(self):  # oplist.jl:106-113
        rng = 1
        for s in self.sites:
            rng = max(rng, s[-1] - s[0] + 1)
        return rng

    def siteindexs(self, site):  # oplist.jl:121-129 (0-based positions into the list)
        return [i for i, s in enumerate(self.sites) if s[0] == site]

    def totensor(self, st, idx):
        """oplist.jl:134-156.  Dense operator (o1,i1,o2,i2,...) of term ``idx``
        (0-based position) over ``min(siterange, N-site+1)`` sites."""
        ops, sites = self.ops[idx], self.sites[idx]
        rng = min(self.siterange(), self.length - sites[0] + 1)
        prod = None
        k = 0
        for j in range(rng):
            if sites[0] + j in sites:
                o = st.op(ops[k])
                k += 1
            else:
                o = st.op("id")
            prod = o if prod is None else np.multiply.outer(prod, o)
        return self.coeffs[idx] * prod

    def sitetensor(self, st, site):
        """oplist.jl:164-183.  Sum of all terms that start at ``site`` or None."""
        if site < 0 or site > self.length:
            raise ValueError("Site index is out of range.")
        idxs = self.siteindexs(site)
        if not idxs:
            return None
        ten = None
        for i in idxs:
            t = self.totensor(st, i)
            ten = t if ten is None else ten + t
        return ten
