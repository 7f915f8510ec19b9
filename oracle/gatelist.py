"""Gate lists, Trotterisation and gate application -- restates
/root/reference/src/structures/mps/gatelist.jl:8-227.

Gate layout (out1, in1, out2, in2, ...) (gatelist.jl:100,149).  Sites 1-based."""
import numpy as np
from .tensors import contract, moveidx, tensor_exp


def gatesize(gate):  # gatelist.jl:28-32
    if gate.ndim % 2 == 1:
        raise ValueError("The gate must have even number of dimensions.")
    return gate.ndim // 2


class GateList:
    """gatelist.jl:8-14: rows of non-overlapping gates."""

    def __init__(self, length):
        self.length = length
        self.sites, self.gates = [], []

    def __len__(self):
        return self.length

    def add(self, sites, gates):  # gatelist.jl:40-63
        if len(sites) != len(gates):
            raise ValueError("The site and gate list must be the same length.")
        perm = np.argsort(sites, kind="stable")
        sites = [int(sites[p]) for p in perm]
        gates = [gates[p] for p in perm]
        for i in range(len(sites) - 1):
            if sites[i] + gatesize(gates[i]) - 1 >= sites[i + 1]:
                raise ValueError("Gates in one sequence cannot be shared by sites.")
        if sites and sites[-1] + gatesize(gates[-1]) - 1 > self.length:
            raise ValueError("The bond gates cannot exceed the length of the lattice.")
        self.sites.append(sites)
        self.gates.append(gates)


def trotterize(st, ops, dt, order=2, evol="imag"):
    """gatelist.jl:75-121.  NB imaginary time uses exp(+dt*h): callers pass -H."""
    gl = GateList(len(ops))
    rng = ops.siterange()
    order = 1 if rng == 1 else order
    if order < 0 or order > 2:
        raise ValueError("Only trotter order 1 and 2 are supported.")
    dt = -1j * dt if evol == "real" else dt
    for i in range(1, rng + 1):
        time = dt / 2 if (i < rng and order == 2) else dt
        gates, sites = [], []
        site = i
        while site <= len(ops):
            gate = ops.sitetensor(st, site)
            if gate is not None:
                gate = tensor_exp(time * gate, [2 * k for k in range(1, gatesize(gate) + 1)])
                gates.append(gate)
                sites.append(site)
            site += rng
        gl.add(sites, gates)
    if order == 2:
        for i in range(1, rng):
            gl.add(gl.sites[rng - i - 1], gl.gates[rng - i - 1])
    return gl


def applygate(psi, site, gate, direction=False, error=True, **kw):
    """gatelist.jl:137-171.  Merge the sites, contract the gate on the first
    physical index of each site, split again with a truncated SVD."""
    rng = gatesize(gate)
    r = psi.rank
    prod = psi[site]
    for i in range(1, rng):
        prod = contract(prod, psi[site + i], prod.ndim, 1)
    prod = contract(prod, gate, [2 + r * (i - 1) for i in range(1, rng + 1)], [2 * i for i in range(1, rng + 1)])
    for i in range(1, rng + 1):
        prod = moveidx(prod, 2 + (r - 1) * rng + i, 2 + r * (i - 1))
    psi.replacesites(prod, site, direction, False, **kw)
    if error:
        pe = psi[site]
        for i in range(1, rng):
            pe = contract(pe, psi[site + i], pe.ndim, 1)
        return abs(np.vdot(prod, pe)) ** 2
    return 0


def applygates(psi, gates, error=False, **kw):
    """gatelist.jl:191-227 (``applygates!`` passes error=false)."""
    err = 1
    for row in range(len(gates.gates)):
        rsites, rgates = gates.sites[row], gates.gates[row]
        firstsite = rsites[0]
        lastsite = rsites[-1] + gatesize(rgates[-1]) - 1
        direction = not (abs(psi.center - firstsite) < abs(psi.center - lastsite))
        n = len(rgates)
        for i in range(1, n + 1):
            g = n + 1 - i if direction else i
            ctr = rsites[g - 1] + gatesize(rgates[g - 1]) - 1 if direction else rsites[g - 1]
            psi.movecenter(ctr, **kw)
            err *= applygate(psi, rsites[g - 1], rgates[g - 1], direction, error=error, **kw)
    return err
