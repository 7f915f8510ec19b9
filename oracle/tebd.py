"""TEBD driver -- restates /root/reference/src/algorithms/mps/tebd.jl:3-100, including the projector branch (:22-41, :67-73:
every ``projection_every`` steps psi <- vmps(psi, -P_1 psi, -P_2 psi, ...)) and MPSProjector (structures/mps/projector.jl)."""
import numpy as np
from .gatelist import trotterize, applygates
from .gmps import inner


class MPSProjector:
    """projector.jl:1-50: |right><left| / constant, constant = <left|right> unless given."""

    rank = 2

    def __init__(self, left, right=None, constant=0.0):
        right = left if right is None else right
        if left.dim != right.dim:
            raise ValueError("The MPS must share the same physical dimension.")
        if len(left) != len(right):
            raise ValueError("The MPS must be the same length.")
        self.dim, self.leftMPS, self.rightMPS = left.dim, left, right
        self.constant = overlap(left, right) if constant == 0.0 else constant

    def apply(self, psi):          # projector.jl:41-47, applyMPO(O, psi) = (<left|psi> / c) * right
        return self.rightMPS.scale(overlap(self.leftMPS, psi) / self.constant)


def overlap(phi, psi):
    """inner(phi, psi) (mpo.jl:181-217 for two MPS): <phi|psi>."""
    prod = np.ones((1, 1), dtype=np.complex128)
    for i in range(1, len(psi) + 1):
        prod = np.tensordot(np.conj(phi[i]), np.tensordot(prod, psi[i], axes=([1], [0])), axes=([0, 1], [0, 1]))
    return prod[0, 0]


def project_out(psi, projs, cutoff, maxdim):
    """tebd.jl:68-73 (and :36-41): psi <- vmps(psi, -1*(P_1*psi), ...)."""
    from .vmps import vmps
    from .mpo import applyMPO
    psis = [psi]
    for P in projs:
        Ppsi = P.apply(psi) if isinstance(P, MPSProjector) else applyMPO(P, psi)
        psis.append(Ppsi.scale(-1))
    return vmps(*psis, cutoff=cutoff, maxdim=maxdim)


def tebd(st, psi, H, dt, tmax, save, observers=(), projectors=(), cutoff=1e-12, maxdim=0, mindim=1,
         evol="imag", order=2, norm=0.0, verbose=False, projection_every=10, variational_cutoff=None):
    if st.dim != psi.dim:
        raise ValueError("Sitetypes and MPS physical dimensions do not match.")
    if len(psi) != len(H):
        raise ValueError("OpList and MPS have different lengths.")
    gates = trotterize(st, H, dt, evol=evol, order=order)
    variational_cutoff = cutoff if variational_cutoff is None else variational_cutoff
    projs = []
    for proj in projectors:                                   # tebd.jl:26-35
        if isinstance(proj, MPSProjector):
            projs.append(proj)
        elif proj.rank == 1:
            projs.append(MPSProjector(proj, proj))
        elif proj.rank == 2:
            projs.append(proj)
        else:
            raise ValueError("Only MPS, MPO and MPSProjectors are supported as projectors.")
    if projs:                                                 # tebd.jl:36-41
        psi = project_out(psi, projs, variational_cutoff, maxdim)
    nsteps = int(round(tmax / dt))
    save = dt if save < dt else save
    nsave = int(round(save / dt))
    normal = float(norm)
    energy = float(np.real(np.sum(inner(st, psi, H, psi))))
    for ob in observers:
        ob.measure(0.0, psi, normal, energy)
    converged = False
    step = 0
    while not converged:
        applygates(psi, gates, mindim=mindim, maxdim=maxdim, cutoff=cutoff)
        if projs and (step + 1) % projection_every == 0:      # tebd.jl:67-73
            psi = project_out(psi, projs, variational_cutoff, maxdim)
        psinorm = np.log(np.real(psi.norm()))
        normal += psinorm
        psi.normalize()
        step += 1
        if step >= nsteps:
            converged = True
        if step % nsave == 0:
            if verbose:
                print("time=%.4f, energy=%.12f, maxbonddim=%d" % (step * dt, energy, psi.maxbonddim()))
            energy = float(np.real(np.sum(inner(st, psi, H, psi))))
            for ob in observers:
                ob.measure(step * dt, psi, normal, energy)
                converged = converged or ob.checkdone()
    return psi, energy


class TEBDOperators:
    """tebd.jl:199-219."""

    def __init__(self, st, oplist):
        self.times, self.measurements, self.oplist, self.st = [], [], oplist, st

    def measure(self, time, psi, norm, energy):
        self.times.append(time)
        self.measurements.append(inner(self.st, psi, self.oplist, psi))

    def checkdone(self):
        return False
