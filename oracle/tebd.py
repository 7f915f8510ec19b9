"""TEBD driver -- restates /root/reference/src/algorithms/mps/tebd.jl:3-100
(projector branch :22-41,67-73 is out of scope and not restated)."""
import numpy as np
from .gatelist import trotterize, applygates
from .gmps import inner


def tebd(st, psi, H, dt, tmax, save, observers=(), cutoff=1e-12, maxdim=0, mindim=1,
         evol="imag", order=2, norm=0.0, verbose=False):
    if st.dim != psi.dim:
        raise ValueError("Sitetypes and MPS physical dimensions do not match.")
    if len(psi) != len(H):
        raise ValueError("OpList and MPS have different lengths.")
    gates = trotterize(st, H, dt, evol=evol, order=order)
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
