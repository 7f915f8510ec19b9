"""Variational MPS optimisation -- restates /root/reference/src/algorithms/mps/vmps.jl.

Finds the MPS ``psi`` that maximises the overlap with a sum of projections
``Vs`` (a ProjMPSSum of ProjMPS(psi_k, psi)); used by the reference to compress
sums of MPSs and to apply TEBD projectors (tebd.jl:36-41)."""
import numpy as np
from .projmps import ProjMPS, ProjMPSSum


def vmps_sweeps(psi, Vs, minsweeps=2, maxsweeps=200, tol=1e-10, numconverges=3, verbose=False,
                nsites=2, cutoff=1e-12, maxdim=1000, mindim=1, history=None):
    """vmps.jl:1-92.  Per bond: movecenter!(Vs), vec = conj(project(Vs, A0)),
    replacesites!(psi, vec, site1, direction) (no normalisation).  The cost is
    norm(psi)^2 - 2 |calculate(Vs)| (vmps.jl:16-23)."""

    def calculatecost():
        normal = psi.norm() ** 2
        projcost = Vs.calculate()
        return normal - 2 * abs(projcost)

    def diff(x, y):
        return abs(x - y) if abs(x) < 1e-10 else abs((x - y) / x)

    lastcost = calculatecost()
    D = psi.maxbonddim()
    lastD = D
    direction = False
    converged = False
    convergedsweeps = sweeps = 0
    N = len(psi)
    while not converged:
        for j in range(1, N + 2 - nsites):
            site = N + 1 - j if direction else j
            site1 = site + 1 - nsites if direction else site
            Vs.movecenter(site)
            A0 = psi[site1]
            for i in range(1, nsites):
                A0 = np.tensordot(A0, psi[site1 + i], axes=([A0.ndim - 1], [0]))
            vec = np.conj(Vs.project(A0, direction, nsites))
            psi.replacesites(vec, site1, direction, cutoff=cutoff, maxdim=maxdim, mindim=mindim)
        Vs.movecenter(1 if direction else N)
        direction = not direction
        sweeps += 1
        D = psi.maxbonddim()
        cost = calculatecost()
        if sweeps >= minsweeps:
            difference = diff(cost, lastcost)
            if difference < tol and lastD == D:
                convergedsweeps += 1
            # vmps.jl:75 ``convergedsweeps == 0`` is a comparison, not an assignment:
            # the counter is never reset, which is restated here as is.
            if convergedsweeps >= numconverges:
                converged = True
            if sweeps >= maxsweeps and maxsweeps != 0:
                converged = True
        lastcost = cost
        lastD = D
        if history is not None:
            history.append((sweeps, complex(cost), D))
        if verbose:
            print("Sweep=%d, energy=%.12f, maxbonddim=%d" % (sweeps, np.real(cost), D))
    return psi


def vmps(*psis, **kw):
    """vmps.jl:95-105: psi0 = deepcopy(psis[1]) with its centre at site 1, Vs =
    ProjMPSSum([ProjMPS(psi_k, psi0)])."""
    psi0 = psis[0].copy()
    psi0.movecenter(1)
    center = kw.pop("center", 1)
    Vs = ProjMPSSum([ProjMPS([p, psi0], rank=1) for p in psis], center=center)
    return vmps_sweeps(psi0, Vs, **kw)
