"""Two-site DMRG driver -- restates /root/reference/src/algorithms/mps/dmrg.jl."""
import numpy as np
from .projmps import ProjMPS, ProjMPSSum
from .lanczos import eigsolve_lowest


def dmrg_sweeps(psi, Hs, nsites=2, krylovdim=3, kryloviter=2, minsweeps=1, maxsweeps=1000,
                tol=1e-10, tolgrad=1e-5, numconverges=4, verbose=False, cutoff=1e-12,
                maxdim=1000, mindim=1, optimal_order=False, history=None):
    """dmrg.jl:1-98.  ``optimal_order`` switches the matvec to the flop-optimal
    contraction order (same result up to rounding)."""
    cost = Hs.calculate()
    lastcost = cost
    D = psi.maxbonddim()
    lastD = D
    grad = 0.0
    direction = False
    converged = False
    convergedsweeps = convergedgrad = sweeps = 0
    N = len(psi)
    while not converged:
        for j in range(1, N + 2 - nsites):
            site = N + 1 - j if direction else j
            site1 = site + 1 - nsites if direction else site
            Hs.movecenter(site)
            A0 = psi[site1]
            for i in range(1, nsites):
                A0 = np.tensordot(A0, psi[site1 + i], axes=([A0.ndim - 1], [0]))
            if optimal_order and nsites == 2:
                heff = lambda x: Hs.product_optimal(x, direction)
            else:
                heff = lambda x: Hs.product(x, direction, nsites)
            eig, vec, _ = eigsolve_lowest(heff, A0, krylovdim=krylovdim, maxiter=kryloviter, tol=1e-14)
            cost = eig
            psi.replacesites(vec, site1, direction, True, cutoff=cutoff, maxdim=maxdim, mindim=mindim)
        Hs.movecenter(1 if direction else N)
        direction = not direction
        sweeps += 1
        D = psi.maxbonddim()

        def diff(x, y):
            return abs(x - y) if abs(x) < 1e-10 else abs((x - y) / x)
        if sweeps >= minsweeps:
            dd = diff(cost, lastcost)
            convergedsweeps = convergedsweeps + 1 if (dd < tol and lastD == D) else 0
            with np.errstate(divide='ignore', invalid='ignore'):
                g = abs(np.float64(dd - grad) / np.float64(dd + grad))
            convergedgrad = convergedgrad + 1 if (g < tolgrad and lastD == D) else 0
            if max(convergedsweeps, convergedgrad) >= numconverges:
                converged = True
            if sweeps >= maxsweeps and maxsweeps != 0:
                converged = True
        grad = abs(diff(cost, lastcost))
        lastcost = cost
        lastD = D
        if history is not None:
            history.append((sweeps, float(np.real(cost)), D))
        if verbose:
            print("Sweep=%d, energy=%.12f, maxbonddim=%d" % (sweeps, np.real(cost), D))
    return psi, cost


def dmrg(psi, *Hs, coeffs=None, **kw):
    """dmrg.jl:128-154 front-end: centre to site 1, build ProjMPS(psi,H,psi; rank=2)."""
    if psi.rank != 1:
        raise ValueError("Psi must be a GMPS of rank 1 (vector).")
    if len(Hs) == 0:
        raise ValueError("You must provide atleast one MPS/MPO for the Hamiltonian.")
    psi.movecenter(1)
    coeffs = [1.0] * len(Hs) if coeffs is None else coeffs
    projs = []
    for H, c in zip(Hs, coeffs):
        if len(H) != len(psi) or H.dim != psi.dim:
            raise ValueError("GMPS must share the same properties.")
        if H.rank == 2:
            projs.append(ProjMPS([psi, H, psi], rank=2, coeff=c))
        elif H.rank == 1:
            projs.append(ProjMPS([H, psi], rank=2, squared=True, coeff=c))
        else:
            raise ValueError("Hamiltonian must be composed of MPOs (rank 2) or MPSs (rank 1).")
    Hsum = ProjMPSSum(projs)
    Hsum.movecenter(1)
    return dmrg_sweeps(psi, Hsum, **kw)
