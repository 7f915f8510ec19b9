"""Front-end of tebd(st, psi, H, dt, tmax, save, observers; kwargs...) (/root/reference/src/algorithms/mps/tebd.jl:3-100) for
operator lists given as ``terms`` = [(ops, sites, coeff)] with d x d operator matrices: Trotterisation on the host
(gatelist.jl:75-121 + oplist.jl:134-183; 2^2 x 2^2 matrix exponentials), everything that touches the MPS on the device
(tn_apply_gates, tn_mps_norm, tn_mps_normalize, tn_inner_oplist).  Interaction range <= 2 sites (the C ABI applies one- and
two-site gates); the projector branch (tebd.jl:22-41,67-73) goes through tnb200.vmps."""
import numpy as np
import scipy.linalg as sla

from . import _lib
from .api import GateList, applygates, inner


def _sorted_term(ops, sites):
    order = np.argsort(np.asarray(sites))
    return [np.asarray(ops[j], dtype=np.complex128) for j in order], [int(sites[j]) for j in order]


def siterange(terms):
    """oplist.jl:106-113"""
    rng = 1
    for _, sites, _ in terms:
        rng = max(rng, max(sites) - min(sites) + 1)
    return rng


def trotterize(N, d, terms, dt, order=2, evol="imag"):
    """trotterize(st, ops, dt; order, evol): gatelist.jl:75-121.  Returns (rows of first sites, rows of gate tensors) with gates
    laid out (out1, in1[, out2, in2]).  Like the reference, imaginary time exponentiates +dt*h (callers pass -H), and a term that
    starts on the last site of a range-2 list still becomes a two-site-range operator truncated to the one site that exists
    (oplist.jl:138), i.e. a one-site gate."""
    rng = siterange(terms)
    if rng > 2:
        raise _lib.TNError("trotterize: interaction range > 2 sites is not supported by the device gate kernels")
    order = 1 if rng == 1 else order
    if order not in (1, 2):
        raise _lib.TNError("Only trotter order 1 and 2 are supported.")
    t = -1j * dt if evol == "real" else dt
    ident = np.eye(d, dtype=np.complex128)
    start = {}
    for ops, sites, coeff in terms:
        o, s = _sorted_term(ops, sites)
        span = min(rng, N - s[0] + 1)
        mats = [o[s.index(s[0] + j)] if (s[0] + j) in s else ident for j in range(span)]
        m = mats[0]
        for x in mats[1:]:
            m = np.kron(m, x)                         # rows (o1,o2), cols (i1,i2), first site most significant
        start[s[0]] = start.get(s[0], 0) + complex(coeff) * m
    rows_s, rows_g = [], []
    for i in range(1, rng + 1):
        time = t / 2 if (i < rng and order == 2) else t
        ss, gg = [], []
        site = i
        while site <= N:
            h = start.get(site)
            if h is not None:
                u = sla.expm(time * h)
                if u.shape[0] == d * d:
                    u = u.reshape(d, d, d, d).transpose(0, 2, 1, 3)      # (o1,o2,i1,i2) -> (o1,i1,o2,i2)
                ss.append(site)
                gg.append(np.ascontiguousarray(u))
            site += rng
        rows_s.append(ss)
        rows_g.append(gg)
    if order == 2:
        for i in range(1, rng):
            rows_s.append(rows_s[rng - i - 1])
            rows_g.append(rows_g[rng - i - 1])
    return rows_s, rows_g


def tebd(psi, terms, dt, tmax, save, observers=(), cutoff=1e-12, maxdim=0, mindim=1, evol="imag", order=2, norm=0.0, verbose=False):
    """tebd.jl:3-100 without projectors.  ``psi`` is a device GMPS, ``terms`` the Hamiltonian as passed to the reference
    (i.e. -H for imaginary time).  Observers are objects with measure(time, psi, norm, energy) and checkdone().
    Returns (psi, energy) like the oracle's restatement."""
    N, d = len(psi), psi.dim
    rs, rg = trotterize(N, d, terms, dt, order=order, evol=evol)
    gates = GateList(d, rs, rg, ctx=psi.ctx)
    nsteps = int(round(tmax / dt))
    save = dt if save < dt else save
    nsave = int(round(save / dt))
    normal = float(norm)
    energy = float(np.real(np.sum(inner(psi, terms, psi))))
    for ob in observers:
        ob.measure(0.0, psi, normal, energy)
    converged = False
    step = 0
    while not converged:
        applygates(psi, gates, cutoff=cutoff, maxdim=maxdim, mindim=mindim)
        normal += float(np.log(np.real(psi.norm())))
        psi.normalize()
        step += 1
        if step >= nsteps:
            converged = True
        if step % nsave == 0:
            if verbose:
                print("time=%.4f, energy=%.12f, maxbonddim=%d" % (step * dt, energy, psi.maxbonddim()))
            energy = float(np.real(np.sum(inner(psi, terms, psi))))
            for ob in observers:
                ob.measure(step * dt, psi, normal, energy)
                converged = converged or ob.checkdone()
    return psi, energy
