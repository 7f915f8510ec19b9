"""Quantum-jump Monte Carlo -- restates /root/reference/src/algorithms/mps/qjmc.jl:9-220.

The reference draws from Julia's unseeded global RNG (qjmc.jl:64,97,99); here
the uniforms come from a caller-supplied generator/array so that a trajectory
is reproducible: per step u0 (drawn and unused in classical mode, :64),
u1 (jump test, :97) and, on a jump, u2 (channel pick, :99)."""
import numpy as np
from .oplist import OpList
from .gatelist import trotterize, applygates
from .gmps import applyop
from .projmps import ProjMPS


def qjmc_gates(st, H, jumpops, dt, **kw):
    """qjmc.jl:9-26: H_eff = -iH - 1/2 sum c^2 L^dag L, Trotterised."""
    escapeops = OpList(jumpops.length)
    for i in range(len(jumpops.ops)):
        names = [st.opprod([st.dag(o), o]) for o in jumpops.ops[i]]
        escapeops.add(names, jumpops.sites[i], jumpops.coeffs[i] ** 2)
    Heff = (-1j * H) + (-0.5 * escapeops)
    return Heff, trotterize(st, Heff, dt, **kw)


def qjmc_emission_rates(st, psi, jumpops):
    """qjmc.jl:170-220: rates_k = |c_k^2 <psi| L_k^dag L_k |psi>| via overlap blocks."""
    proj = ProjMPS([psi, psi])
    rates = np.zeros(len(jumpops.ops), dtype=np.complex128)
    for i in range(1, len(psi) + 1):
        proj.movecenter(i)
        for idx in jumpops.siteindexs(i):
            sites, ops = jumpops.sites[idx], jumpops.ops[idx]
            coeff = jumpops.coeffs[idx] ** 2
            rng = sites[-1] - sites[0] + 1
            prod = proj.block(sites[0] - 1)
            right = proj.block(sites[-1] + 1)
            k = 0
            for j in range(1, rng + 1):
                site = i + j - 1
                A = psi[site]
                if site in sites:
                    O = st.op(ops[k])
                    k += 1
                else:
                    O = st.op("id")
                # <A| O^dag O |A> sandwiched between the blocks (qjmc.jl:206-209)
                OA = np.einsum('st,btc->bsc', O, A)
                # prod(c,d) = sum_{a,b,s} prod(a,b) conj(OA)(a,s,c) OA(b,s,d): two pairwise contractions like the reference's contract()
                prod = np.tensordot(np.conj(OA), np.tensordot(prod, OA, axes=([1], [0])), axes=([0, 1], [0, 1]))
            rates[idx] = coeff * np.einsum('ab,ab->', prod, right)
    return np.abs(rates)


def qjmc_simulation(st, psi, H, jumpops, tmax, dt, observers=(), uniforms=None, save=0, cutoff=1e-12,
                    mindim=1, maxdim=0, classical=True, verbose=False, **kw):
    """qjmc.jl:28-167: the default classical branch (:88-112) and the norm-based branch (classical=false, :65-87).
    ``uniforms`` is a callable returning the next U(0,1) sample."""
    if uniforms is None:
        g = np.random.default_rng(0)
        uniforms = g.random
    save = dt if save == 0 else save
    steps = int(np.ceil(round(tmax / dt, 5)))
    savesteps = int(np.ceil(round(save / dt)))
    jumps, jumptimes = [], []
    _, gates = qjmc_gates(st, H, jumpops, dt, **kw)
    time = 0.0
    for ob in observers:
        ob.measure(time, psi, jumps, jumptimes)
    for i in range(1, steps + 1):
        applygates(psi, gates, mindim=mindim, maxdim=maxdim, cutoff=cutoff)
        r0 = uniforms()                 # qjmc.jl:64 (unused in classical mode)
        if not classical:               # qjmc.jl:65-87: jump when the decayed norm^2 falls below r
            prob = np.real(psi.norm() ** 2)
            psi.normalize()
            if r0 > prob:
                rates = qjmc_emission_rates(st, psi, jumpops)
                r = uniforms()
                cs = np.cumsum(rates) / np.sum(rates)
                idx = int(np.nonzero(r < cs)[0][0])
                psi.movecenter(1)
                applyop(st, psi, jumpops.ops[idx], jumpops.sites[idx])
                psi.movecenter(len(psi))
                psi.movecenter(1, cutoff=cutoff, maxdim=maxdim, mindim=mindim)
                psi.normalize()
                jumps.append(idx + 1)
                jumptimes.append(time + dt)
            time += dt
            if i % savesteps == 0:
                for ob in observers:
                    ob.measure(time, psi, jumps, jumptimes)
            continue
        psi.normalize()
        rates = qjmc_emission_rates(st, psi, jumpops)
        er = np.sum(rates)
        prob = np.exp(-er * dt)
        if uniforms() > prob:
            r = uniforms()
            cs = np.cumsum(rates) / er
            idx = int(np.nonzero(r < cs)[0][0])
            psi.movecenter(1)
            applyop(st, psi, jumpops.ops[idx], jumpops.sites[idx])
            psi.movecenter(len(psi))
            psi.movecenter(1, cutoff=cutoff, maxdim=maxdim, mindim=mindim)
            psi.normalize()
            jumps.append(idx + 1)       # 1-based channel index like the reference
            jumptimes.append(time + dt)
        time += dt
        if i % savesteps == 0:
            if verbose:
                print("time=%.5f, jumps=%d, maxbonddim=%d" % (time, len(jumps), psi.maxbonddim()))
            for ob in observers:
                ob.measure(time, psi, jumps, jumptimes)
    return jumps, jumptimes


class QJMCOperators:
    """qjmc.jl:238-264."""

    def __init__(self, oplist, st):
        self.times, self.measurements, self.oplist, self.st = [], [], oplist, st

    def measure(self, time, psi, jumps, jumptimes):
        from .gmps import inner
        self.times.append(time)
        self.measurements.append(inner(self.st, psi, self.oplist, psi))
