"""Front-end of tebd(st, psi, H, dt, tmax, save, observers; kwargs...) (/root/reference/src/algorithms/mps/tebd.jl:3-100) for
operator lists given as ``terms`` = [(ops, sites, coeff)] with d x d operator matrices: Trotterisation on the host
(gatelist.jl:75-121 + oplist.jl:134-183; 2^2 x 2^2 matrix exponentials), everything that touches the MPS on the device
(tn_apply_gates, tn_mps_norm, tn_mps_normalize, tn_inner_oplist).  Interaction range <= 2 sites (the C ABI applies one- and
two-site gates); the projector branch (tebd.jl:22-41,67-73) goes through tnb200.vmps."""
import ctypes as C

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


class MPSProjector:
    """projector.jl:1-50: |right><left| / constant with constant = <left|right> unless given (device MPSs)."""

    rank = 2

    def __init__(self, left, right=None, constant=0.0):
        right = left if right is None else right
        if left.dim != right.dim or len(left) != len(right):
            raise _lib.TNError("The MPS must share the same physical dimension and length.")
        self.dim, self.leftMPS, self.rightMPS = left.dim, left, right
        self.constant = left.overlap(right) if constant == 0.0 else constant

    def apply(self, psi):                       # projector.jl:41-47
        return self.rightMPS.scale(self.leftMPS.overlap(psi) / self.constant)


def project_out(psi, projs, cutoff, maxdim):
    """tebd.jl:68-73 (and :36-41): psi <- vmps(psi, -1*(P_1*psi), ...); MPO projectors go through applyMPO, MPS projectors through
    their overlap with psi; the variational sum runs on the device (tn_vmps_sweep)."""
    from .api import applyMPO, vmps
    psis = [psi]
    for P in projs:
        Ppsi = P.apply(psi) if isinstance(P, MPSProjector) else applyMPO(P, psi)
        psis.append(Ppsi.scale(-1))
    return vmps(*psis, cutoff=cutoff, maxdim=maxdim)


def tebd(psi, terms, dt, tmax, save, observers=(), projectors=(), cutoff=1e-12, maxdim=0, mindim=1, evol="imag", order=2, norm=0.0,
         verbose=False, projection_every=10, variational_cutoff=None):
    """tebd.jl:3-100, including the projector branch (:22-41, :67-73; the returned psi is then a new device MPS).  ``psi`` is a device GMPS, ``terms`` the Hamiltonian as passed to the reference
    (i.e. -H for imaginary time).  Observers are objects with measure(time, psi, norm, energy) and checkdone().
    Returns (psi, energy) like the oracle's restatement."""
    N, d = len(psi), psi.dim
    rs, rg = trotterize(N, d, terms, dt, order=order, evol=evol)
    gates = GateList(d, rs, rg, ctx=psi.ctx)
    variational_cutoff = cutoff if variational_cutoff is None else variational_cutoff
    projs = []
    for proj in projectors:                                   # tebd.jl:26-35
        if isinstance(proj, MPSProjector) or proj.rank == 2:
            projs.append(proj)
        elif proj.rank == 1:
            projs.append(MPSProjector(proj, proj))
        else:
            raise _lib.TNError("Only MPS, MPO and MPSProjectors are supported as projectors.")
    if projs:                                                 # tebd.jl:36-41
        psi = project_out(psi, projs, variational_cutoff, maxdim)
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
        if projs and (step + 1) % projection_every == 0:      # tebd.jl:67-73
            psi = project_out(psi, projs, variational_cutoff, maxdim)
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


# ---------------------------------------------------------------------------------------------
# Observers (host-side bookkeeping, same convergence rules as the reference)
# ---------------------------------------------------------------------------------------------
class TEBDNorm:
    """tebd.jl:124-152: records the accumulated log-norm; converged when its slope stops changing."""

    def __init__(self, tol=1e-10):
        self.times, self.measurements, self.tol = [], [], tol

    def measure(self, time, psi, norm, energy):
        self.times.append(time)
        self.measurements.append(norm)

    def checkdone(self):
        if len(self.times) < 3:
            return False
        m, t = self.measurements, self.times
        e1 = (m[-1] - m[-2]) / (t[-1] - t[-2])
        e2 = (m[-2] - m[-3]) / (t[-2] - t[-3])
        if abs(e2 + e1) < 1e-8:
            return abs(e2 - e1) < self.tol
        return abs((e2 - e1) / (0.5 * (e2 + e1))) < self.tol


class TEBDEnergy:
    """tebd.jl:161-187: records the energy; converged on its relative change."""

    def __init__(self, tol=1e-10):
        self.times, self.measurements, self.tol = [], [], tol

    def measure(self, time, psi, norm, energy):
        self.times.append(time)
        self.measurements.append(energy)

    def checkdone(self):
        if len(self.times) < 3:
            return False
        e1, e2 = self.measurements[-1], self.measurements[-2]
        if 0.5 * abs(e2 + e1) < 1e-8:
            return abs(e2 - e1) < self.tol
        return abs(2 * (e2 - e1) / (e2 + e1)) < self.tol


class TEBDOperators:
    """tebd.jl:195-219: inner(st, psi, oplist, psi) at every save point (device-side tn_inner_oplist)."""

    def __init__(self, terms):
        self.times, self.measurements, self.terms = [], [], terms

    def measure(self, time, psi, norm, energy):
        self.times.append(time)
        self.measurements.append(inner(psi, self.terms, psi))

    def checkdone(self):
        return False


class QJMCOperators:
    """qjmc.jl:238-264"""

    def __init__(self, terms):
        self.times, self.measurements, self.terms = [], [], terms

    def measure(self, time, psi, jumps, jumptimes):
        self.times.append(time)
        self.measurements.append(inner(psi, self.terms, psi))


class QJMCActivity:
    """qjmc.jl:268-289 (without the HDF5 dump)"""

    def __init__(self):
        self.time, self.jumps = 0.0, 0

    def measure(self, time, psi, jumps, jumptimes):
        self.time, self.jumps = time, len(jumps)


class QJMCEntropy:
    """qjmc.jl:297-310: entanglement entropy of every bond at every save point"""

    def __init__(self):
        self.times, self.measurements = [], []

    def measure(self, time, psi, jumps, jumptimes):
        self.times.append(time)
        self.measurements.append([psi.entropy(i) for i in range(1, len(psi))])


def qjmc(psi, gates, jump_sites, jump_ops, jump_coeffs, tmax, dt, observers=(), save=0, uniforms=None, rng=None, cutoff=1e-12, maxdim=0,
         mindim=1, classical=True):
    """qjmc_simulation(st, psi, H, jumpops, tmax, dt, observers; save, ...) (qjmc.jl:28-167) with observers: the trajectory runs on the
    device in chunks of ``savesteps`` steps (tn_qjmc_run), the observers are called between the chunks with psi still on the device.
    ``gates`` is a device GateList of the Trotterised effective Hamiltonian (qjmc_gates, qjmc.jl:9-26).  Random numbers: 3 uniforms per
    step, from ``uniforms`` (array of 3 * steps) or drawn here from ``rng`` (numpy Generator; default seed 0), so that a chunked run
    and a single call see the same stream.  Returns (jumps, jumptimes)."""
    from .api import qjmc_simulation
    save = dt if save == 0 else save
    steps = int(np.ceil(round(tmax / dt, 5)))
    savesteps = max(1, int(np.ceil(round(save / dt))))
    if uniforms is None:
        rng = np.random.default_rng(0) if rng is None else rng
        uniforms = rng.random(3 * steps)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    if uniforms.size < 3 * steps:
        raise _lib.TNError("need 3 uniforms per step")
    jumps, jumptimes = [], []
    for ob in observers:
        ob.measure(0.0, psi, jumps, jumptimes)
    done = 0
    while done < steps:
        n = min(savesteps, steps - done)
        j, t, _ = qjmc_simulation(psi, gates, jump_sites, jump_ops, jump_coeffs, n, dt, uniforms=uniforms[3 * done:3 * (done + n)],
                                  cutoff=cutoff, maxdim=maxdim, mindim=mindim, classical=classical)
        jumps.extend(j)
        jumptimes.extend(done * dt + x for x in t)
        done += n
        if done % savesteps == 0:
            for ob in observers:
                ob.measure(done * dt, psi, jumps, jumptimes)
    return jumps, jumptimes


# ---------------------------------------------------------------------------------------------
# Infinite TEBD (Vidal form, two-site cell): algorithms/mps/itebd.jl
# ---------------------------------------------------------------------------------------------
class IGMPS:
    """Device-resident iGMPS of rank 1 (structures/mps/igmps.jl:8-17): ``tensors[i]`` (D_{i-1}, d, D_i), ``singulars[i]`` to the
    left of ``tensors[i]``, periodic two-site cell."""

    def __init__(self, dim, tensors, singulars=None, ctx=None):
        from .api import Context, _f
        self.ctx = ctx or Context.default()
        self.lib = self.ctx.lib
        ts = [_f(t) for t in tensors]
        L = len(ts)
        if singulars is None:
            singulars = [np.ones(t.shape[0]) for t in ts]
        ss = [np.ascontiguousarray(s, dtype=np.float64) for s in singulars]
        if any(t.ndim != 3 for t in ts) or any(s.shape != (t.shape[0],) for s, t in zip(ss, ts)):
            raise _lib.TNError("iGMPS: every cell tensor needs 3 indices and one singular value per left-bond state")
        dims = np.array([t.shape for t in ts], dtype=np.int64).reshape(L, 3)
        tp = (C.c_void_p * L)(*[t.ctypes.data for t in ts])
        sp = (C.c_void_p * L)(*[s.ctypes.data for s in ss])
        h = C.c_void_p()
        _lib.check(self.lib.tn_imps_create(self.ctx.h, int(dim), L, dims.ctypes.data_as(C.POINTER(C.c_int64)), tp, sp, C.byref(h)))
        self.h, self.dim, self._L = h, int(dim), L

    @classmethod
    def product(cls, length, A, ctx=None):
        """iMPS(length, A): igmps.jl:53-59"""
        A = np.asarray(A, dtype=np.complex128)
        return cls(A.shape[0], [A.reshape(1, -1, 1) for _ in range(length)], ctx=ctx)

    def __len__(self):
        return self._L

    def __del__(self):
        try:
            if self.h:
                self.lib.tn_imps_free(self.h)
                self.h = None
        except Exception:
            pass

    def dims(self):
        d = np.zeros((self._L, 3), dtype=np.int64)
        _lib.check(self.lib.tn_imps_dims(self.h, d.ctypes.data_as(C.POINTER(C.c_int64))))
        return d

    def maxbonddim(self):                        # igmps.jl:39
        return int(self.dims()[:, 0].max())

    def site(self, i):
        """(tensors[i], singulars[i], norms[i]) downloaded from the device (1-based)"""
        shape = tuple(int(x) for x in self.dims()[i - 1])
        t = np.zeros(shape, dtype=np.complex128, order='F')
        s = np.zeros(shape[0])
        n = C.c_double()
        _lib.check(self.lib.tn_imps_download(self.h, int(i), t.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)))
        return t, s, n.value

    def apply_gate(self, gate, nsteps=1, cutoff=1e-12, maxdim=0, mindim=1):
        """nsteps passes of _itebd_apply_gates_mps! (itebd.jl:71-119) with the two-site gate (o1,i1,o2,i2)"""
        from .api import _f, Trunc
        g = _f(gate)
        if g.shape != (self.dim,) * 4:
            raise _lib.TNError("iTEBD: the gate must have shape (d, d, d, d) for a two-site cell")
        _lib.check(self.lib.tn_itebd_apply_gate(self.h, g.ctypes.data_as(C.c_void_p), int(nsteps), Trunc(cutoff, maxdim, mindim)))

    def bond_energy(self, h2):
        """<h> per bond from the Vidal form (exact for a canonical cell): mean over both bonds of <Theta|h|Theta>/<Theta|Theta>,
        Theta = S_a G_a S_b G_b S_a.  Host-side on the downloaded cell (chi^3 d^2 work)."""
        (ta, sa, _), (tb, sb, _) = self.site(1), self.site(2)
        out = []
        for (A, Sa), (B, Sb) in (((ta, sa), (tb, sb)), ((tb, sb), (ta, sa))):
            th = np.einsum('l,lsm,m,mtr,r->lstr', Sa, A, Sb, B, Sa)
            out.append(np.vdot(th, np.einsum('sutv,lutr->lstr', h2, th)) / np.vdot(th, th))
        return complex(np.mean(out))


def itebd(psi, terms, dt, tmax, observers=(), cutoff=1e-12, maxdim=0, mindim=1, evol="imag", save_time=None):
    """itebd(st, psi, H, dt, tmax, observers; kwargs...) (itebd.jl:3-68) for a two-site cell: the gate is exp(+-dt * h) with h the sum of
    the terms that start on site 1 of the cell (sitetensor(H, st, 1), itebd.jl:19-20; callers pass -H for imaginary time).  Observers are
    called as measure(time, psi, energy) at multiples of ``save_time`` with energy = bond_energy of h (the reference measures through
    infinite environments, igmps.jl:155-340, which stay out of scope)."""
    d, L = psi.dim, len(psi)
    if L != 2:
        raise _lib.TNError("iTEBD on the device supports a two-site unit cell")
    ident = np.eye(d, dtype=np.complex128)
    h = np.zeros((d * d, d * d), dtype=np.complex128)
    for ops, sites, coeff in terms:
        o, s = _sorted_term(ops, sites)
        if s[0] != 1:
            continue
        if s[-1] > 2:
            raise _lib.TNError("iTEBD: terms must fit in the two-site cell")
        mats = [o[s.index(q)] if q in s else ident for q in (1, 2)]
        h = h + complex(coeff) * np.kron(mats[0], mats[1])
    u = sla.expm((-1j if evol == "real" else 1) * dt * h)
    gate = np.ascontiguousarray(u.reshape(d, d, d, d).transpose(0, 2, 1, 3))
    h2 = h.reshape(d, d, d, d).transpose(0, 2, 1, 3)
    nsteps = int(round(tmax / dt))
    every = max(1, int(round((save_time or dt) / dt)))
    energy = None
    for ob in observers:
        ob.measure(0.0, psi, psi.bond_energy(h2))
    done = 0
    while done < nsteps:
        n = min(every, nsteps - done)
        psi.apply_gate(gate, n, cutoff=cutoff, maxdim=maxdim, mindim=mindim)
        done += n
        if observers:
            energy = psi.bond_energy(h2)
            for ob in observers:
                ob.measure(done * dt, psi, energy)
            if any(getattr(ob, "checkdone", lambda: False)() for ob in observers):
                break
    return psi


# ---------------------------------------------------------------------------------------------
# Thermal-state measurement of examples/thermal.jl: energy = trace(H, adjoint(U), U) / trace(adjoint(U), U) (mpo.jl:229-252)
# ---------------------------------------------------------------------------------------------
def doubled_state_tensors(U_tensors):
    """An MPO tensor (w_l, s, t, w_r) is, byte for byte, an MPS tensor (w_l, (s,t), w_r) of physical dimension d^2 (s fastest)."""
    return [np.reshape(np.asfortranarray(t), (t.shape[0], t.shape[1] * t.shape[2], t.shape[3]), order='F') for t in U_tensors]


def doubled_operator_tensors(H_tensors):
    """MPO of 1 (x) H^T on the doubled sites: H'(w, (s,t), (s',t'), w') = delta(s,s') M(w, t', t, w'), so that
    tr(H U^dag U) = <<U| H' |U>> for the vectorised U (H then acts on U's second physical index)."""
    out = []
    for M in H_tensors:
        M = np.asarray(M, dtype=np.complex128)
        w, d, _, w2 = M.shape
        X = np.einsum('su,wvtx->wstuvx', np.eye(d), M)
        out.append(np.reshape(X, (w, d * d, d * d, w2), order='F'))
    return out


def thermal_energy(U, H_tensors):
    """trace(H, adjoint(U), U) / trace(adjoint(U), U) for a device-resident MPO ``U`` (the state exp(-beta H / 2) of
    examples/thermal.jl) and the host tensors of the Hamiltonian MPO, evaluated as an ordinary expectation value on the doubled
    sites with the environment kernels (tn_env_create / tn_env_calculate at physical dimension d^2).  The energy is invariant
    under rescaling U, so the unnormalised evolved MPO can be passed as is."""
    from .api import GMPS, ProjMPS
    if U.rank != 2:
        raise _lib.TNError("thermal_energy: U must be an MPO (rank 2)")
    d2 = U.dim * U.dim
    Ud = GMPS(1, d2, doubled_state_tensors(U.tensors), 0, U.ctx)
    Hd = GMPS(2, d2, doubled_operator_tensors(H_tensors), 0, U.ctx)
    num = ProjMPS(Ud, Hd, Ud, center=1).calculate()
    den = ProjMPS(Ud, None, Ud, center=1).calculate()
    return num / den
