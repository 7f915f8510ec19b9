"""Host-side mirror of the reference's hot-path API on top of the C ABI.

Reference interface mirrored (all under /root/reference/src):
  GMPS                structures/mps/gmps.jl:8-13 (+ abstractmps.jl accessors)
  ProjMPS             structures/mps/projmps.jl, abstractprojmps.jl
  GateList/applygates structures/mps/gatelist.jl
  svd                 tensors.jl:168-227
  dmrg / tebd / qjmc_simulation   algorithms/mps/{dmrg,tebd,qjmc}.jl
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, tn_cplx, tn_trunc_t, tn_lanczos_t, tn_idx2_t


def _f(a):
    """complex128, Fortran (Julia) memory order."""
    return np.asfortranarray(a, dtype=np.complex128)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def Trunc(cutoff=0.0, maxdim=0, mindim=1):
    """kwargs of svd(): tensors.jl:170-172."""
    return tn_trunc_t(float(cutoff), int(maxdim), int(mindim))


class Context:
    """One GPU + one CUDA stream (tn_ctx)."""
    _default = None

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.tn_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.device = device

    @classmethod
    def default(cls):
        if cls._default is None:
            cls._default = cls(0)
        return cls._default

    def sync(self):
        check(self.lib.tn_sync(self.h))

    def stream(self):
        s = C.c_void_p()
        check(self.lib.tn_ctx_stream(self.h, C.byref(s)))
        return s.value or 0

    def counters(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(self.lib.tn_counters(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(launches=a.value, matvecs=b.value, svds=c.value)

    def close(self):
        if self.h:
            self.lib.tn_ctx_destroy(self.h)
            self.h = None


class GMPS:
    """Device-resident generalised MPS (rank 1) / MPO (rank 2): gmps.jl:8-13."""

    def __init__(self, rank, dim, tensors, center=0, ctx=None):
        self.ctx = ctx or Context.default()
        self.lib = self.ctx.lib
        ts = [_f(t) for t in tensors]
        N = len(ts)
        if rank not in (1, 2) or any(t.ndim != rank + 2 for t in ts):
            raise _lib.TNError("GMPS rank must be 1 (MPS) or 2 (MPO) and every site tensor must have rank+2 indices")
        dims = np.array([t.shape for t in ts], dtype=np.int64).reshape(N, rank + 2)
        ptrs = (C.c_void_p * N)(*[t.ctypes.data for t in ts])
        h = C.c_void_p()
        check(self.lib.tn_mps_upload(self.ctx.h, rank, dim, N, dims.ctypes.data_as(C.POINTER(C.c_int64)), ptrs, int(center), C.byref(h)))
        self.h = h
        self.rank, self.dim, self._N = rank, dim, N

    @classmethod
    def from_host(cls, psi, ctx=None):
        """From any object with .rank/.dim/.tensors/.center (e.g. the oracle's GMPS)."""
        return cls(psi.rank, psi.dim, psi.tensors, psi.center, ctx)

    def __len__(self):
        return self._N

    def __del__(self):
        try:
            if self.h:
                self.lib.tn_mps_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def center(self):
        c = C.c_int32()
        check(self.lib.tn_mps_info(self.h, None, None, None, C.byref(c)))
        return c.value

    @center.setter
    def center(self, v):
        check(self.lib.tn_mps_set_center(self.h, int(v)))

    def dims(self):
        d = np.zeros((self._N, self.rank + 2), dtype=np.int64)
        check(self.lib.tn_mps_dims(self.h, d.ctypes.data_as(C.POINTER(C.c_int64))))
        return d

    def __getitem__(self, i):
        """psi[i] (1-based): downloads the site tensor."""
        shape = tuple(int(x) for x in self.dims()[i - 1])
        out = np.zeros(shape, dtype=np.complex128, order='F')
        check(self.lib.tn_mps_download_site(self.h, int(i), _ptr(out)))
        return out

    def __setitem__(self, i, x):
        x = _f(x)
        d = np.array(x.shape, dtype=np.int64)
        check(self.lib.tn_mps_upload_site(self.h, int(i), d.ctypes.data_as(C.POINTER(C.c_int64)), _ptr(x)))

    @property
    def tensors(self):
        return [self[i] for i in range(1, self._N + 1)]

    @classmethod
    def _wrap(cls, like, h):
        out = cls.__new__(cls)
        out.ctx, out.lib, out.h, out.rank, out.dim, out._N = like.ctx, like.lib, h, like.rank, like.dim, like._N
        return out

    def copy(self):                               # deepcopy(psi): abstractmps.jl:95-96 (device to device)
        h = C.c_void_p()
        check(self.lib.tn_mps_copy(self.h, C.byref(h)))
        return GMPS._wrap(self, h)

    def scale(self, a):                           # psi * a: abstractmps.jl:99-109 (a scaled copy, like the reference)
        out = self.copy()
        a = complex(a)
        check(self.lib.tn_mps_scale(out.h, tn_cplx(a.real, a.imag)))
        return out

    def overlap(self, psi):
        """inner(self, psi) = <self|psi> (mpo.jl:181-217 for two MPS) through the overlap environment."""
        return ProjMPS(self, None, psi, center=1).calculate()

    def bonddim(self, site):                      # abstractmps.jl:62-65
        if site < 1 or site > self._N:
            return None
        return int(self.dims()[site][0])

    def maxbonddim(self):                         # abstractmps.jl:71-77
        v = C.c_int64()
        check(self.lib.tn_mps_maxbonddim(self.h, C.byref(v)))
        return v.value

    def norm(self):                               # gmps.jl:29-37
        v = tn_cplx()
        check(self.lib.tn_mps_norm(self.h, C.byref(v)))
        return complex(v.re, v.im)

    def normalize(self):                          # gmps.jl:46-51
        check(self.lib.tn_mps_normalize(self.h))

    def movecenter(self, idx, cutoff=0.0, maxdim=0, mindim=1):   # gmps.jl:90-112
        check(self.lib.tn_mps_movecenter(self.h, int(idx), Trunc(cutoff, maxdim, mindim)))

    def _span_size(self, site, nsites):
        """number of elements of the tensor spanning sites site .. site+nsites-1 (host buffers are checked before they cross the ABI)"""
        if not (1 <= site and site + nsites - 1 <= self._N):
            raise _lib.TNError("site out of range")
        dims = self.dims()
        return int(dims[site - 1][0]) * (self.dim ** self.rank) ** nsites * int(dims[site + nsites - 2][-1])

    def replacesites(self, A, site, direction=False, normalize=False, cutoff=0.0, maxdim=0, mindim=1):   # gmps.jl:199-267
        A = _f(A)
        if A.size != self._span_size(int(site), 2):
            raise _lib.TNError("replacesites: A does not have the size of the two-site tensor at this site")
        check(self.lib.tn_mps_replacesites(self.h, _ptr(A), int(site), int(bool(direction)), int(bool(normalize)),
                                           Trunc(cutoff, maxdim, mindim)))

    def applyop(self, site, op):                  # mps.jl:141-152
        op = _f(op)
        if op.shape != (self.dim, self.dim):
            raise _lib.TNError("applyop: the operator must be d x d")
        check(self.lib.tn_mps_applyop(self.h, int(site), _ptr(op)))

    def spectrum(self, site):                     # singular values across bond (site, site+1); gmps.jl:184-189
        cap = int(self.dims().max()) * self.dim ** self.rank + 8
        out = np.zeros(cap)
        k = C.c_int64()
        check(self.lib.tn_mps_bond_spectrum(self.h, int(site), out.ctypes.data_as(C.POINTER(C.c_double)), cap, C.byref(k)))
        return out[:k.value].copy()

    def entropy(self, site):                      # gmps.jl:184-189
        s2 = self.spectrum(site) ** 2
        return float(-np.sum(s2 * np.log(s2)))

    def expect(self, ops, sites):
        """<psi|O_k|psi> for single-site operators O_k (d x d) at 1-based sites."""
        ops = np.ascontiguousarray(np.stack([_f(o).T for o in ops]))   # each d x d block column-major
        st = np.asarray(sites, dtype=np.int32)
        out = np.zeros(len(st), dtype=np.complex128)
        check(self.lib.tn_expect_local(self.h, len(st), st.ctypes.data_as(C.POINTER(C.c_int32)), _ptr(ops), _ptr(out)))
        return out


def applyMPO(O, psi, cutoff=0.0, maxdim=0, mindim=1):
    """applyMPO(O, psi; kwargs...) = O * psi (mpo.jl:105-143): a new device MPS with its centre at site 1."""
    if O.rank != 2 or psi.rank != 1:
        raise _lib.TNError("Unallowed combinations of MPS ranks.")
    h = C.c_void_p()
    check(psi.lib.tn_mpo_apply(O.h, psi.h, Trunc(cutoff, maxdim, mindim), C.byref(h)))
    return GMPS._wrap(psi, h)


class ProjMPS:
    """ProjMPS(bra, [mpo,] ket; rank, squared, coeff, center) environment cache: projmps.jl:1-42.
    ``squared=True`` (mpo must be None) is the rank-1 projector penalty ProjMPS(V, psi; rank=2, squared=true)
    that dmrg builds for MPS arguments (dmrg.jl:144-145)."""

    def __init__(self, bra, mpo, ket, coeff=1.0, center=1, squared=False):
        self.ctx = ket.ctx
        self.lib = ket.lib
        self.bra, self.mpo, self.ket = bra, mpo, ket
        self.squared = bool(squared)
        h = C.c_void_p()
        co = complex(coeff)
        if self.squared:
            if mpo is not None:
                raise _lib.TNError("a squared projection has no MPO layer")
            check(self.lib.tn_env_create_squared(self.ctx.h, bra.h, ket.h, tn_cplx(co.real, co.imag), int(center), C.byref(h)))
        else:
            check(self.lib.tn_env_create(self.ctx.h, bra.h, mpo.h if mpo is not None else None, ket.h,
                                         tn_cplx(co.real, co.imag), int(center), C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h:
                self.lib.tn_env_free(self.h)
                self.h = None
        except Exception:
            pass

    def __len__(self):
        return len(self.ket)

    @property
    def center(self):
        c = C.c_int32()
        check(self.lib.tn_env_center(self.h, C.byref(c)))
        return c.value

    def buildleft(self, idx):
        check(self.lib.tn_env_buildleft(self.h, int(idx)))

    def buildright(self, idx):
        check(self.lib.tn_env_buildright(self.h, int(idx)))

    def movecenter(self, idx):
        check(self.lib.tn_env_movecenter(self.h, int(idx)))

    def block(self, idx):
        d = np.zeros(3, dtype=np.int64)
        check(self.lib.tn_env_block_dims(self.h, int(idx), d.ctypes.data_as(C.POINTER(C.c_int64))))
        out = np.zeros(tuple(int(x) for x in d), dtype=np.complex128, order='F')
        check(self.lib.tn_env_block_download(self.h, int(idx), _ptr(out)))
        return out

    def setblock(self, idx, x):
        """projV[idx] = x: abstractprojmps.jl:49-52."""
        x = _f(x)
        d = np.array(x.shape, dtype=np.int64)
        check(self.lib.tn_env_block_upload(self.h, int(idx), d.ctypes.data_as(C.POINTER(C.c_int64)), _ptr(x)))

    def product(self, A, direction=False, nsites=2, out=None):
        """H_eff * A for the two sites at the centre: projmps.jl:103-145 (rank-2 branch).
        ``out`` may be a caller-owned (e.g. pinned) complex128 Fortran-ordered buffer."""
        if nsites not in (1, 2):
            raise _lib.TNError("nsites must be 1 or 2")
        A = _f(A)
        if A.shape != self._local_shape(direction, nsites):
            raise _lib.TNError(f"product: A has shape {A.shape}, the sites at the centre need {self._local_shape(direction, nsites)}")
        if out is None:
            out = np.zeros(A.shape, dtype=np.complex128, order='F')
        elif not (out.dtype == np.complex128 and out.flags.f_contiguous and out.size == A.size):
            raise _lib.TNError("out must be a complex128 Fortran-contiguous array of A's size")
        if nsites == 2 and not self.squared:
            check(self.lib.tn_env_product(self.h, _ptr(A), int(bool(direction)), _ptr(out)))
        else:   # one-site rank-2 branch / squared branch (projmps.jl:135-143)
            check(self.lib.tn_env_product_n(self.h, _ptr(A), int(bool(direction)), int(nsites), _ptr(out)))
        return out

    def _local_shape(self, direction, nsites):
        site = self.center - nsites + 1 if direction else self.center
        if not (1 <= site and site + nsites - 1 <= len(self.ket)):
            raise _lib.TNError("the sites fall outside the chain")
        dims = self.ket.dims()
        return (int(dims[site - 1][0]),) + (self.ket.dim,) * nsites + (int(dims[site + nsites - 2][-1]),)

    def project(self, A=None, direction=False, nsites=2):
        """project(projV, A, direction, nsites): projmps.jl:153-185 (A is unused, as in the reference)."""
        out = np.zeros(self._local_shape(direction, nsites), dtype=np.complex128, order='F')
        check(self.lib.tn_env_project(self.h, int(bool(direction)), int(nsites), _ptr(out)))
        return out

    def calculate(self):
        v = tn_cplx()
        check(self.lib.tn_env_calculate(self.h, C.byref(v)))
        return complex(v.re, v.im)

    def eigsolve(self, A0, direction=False, krylovdim=3, maxiter=2, tol=1e-14):
        """KrylovKit eigsolve(Heff, A0, 1, :SR; ...) as called at dmrg.jl:51-53."""
        A0 = _f(A0)
        if A0.shape != self._local_shape(direction, 2):
            raise _lib.TNError(f"eigsolve: A0 has shape {A0.shape}, the sites at the centre need {self._local_shape(direction, 2)}")
        out = np.zeros(A0.shape, dtype=np.complex128, order='F')
        e, n = C.c_double(), C.c_int32()
        check(self.lib.tn_eigsolve(self.h, _ptr(A0), int(bool(direction)), tn_lanczos_t(krylovdim, maxiter, tol), C.byref(e), _ptr(out), C.byref(n)))
        return e.value, out, n.value


class ProjMPSSum:
    """ProjMPSSum(projVs; center): projmpssum.jl:1-108.  Members must share the ket MPS."""

    def __init__(self, projs, center=1):
        self.projs = list(projs)
        if not self.projs:
            raise _lib.TNError("ProjMPSSum needs at least one projection")
        self.ctx = self.projs[0].ctx
        self.lib = self.projs[0].lib
        self.ket = self.projs[0].ket
        arr = (C.c_void_p * len(self.projs))(*[p.h.value for p in self.projs])
        h = C.c_void_p()
        check(self.lib.tn_envsum_create(self.ctx.h, len(self.projs), arr, int(center), C.byref(h)))
        self.h = h
        self.center = int(center)

    def __del__(self):
        try:
            if self.h:
                self.lib.tn_envsum_free(self.h)
                self.h = None
        except Exception:
            pass

    def __len__(self):
        return len(self.ket)

    def movecenter(self, idx):
        check(self.lib.tn_envsum_movecenter(self.h, int(idx)))
        self.center = int(idx)

    def calculate(self):
        v = tn_cplx()
        check(self.lib.tn_envsum_calculate(self.h, C.byref(v)))
        return complex(v.re, v.im)

    def product(self, A, direction=False, nsites=2):
        A = _f(A)
        if A.shape != self.projs[0]._local_shape(direction, nsites):
            raise _lib.TNError(f"product: A has shape {A.shape}, the sites at the centre need {self.projs[0]._local_shape(direction, nsites)}")
        out = np.zeros(A.shape, dtype=np.complex128, order='F')
        check(self.lib.tn_envsum_product(self.h, _ptr(A), int(bool(direction)), int(nsites), _ptr(out)))
        return out

    def project(self, A=None, direction=False, nsites=2):
        site = self.center - nsites + 1 if direction else self.center
        dims = self.ket.dims()
        shape = (int(dims[site - 1][0]),) + (self.ket.dim,) * nsites + (int(dims[site + nsites - 2][-1]),)
        out = np.zeros(shape, dtype=np.complex128, order='F')
        check(self.lib.tn_envsum_project(self.h, int(bool(direction)), int(nsites), _ptr(out)))
        return out


class GateList:
    """Device copy of a GateList (gatelist.jl:8-13): rows of (site, gate tensor)."""

    def __init__(self, d, sites, gates, ctx=None):
        self.ctx = ctx or Context.default()
        self.lib = self.ctx.lib
        flat = [(int(s), _f(g)) for rs, rg in zip(sites, gates) for s, g in zip(rs, rg)]
        counts = np.array([len(r) for r in sites], dtype=np.int32)
        st = np.array([s for s, _ in flat], dtype=np.int32)
        ns = np.array([g.ndim // 2 for _, g in flat], dtype=np.int32)
        self._keep = [g for _, g in flat]
        ptrs = (C.c_void_p * len(flat))(*[g.ctypes.data for g in self._keep])
        h = C.c_void_p()
        check(self.lib.tn_gates_upload(self.ctx.h, int(d), len(counts), counts.ctypes.data_as(C.POINTER(C.c_int32)),
                                       st.ctypes.data_as(C.POINTER(C.c_int32)), ns.ctypes.data_as(C.POINTER(C.c_int32)), ptrs, C.byref(h)))
        self.h = h
        self.length = None

    @classmethod
    def from_host(cls, d, gl, ctx=None):
        """From any object with .sites / .gates rows (e.g. the oracle's trotterize output)."""
        return cls(d, gl.sites, gl.gates, ctx)

    def __del__(self):
        try:
            if self.h:
                self.lib.tn_gates_free(self.h)
                self.h = None
        except Exception:
            pass


def svd(x, idx, cutoff=0.0, maxdim=0, mindim=1, ctx=None, return_sweeps=False):
    """svd(x, idx; cutoff, maxdim, mindim): tensors.jl:168-227.  Returns U, S (dense
    k x k diagonal matrix) and V = V^H (k x dim(idx)); in U the new bond replaces idx."""
    ctx = ctx or Context.default()
    x = np.asarray(x, dtype=np.complex128)
    nd = x.ndim
    if idx == -1:
        idx = nd
    rest = [i for i in range(nd) if i != idx - 1]
    rest_shape = tuple(x.shape[i] for i in rest)
    m, n = int(np.prod(rest_shape)), x.shape[idx - 1]
    y = _f(np.reshape(np.transpose(x, rest + [idx - 1]), (m, n), order='F'))
    kmax = min(m, n)
    U = np.zeros((m, kmax), dtype=np.complex128, order='F')
    Vh = np.zeros(kmax * n, dtype=np.complex128)
    S = np.zeros(kmax)
    k, sw = C.c_int64(), C.c_int32()
    check(ctx.lib.tn_svd_trunc(ctx.h, _ptr(y), m, n, Trunc(cutoff, maxdim, mindim), _ptr(U), S.ctypes.data_as(C.POINTER(C.c_double)),
                               _ptr(Vh), C.byref(k), C.byref(sw)))
    k = k.value
    U = np.reshape(U.reshape(-1, order='F')[:m * k], (m, k), order='F')
    Vh = np.reshape(Vh[:k * n], (k, n), order='F')
    Ut = np.moveaxis(np.reshape(U, rest_shape + (k,), order='F'), -1, idx - 1)
    out = (np.ascontiguousarray(Ut), np.diag(S[:k]).astype(np.complex128), Vh)
    return out + (sw.value,) if return_sweeps else out


def svd_split(y, side, cutoff=0.0, maxdim=0, mindim=1, ctx=None, repeat=1):
    """The factorisation inside replacesites! / moveleft! / moveright! (gmps.jl:60-82, 218-256) of a matrix ``y`` (m x n):
    side = 1 returns (U, S, S V^H) with U orthonormal, side = 2 returns (U S, S, V^H) with V^H orthonormal.  Also returns the
    number of Jacobi sweeps and the device time in ms of the last of ``repeat`` runs (factorisation + gathers, no PCIe)."""
    ctx = ctx or Context.default()
    y = _f(np.asarray(y, dtype=np.complex128))
    m, n = y.shape
    kmax = min(m, n)
    U = np.zeros(m * kmax, dtype=np.complex128)
    Vh = np.zeros(kmax * n, dtype=np.complex128)
    S = np.zeros(kmax)
    k, sw, ms = C.c_int64(), C.c_int32(), C.c_double()
    check(ctx.lib.tn_svd_trunc_split(ctx.h, _ptr(y), m, n, Trunc(cutoff, maxdim, mindim), int(side), _ptr(U), S.ctypes.data_as(C.POINTER(C.c_double)),
                                     _ptr(Vh), C.byref(k), C.byref(sw), int(repeat), C.byref(ms)))
    k = k.value
    return (np.reshape(U[:m * k], (m, k), order='F'), S[:k].copy(), np.reshape(Vh[:k * n], (k, n), order='F'), sw.value, ms.value)


def svd_batched(mats, cutoff=0.0, maxdim=0, mindim=1, ctx=None):
    """Truncated SVDs (tensors.jl:168-227 on matrices, idx = last) of a stack ``mats[b]`` (B, m, n) of same-shape matrices in one
    batched factorisation (tn_svd_trunc_batched).  Returns a list of (U (m x k_b), S (k_b,), Vh (k_b x n))."""
    ctx = ctx or Context.default()
    mats = np.asarray(mats, dtype=np.complex128)
    B, m, n = mats.shape
    flat = np.ascontiguousarray(np.stack([np.reshape(x, -1, order='F') for x in mats]))
    kmax = min(m, n)
    U = np.zeros((B, m * kmax), dtype=np.complex128)
    Vh = np.zeros((B, kmax * n), dtype=np.complex128)
    S = np.zeros((B, kmax))
    k = np.zeros(B, dtype=np.int64)
    sw = C.c_int32()
    check(ctx.lib.tn_svd_trunc_batched(ctx.h, B, _ptr(flat), m, n, Trunc(cutoff, maxdim, mindim), _ptr(U), S.ctypes.data_as(C.POINTER(C.c_double)),
                                       _ptr(Vh), k.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(sw)))
    out = []
    for b in range(B):
        kb = int(k[b])
        out.append((np.reshape(U[b, :m * kb], (m, kb), order='F'), S[b, :kb].copy(), np.reshape(Vh[b, :kb * n], (kb, n), order='F')))
    return out


def contract_strided(M, N, K, A, am, ak, conjA, B, bk, bn, conjB, c_elems, cm, cn, alpha=1.0, ctx=None):
    """Debug/parity entry to the strided contraction kernel (tensors.jl:9-18 in GEMM form).
    A, B are flat complex128 buffers; index descriptors are (n0, s0, s1) triples."""
    ctx = ctx or Context.default()
    A = np.ascontiguousarray(A, dtype=np.complex128).reshape(-1)
    B = np.ascontiguousarray(B, dtype=np.complex128).reshape(-1)
    Cc = np.zeros(int(c_elems), dtype=np.complex128)
    al = complex(alpha)
    check(ctx.lib.tn_contract_strided(ctx.h, M, N, K, _ptr(A), A.size, tn_idx2_t(*am), tn_idx2_t(*ak), int(conjA),
                                      _ptr(B), B.size, tn_idx2_t(*bk), tn_idx2_t(*bn), int(conjB),
                                      _ptr(Cc), Cc.size, tn_idx2_t(*cm), tn_idx2_t(*cn), tn_cplx(al.real, al.imag)))
    return Cc


# ---------------------------------------------------------------------------------------------
# drivers
# ---------------------------------------------------------------------------------------------
def dmrg(psi, *Hs, coeffs=None, coeff=None, nsites=2, krylovdim=3, kryloviter=2, minsweeps=1, maxsweeps=1000, tol=1e-10,
         tolgrad=1e-5, numconverges=4, verbose=False, cutoff=1e-12, maxdim=1000, mindim=1, history=None):
    """dmrg(psi, Hs::GMPS...; kwargs...): algorithms/mps/dmrg.jl:1-154.  Every rank-2 argument is an MPO term
    ProjMPS(psi, H, psi; rank=2, coeff), every rank-1 argument a penalty ProjMPS(V, psi; rank=2, squared=true, coeff)
    (:136-147).  The sweep body (:35-63) runs on the device (tn_dmrg_sweep for one MPO with nsites=2, tn_dmrg_sweep_sum
    otherwise); the convergence logic (:66-89) is host-side."""
    if psi.rank != 1:
        raise _lib.TNError("Psi must be a GMPS of rank 1 (vector).")
    if nsites not in (1, 2):
        raise _lib.TNError("nsites must be 1 or 2")
    if len(Hs) == 0:
        raise _lib.TNError("You must provide atleast one MPS/MPO for the Hamiltonian.")
    if coeffs is None:
        coeffs = [1.0 if coeff is None else coeff] * len(Hs)
    if len(coeffs) != len(Hs):
        raise _lib.TNError("coeffs must have one entry per Hamiltonian term")
    for H in Hs:
        if len(H) != len(psi) or H.dim != psi.dim:
            raise _lib.TNError("GMPS must share the same properties.")
        if H.rank not in (1, 2):
            raise _lib.TNError("Hamiltonian must be composed of MPOs (rank 2) or MPSs (rank 1).")
    lib = psi.lib
    psi.movecenter(1)
    projs = [ProjMPS(psi, H, psi, coeff=c, center=1) if H.rank == 2 else ProjMPS(H, None, psi, coeff=c, center=1, squared=True)
             for H, c in zip(Hs, coeffs)]
    single = len(projs) == 1 and nsites == 2 and not projs[0].squared
    Hs = projs[0] if single else ProjMPSSum(projs, center=1)
    cost = Hs.calculate()
    lastcost = cost
    D = psi.maxbonddim()
    lastD = D
    grad = 0.0
    direction = False
    converged = False
    convergedsweeps = convergedgrad = sweeps = 0
    lz = tn_lanczos_t(krylovdim, kryloviter, 1e-14)
    tr = Trunc(cutoff, maxdim, mindim)
    while not converged:
        e, mb = C.c_double(), C.c_int64()
        if single:
            check(lib.tn_dmrg_sweep(psi.h, Hs.h, int(direction), lz, tr, C.byref(e), C.byref(mb)))
        else:
            check(lib.tn_dmrg_sweep_sum(psi.h, Hs.h, int(direction), int(nsites), lz, tr, C.byref(e), C.byref(mb)))
        cost = e.value
        direction = not direction
        sweeps += 1
        D = mb.value

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


def vmps(*psis, minsweeps=2, maxsweeps=200, tol=1e-10, numconverges=3, verbose=False, nsites=2, cutoff=1e-12, maxdim=1000,
         mindim=1, history=None):
    """vmps(psis::GMPS...; kwargs...): algorithms/mps/vmps.jl:1-105.  psi0 = copy of psis[0] with its centre at site 1,
    Vs = ProjMPSSum([ProjMPS(psi_k, psi0)]); the sweep body (:36-62) runs on the device (tn_vmps_sweep), the cost
    norm(psi)^2 - 2|calculate(Vs)| and the convergence counters (:64-83) host-side."""
    first = psis[0]
    psi = GMPS(first.rank, first.dim, first.tensors, first.center, first.ctx)     # deepcopy(psis[1]) (vmps.jl:97)
    if psi.rank != 1:
        raise _lib.TNError("vmps: rank-1 MPS only")
    psi.movecenter(1)
    Vs = ProjMPSSum([ProjMPS(p, None, psi, center=1) for p in psis], center=1)
    return vmps_sweeps(psi, Vs, minsweeps=minsweeps, maxsweeps=maxsweeps, tol=tol, numconverges=numconverges, verbose=verbose,
                       nsites=nsites, cutoff=cutoff, maxdim=maxdim, mindim=mindim, history=history)


def vmps_sweeps(psi, Vs, minsweeps=2, maxsweeps=200, tol=1e-10, numconverges=3, verbose=False, nsites=2, cutoff=1e-12,
                maxdim=1000, mindim=1, history=None):
    """vmps(psi, Vs::AbstractProjMPS; kwargs...): vmps.jl:1-92 for a ProjMPSSum of ProjMPS(psi_k, [H,] psi)."""
    lib = psi.lib
    tr = Trunc(cutoff, maxdim, mindim)

    def calculatecost():                      # vmps.jl:16-23
        return psi.norm() ** 2 - 2 * abs(Vs.calculate())

    def diff(x, y):
        return abs(x - y) if abs(x) < 1e-10 else abs((x - y) / x)
    lastcost = calculatecost()
    lastD = psi.maxbonddim()
    direction = False
    converged = False
    convergedsweeps = sweeps = 0
    while not converged:
        mb = C.c_int64()
        check(lib.tn_vmps_sweep(psi.h, Vs.h, int(direction), int(nsites), tr, C.byref(mb)))
        Vs.center = 1 if direction else len(psi)
        direction = not direction
        sweeps += 1
        D = mb.value
        cost = calculatecost()
        if sweeps >= minsweeps:
            if diff(cost, lastcost) < tol and lastD == D:
                convergedsweeps += 1          # never reset: vmps.jl:75 compares instead of assigning
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


def applygates(psi, gates, cutoff=0.0, maxdim=0, mindim=1, error=False):
    """applygates!(psi, gates; kwargs...): gatelist.jl:225-227; with ``error=True`` applygates(...; error=true) (:191-223), which
    returns the product of the two-site gates' truncation fidelities."""
    if not error:
        check(psi.lib.tn_apply_gates(psi.h, gates.h, Trunc(cutoff, maxdim, mindim)))
        return None
    f = C.c_double()
    check(psi.lib.tn_apply_gates_fidelity(psi.h, gates.h, Trunc(cutoff, maxdim, mindim), C.byref(f)))
    return f.value


def tebd(psi, gates, nsteps, energy_fn=None, nsave=1, cutoff=1e-12, maxdim=0, mindim=1, norm=0.0, observers=()):
    """tebd loop body, algorithms/mps/tebd.jl:62-96, on pre-Trotterised gates: applygates!,
    log-norm accumulation, normalize!; ``energy_fn(psi)`` / observers are called every nsave steps."""
    normal = float(norm)
    energy = energy_fn(psi) if energy_fn else None
    for step in range(1, nsteps + 1):
        applygates(psi, gates, cutoff=cutoff, maxdim=maxdim, mindim=mindim)
        normal += float(np.log(np.real(psi.norm())))
        psi.normalize()
        if step % nsave == 0:
            if energy_fn:
                energy = energy_fn(psi)
            for ob in observers:
                ob(step, psi, normal, energy)
    return psi, energy, normal


def qjmc_simulation(psi, gates, jump_sites, jump_ops, jump_coeffs, steps, dt, uniforms=None, seed=0, trajectory=0,
                    obs_op=None, save_every=1, cutoff=1e-12, maxdim=0, mindim=1, classical=True):
    """qjmc_simulation loop, algorithms/mps/qjmc.jl:59-164 (classical=true :88-112 or the norm-based branch :65-87), one trajectory.
    Returns (jumps, jumptimes, observable array [nsaves, N])."""
    lib = psi.lib
    nj = len(jump_sites)
    js = np.asarray(jump_sites, dtype=np.int32)
    jo = np.ascontiguousarray(np.stack([_f(o).T for o in jump_ops]))
    jc = np.asarray(jump_coeffs, dtype=np.float64)
    un = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
    if un is not None and un.size < 3 * steps:
        raise _lib.TNError("need 3 uniforms per step")
    N = len(psi)
    nsaves = steps // save_every if obs_op is not None else 0
    obs = np.zeros((max(nsaves, 1), N), dtype=np.complex128)
    oo = None if obs_op is None else np.ascontiguousarray(_f(obs_op).T)
    cap = steps + 1
    jumps = np.zeros(cap, dtype=np.int32)
    times = np.zeros(cap)
    njumps = C.c_int32()
    check(lib.tn_qjmc_run(psi.h, gates.h, nj, js.ctypes.data_as(C.POINTER(C.c_int32)), _ptr(jo), jc.ctypes.data_as(C.POINTER(C.c_double)),
                          int(steps), float(dt), Trunc(cutoff, maxdim, mindim),
                          None if un is None else un.ctypes.data_as(C.POINTER(C.c_double)), int(seed), int(trajectory),
                          None if oo is None else _ptr(oo), int(save_every), _ptr(obs), jumps.ctypes.data_as(C.POINTER(C.c_int32)),
                          times.ctypes.data_as(C.POINTER(C.c_double)), cap, C.byref(njumps), int(bool(classical))))
    n = min(njumps.value, cap)
    return list(jumps[:n]), list(times[:n]), obs[:nsaves]


def qjmc_ensemble(tensors, center, gate_sites, gate_tensors, jump_sites, jump_ops, jump_coeffs, steps, dt, traj_ids, workers=8,
                  device=0, seed=0, obs_op=None, save_every=1, cutoff=1e-12, maxdim=0, mindim=1, classical=True):
    """Many independent ``qjmc_simulation`` trajectories (the loop a user writes around algorithms/mps/qjmc.jl:28) from one
    initial MPS, distributed over ``workers`` host threads / CUDA streams inside the library (tn_qjmc_ensemble).
    Returns (njumps [T], jumps [T, steps+1], jumptimes [T, steps+1], observable [T, nsaves, N])."""
    lib = _lib.load()
    ts = [_f(t) for t in tensors]
    N, d = len(ts), ts[0].shape[1]
    dims = np.array([t.shape for t in ts], dtype=np.int64).reshape(N, 3)
    sptr = (C.c_void_p * N)(*[t.ctypes.data for t in ts])
    flat = [(int(s), _f(g)) for rs, rg in zip(gate_sites, gate_tensors) for s, g in zip(rs, rg)]
    counts = np.array([len(r) for r in gate_sites], dtype=np.int32)
    gs = np.array([s for s, _ in flat], dtype=np.int32)
    gn = np.array([g.ndim // 2 for _, g in flat], dtype=np.int32)
    gptr = (C.c_void_p * len(flat))(*[g.ctypes.data for _, g in flat])
    js = np.asarray(jump_sites, dtype=np.int32)
    jo = np.ascontiguousarray(np.stack([_f(o).T for o in jump_ops]))
    jc = np.asarray(jump_coeffs, dtype=np.float64)
    ids = np.ascontiguousarray(traj_ids, dtype=np.uint64)
    T = len(ids)
    nsaves = steps // save_every if obs_op is not None else 0
    obs = np.zeros((T, max(nsaves, 1), N), dtype=np.complex128)
    oo = None if obs_op is None else np.ascontiguousarray(_f(obs_op).T)
    cap = steps + 1
    njumps = np.zeros(T, dtype=np.int32)
    jumps = np.zeros((T, cap), dtype=np.int32)
    times = np.zeros((T, cap))
    i32p, f64p = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    check(lib.tn_qjmc_ensemble(int(device), int(workers), T, ids.ctypes.data_as(C.POINTER(C.c_uint64)), d, N,
                               dims.ctypes.data_as(C.POINTER(C.c_int64)), sptr, int(center), len(counts), counts.ctypes.data_as(i32p),
                               gs.ctypes.data_as(i32p), gn.ctypes.data_as(i32p), gptr, len(js), js.ctypes.data_as(i32p), _ptr(jo),
                               jc.ctypes.data_as(f64p), int(steps), float(dt), Trunc(cutoff, maxdim, mindim), int(seed),
                               None if oo is None else _ptr(oo), int(save_every), _ptr(obs), njumps.ctypes.data_as(i32p),
                               jumps.ctypes.data_as(i32p), times.ctypes.data_as(f64p), cap, int(bool(classical))))
    return njumps, jumps, times, obs[:, :nsaves]


def inner(psi, terms, phi=None):
    """inner(st, psi, oplist, phi): mps.jl:87-134.  ``terms`` is a list of (ops, sites, coeff) with ``ops`` a list of
    d x d matrices and ``sites`` the matching 1-based sites (any order; sorted here like OpList.add does, oplist.jl:35-40).
    Returns the complex array coeff_t * <psi| O_t |phi> (phi defaults to psi)."""
    phi = psi if phi is None else phi
    nops, sites, mats, coeffs = [], [], [], []
    for ops, st, co in terms:
        order = np.argsort(np.asarray(st))
        nops.append(len(order))
        for j in order:
            sites.append(int(st[j]))
            mats.append(np.ascontiguousarray(_f(ops[j]).T))      # d x d block, column-major
        coeffs.append(complex(co))
    n = len(nops)
    out = np.zeros(n, dtype=np.complex128)
    if n == 0:
        return out
    nn = np.asarray(nops, dtype=np.int32)
    ss = np.asarray(sites, dtype=np.int32)
    mm = np.ascontiguousarray(np.stack(mats))
    cc = np.asarray(coeffs, dtype=np.complex128)
    i32p = C.POINTER(C.c_int32)
    check(psi.lib.tn_inner_oplist(psi.h, phi.h, n, nn.ctypes.data_as(i32p), ss.ctypes.data_as(i32p), _ptr(mm), _ptr(cc), _ptr(out)))
    return out
