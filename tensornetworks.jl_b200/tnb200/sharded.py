"""Multi-GPU paths of SURVEY.md 8(e): one process per GPU, torch.distributed for the plumbing.

1. MPO-bond-sharded H_eff matvec (large chi, config C5).  Rank g owns L(:, w in g, :) and R(:, w2 in g, :);
   Theta is replicated.
     stage 1  T1_g = L_g . Theta                                    (chi^3 d^2 w / G complex MACs, local)
              T2p(a,s1,s2,b',w2) = sum_{w in g,s1',s2'} T1_g W      (partial sums for ALL w2, small)
     exchange reduce_scatter over w2 (w2 is the slowest index, so rank r's chunk is contiguous)
     stage 2  out_p = T2_g . R_g                                    (chi^3 d^2 w / G, local)
     exchange all_reduce(out_p)
   The reference (projmps.jl:107-134) is single-process; results equal the unsharded matvec up to rounding.
2. QJMC ensembles: trajectory t -> rank t mod G, no collective during the evolution, one gather at the end.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, tn_cplx, tn_idx2_t

BIG = 1 << 40


def shard_range(n, rank, world):
    """Contiguous block partition of range(n) (first n % world ranks get one extra)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def dense_w(M1, M2):
    """W[(w,s1',s2'),(s1,s2,w2)] = sum_w1 M1(w,s1,s1',w1) M2(w1,s2,s2',w2) (tiny, host-side)."""
    W = np.einsum('wstx,xuvy->wtvsuy', M1, M2)
    w, d = M1.shape[0], M1.shape[1]
    return np.reshape(W, (w * d * d, d * d * M2.shape[3]), order='F')


class HeffSharded:
    """MPO-bond-sharded H_eff application over several GPUs of THIS process, entirely behind the C ABI (tn_heff_sharded_*: one stream
    per device, NCCL communicators owned by the library) -- what a ccall caller uses.  L (chi, w, chi), R (chi2, w2, chi2),
    M1 (w, d, d, w1), M2 (w1, d, d, w2) are host arrays; apply(theta) takes and returns host arrays (projmps.jl:103-145)."""

    def __init__(self, L, R, M1, M2, devices, coeff=1.0):
        self.lib = _lib.load()
        L, R, M1, M2 = (np.asfortranarray(np.asarray(x, dtype=np.complex128)) for x in (L, R, M1, M2))
        chi, w, _ = L.shape
        chi2, w2, _ = R.shape
        d, w1 = M1.shape[1], M1.shape[3]
        self.shape_in = (chi, d, d, chi2)
        dev = np.asarray(list(devices), dtype=np.int32)
        h = C.c_void_p()
        cf = complex(coeff)
        check(self.lib.tn_heff_sharded_create(len(dev), dev.ctypes.data_as(C.POINTER(C.c_int32)), chi, chi2, d, w, w1, w2,
                                              C.c_void_p(L.ctypes.data), C.c_void_p(R.ctypes.data), C.c_void_p(M1.ctypes.data), C.c_void_p(M2.ctypes.data),
                                              tn_cplx(cf.real, cf.imag), C.byref(h)))
        self.h = h

    def apply(self, theta, out=None):
        theta = np.asfortranarray(np.asarray(theta, dtype=np.complex128))
        assert theta.shape == self.shape_in
        if out is None:
            out = np.empty(self.shape_in, dtype=np.complex128, order='F')
        check(self.lib.tn_heff_sharded_apply(self.h, C.c_void_p(theta.ctypes.data), C.c_void_p(out.ctypes.data)))
        return out

    def __del__(self):
        try:
            if self.h:
                self.lib.tn_heff_sharded_free(self.h)
                self.h = None
        except Exception:
            pass


class GpuContractor:
    """C = alpha * A B + beta * C on device buffers through the C ABI (tn_contract_strided_dev)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def __call__(self, M, N, K, A, am, ak, B, bk, bn, Cc, cm, cn, beta=0.0):
        check(self.ctx.lib.tn_contract_strided_dev(self.ctx.h, M, N, K, C.c_void_p(A.data_ptr()), tn_idx2_t(*am), tn_idx2_t(*ak), 0,
                                                   C.c_void_p(B.data_ptr()), tn_idx2_t(*bk), tn_idx2_t(*bn), 0,
                                                   C.c_void_p(Cc.data_ptr()), tn_idx2_t(*cm), tn_idx2_t(*cn), tn_cplx(1.0, 0.0), tn_cplx(beta, 0.0)))

    def sync(self):
        self.ctx.sync()


class ShardedHeff:
    """H_eff matvec sharded over the MPO bond.  All tensors are flat torch complex128 buffers in Julia
    (column-major) order.  ``contract`` is the strided-contraction backend (GpuContractor on the GPU)."""

    def __init__(self, L, R, M1, M2, rank, world, contract, device, dist=None, pad_w2=True):
        import torch
        self.torch, self.dist, self.rank, self.world, self.contract = torch, dist, rank, world, contract
        chi_a, w, chi_b = L.shape
        chi_a2, w2, chi_b2 = R.shape
        d = M1.shape[1]
        self.dims = (chi_a, chi_b, chi_a2, chi_b2, d, w, w2)
        self.w2_pad = ((w2 + world - 1) // world) * world        # equal chunks for reduce_scatter
        self.w2g = self.w2_pad // world
        lo, hi = shard_range(w, rank, world)
        self.wg = hi - lo
        Wfull = dense_w(M1, M2)                                   # (w d^2, d^2 w2)
        Wg = Wfull.reshape(w, d * d, d * d * w2, order='F')[lo:hi].reshape(self.wg * d * d, d * d * w2, order='F')
        r_lo, r_hi = rank * self.w2g, min((rank + 1) * self.w2g, w2)
        Rg = np.zeros((chi_a2, self.w2g, chi_b2), dtype=np.complex128)
        if r_hi > r_lo:
            Rg[:, :r_hi - r_lo, :] = R[:, r_lo:r_hi, :]

        def dev(x):
            return torch.from_numpy(np.ascontiguousarray(np.reshape(x, -1, order='F'))).to(device)
        self.Lg, self.Rg, self.Wg = dev(L[:, lo:hi, :]), dev(Rg), dev(Wg)
        n_t1 = chi_a * max(self.wg, 1) * d * d * chi_b2
        self.T1 = torch.zeros(n_t1, dtype=torch.complex128, device=device)
        self.T2p = torch.zeros(chi_a * d * d * chi_b2 * self.w2_pad, dtype=torch.complex128, device=device)
        self.T2g = torch.zeros(chi_a * d * d * chi_b2 * self.w2g, dtype=torch.complex128, device=device)
        self.out = torch.zeros(chi_a * d * d * chi_a2, dtype=torch.complex128, device=device)

    def apply(self, theta):
        """theta: flat complex128 buffer of Theta(chi_b, d, d, chi_b2).  Returns H_eff*Theta (chi_a, d, d, chi_a2), flat."""
        torch, dist = self.torch, self.dist
        chi_a, chi_b, chi_a2, chi_b2, d, w, w2 = self.dims
        d2, wg, w2g = d * d, self.wg, self.w2g
        ct = self.contract
        if wg > 0:
            # T1[(a,wg),(s1',s2',b')] = L_g[(a,wg),b] Theta[b,(s1',s2',b')]
            ct(chi_a * wg, d2 * chi_b2, chi_b, self.Lg, (BIG, 1, 0), (BIG, chi_a * wg, 0), theta, (BIG, 1, 0), (BIG, chi_b, 0),
               self.T1, (BIG, 1, 0), (BIG, chi_a * wg, 0))
            # T2p(a,s1,s2,b',w2) = sum_{(wg,s1',s2')} T1(a,(wg,s1',s2'),b') Wg[(wg,s1',s2'),(s1,s2,w2)]
            ct(chi_a * chi_b2, d2 * w2, wg * d2, self.T1, (chi_a, 1, chi_a * wg * d2), (BIG, chi_a, 0), self.Wg, (BIG, 1, 0), (BIG, wg * d2, 0),
               self.T2p, (chi_a, 1, chi_a * d2), (d2, chi_a, chi_a * d2 * chi_b2))
        else:
            self.T2p.zero_()
        ct.sync()
        if self.world > 1:
            dist.reduce_scatter_tensor(torch.view_as_real(self.T2g), torch.view_as_real(self.T2p), op=dist.ReduceOp.SUM)
        else:
            self.T2g.copy_(self.T2p[:self.T2g.numel()])
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        # out_p[(a,s1,s2),a'] = sum_{(b',w2g)} T2g[(a,s1,s2),(b',w2g)] R_g(a',w2g,b')
        ct(chi_a * d2, chi_a2, chi_b2 * w2g, self.T2g, (BIG, 1, 0), (BIG, chi_a * d2, 0), self.Rg, (chi_b2, chi_a2 * w2g, chi_a2), (BIG, 1, 0),
           self.out, (BIG, 1, 0), (BIG, chi_a * d2, 0))
        ct.sync()
        if self.world > 1:
            dist.all_reduce(torch.view_as_real(self.out), op=dist.ReduceOp.SUM)
            if torch.cuda.is_available():
                torch.cuda.synchronize()
        return self.out


    # ---- pipelined variant: the reduce_scatter of slice j runs under the chi^3 GEMM of slice j+1 -------------------------
    def apply_pipelined(self, theta, nslices=4, lib_stream=None):
        """Same result as :meth:`apply`, with Theta's right bond cut into ``nslices`` slices: stage 1+2 of slice j+1 is enqueued
        while the reduce_scatter of slice j is in flight on a communication stream, and stage 3 accumulates slice by slice
        (out += T2g_j . R_g[:, :, slice j]).  At chi = 2048, w = 24 on 8 GPUs the reduce_scatter payload is 6.4 GB per rank
        (~25 % of the unpipelined matvec).  ``lib_stream``: the library's CUDA stream handle (Context.stream()); without a GPU
        (gloo tests) the slices simply run one after the other."""
        torch, dist = self.torch, self.dist
        chi_a, chi_b, chi_a2, chi_b2, d, w, w2 = self.dims
        d2, wg, w2g = d * d, self.wg, self.w2g
        ct = self.contract
        nslices = max(1, min(int(nslices), chi_b2))
        cuts = [chi_b2 * j // nslices for j in range(nslices + 1)]
        gpu = torch.cuda.is_available() and lib_stream is not None
        if gpu:
            lib = torch.cuda.ExternalStream(int(lib_stream))
            if not hasattr(self, "_comm"):
                self._comm = torch.cuda.Stream()
            comm = self._comm
        if not hasattr(self, "_pipe") or self._pipe[0] != cuts:
            mk = lambda n: torch.zeros(max(n, 1), dtype=torch.complex128, device=self.out.device)
            bufs = [(mk(chi_a * d2 * (cuts[j + 1] - cuts[j]) * self.w2_pad), mk(chi_a * d2 * (cuts[j + 1] - cuts[j]) * w2g)) for j in range(nslices)]
            self._pipe = (cuts, bufs)
            if gpu:
                torch.cuda.synchronize()
        bufs = self._pipe[1]

        def stage12(j):
            b0, nb = cuts[j], cuts[j + 1] - cuts[j]
            T2p = bufs[j][0]
            if wg > 0:
                th = theta[chi_b * d2 * b0:]
                ct(chi_a * wg, d2 * nb, chi_b, self.Lg, (BIG, 1, 0), (BIG, chi_a * wg, 0), th, (BIG, 1, 0), (BIG, chi_b, 0),
                   self.T1, (BIG, 1, 0), (BIG, chi_a * wg, 0))
                ct(chi_a * nb, d2 * w2, wg * d2, self.T1, (chi_a, 1, chi_a * wg * d2), (BIG, chi_a, 0), self.Wg, (BIG, 1, 0), (BIG, wg * d2, 0),
                   T2p, (chi_a, 1, chi_a * d2), (d2, chi_a, chi_a * d2 * nb))
            elif gpu:
                with torch.cuda.stream(lib):                   # ordered like the GEMMs it stands in for
                    T2p.zero_()
            else:
                T2p.zero_()

        def exchange(j):
            T2p, T2g = bufs[j]
            if self.world > 1:
                dist.reduce_scatter_tensor(torch.view_as_real(T2g), torch.view_as_real(T2p), op=dist.ReduceOp.SUM)
            else:
                T2g.copy_(T2p[:T2g.numel()])

        def stage3(j):
            b0, nb = cuts[j], cuts[j + 1] - cuts[j]
            Rg = self.Rg[chi_a2 * w2g * b0:]
            ct(chi_a * d2, chi_a2, nb * w2g, bufs[j][1], (BIG, 1, 0), (BIG, chi_a * d2, 0), Rg, (nb, chi_a2 * w2g, chi_a2), (BIG, 1, 0),
               self.out, (BIG, 1, 0), (BIG, chi_a * d2, 0), beta=0.0 if j == 0 else 1.0)

        if not gpu:
            for j in range(nslices):
                stage12(j)
                ct.sync()
                exchange(j)
                stage3(j)
            ct.sync()
        else:
            done12 = [torch.cuda.Event() for _ in range(nslices)]
            donex = [torch.cuda.Event() for _ in range(nslices)]
            for j in range(nslices):
                stage12(j)                                     # library stream
                done12[j].record(lib)
                with torch.cuda.stream(comm):                  # collective ordered after stage 2 of this slice only
                    comm.wait_event(done12[j])
                    exchange(j)
                    donex[j].record(comm)
                if j >= 1:                                     # stage 3 of the previous slice: its exchange overlapped stage 1+2 of slice j
                    lib.wait_event(donex[j - 1])
                    stage3(j - 1)
            lib.wait_event(donex[nslices - 1])
            stage3(nslices - 1)
            fin = torch.cuda.Event()
            fin.record(lib)
            torch.cuda.current_stream().wait_event(fin)
        if self.world > 1:
            dist.all_reduce(torch.view_as_real(self.out), op=dist.ReduceOp.SUM)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        return self.out


class _BalancedPipelineMixin:
    """apply_pipelined for BalancedShardedHeff: the balanced split of both chi^3 stages combined with the overlap of the
    reduce_scatter of slice j with stage 1+2 of slice j+1 (see ShardedHeff.apply_pipelined)."""

    def _prepare_slices(self, nslices):
        torch = self.torch
        ca, cb, ca2, cb2, d, w, w2 = self.dims
        d2 = d * d
        cuts = [cb2 * j // nslices for j in range(nslices + 1)]
        if getattr(self, "_bp", None) is not None and self._bp[0] == cuts:
            return self._bp
        z = lambda n: torch.zeros(max(int(n), 1), dtype=torch.complex128, device=self.device)
        per = []
        for j in range(nslices):
            b0, nb = cuts[j], cuts[j + 1] - cuts[j]
            ktot = nb * w2
            c = (ktot + self.world - 1) // self.world
            k0, k1 = min(self.rank * c, ktot), min((self.rank + 1) * c, ktot)
            Rk = np.reshape(np.transpose(self._Rhost[:, :, b0:b0 + nb], (2, 1, 0)), (ktot, ca2), order='F')     # rows k = b'_local + nb * w2
            Rg = np.zeros((c, ca2), dtype=np.complex128)
            Rg[:k1 - k0] = Rk[k0:k1]
            per.append(dict(b0=b0, nb=nb, c=c, Rg=self._dev(Rg), T2p=z(ca * d2 * c * self.world), T2g=z(ca * d2 * c)))
        self._bp = (cuts, per)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        return self._bp

    def apply_pipelined(self, theta, nslices=4, lib_stream=None):
        torch, dist = self.torch, self.dist
        ca, cb, ca2, cb2, d, w, w2 = self.dims
        d2, ct = d * d, self.contract
        nslices = max(1, min(int(nslices), cb2))
        _, per = self._prepare_slices(nslices)
        ld1 = ca * self.nw
        gpu = torch.cuda.is_available() and lib_stream is not None
        if gpu:
            lib = torch.cuda.ExternalStream(int(lib_stream))
            if not hasattr(self, "_comm"):
                self._comm = torch.cuda.Stream()
            comm = self._comm

        def stage12(p):
            b0, nb = p["b0"], p["nb"]
            if self.mloc > 0:
                ct(self.mloc, d2 * nb, cb, self.Lg, (BIG, 1, 0), (BIG, self.mloc, 0), theta[cb * d2 * b0:], (BIG, 1, 0), (BIG, cb, 0),
                   self.T1[self.r0:], (BIG, 1, 0), (BIG, ld1, 0))
                ct(ca * nb, d2 * w2, self.nw * d2, self.T1, (ca, 1, ld1 * d2), (BIG, ca, 0), self.Wg, (BIG, 1, 0), (BIG, self.nw * d2, 0),
                   p["T2p"], (ca, 1, ca * d2), (d2, ca, ca * d2 * nb))
            elif gpu:
                with torch.cuda.stream(lib):
                    p["T2p"].zero_()
            else:
                p["T2p"].zero_()

        def exchange(p):
            if self.world > 1:
                dist.reduce_scatter_tensor(torch.view_as_real(p["T2g"]), torch.view_as_real(p["T2p"]), op=dist.ReduceOp.SUM)
            else:
                p["T2g"].copy_(p["T2p"][:p["T2g"].numel()])

        def stage3(j, p):
            ct(ca * d2, ca2, p["c"], p["T2g"], (BIG, 1, 0), (BIG, ca * d2, 0), p["Rg"], (BIG, 1, 0), (BIG, p["c"], 0),
               self.out, (BIG, 1, 0), (BIG, ca * d2, 0), beta=0.0 if j == 0 else 1.0)

        if not gpu:
            for j, p in enumerate(per):
                stage12(p)
                ct.sync()
                exchange(p)
                stage3(j, p)
            ct.sync()
        else:
            done12 = [torch.cuda.Event() for _ in per]
            donex = [torch.cuda.Event() for _ in per]
            for j, p in enumerate(per):
                stage12(p)
                done12[j].record(lib)
                with torch.cuda.stream(comm):
                    comm.wait_event(done12[j])
                    exchange(p)
                    donex[j].record(comm)
                if j >= 1:
                    lib.wait_event(donex[j - 1])
                    stage3(j - 1, per[j - 1])
            lib.wait_event(donex[-1])
            stage3(len(per) - 1, per[-1])
            fin = torch.cuda.Event()
            fin.record(lib)
            torch.cuda.current_stream().wait_event(fin)
        if self.world > 1:
            dist.all_reduce(torch.view_as_real(self.out), op=dist.ReduceOp.SUM)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        return self.out


class BalancedShardedHeff(_BalancedPipelineMixin):
    """ShardedHeff with the two chi^3 stages split evenly for ANY MPO bond dimension (w = 20 on 8 GPUs is a 3/3/3/3/2/2/2/2 split of
    whole bond values, an 83 % ceiling).  Stage 1 shards the fused row index (a, w) of L -- rank g owns rows [m0, m1) of the
    (chi w) x chi matrix, whatever bond values they straddle: its GEMM writes them at their natural place in a zeroed
    T1(a, w_first..w_last, s1', s2', b'), so stage 2 is the ordinary small contraction over those bond values.  Stage 3 shards the
    fused contraction index k = (b', w2): the reduce_scatter hands out equal flat chunks of T2p(a, s1, s2, [b', w2]) and every rank
    holds the matching rows of R in (k, a') layout.  Same collectives as ShardedHeff."""

    def __init__(self, L, R, M1, M2, rank, world, contract, device, dist=None):
        import torch
        self.torch, self.dist, self.rank, self.world, self.contract = torch, dist, rank, world, contract
        ca, w, cb = L.shape
        ca2, w2, cb2 = R.shape
        d = M1.shape[1]
        d2 = d * d
        self.dims = (ca, cb, ca2, cb2, d, w, w2)
        mtot = ca * w
        m0, m1 = (mtot * rank) // world, (mtot * (rank + 1)) // world
        self.mloc = m1 - m0
        w_first = m0 // ca if self.mloc > 0 else 0
        w_last = (m1 - 1) // ca if self.mloc > 0 else 0
        self.nw = w_last - w_first + 1
        self.r0 = m0 - ca * w_first
        ktot = cb2 * w2
        self.c = (ktot + world - 1) // world
        k0, k1 = min(rank * self.c, ktot), min((rank + 1) * self.c, ktot)
        Lmat = np.reshape(L, (mtot, cb), order='F')
        Wfull = dense_w(M1, M2)
        Wg = Wfull.reshape(w, d2, d2 * w2, order='F')[w_first:w_last + 1].reshape(self.nw * d2, d2 * w2, order='F')
        Rk = np.reshape(np.transpose(R, (2, 1, 0)), (ktot, ca2), order='F')        # rows k = b' + chi * w2
        Rg = np.zeros((self.c, ca2), dtype=np.complex128)
        Rg[:k1 - k0] = Rk[k0:k1]

        def dev(x):
            return torch.from_numpy(np.ascontiguousarray(np.reshape(x, -1, order='F'))).to(device)
        self.Lg, self.Wg, self.Rg = dev(np.asfortranarray(Lmat[m0:m1])), dev(Wg), dev(Rg)
        self._Rhost, self._dev, self.device = R, dev, device          # the pipelined variant re-slices R per Theta slice
        z = lambda n: torch.zeros(max(int(n), 1), dtype=torch.complex128, device=device)
        self.T1 = z(ca * self.nw * d2 * cb2)           # rows outside [r0, r0 + mloc) stay zero
        self.T2p = z(ca * d2 * self.c * world)
        self.T2g = z(ca * d2 * self.c)
        self.out = z(ca * d2 * ca2)

    def apply(self, theta):
        torch, dist = self.torch, self.dist
        ca, cb, ca2, cb2, d, w, w2 = self.dims
        d2, ct = d * d, self.contract
        ld1 = ca * self.nw
        if self.mloc > 0:
            # T1[r0 + m, (s1',s2',b')] = L_g[m, b] Theta[b, (s1',s2',b')]
            ct(self.mloc, d2 * cb2, cb, self.Lg, (BIG, 1, 0), (BIG, self.mloc, 0), theta, (BIG, 1, 0), (BIG, cb, 0),
               self.T1[self.r0:], (BIG, 1, 0), (BIG, ld1, 0))
            # T2p(a,s1,s2,b',w2) = sum_{(w,s1',s2')} T1(a,(w,s1',s2'),b') W_g[(w,s1',s2'),(s1,s2,w2)]
            ct(ca * cb2, d2 * w2, self.nw * d2, self.T1, (ca, 1, ld1 * d2), (BIG, ca, 0), self.Wg, (BIG, 1, 0), (BIG, self.nw * d2, 0),
               self.T2p, (ca, 1, ca * d2), (d2, ca, ca * d2 * cb2))
        else:
            self.T2p.zero_()
        ct.sync()
        if self.world > 1:
            dist.reduce_scatter_tensor(torch.view_as_real(self.T2g), torch.view_as_real(self.T2p), op=dist.ReduceOp.SUM)
        else:
            self.T2g.copy_(self.T2p[:self.T2g.numel()])
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        # out[(a,s1,s2), a'] = sum_k T2g[(a,s1,s2), k] R_g[k, a']
        ct(ca * d2, ca2, self.c, self.T2g, (BIG, 1, 0), (BIG, ca * d2, 0), self.Rg, (BIG, 1, 0), (BIG, self.c, 0),
           self.out, (BIG, 1, 0), (BIG, ca * d2, 0))
        ct.sync()
        if self.world > 1:
            dist.all_reduce(torch.view_as_real(self.out), op=dist.ReduceOp.SUM)
            if torch.cuda.is_available():
                torch.cuda.synchronize()
        return self.out


# ---------------------------------------------------------------------------------------------
# QJMC ensembles
# ---------------------------------------------------------------------------------------------
def my_trajectories(ntraj, rank, world):
    """Round-robin ownership: trajectory t runs on rank t mod world."""
    return list(range(rank, ntraj, world))


def run_ensemble(run_one, ntraj, rank=0, world=1, dist=None, workers=1):
    """Run ``run_one(t) -> picklable result`` for every owned trajectory (optionally on ``workers`` host
    threads, each expected to use its own tn_ctx/stream) and gather {t: result} on every rank."""
    mine = my_trajectories(ntraj, rank, world)
    results = {}
    if workers <= 1:
        for t in mine:
            results[t] = run_one(t)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=workers) as ex:
            for t, r in zip(mine, ex.map(run_one, mine)):
                results[t] = r
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, results)
        merged = {}
        for g in gathered:
            merged.update(g)
        results = merged
    return results


# ---------------------------------------------------------------------------------------------
# MPO-bond-sharded environments and DMRG sweep (SURVEY 8(e), row "DMRG matvec at large chi (A5) + environments (A4)")
#
# Rank g owns the slice w in chunk(g) of EVERY environment block (chi_bra, w, chi_ket), so a block set that does not
# fit one GPU (C5: ~390 GB) is spread over the box; the site tensors, Theta, the Lanczos vectors and the truncated SVD
# are replicated.  chunk(g) = [g*c, min((g+1)*c, w)) with c = ceil(w / G): the equal chunks reduce_scatter needs (the
# largest chunk -- the critical path -- is the same as for a balanced split).
#   buildleft  (projmps.jl:50-66):  X1_g = L_g . A  ->  X2p = X1_g . M[w in g]  (partial sums for ALL w')
#                                    -> reduce_scatter over w' -> L'_g = conj(A) . X2_g
#   buildright (projmps.jl:74-95):  mirror image with a reduce_scatter over the left MPO bond
#   product    (projmps.jl:107-134): as ShardedHeff above, blocks taken from the store, all_reduce of the result
# The bond loop of dmrg.jl:35-63 runs in this process; the dense pieces go through a ``backend``:
#   GpuBackend -- C ABI (tn_contract_strided_dev, tn_eigsolve_fn, tn_mps_replacesites_dev, ...) + torch.distributed/NCCL.
# The CPU tests drive the same code with a NumPy stand-in backend over gloo (tests/test_sharded_cpu.py).
# ---------------------------------------------------------------------------------------------
def chunk_range(n, rank, world):
    """Equal-chunk partition of range(n): chunk size ceil(n / world); trailing ranks may be empty."""
    c = (n + world - 1) // world
    return min(rank * c, n), min((rank + 1) * c, n)


def _i1(stride):
    return (BIG, int(stride), 0)


def _i2(n0, s0, s1):
    return (int(n0), int(s0), int(s1))


class ShardedProjMPS:
    """ProjMPS(psi, H, psi; rank=2) with every block sharded over the MPO bond (abstractprojmps.jl:33-81 protocol).

    ``psi`` is the backend's MPS handle (replicated), ``mpo_host`` the list of MPO site tensors (w_l, d, d, w_r) as NumPy
    arrays (tiny, host-side), ``be`` the backend.  Blocks are flat column-major buffers (chi_bra, w_g, chi_ket)."""

    def __init__(self, psi, mpo_host, be, rank=0, world=1, dist=None, center=1, coeff=1.0):
        self.psi, self.be, self.rank, self.world, self.dist = psi, be, rank, world, dist
        self.mpo = [np.asarray(m, dtype=np.complex128) for m in mpo_host]
        self.N = len(self.mpo)
        self.d = self.mpo[0].shape[1]
        self.coeff = complex(coeff)
        self.blocks = [None] * self.N          # per site: (buffer, (chi_bra, w_g, chi_ket)) or None
        self.center = 0
        self._mslice = {}
        self.movecenter(center)

    # ---- helpers -------------------------------------------------------------------------------------------------
    def _wrange(self, w):
        return chunk_range(w, self.rank, self.world)

    def _wchunk(self, w):
        return (w + self.world - 1) // self.world

    def _edge(self):
        """ones(1,1,1) (abstractprojmps.jl:33-35): rank 0 owns the single MPO-bond value."""
        wg = 1 if self.rank == 0 else 0
        buf = self.be.new(max(wg, 1))
        if wg:
            self.be.fill_ones(buf, 1)
        return buf, (1, wg, 1)

    def block(self, idx):
        if idx < 1 or idx > self.N:
            return self._edge()
        b = self.blocks[idx - 1]
        if b is None:
            raise _lib.TNError("environment block has not been built")
        return b

    def _mpo_dev(self, key, arr):
        """flat device copy of a (sliced) MPO tensor, cached"""
        if key not in self._mslice:
            self._mslice[key] = self.be.from_host(np.reshape(arr, -1, order='F'))
        return self._mslice[key]

    def _reduce_scatter(self, full, chunk_elems):
        """sum over ranks of ``full`` (world equal chunks, slowest index = the sharded MPO bond); returns this rank's chunk"""
        if self.world == 1:
            return full
        out = self.be.new(chunk_elems)
        self.be.sync()
        self.be.reduce_scatter(out, full, self.dist)
        return out

    # ---- block updates -------------------------------------------------------------------------------------------
    def buildleft(self, idx):
        be, d = self.be, self.d
        Lb, (ca, wg, cb) = self.block(idx - 1)
        A, adims = be.site(self.psi, idx)
        ca_, _, cb2 = adims
        ca2 = cb2                                             # bra == ket
        M = self.mpo[idx - 1]
        w, w2 = M.shape[0], M.shape[3]
        lo, hi = self._wrange(w)
        assert hi - lo == wg and ca_ == ca == cb, "buildleft: block / site dimension mismatch"
        c2 = self._wchunk(w2)
        w2pad = c2 * self.world
        X2p = be.new(ca * d * cb2 * w2pad)                    # (a, s, b', w'_pad), zero-initialised
        if wg > 0:
            # X1[(a,wg),(s',b')] = L_g[(a,wg),b] A[b,(s',b')]
            X1 = be.new(ca * wg * d * cb2)
            be.contract(ca * wg, d * cb2, cb, Lb, _i1(1), _i1(ca * wg), 0, A, _i1(1), _i1(cb), 0, X1, _i1(1), _i1(ca * wg))
            # X2p(a,s,b',w') = sum_{(wg,s')} X1(a,(wg,s'),b') M_g(wg,s,s',w'); rows m = (a,b'), n = (s,w')
            Mg = self._mpo_dev(("L", idx, lo, hi), M[lo:hi])
            be.contract(ca * cb2, d * w2, wg * d, X1, _i2(ca, 1, ca * wg * d), _i1(ca), 0,
                        Mg, _i2(wg, 1, wg * d), _i2(d, wg, wg * d * d), 0,
                        X2p, _i2(ca, 1, ca * d), _i2(d, ca, ca * d * cb2))
        X2g = self._reduce_scatter(X2p, ca * d * cb2 * c2)    # (a, s, b', c2); first w2g columns valid
        lo2, hi2 = self._wrange(w2)
        w2g = hi2 - lo2
        out = be.new(max(ca2 * w2g * cb2, 1))
        if w2g > 0:
            # L'_g(a',w'g,b') = sum_{(a,s)} conj(A[(a,s),a']) X2g[(a,s),(b',w'g)]
            be.contract(ca2, cb2 * w2g, ca * d, A, _i1(ca * d), _i1(1), 1, X2g, _i1(1), _i1(ca * d), 0,
                        out, _i1(1), _i2(cb2, ca2 * w2g, ca2))
        be.sync()
        self.blocks[idx - 1] = (out, (ca2, w2g, cb2))

    def buildright(self, idx):
        be, d = self.be, self.d
        Rb, (ca, wg, cb) = self.block(idx + 1)
        A, adims = be.site(self.psi, idx)
        cbl, _, cb_ = adims
        cal = cbl
        M = self.mpo[idx - 1]
        wl, w = M.shape[0], M.shape[3]
        lo, hi = self._wrange(w)
        assert hi - lo == wg and cb_ == cb == ca, "buildright: block / site dimension mismatch"
        cl = self._wchunk(wl)
        wlpad = cl * self.world
        Y2p = be.new(cbl * ca * d * wlpad)                    # (b_l, a, s, w_l_pad)
        if wg > 0:
            # Y1(b_l,s',wg,a) = sum_b A[(b_l,s'),b] R_g[(a,wg),b]; n = (a,wg) written transposed
            Y1 = be.new(cbl * d * wg * ca)
            be.contract(cbl * d, ca * wg, cb, A, _i1(1), _i1(cbl * d), 0, Rb, _i1(ca * wg), _i1(1), 0,
                        Y1, _i1(1), _i2(ca, cbl * d * wg, cbl * d))
            # Y2p(b_l,a,s,w_l) = sum_{(s',wg)} Y1(b_l,(s',wg),a) M_g(w_l,s,s',wg); rows m = (b_l,a), n = (s,w_l)
            Mg = self._mpo_dev(("R", idx, lo, hi), M[:, :, :, lo:hi])
            be.contract(cbl * ca, d * wl, d * wg, Y1, _i2(cbl, 1, cbl * d * wg), _i1(cbl), 0,
                        Mg, _i1(wl * d), _i2(d, wl, 1), 0, Y2p, _i1(1), _i1(cbl * ca))
        Y2g = self._reduce_scatter(Y2p, cbl * ca * d * cl)    # (b_l, a, s, cl)
        lol, hil = self._wrange(wl)
        wlg = hil - lol
        out = be.new(max(cal * wlg * cbl, 1))
        if wlg > 0:
            # R'_g(a_l,wlg,b_l) = sum_{(a,s)} conj(A)(a_l,s,a) Y2g(b_l,(a,s),wlg); k = (a,s) with a fastest
            be.contract(cal, cbl * wlg, ca * d, A, _i1(1), _i2(ca, cal * d, cal), 1,
                        Y2g, _i1(cbl), _i2(cbl, 1, cbl * ca * d), 0, out, _i1(1), _i2(cbl, cal * wlg, cal))
        be.sync()
        self.blocks[idx - 1] = (out, (cal, wlg, cbl))

    def movecenter(self, idx):                                # abstractprojmps.jl:60-81
        N = self.N
        if idx < 1 or idx > N:
            raise _lib.TNError("The index is out of range.")
        if self.center == 0:
            for i in range(1, idx):
                self.buildleft(i)
            for i in range(1, N - idx + 1):
                self.buildright(N + 1 - i)
        elif idx > self.center:
            for i in range(1, idx - self.center + 1):
                self.buildleft(self.center - 1 + i)
        elif idx < self.center:
            for i in range(1, self.center - idx + 1):
                self.buildright(self.center + 1 - i)
        self.center = idx

    # ---- H_eff application -----------------------------------------------------------------------------------------
    def prepare(self, site):
        """per bond: W_g = M1[w in g] . M2 and the work buffers of the two-site product for sites (site, site+1)"""
        be, d = self.be, self.d
        d2 = d * d
        Lb, (ca, wg, cb) = self.block(site - 1)
        Rb, (ca2, w2g, cb2) = self.block(site + 2)
        M1, M2 = self.mpo[site - 1], self.mpo[site]
        w, w2 = M1.shape[0], M2.shape[3]
        lo, hi = self._wrange(w)
        c2 = self._wchunk(w2)
        p = dict(site=site, ca=ca, wg=wg, cb=cb, ca2=ca2, w2g=w2g, cb2=cb2, w2=w2, c2=c2, L=Lb, R=Rb)
        if wg > 0:
            p["W"] = be.from_host(np.reshape(dense_w(M1[lo:hi], M2), -1, order='F'))
            p["T1"] = be.new(ca * wg * d2 * cb2)
        p["T2p"] = be.new(ca * d2 * cb2 * c2 * self.world)
        p["out"] = be.new(ca * d2 * ca2)
        self._prep = p

    def product(self, theta, out):
        """out = coeff * H_eff . theta for the prepared bond; theta / out are backend buffers (chi, d, d, chi), flat."""
        be, d2, p = self.be, self.d * self.d, self._prep
        ca, wg, cb, ca2, w2g, cb2, w2, c2 = (p[k] for k in ("ca", "wg", "cb", "ca2", "w2g", "cb2", "w2", "c2"))
        if wg > 0:
            # T1[(a,wg),(s1',s2',b')] = L_g[(a,wg),b] theta[b,(s1',s2',b')]
            be.contract(ca * wg, d2 * cb2, cb, p["L"], _i1(1), _i1(ca * wg), 0, theta, _i1(1), _i1(cb), 0, p["T1"], _i1(1), _i1(ca * wg))
            # T2p(a,s1,s2,b',w2) = sum_{(wg,s1',s2')} T1(a,(wg,s1',s2'),b') W_g[(wg,s1',s2'),(s1,s2,w2)]
            be.contract(ca * cb2, d2 * w2, wg * d2, p["T1"], _i2(ca, 1, ca * wg * d2), _i1(ca), 0, p["W"], _i1(1), _i1(wg * d2), 0,
                        p["T2p"], _i2(ca, 1, ca * d2), _i2(d2, ca, ca * d2 * cb2))
        T2g = self._reduce_scatter(p["T2p"], ca * d2 * cb2 * c2)
        res = p["out"]
        if w2g > 0:
            # out_p[(a,s1,s2),a'] = coeff * sum_{(b',w2g)} T2g[(a,s1,s2),(b',w2g)] R_g(a',w2g,b')
            be.contract(ca * d2, ca2, cb2 * w2g, T2g, _i1(1), _i1(ca * d2), 0, p["R"], _i2(cb2, ca2 * w2g, ca2), _i1(1), 0,
                        res, _i1(1), _i1(ca * d2), alpha=self.coeff)
        else:
            be.zero(res)
        be.sync()
        if self.world > 1:
            be.all_reduce(res, self.dist)
        be.copy(out, res, ca * d2 * ca2)

    def calculate(self):
        """<psi|H|psi> at the centre (projmps.jl:192-216): left block extended over the centre site, closed with the right block."""
        site = self.center
        saved = self.blocks[site - 1]
        self.blocks[site - 1] = None
        self.buildleft(site)
        tmp, tdims = self.blocks[site - 1]
        self.blocks[site - 1] = saved
        Rb, rdims = self.block(site + 1)
        assert tdims == rdims
        n = tdims[0] * tdims[1] * tdims[2]
        v = self.be.dotu(tmp, Rb, n) if n > 0 else 0.0
        if self.world > 1:
            v = self.be.all_reduce_scalar(v, self.dist)
        return self.coeff * v


def sharded_dmrg(psi, mpo_host, be, rank=0, world=1, dist=None, krylovdim=3, kryloviter=2, minsweeps=1, maxsweeps=1000, tol=1e-10,
                 tolgrad=1e-5, numconverges=4, verbose=False, cutoff=1e-12, maxdim=1000, mindim=1, coeff=1.0, history=None,
                 svd_engine=None):
    """dmrg(psi, H; nsites=2, kwargs...) (algorithms/mps/dmrg.jl:1-154) with the environments and the H_eff application
    sharded over the MPO bond on ``world`` ranks.  Every rank runs this function with the same arguments and ends with the
    same psi: after each bond the two new site tensors of rank 0 are broadcast, so rounding differences between the
    replicated SVDs cannot make the replicas drift apart.  ``svd_engine`` (GpuSvdEngine): distribute the Jacobi sweeps of the
    truncated SVD over the ranks as well (dist_svd_replacesites) instead of factorising redundantly on every rank."""
    N = be.length(psi)
    be.movecenter(psi, 1)
    Hs = ShardedProjMPS(psi, mpo_host, be, rank, world, dist, center=1, coeff=coeff)
    cost = Hs.calculate()
    lastcost = cost
    lastD = D = be.maxbonddim(psi)
    grad = 0.0
    direction = False
    converged = False
    convergedsweeps = convergedgrad = sweeps = 0
    d = Hs.d
    while not converged:
        for j in range(1, N):
            site = N + 1 - j if direction else j
            site1 = site - 1 if direction else site
            Hs.movecenter(site)
            A, (cl, _, cm) = be.site(psi, site1)
            B, (_, _, cr) = be.site(psi, site1 + 1)
            n = cl * d * d * cr
            th0, th1 = be.new(n), be.new(n)
            be.contract(cl * d, d * cr, cm, A, _i1(1), _i1(cl * d), 0, B, _i1(1), _i1(cm), 0, th0, _i1(1), _i1(cl * d))
            Hs.prepare(site1)
            cost = be.eigsolve(Hs.product, th0, th1, n, krylovdim, kryloviter, 1e-14)
            if svd_engine is not None:
                dist_svd_replacesites(be, svd_engine, psi, th1, site1, direction, True, cutoff, maxdim, mindim, rank, world, dist)
            else:
                be.replacesites(psi, th1, site1, direction, True, cutoff, maxdim, mindim)
            if world > 1:
                be.broadcast_site(psi, site1, dist)
                be.broadcast_site(psi, site1 + 1, dist)
        Hs.movecenter(1 if direction else N)
        direction = not direction
        sweeps += 1
        D = be.maxbonddim(psi)

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
        if verbose and rank == 0:
            print("Sweep=%d, energy=%.12f, maxbonddim=%d" % (sweeps, np.real(cost), D))
    return psi, cost


class _Raw:
    """library-owned device memory seen through the backend's buffer protocol (data_ptr only)"""

    def __init__(self, ptr):
        self.ptr = int(ptr)

    def data_ptr(self):
        return self.ptr


class GpuBackend:
    """Dense pieces of the sharded sweep on one GPU: contractions, Lanczos, truncated SVD through the C ABI; buffers are flat
    torch complex128 CUDA tensors (or library memory wrapped in _Raw); collectives through torch.distributed (NCCL)."""

    def __init__(self, ctx, device):
        import torch
        self.torch, self.ctx, self.lib, self.device = torch, ctx, ctx.lib, device

    # buffers.  torch fills / copies run on torch's stream, the library on its own non-blocking stream: every torch-side
    # write is completed (device synchronize) before the buffer is handed to the library.
    def new(self, n):
        b = self.torch.zeros(max(int(n), 1), dtype=self.torch.complex128, device=self.device)
        self.torch.cuda.synchronize()
        return b

    def from_host(self, flat):
        b = self.torch.from_numpy(np.ascontiguousarray(flat, dtype=np.complex128)).to(self.device)
        self.torch.cuda.synchronize()
        return b

    def fill_ones(self, buf, n):
        buf[:n] = 1.0
        self.torch.cuda.synchronize()

    def zero(self, buf):
        buf.zero_()
        self.torch.cuda.synchronize()

    def copy(self, dst, src, n):
        self.torch.cuda.synchronize()
        check(self.lib.tn_memcpy_dev(self.ctx.h, C.c_void_p(dst.data_ptr()), C.c_void_p(src.data_ptr()), int(n) * 16))

    def sync(self):
        self.ctx.sync()
        self.torch.cuda.synchronize()

    def dotu(self, a, b, n):
        self.sync()
        return complex(self.torch.sum(a[:n] * b[:n]).item())

    # dense kernels
    def contract(self, M, N, K, A, am, ak, conjA, B, bk, bn, conjB, Cc, cm, cn, alpha=1.0, beta=0.0):
        al = complex(alpha)
        check(self.lib.tn_contract_strided_dev(self.ctx.h, M, N, K, C.c_void_p(A.data_ptr()), tn_idx2_t(*am), tn_idx2_t(*ak), int(conjA),
                                               C.c_void_p(B.data_ptr()), tn_idx2_t(*bk), tn_idx2_t(*bn), int(conjB),
                                               C.c_void_p(Cc.data_ptr()), tn_idx2_t(*cm), tn_idx2_t(*cn), tn_cplx(al.real, al.imag),
                                               tn_cplx(float(beta), 0.0)))

    def eigsolve(self, apply, th0, th1, n, krylovdim, maxiter, tol):
        from ._lib import tn_lanczos_t, APPLY_FN
        err = []

        def cb(_user, pin, pout):
            try:
                apply(_Raw(pin), _Raw(pout))
                return 0
            except Exception as e:          # never unwind through the C frames
                err.append(e)
                return 1
        e, nops = C.c_double(), C.c_int32()
        self.sync()
        st = self.lib.tn_eigsolve_fn(self.ctx.h, int(n), C.c_void_p(th0.data_ptr()), C.c_void_p(th1.data_ptr()),
                                     tn_lanczos_t(krylovdim, maxiter, tol), APPLY_FN(cb), None, C.byref(e), C.byref(nops))
        if err:
            raise err[0]
        check(st)
        return e.value

    # MPS handle (tnb200.GMPS)
    def length(self, psi):
        return len(psi)

    def movecenter(self, psi, idx):
        psi.movecenter(idx)

    def maxbonddim(self, psi):
        return psi.maxbonddim()

    def site(self, psi, i):
        p = C.c_void_p()
        check(self.lib.tn_mps_site_ptr(psi.h, int(i), C.byref(p)))
        return _Raw(p.value), tuple(int(x) for x in psi.dims()[i - 1])

    def replacesites(self, psi, theta, site, direction, normalize, cutoff, maxdim, mindim):
        from ._lib import tn_trunc_t
        self.sync()
        check(self.lib.tn_mps_replacesites_dev(psi.h, C.c_void_p(theta.data_ptr()), int(site), int(bool(direction)), int(bool(normalize)),
                                               tn_trunc_t(float(cutoff), int(maxdim), int(mindim))))

    # collectives
    def reduce_scatter(self, out, full, dist):
        t = self.torch
        dist.reduce_scatter_tensor(t.view_as_real(out), t.view_as_real(full), op=dist.ReduceOp.SUM)
        t.cuda.synchronize()

    def all_reduce(self, buf, dist):
        t = self.torch
        dist.all_reduce(t.view_as_real(buf), op=dist.ReduceOp.SUM)
        t.cuda.synchronize()

    def all_reduce_scalar(self, v, dist):
        t = self.torch
        x = t.tensor([complex(v).real, complex(v).imag], dtype=t.float64, device=self.device)
        dist.all_reduce(x, op=dist.ReduceOp.SUM)
        return complex(x[0].item(), x[1].item())

    def broadcast_site(self, psi, i, dist):
        raw, dims = self.site(psi, i)
        n = int(np.prod(dims))
        buf = self.new(n)
        check(self.lib.tn_memcpy_dev(self.ctx.h, C.c_void_p(buf.data_ptr()), C.c_void_p(raw.data_ptr()), n * 16))
        dist.broadcast(self.torch.view_as_real(buf), src=0)
        self.torch.cuda.synchronize()
        check(self.lib.tn_memcpy_dev(self.ctx.h, C.c_void_p(raw.data_ptr()), C.c_void_p(buf.data_ptr()), n * 16))


# ---------------------------------------------------------------------------------------------
# Distributed one-sided block Jacobi (the replicated truncated SVD is the bottleneck of the sharded sweep at large chi:
# DESIGN.md section 9).  The QR preconditioner and the finish (norms, sort, truncation rule, gathers) stay replicated; the
# Jacobi sweeps on Z = [W; V] are distributed over the column blocks:
#   * the nb column blocks (32 columns each) are grouped into 2G super-blocks; rank r works on two of them at a time;
#   * a sweep = (a) the pairs INSIDE every super-block (each rank: its two), (b) 2G-1 meetings of super-block pairs in
#     round-robin (circle) order; in a meeting the rank rotates all cross pairs of its two super-blocks in k perfect matchings
#     (k = blocks per super-block), every matching one launch of the Gram / EVD / rotation kernels over its pairs;
#   * between meetings the super-blocks whose owner changes travel point-to-point (each is one contiguous slab of Z);
#   * convergence: all_reduce(max) of the largest normalised off-diagonal Gram entry seen in the sweep;
#   * at the end every rank all-gathers the super-blocks so that Z is complete everywhere and the replicated finish runs.
# Every pair of column blocks meets exactly once per sweep, as in the single-GPU round-robin order (tools/jacobi_dist_emul.py:
# same number of sweeps +- 1).  The ``engine`` does the dense work: GpuSvdEngine (C ABI: tn_svd_dist_begin / _step / _finish) or
# the NumPy stand-in of the gloo tests.
# ---------------------------------------------------------------------------------------------
def rr_pairs(n, step):
    """circle-method round robin: the n/2 disjoint pairs of step ``step`` (n even, 0 <= step < n-1)"""
    out = []
    for k in range(n // 2):
        if n == 2:
            a, b = 0, 1
        elif k == 0:
            a, b = n - 1, step
        else:
            a, b = (step + k) % (n - 1), (step - k + n - 1) % (n - 1)
        out.append((min(a, b), max(a, b)))
    return out


def dist_jacobi_plan(nb, world):
    """Static plan for nb column blocks on ``world`` ranks: (k, nsb, meetings) with k blocks per super-block, nsb = 2*world
    super-blocks and meetings[t] = list over ranks of the super-block pair (P, Q) the rank works on in meeting t.
    nb must be a multiple of 2*world (the caller falls back to fewer ranks otherwise)."""
    nsb = 2 * world
    if nb % nsb != 0:
        raise _lib.TNError("distributed Jacobi: the number of column blocks must be a multiple of 2 * world")
    k = nb // nsb
    meetings = [rr_pairs(nsb, t) for t in range(nsb - 1)]
    return k, nsb, meetings


def usable_world(nb, world):
    """largest number of ranks g <= world with nb % (2 g) == 0 (ranks >= g idle during the sweeps)"""
    g = world
    while g > 1 and nb % (2 * g) != 0:
        g -= 1
    return g


def dist_jacobi_sweeps(engine, nb, tol, rank, world, dist, max_sweeps=60):
    """Runs the distributed sweeps on the engine's Z (already prepared by engine.begin on every rank with identical content).
    Returns the number of sweeps.  On return Z is complete and identical on every rank."""
    g = usable_world(nb, world)
    if g == 1:                                  # nothing to distribute: every rank runs the whole sweep redundantly
        sweeps = 0
        for _ in range(max_sweeps):
            off = 0.0
            for st in range(nb - 1):
                off = max(off, engine.step(rr_pairs(nb, st)))
            sweeps += 1
            if off <= tol:
                break
        return sweeps
    k, nsb, meetings = dist_jacobi_plan(nb, g)
    active = rank < g
    owner = {}                                  # super-block -> rank holding its current version
    for r, (P, Q) in enumerate(meetings[0]):
        owner[P] = r
        owner[Q] = r

    def blocks(sb):
        return range(sb * k, (sb + 1) * k)

    def move(sb, src, dst):
        """super-block ``sb`` travels src -> dst (one contiguous slab of Z)"""
        if src == dst:
            return []
        ops = []
        if rank == src:
            ops.append(("send", sb, dst))
        if rank == dst:
            ops.append(("recv", sb, src))
        return ops

    sweeps = 0
    for _ in range(max_sweeps):
        off = 0.0
        for t, meeting in enumerate(meetings):
            # who needs what for this meeting
            ops = []
            for r, (P, Q) in enumerate(meeting):
                for sb in (P, Q):
                    ops += move(sb, owner[sb], r)
                    owner[sb] = r
            engine.exchange(ops, k, dist)
            if active:
                P, Q = meeting[rank]
                if t == 0 and k > 1:            # (a) pairs inside the two super-blocks this rank holds at the start of the sweep
                    kk = k if k % 2 == 0 else k + 1
                    for st in range(kk - 1):
                        pairs = []
                        for sb in (P, Q):
                            pairs += [(sb * k + a, sb * k + b) for a, b in rr_pairs(kk, st) if a < k and b < k]
                        if pairs:
                            off = max(off, engine.step(pairs))
                for shift in range(k):          # (b) cross pairs of the meeting: k perfect matchings
                    off = max(off, engine.step([(P * k + i, Q * k + (i + shift) % k) for i in range(k)]))
        off = engine.all_reduce_max(off, dist)
        sweeps += 1
        if off <= tol:
            break
    # make Z complete everywhere: every super-block is broadcast from its last owner
    engine.gather_all([(sb, owner[sb]) for sb in range(nsb)], k, dist)
    return sweeps


class GpuSvdEngine:
    """Dense side of dist_jacobi_sweeps on one GPU: tn_svd_dist_begin / _step / _finish on the context's SVD workspace; column
    super-blocks travel through torch staging buffers (tn_memcpy_dev <-> NCCL point-to-point / broadcast)."""

    JB = 32

    def __init__(self, ctx, device):
        import torch
        self.torch, self.ctx, self.lib, self.device = torch, ctx, ctx.lib, device
        self.Z = self.ldz = self.zrows = self.nb = self.tol = None

    def begin(self, mat, m, n):
        """mat: object with data_ptr() of the m x n column-major matrix on the device.  Returns (nblocks, tol)."""
        Z, ldz, zr, nb, tol = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int32(), C.c_double()
        check(self.lib.tn_svd_dist_begin(self.ctx.h, C.c_void_p(mat.data_ptr()), int(m), int(n), C.byref(Z), C.byref(ldz), C.byref(zr),
                                         C.byref(nb), C.byref(tol)))
        self.Z, self.ldz, self.zrows, self.nb, self.tol = Z.value, ldz.value, zr.value, nb.value, tol.value
        return self.nb, self.tol

    def step(self, pairs):
        arr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1))
        off = C.c_double()
        check(self.lib.tn_svd_dist_step(self.ctx.h, arr.ctypes.data_as(C.POINTER(C.c_int32)), len(pairs), C.byref(off)))
        return off.value

    def finish(self, cutoff, maxdim, mindim, sweeps):
        from ._lib import tn_trunc_t
        k = C.c_int64()
        check(self.lib.tn_svd_dist_finish(self.ctx.h, tn_trunc_t(float(cutoff), int(maxdim), int(mindim)), int(sweeps), C.byref(k)))
        return k.value

    # -- column super-blocks: slab sb of k blocks = k*32*ldz contiguous elements of Z
    def _slab(self, sb, k):
        n = k * self.JB * self.ldz
        return self.Z + 16 * sb * n, n

    def _stage_out(self, sb, k):
        ptr, n = self._slab(sb, k)
        t = self.torch.empty(2 * n, dtype=self.torch.float64, device=self.device)
        self.torch.cuda.synchronize()
        check(self.lib.tn_memcpy_dev(self.ctx.h, C.c_void_p(t.data_ptr()), C.c_void_p(ptr), 16 * n))
        return t

    def _stage_in(self, sb, k, t):
        ptr, n = self._slab(sb, k)
        self.torch.cuda.synchronize()
        check(self.lib.tn_memcpy_dev(self.ctx.h, C.c_void_p(ptr), C.c_void_p(t.data_ptr()), 16 * n))

    def exchange(self, ops, k, dist):
        if not ops:
            return
        reqs, recvs = [], []
        for kind, sb, peer in ops:
            if kind == "send":
                reqs.append(dist.P2POp(dist.isend, self._stage_out(sb, k), peer))
            else:
                _, n = self._slab(sb, k)
                t = self.torch.empty(2 * n, dtype=self.torch.float64, device=self.device)
                reqs.append(dist.P2POp(dist.irecv, t, peer))
                recvs.append((sb, t))
        for w in dist.batch_isend_irecv(reqs):
            w.wait()
        self.torch.cuda.synchronize()
        for sb, t in recvs:
            self._stage_in(sb, k, t)

    def all_reduce_max(self, v, dist):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_all(self, owners, k, dist):
        for sb, src in owners:
            t = self._stage_out(sb, k)
            dist.broadcast(t, src=src)
            self.torch.cuda.synchronize()
            self._stage_in(sb, k, t)


def dist_svd_replacesites(be, engine, psi, theta, site, direction, normalize, cutoff, maxdim, mindim, rank, world, dist):
    """replacesites!(psi, theta, site, direction, normalize; trunc) (gmps.jl:215-266) with the Jacobi sweeps of the truncated SVD
    distributed over the ranks; every rank ends with the same two site tensors (the finish runs on identical, complete factors)."""
    _, (cl, d, _) = be.site(psi, site)
    _, (_, _, cr) = be.site(psi, site + 1)
    rows, cols = cl * d, d * cr
    be.sync()
    nb, tol = engine.begin(theta, rows, cols)
    sweeps = dist_jacobi_sweeps(engine, nb, tol, rank, world, dist)
    engine.finish(cutoff, maxdim, mindim, sweeps)
    check(be.lib.tn_mps_replacesites_factored(psi.h, int(site), int(bool(direction)), int(bool(normalize))))
    return sweeps
