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
