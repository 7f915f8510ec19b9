"""Test-only NumPy stand-in for tnb200.sharded.GpuBackend: the same backend protocol (strided contraction semantics of
tn_contract_strided_dev, eigsolve with a caller-supplied map, replacesites, collectives) on CPU torch tensors + the oracle,
so that the sharding algebra of ShardedProjMPS / sharded_dmrg can run under gloo without a GPU.  Never used by the product."""
import numpy as np
import torch

import oracle


class StandinBackend:
    def __init__(self):
        self.ncontract = 0

    # buffers: flat torch complex128 CPU tensors
    def new(self, n):
        return torch.zeros(max(int(n), 1), dtype=torch.complex128)

    def from_host(self, flat):
        return torch.from_numpy(np.ascontiguousarray(flat, dtype=np.complex128).copy())

    def fill_ones(self, buf, n):
        buf[:n] = 1.0

    def zero(self, buf):
        buf.zero_()

    def copy(self, dst, src, n):
        dst[:n] = src[:n]

    def sync(self):
        pass

    def dotu(self, a, b, n):
        return complex(torch.sum(a[:n] * b[:n]).item())

    @staticmethod
    def _offs(n, ix):
        n0, s0, s1 = ix
        i = np.arange(n)
        return (i % n0) * s0 + (i // n0) * s1

    def contract(self, M, N, K, A, am, ak, conjA, B, bk, bn, conjB, Cc, cm, cn, alpha=1.0, beta=0.0):
        a, b, c = A.numpy(), B.numpy(), Cc.numpy()
        Am = a[self._offs(M, am)[:, None] + self._offs(K, ak)[None, :]]
        Bm = b[self._offs(K, bk)[:, None] + self._offs(N, bn)[None, :]]
        if conjA:
            Am = Am.conj()
        if conjB:
            Bm = Bm.conj()
        idx = self._offs(M, cm)[:, None] + self._offs(N, cn)[None, :]
        assert len(np.unique(idx)) == M * N
        c[idx] = alpha * (Am @ Bm) + (beta * c[idx] if beta != 0 else 0)
        self.ncontract += 1

    def eigsolve(self, apply, th0, th1, n, krylovdim, maxiter, tol):
        def heff(x):
            xin = torch.from_numpy(np.reshape(x, -1, order='F').copy())
            out = torch.zeros(n, dtype=torch.complex128)
            apply(xin, out)
            return out.numpy().reshape(x.shape, order='F')
        eig, vec, _ = oracle.eigsolve_lowest(heff, th0.numpy()[:n].copy(), krylovdim=krylovdim, maxiter=maxiter, tol=tol)
        th1[:n] = torch.from_numpy(np.reshape(vec, -1, order='F').copy())
        return eig

    # MPS handle = oracle.GMPS
    def length(self, psi):
        return len(psi)

    def movecenter(self, psi, idx):
        psi.movecenter(idx)

    def maxbonddim(self, psi):
        return psi.maxbonddim()

    def site(self, psi, i):
        t = psi[i]
        return torch.from_numpy(np.reshape(t, -1, order='F').copy()), tuple(t.shape)

    def replacesites(self, psi, theta, site, direction, normalize, cutoff, maxdim, mindim):
        cl, d, cr = psi[site].shape[0], psi.dim, psi[site + 1].shape[2]
        th = theta.numpy()[:cl * d * d * cr].reshape((cl, d, d, cr), order='F')
        psi.replacesites(th, site, direction, normalize, cutoff=cutoff, maxdim=maxdim, mindim=mindim)

    # collectives (gloo)
    def reduce_scatter(self, out, full, dist):
        dist.reduce_scatter_tensor(torch.view_as_real(out), torch.view_as_real(full), op=dist.ReduceOp.SUM)

    def all_reduce(self, buf, dist):
        dist.all_reduce(torch.view_as_real(buf), op=dist.ReduceOp.SUM)

    def all_reduce_scalar(self, v, dist):
        x = torch.tensor([complex(v).real, complex(v).imag], dtype=torch.float64)
        dist.all_reduce(x, op=dist.ReduceOp.SUM)
        return complex(x[0].item(), x[1].item())

    def broadcast_site(self, psi, i, dist):
        t = psi[i]
        buf = torch.from_numpy(np.reshape(t, -1, order='F').copy())
        dist.broadcast(torch.view_as_real(buf), src=0)
        psi[i] = buf.numpy().reshape(t.shape, order='F')
