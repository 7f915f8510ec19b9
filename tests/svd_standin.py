"""Test-only NumPy engine for tnb200.sharded.dist_jacobi_sweeps: Z = [A; I] in one array (column blocks of 32), a step = exact
diagonalisation of each listed pair's 64 x 64 Gram block (what the device's Gram / EVD / rotation launches converge to), exchange
over torch.distributed (gloo).  Never used by the product."""
import numpy as np
import torch

JB = 32


def _rr(n, step):
    ps, qs = [], []
    for k in range(n // 2):
        if k == 0:
            a, b = n - 1, step
        else:
            a, b = (step + k) % (n - 1), (step - k + n - 1) % (n - 1)
        ps.append(min(a, b))
        qs.append(max(a, b))
    return np.array(ps), np.array(qs)


def evd_jacobi(G, tol, max_sweeps=12):
    """two-sided cyclic Jacobi by plane rotations (what the device's EVD kernel does): keeps the RELATIVE accuracy of tiny columns,
    which a LAPACK eigh of the graded Gram block cannot"""
    n = G.shape[0]
    G = G.copy()
    J = np.eye(n, dtype=complex)
    for _ in range(max_sweeps):
        rotated = False
        for step in range(n - 1):
            ps, qs = _rr(n, step)
            a, b, c = G[ps, ps].real, G[qs, qs].real, G[ps, qs]
            absc = np.abs(c)
            act = (absc > tol * np.sqrt(np.abs(a * b))) & (absc > 0)
            if not act.any():
                continue
            rotated = True
            zeta = np.where(act, (b - a) / (2 * np.where(act, absc, 1)), 0)
            t = np.where(zeta >= 0, 1.0, -1.0) / (np.abs(zeta) + np.sqrt(1 + zeta ** 2))
            cs = np.where(act, 1 / np.sqrt(1 + t * t), 1.0)
            sn = np.where(act, cs * t * c / np.where(act, absc, 1), 0)
            for M in (G, J):
                x, y = M[:, ps].copy(), M[:, qs].copy()
                M[:, ps] = cs * x - y * np.conj(sn)
                M[:, qs] = x * sn + cs * y
            x, y = G[ps, :].copy(), G[qs, :].copy()
            G[ps, :] = cs[:, None] * x - sn[:, None] * y
            G[qs, :] = np.conj(sn)[:, None] * x + cs[:, None] * y
            G[ps, qs] = 0
            G[qs, ps] = 0
        if not rotated:
            break
    return J


class NumpySvdEngine:
    def __init__(self, A):
        A = np.asarray(A, dtype=np.complex128)
        m, n = A.shape
        assert m >= n
        self.m, self.n = m, n
        self.npad = ((n + 2 * JB - 1) // (2 * JB)) * 2 * JB
        self.Z = np.zeros((m + self.npad, self.npad), dtype=np.complex128, order='F')
        self.Z[:m, :n] = A
        self.Z[m:, :] = np.eye(self.npad)
        self.nb = self.npad // JB
        self.tol = 3.0 * np.sqrt(m) * 2.220446049250313e-16
        self.nsteps = 0

    def step(self, pairs):
        off = 0.0
        for p, q in pairs:
            cols = np.r_[p * JB:(p + 1) * JB, q * JB:(q + 1) * JB]
            P = self.Z[:self.m, cols]
            G = P.conj().T @ P
            G = (G + G.conj().T) / 2
            dg = np.sqrt(np.abs(np.diag(G).real))
            den = np.outer(dg, dg)
            R = np.where(den > 0, np.abs(G) / np.where(den > 0, den, 1), 0.0)
            np.fill_diagonal(R, 0)
            o = R.max()
            off = max(off, o)
            if o <= self.tol:
                continue
            J = evd_jacobi(G, self.tol)
            self.Z[:, cols] = self.Z[:, cols] @ J
        self.nsteps += 1
        return float(off)

    def _slab(self, sb, k):
        return self.Z[:, sb * k * JB:(sb + 1) * k * JB]

    def exchange(self, ops, k, dist):
        if not ops:
            return
        reqs, recvs = [], []
        for kind, sb, peer in ops:
            if kind == "send":
                t = torch.from_numpy(np.ascontiguousarray(self._slab(sb, k).reshape(-1, order='F')).view(np.float64).copy())
                reqs.append(dist.P2POp(dist.isend, t, peer, tag=sb))
            else:
                t = torch.zeros(2 * self._slab(sb, k).size, dtype=torch.float64)
                reqs.append(dist.P2POp(dist.irecv, t, peer, tag=sb))
                recvs.append((sb, t))
        for w in dist.batch_isend_irecv(reqs):
            w.wait()
        for sb, t in recvs:
            self._slab(sb, k)[...] = t.numpy().view(np.complex128).reshape(self._slab(sb, k).shape, order='F')

    def all_reduce_max(self, v, dist):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_all(self, owners, k, dist):
        for sb, src in owners:
            t = torch.from_numpy(np.ascontiguousarray(self._slab(sb, k).reshape(-1, order='F')).view(np.float64).copy())
            dist.broadcast(t, src=src)
            self._slab(sb, k)[...] = t.numpy().view(np.complex128).reshape(self._slab(sb, k).shape, order='F')

    def singular_values(self):
        return np.sort(np.linalg.norm(self.Z[:self.m, :self.n], axis=0))[::-1]
