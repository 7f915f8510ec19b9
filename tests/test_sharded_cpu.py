"""Multi-process host logic of the sharded paths on CPU (gloo, world_size 2): the MPO-bond-sharded matvec's
slicing / chunk layout / reduce-scatter + all-reduce algebra (with a NumPy stand-in for the GPU contraction
kernel, test-only) equals the oracle's unsharded matvec; QJMC trajectory ownership + gather."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyContractor:
    """Test-only stand-in with the same strided semantics as tn_contract_strided_dev."""

    @staticmethod
    def _offs(n, ix):
        n0, s0, s1 = ix
        i = np.arange(n)
        return (i % n0) * s0 + (i // n0) * s1 if n0 < n else i * s0

    def __call__(self, M, N, K, A, am, ak, B, bk, bn, Cc, cm, cn, beta=0.0):
        a, b, c = A.numpy(), B.numpy(), Cc.numpy()
        Am = a[self._offs(M, am)[:, None] + self._offs(K, ak)[None, :]]
        Bm = b[self._offs(K, bk)[:, None] + self._offs(N, bn)[None, :]]
        idx = self._offs(M, cm)[:, None] + self._offs(N, cn)[None, :]
        c[idx] = Am @ Bm + beta * c[idx]

    def sync(self):
        pass


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tensornetworks.jl_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tnb200.sharded import ShardedHeff, run_ensemble
    import oracle
    rng = np.random.default_rng(0)

    def crandn(*s):
        return rng.standard_normal(s) + 1j * rng.standard_normal(s)
    chi, d, w, w1, w2 = 6, 2, 5, 4, 3          # w2 = 3 is not divisible by world = 2: exercises the padding
    L, R = crandn(chi, w, chi + 1), crandn(chi + 2, w2, chi)
    M1, M2 = crandn(w, d, d, w1), crandn(w1, d, d, w2)
    theta = crandn(chi + 1, d, d, chi)
    want = np.einsum('awb,wstx,xuvy,btvc,eyc->asue', L, M1, M2, theta, R)
    sh = ShardedHeff(L, R, M1, M2, rank, world, NumpyContractor(), "cpu", dist)
    out = sh.apply(torch.from_numpy(np.reshape(theta, -1, order='F').copy()))
    got = out.numpy().reshape(chi, d, d, chi + 2, order='F')
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    res = run_ensemble(lambda t: (t, t * t), 7, rank, world, dist)
    q.put((rank, err, sorted(res.items())))
    dist.destroy_process_group()


def test_sharded_matvec_and_ensemble_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, err, res in outs:
        assert err < 1e-13, (rank, err)
        assert res == [(t, (t, t * t)) for t in range(7)]


def test_shard_helpers():
    sys.path.insert(0, os.path.join(ROOT, "tensornetworks.jl_b200"))
    from tnb200.sharded import shard_range, my_trajectories
    assert [shard_range(20, r, 8) for r in range(8)] == [(0, 3), (3, 6), (6, 9), (9, 12), (12, 14), (14, 16), (16, 18), (18, 20)]
    assert sorted(sum([my_trajectories(10, r, 4) for r in range(4)], [])) == list(range(10))
