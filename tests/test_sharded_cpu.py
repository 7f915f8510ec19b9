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
    for ns in (2, 3, 6):        # pipelined variant: Theta's right bond in slices, stage 3 accumulating slice by slice
        out = sh.apply_pipelined(torch.from_numpy(np.reshape(theta, -1, order='F').copy()), nslices=ns)
        got = out.numpy().reshape(chi, d, d, chi + 2, order='F')
        err = max(err, np.linalg.norm(got - want) / np.linalg.norm(want))
    from tnb200.sharded import BalancedShardedHeff
    bal = BalancedShardedHeff(L, R, M1, M2, rank, world, NumpyContractor(), "cpu", dist)      # rows of (a,w) / flat (b',w2) chunks
    got = bal.apply(torch.from_numpy(np.reshape(theta, -1, order='F').copy())).numpy().reshape(chi, d, d, chi + 2, order='F')
    err = max(err, np.linalg.norm(got - want) / np.linalg.norm(want))
    for ns in (2, 5):                                  # balanced + pipelined
        got = bal.apply_pipelined(torch.from_numpy(np.reshape(theta, -1, order='F').copy()), nslices=ns).numpy().reshape(chi, d, d, chi + 2, order='F')
        err = max(err, np.linalg.norm(got - want) / np.linalg.norm(want))
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


# ---------------------------------------------------------------------------------------------
# MPO-bond-sharded environments + full DMRG sweep (ShardedProjMPS / sharded_dmrg) under gloo
# ---------------------------------------------------------------------------------------------
def _dmrg_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tensornetworks.jl_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from gpu_util import random_complex_mps, random_mpo
    from models import xxz
    from sharded_standin import StandinBackend
    from tnb200.sharded import ShardedProjMPS, sharded_dmrg
    be = StandinBackend()
    # (i) blocks / calculate / product of a random complex MPO with w = 5 (chunks 3+2 at world 2, 2+2+1 at world 3; the edge
    #     bonds w = 1 leave every rank but 0 empty)
    rng = np.random.default_rng(0)
    N = 6
    psi = random_complex_mps(rng, N, 2, 6, center=1)
    H = random_mpo(rng, N, 2, 5)
    S = ShardedProjMPS(psi, [H[i] for i in range(1, N + 1)], be, rank, world, dist, center=3, coeff=0.7 - 0.1j)
    P = oracle.ProjMPS([psi, H, psi], rank=2, center=3, coeff=0.7 - 0.1j)
    errs = []
    for i in (1, 2, 4, 5, 6):
        buf, dims = S.block(i)
        full = P.block(i)
        c = (full.shape[1] + world - 1) // world
        lo, hi = min(rank * c, full.shape[1]), min((rank + 1) * c, full.shape[1])
        assert dims == (full.shape[0], hi - lo, full.shape[2])
        if hi > lo:
            got = buf.numpy()[:int(np.prod(dims))].reshape(dims, order='F')
            errs.append(np.linalg.norm(got - full[:, lo:hi, :]) / np.linalg.norm(full[:, lo:hi, :]))
    cal = abs(S.calculate() - P.calculate()) / abs(P.calculate())
    th = rng.standard_normal((psi[3].shape[0], 2, 2, psi[4].shape[2])) + 1j * rng.standard_normal((psi[3].shape[0], 2, 2, psi[4].shape[2]))
    S.prepare(3)
    out = torch.zeros(th.size, dtype=torch.complex128)
    S.product(torch.from_numpy(th.reshape(-1, order='F').copy()), out)
    want = P.product(th, False, 2)
    perr = np.linalg.norm(out.numpy().reshape(th.shape, order='F') - want) / np.linalg.norm(want)
    # (ii) full sweeps on the Heisenberg chain (w = 5) against the oracle's unsharded dmrg
    sh = oracle.spinhalf()
    M = oracle.MPO(sh, xxz(8, 1.0))
    p0 = oracle.randomMPS(2, 8, 4, np.random.default_rng(1))
    ho, hs = [], []
    oracle.dmrg(p0.copy(), M, maxdim=16, maxsweeps=4, history=ho)
    ps, _ = sharded_dmrg(p0.copy(), [M[i] for i in range(1, 9)], be, rank, world, dist, maxdim=16, maxsweeps=4, history=hs)
    q.put((rank, max(errs, default=0.0), cal, perr, ho, hs, [t.tobytes() for t in ps.tensors]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_environments_and_dmrg_sweep(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_dmrg_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    states = []
    for rank, berr, cal, perr, ho, hs, tens in outs:
        assert berr < 1e-13 and cal < 1e-12 and perr < 1e-13, (rank, berr, cal, perr)
        assert len(ho) == len(hs)
        for a, b in zip(ho, hs):
            assert a[2] == b[2] and abs(a[1] - b[1]) < 1e-10 * abs(a[1]), (rank, a, b)
        states.append(tens)
    # the replicas hold bitwise identical site tensors (rank 0's are broadcast after every bond)
    assert all(s == states[0] for s in states[1:])


def test_chunk_range():
    sys.path.insert(0, os.path.join(ROOT, "tensornetworks.jl_b200"))
    from tnb200.sharded import chunk_range
    assert [chunk_range(20, r, 8) for r in range(8)] == [(0, 3), (3, 6), (6, 9), (9, 12), (12, 15), (15, 18), (18, 20), (20, 20)]
    assert [chunk_range(1, r, 2) for r in range(2)] == [(0, 1), (1, 1)]


# ---------------------------------------------------------------------------------------------
# Distributed one-sided block Jacobi (dist_jacobi_sweeps) under gloo with the NumPy engine
# ---------------------------------------------------------------------------------------------
def _svd_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tensornetworks.jl_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from svd_standin import NumpySvdEngine
    from tnb200.sharded import dist_jacobi_sweeps
    from threadpoolctl import threadpool_limits
    threadpool_limits(1)                                     # tiny matrices: BLAS threads of several ranks only fight each other
    out = []
    for (m, n, kind) in ((128, 128, "randn"), (200, 192, "graded")):      # 4 and 6 column blocks (6 = 2 * 3: all three ranks busy at world 3)
        rng = np.random.default_rng(m + n)                  # same matrix on every rank
        A = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
        if kind == "graded":
            u, s, vh = np.linalg.svd(A, full_matrices=False)
            A = (u * np.exp(-np.arange(n) * 20.0 / n)) @ vh
        eng = NumpySvdEngine(A)
        sweeps = dist_jacobi_sweeps(eng, eng.nb, eng.tol, rank, world, dist)
        so = np.linalg.svd(A, compute_uv=False)
        s = eng.singular_values()
        V = eng.Z[eng.m:, :]
        rec = np.linalg.norm(A @ V[:n, :n] - eng.Z[:m, :n]) / np.linalg.norm(A)        # W = A V on the real columns
        out.append((m, n, sweeps, eng.nsteps, float(np.max(np.abs(s - so)) / so[0]), float(np.linalg.norm(V.conj().T @ V - np.eye(eng.npad))), float(rec),
                    eng.Z.tobytes()))
    q.put((rank, out))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_jacobi_sweeps(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_svd_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, res in outs.items():
        for (m, n, sweeps, nsteps, sverr, orth, rec, zb), ref in zip(res, outs[0]):
            assert sweeps < 40 and sverr < 1e-12 and orth < 1e-10 and rec < 1e-12, (rank, m, n, sweeps, sverr, orth, rec)
            assert zb == ref[7], "Z differs between ranks after the final gather"
    # the work really is distributed: 128 x 128 has nb = 4 blocks; at G = 2, k = 1: a sweep is 3 meetings of one pair per rank
    m0 = outs[0][0]
    if world == 2:
        assert m0[3] == m0[2] * 3


def test_distributed_jacobi_schedule_visits_every_block_pair_once_per_sweep():
    """Pure schedule check (no arithmetic, no communication): over the ranks, one sweep of dist_jacobi_sweeps rotates every pair of
    column blocks exactly once, a step never touches a block twice, and the moves keep every super-block with exactly one owner."""
    sys.path.insert(0, os.path.join(ROOT, "tensornetworks.jl_b200"))
    from tnb200.sharded import dist_jacobi_sweeps, usable_world

    class Recorder:
        def __init__(self):
            self.steps, self.sends, self.recvs = [], [], []

        def step(self, pairs):
            flat = [b for p in pairs for b in p]
            assert len(flat) == len(set(flat)), "a step touches a block twice"
            self.steps.append(list(pairs))
            return 1.0                       # never converged: exactly max_sweeps sweeps

        def exchange(self, ops, k, dist):
            for kind, sb, peer in ops:
                (self.sends if kind == "send" else self.recvs).append((sb, peer))

        def all_reduce_max(self, v, dist):
            return v

        def gather_all(self, owners, k, dist):
            assert sorted(sb for sb, _ in owners) == list(range(len(owners)))

    for nb, world in ((8, 2), (12, 3), (16, 4), (16, 2), (8, 3), (6, 3), (4, 8)):
        recs = []
        for rank in range(world):
            r = Recorder()
            assert dist_jacobi_sweeps(r, nb, 0.0, rank, world, None, max_sweeps=1) == 1
            recs.append(r)
        g = usable_world(nb, world)
        seen = {}
        ranks = range(g) if g > 1 else range(1)          # g == 1: every rank runs the whole (identical) sweep
        for rank in ranks:
            for st in recs[rank].steps:
                for p, q in st:
                    key = (min(p, q), max(p, q))
                    seen[key] = seen.get(key, 0) + 1
        assert len(seen) == nb * (nb - 1) // 2 and set(seen.values()) == {1}, (nb, world, len(seen))
        # every send has its matching receive
        sends = sorted((sb, src, dst) for src, r in enumerate(recs) for sb, dst in r.sends)
        recvs = sorted((sb, src, dst) for dst, r in enumerate(recs) for sb, src in r.recvs)
        assert sends == recvs
