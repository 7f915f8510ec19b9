"""GPU: the MPO-bond-sharded matvec code path (world size 1: every stage through tn_contract_strided_dev,
collectives skipped) equals the unsharded tn_env_product and the oracle to 1e-13; QJMC ensemble runner with
concurrent host threads/streams gives the same per-trajectory results as sequential runs."""
import numpy as np
import pytest

import oracle
from gpu_util import crandn, relerr

pytestmark = pytest.mark.gpu


def test_sharded_heff_world1_matches_product():
    import torch
    import tnb200
    from tnb200.sharded import ShardedHeff, GpuContractor
    rng = np.random.default_rng(0)
    chi, d, w, w1, w2 = 48, 2, 7, 6, 5
    L, R = crandn(rng, chi, w, chi), crandn(rng, chi, w2, chi)
    M1, M2 = crandn(rng, w, d, d, w1), crandn(rng, w1, d, d, w2)
    theta = crandn(rng, chi, d, d, chi)
    want = np.einsum('awb,wstx,xuvy,btvc,eyc->asue', L, M1, M2, theta, R)
    ctx = tnb200.Context.default()
    sh = ShardedHeff(L, R, M1, M2, 0, 1, GpuContractor(ctx), "cuda")
    th = torch.from_numpy(np.reshape(theta, -1, order='F').copy()).cuda()
    out = sh.apply(th).cpu().numpy().reshape(chi, d, d, chi, order='F')
    assert relerr(out, want) < 1e-13


def test_qjmc_ensemble_threads_match_sequential():
    import tnb200
    from tnb200.sharded import run_ensemble
    from models import tfim
    sh = oracle.spinhalf()
    N, dt, steps = 6, 0.02, 25
    H = tfim(N, 1.0, 2.0, 1.0)
    J = oracle.OpList(N)
    for i in range(1, N + 1):
        J.add("s-", i, np.sqrt(0.9))
    _, gl = oracle.qjmc_gates(sh, H, J, dt)
    psi0 = oracle.productMPS(sh, ["up" if i % 2 else "dn" for i in range(1, N + 1)])
    psi0.movecenter(1)

    def make_runner(shared_ctx):
        def run(t):
            ctx = shared_ctx or tnb200.Context(0)
            g = tnb200.GMPS.from_host(psi0, ctx=ctx)
            gg = tnb200.GateList.from_host(2, gl, ctx=ctx)
            jumps, times, obs = tnb200.qjmc_simulation(g, gg, list(range(1, N + 1)), [sh.op("s-")] * N, [np.sqrt(0.9)] * N, steps, dt,
                                                       seed=7, trajectory=t, obs_op=sh.op("z"), save_every=steps, cutoff=1e-10, maxdim=16)
            return jumps, np.real(obs[-1]).round(10).tolist()
        return run
    seq = run_ensemble(make_runner(tnb200.Context.default()), 6)
    par = run_ensemble(make_runner(None), 6, workers=3)
    assert seq == par
    assert len({tuple(v[0]) for v in seq.values()}) > 1     # different trajectories really differ


def test_sharded_dmrg_two_gpus_nccl_matches_single_gpu():
    """World size 2 over NCCL (one process per GPU, torchrun): the MPO-bond-sharded environments + sweep of the J1-J2 cylinder give
    the single-GPU energies to 1e-10 relative (north_star tolerance); also with the Jacobi sweeps of the SVD distributed over the ranks.
    Needs two visible GPUs (first run on 2 x B200: profiles/r02_sharded_dmrg_2gpu.jsonl)."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in ([], ["--dist-svd"]):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29541",
               os.path.join(root, "tools", "bench_multigpu.py"), "--what", "dmrg", "--lx", "4", "--ly", "4", "--chi", "64", "--sweeps", "2", "--check"] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        rec = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        assert rec["n_gpus"] == 2 and rec["single_gpu_energies"] is not None
        for a, b in zip(rec["energies"], rec["single_gpu_energies"]):
            assert abs(a - b) <= 1e-10 * abs(b), rec


@pytest.mark.parametrize("ngpus", [1, 2])
def test_cabi_sharded_heff_matches_einsum(ngpus):
    """tn_heff_sharded_* (one process, ngpus devices, NCCL inside the library): the MPO-bond-sharded matvec a ccall caller reaches,
    against the five-tensor contraction of projmps.jl:107-134; bond dimensions that do not divide evenly over the devices."""
    import torch
    from tnb200.sharded import HeffSharded
    if torch.cuda.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    rng = np.random.default_rng(3)
    chi, chi2, d, w, w1, w2 = 48, 40, 2, 7, 6, 5
    L, R = crandn(rng, chi, w, chi), crandn(rng, chi2, w2, chi2)
    M1, M2 = crandn(rng, w, d, d, w1), crandn(rng, w1, d, d, w2)
    theta = crandn(rng, chi, d, d, chi2)
    want = (0.5 - 0.25j) * np.einsum('awb,wstx,xuvy,btvc,eyc->asue', L, M1, M2, theta, R)
    sh = HeffSharded(L, R, M1, M2, list(range(ngpus)), coeff=0.5 - 0.25j)
    for _ in range(2):
        out = sh.apply(theta)
        assert relerr(out, want) < 1e-13
