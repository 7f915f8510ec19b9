"""Multi-GPU measurements for the sharded paths; launch with torchrun, one rank per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/bench_multigpu.py --what heff --chi 2048 --w 20
  ... --what dmrg --lx 12 --ly 6 --chi 1024 --sweeps 2      (MPO-bond-sharded environments + DMRG sweep, J1-J2 cylinder)
  ... --what qjmc --sites 32 --chi 64 --traj 64 --steps 5 --workers 4
Prints one JSON line on rank 0.  Times on the device (CUDA events), max over ranks."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tensornetworks.jl_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist
import tnb200
from tnb200.sharded import ShardedHeff, BalancedShardedHeff, GpuContractor, run_ensemble


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="heff")
    ap.add_argument("--chi", type=int, default=2048)
    ap.add_argument("--w", type=int, default=20)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--sites", type=int, default=32)
    ap.add_argument("--traj", type=int, default=32)
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--pipeline", type=int, default=0, help="heff: slices of Theta's right bond for apply_pipelined (0 = plain apply)")
    ap.add_argument("--balanced", action="store_true", help="heff: BalancedShardedHeff (even split of the fused (a,w) rows / (b',w2) contraction index)")
    ap.add_argument("--lx", type=int, default=6)
    ap.add_argument("--ly", type=int, default=4)
    ap.add_argument("--sweeps", type=int, default=2)
    ap.add_argument("--dist-svd", action="store_true", help="dmrg: distribute the Jacobi sweeps of the truncated SVD over the ranks")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = tnb200.Context(local)
    d = 2
    if a.what == "heff":
        rng = np.random.default_rng(0)      # same inputs on every rank
        chi, w = a.chi, a.w
        cr = lambda *s: (rng.standard_normal(s) + 1j * rng.standard_normal(s)) / np.sqrt(s[0])
        L, R = cr(chi, w, chi), cr(chi, w, chi)
        M1, M2 = cr(w, d, d, w), cr(w, d, d, w)
        theta = cr(chi, d, d, chi)
        cls = BalancedShardedHeff if a.balanced else ShardedHeff
        sh = cls(L, R, M1, M2, rank, world, GpuContractor(ctx), "cuda", dist if world > 1 else None)
        th = torch.from_numpy(np.reshape(theta, -1, order='F').copy()).cuda()
        run = (lambda: sh.apply_pipelined(th, a.pipeline, ctx.stream())) if a.pipeline > 0 else (lambda: sh.apply(th))
        for _ in range(2):
            out = run()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(a.steps):
            out = run()
        e1.record(); torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_s = e0.elapsed_time(e1) * 1e-3       # CUDA events on the device (every stage ends with a device sync, so the
        t = torch.tensor([dev_s], dtype=torch.float64, device="cuda")   # events bracket all streams' work); max over ranks below
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item()) / a.steps
        flops = 8.0 * (2.0 * chi ** 3 * d * d * w + 2.0 * chi ** 2 * d ** 3 * w * w)
        err = None
        if a.check and rank == 0:
            want = None
            if chi <= 512:
                T = np.tensordot(L, theta, axes=([2], [0]))
                Wd = np.tensordot(M1, M2, axes=([3], [0]))
                T = np.tensordot(T, Wd, axes=([1, 2, 3], [0, 2, 4]))
                want = np.tensordot(T, R, axes=([4, 1], [1, 2]))
            if want is not None:
                got = out.cpu().numpy().reshape(chi, d, d, chi, order='F')
                err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        chk = float(torch.view_as_real(out).abs().sum().item())
        if rank == 0:
            print(json.dumps({"what": "heff_mpo_bond_sharded", "n_gpus": world, "chi": chi, "w": w, "pipeline_slices": a.pipeline, "balanced": bool(a.balanced), "ms_per_matvec": sec * 1e3,
                              "tflops_total": flops / sec / 1e12, "checksum": chk, "rel_err_vs_einsum": err,
                              "collectives": "NCCL reduce_scatter(T2 over w2) + all_reduce(out)" if world > 1 else "none"}), flush=True)
    elif a.what == "dmrg":
        # Sharded DMRG sweeps on the J1-J2 cylinder (C5 shapes when --lx 12 --ly 6 --chi 4096): every rank holds its w-slice
        # of every environment block; energies must agree with the single-GPU fused sweep (--check, small sizes only).
        from tnb200.sharded import GpuBackend, GpuSvdEngine, sharded_dmrg
        from tnb200.mpo import MPO
        Nn = a.lx * a.ly
        gH = MPO(Nn, d, tnb200.models.j1j2_cylinder_terms(a.lx, a.ly), ctx=ctx)
        mpo_host = gH.tensors
        wmax = max(t.shape[3] for t in mpo_host)
        tens = tnb200.models.random_canonical_mps(Nn, d, a.chi, seed=1)
        psi = tnb200.GMPS(1, d, tens, 1, ctx=ctx)
        be = GpuBackend(ctx, "cuda")
        hist = []
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        # cutoff = 0 keeps the bond dimension at chi, so the shapes are data-independent
        sharded_dmrg(psi, mpo_host, be, rank, world, dist if world > 1 else None, cutoff=0.0, maxdim=a.chi, minsweeps=a.sweeps,
                     maxsweeps=a.sweeps, history=hist, svd_engine=GpuSvdEngine(ctx, "cuda") if a.dist_svd else None)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        t = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ref = None
        if a.check and rank == 0:
            h2 = []
            tnb200.dmrg(tnb200.GMPS(1, d, tens, 1, ctx=ctx), gH, cutoff=0.0, maxdim=a.chi, minsweeps=a.sweeps, maxsweeps=a.sweeps, history=h2)
            ref = [h[1] for h in h2]
        if rank == 0:
            print(json.dumps({"what": "dmrg_mpo_bond_sharded", "n_gpus": world, "lattice": [a.lx, a.ly], "sites": Nn, "chi": a.chi, "w_max": wmax,
                              "sweeps": a.sweeps, "distributed_svd": bool(a.dist_svd), "s_per_sweep": float(t.item()) / a.sweeps, "energies": [h[1] for h in hist],
                              "maxbond": [h[2] for h in hist], "single_gpu_energies": ref,
                              "collectives": "NCCL reduce_scatter (block updates, T2) + all_reduce (H_eff result) + broadcast (new sites)" if world > 1 else "none",
                              "mem_allocated_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    else:
        Nn, chi = a.sites, a.chi
        X, Z, I2, SM = tnb200.models.X, tnb200.models.Z, tnb200.models.I2, tnb200.models.SM
        gamma, dt = 0.1, 5e-3
        # H_eff = -iH - 1/2 sum gamma n  (qjmc.jl:9-26) with H = sum (x + 20 z) + 10 sum zz (examples/qjmc.jl couplings)
        onsite = -1j * (1.0 * X + 20.0 * Z) - 0.5 * gamma * (SM.conj().T @ SM)
        bond = -1j * 10.0 * np.kron(Z, Z)
        ss, gg = tnb200.models.trotter_gates(Nn, onsite, bond, dt, evol="imag", order=2)
        tens = tnb200.models.random_canonical_mps(Nn, d, chi, seed=1)
        import threading
        tl = threading.local()

        def run(t):
            if not hasattr(tl, "ctx"):
                tl.ctx = tnb200.Context(local)
                tl.gates = tnb200.GateList(d, ss, gg, ctx=tl.ctx)
            psi = tnb200.GMPS(1, d, tens, 1, ctx=tl.ctx)
            jumps, times, obs = tnb200.qjmc_simulation(psi, tl.gates, list(range(1, Nn + 1)), [SM] * Nn, [np.sqrt(gamma)] * Nn, a.steps, dt,
                                                       seed=0, trajectory=t, obs_op=Z, save_every=a.steps, cutoff=0.0, maxdim=chi)
            return len(jumps), float(np.real(obs[-1]).sum())
        run_ensemble(run, min(a.workers, a.traj) * world, rank, world, dist if world > 1 else None, workers=a.workers)   # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res = run_ensemble(run, a.traj, rank, world, dist if world > 1 else None, workers=a.workers)
        wall = time.perf_counter() - t0
        t = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
        if rank == 0:
            print(json.dumps({"what": "qjmc_ensemble", "n_gpus": world, "sites": Nn, "chi": chi, "steps_per_traj": a.steps, "trajectories": a.traj,
                              "workers_per_gpu": a.workers, "seconds": sec, "traj_per_s": a.traj / sec, "traj_steps_per_s": a.traj * a.steps / sec,
                              "jumps_total": int(sum(v[0] for v in res.values())), "sum_z_mean": float(np.mean([v[1] for v in res.values()]))}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
