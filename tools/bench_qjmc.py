"""QJMC trajectory throughput (BASELINE.json configs[3]): dissipative Ising chain
H = sum (x + 20 z) + 10 sum zz, jumps sqrt(0.1) s- on every site, dt = 5e-3, steps per trajectory = 20, cutoff = 0,
maxdim = chi, seeded random canonical MPS.  One process per GPU (torchrun for N > 1): trajectory t -> rank t mod G
(no collective during the evolution), inside a rank the library's worker threads / streams (tn_qjmc_ensemble).
  python tools/bench_qjmc.py --sites 64 --chi 256 --traj 16 --steps 20 --workers 16
Prints one JSON line on rank 0; time = max over ranks of the device-synchronised wall clock."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tensornetworks.jl_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist
import tnb200
from tnb200 import models


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=64)
    ap.add_argument("--chi", type=int, default=256)
    ap.add_argument("--traj", type=int, default=16, help="trajectories over ALL ranks")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--workers", type=int, default=16)
    ap.add_argument("--warm", type=int, default=1, help="warm-up steps per worker")
    ap.add_argument("--profile", default="", help="write a per-kernel time summary of the timed region (torch.profiler / CUPTI) to this file")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, chi, d = a.sites, a.chi, 2
    gamma, dt = 0.1, 5e-3
    onsite = -1j * (1.0 * models.X + 20.0 * models.Z) - 0.5 * gamma * (models.SM.conj().T @ models.SM)
    bond = -1j * 10.0 * np.kron(models.Z, models.Z)
    ss, gg = models.trotter_gates(N, onsite, bond, dt, evol="imag", order=2)
    tens = models.random_canonical_mps(N, d, chi, seed=1)
    mine = list(range(rank, a.traj, world))
    args = (tens, 1, ss, gg, list(range(1, N + 1)), [models.SM] * N, [np.sqrt(gamma)] * N)
    kw = dict(workers=a.workers, device=local, seed=0, obs_op=models.Z, cutoff=0.0, maxdim=chi)
    if a.warm:
        tnb200.qjmc_ensemble(*args, a.warm, dt, list(range(10**6, 10**6 + min(a.workers, max(1, len(mine))))), save_every=a.warm, **kw)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = tnb200.Context(local).counters()["launches"]
    t0 = time.perf_counter()
    if a.profile:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            nj, jumps, times, obs = tnb200.qjmc_ensemble(*args, a.steps, dt, mine, save_every=a.steps, **kw)
            torch.cuda.synchronize()
        with open(a.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
    else:
        nj, jumps, times, obs = tnb200.qjmc_ensemble(*args, a.steps, dt, mine, save_every=a.steps, **kw)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    launches = tnb200.Context(local).counters()["launches"] - l0
    t = torch.tensor([wall], dtype=torch.float64, device="cuda")
    stats = torch.tensor([float(nj.sum()), float(np.real(obs[:, -1, :]).sum()), float(len(mine)), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    sec = float(t.item())
    if rank == 0:
        print(json.dumps({"what": "qjmc_ensemble", "n_gpus": world, "sites": N, "chi": chi, "steps_per_traj": a.steps, "trajectories": a.traj,
                          "workers_per_gpu": a.workers, "svd_batching_rounds": os.environ.get("TN_QJMC_BATCH", "1") != "0", "seconds": sec, "traj_per_s": a.traj / sec, "traj_steps_per_s": a.traj * a.steps / sec,
                          "jumps_total": int(stats[0].item()), "mean_sum_z": float(stats[1].item() / max(1.0, stats[2].item())),
                          "gpu_launches": int(stats[3].item()), "scaling": "weak-free (independent trajectories, no data-path collective)"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
