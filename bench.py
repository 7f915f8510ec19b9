#!/usr/bin/env python
"""bench.py -- headline benchmark of the MPS hot path (BASELINE.json metric).

Workload (config.workload = "C2_heff_matvec"): the two-site effective-Hamiltonian matvec
H_eff*Theta of the Heisenberg XXZ chain MPO (w = 5, d = 2) at chi = 1024, complex FP64
(BASELINE.json configs[1]; SURVEY.md 8(d) C2 micro-benchmark: random L, R, Theta, unit-normal
re/im, seed 0).  One "step" = one H_eff application (reference: ProjMPS product,
src/structures/mps/projmps.jl:107-134) = 3 contraction launches on the GPU.

  value      H_eff matvec FP64 TFLOP/s with L, R, Theta resident in HBM (algorithmic flops of the
             flop-optimal order, F_mv = 8*(2 chi^3 d^2 w + 2 chi^2 d^3 w^2)).
  e2e        same metric through the public API call ProjMPS.product(A) with HOST buffers (pinned):
             Theta H2D and result D2H inside the timed region, environments resident (they are the
             state the reference's ProjMPS object carries between calls).
  roofline   dominant kernel tn::zgemm_sk_kernel<4,1,4,4,true> (the two chi^3 contractions), live CUDA events.
  extras     DMRG sweep wall time on the C2 chain at maxdim 64/128/256 and a bounded QJMC ensemble sample at the C4 shapes
             (the other parts of BASELINE.json's metric).
  cpu_baseline / --impl reference: the oracle's restatement of the reference's product() in the
             REFERENCE contraction order, NumPy/OpenBLAS with all host threads, bounded sample.
N > 1: the chi = 1024 matvec does not shard: N independent replicas, scaling "weak".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

if "reference" in sys.argv:
    # the reference arm is a CPU run on all host cores; torchrun exports OMP_NUM_THREADS=1, which NumPy's BLAS would obey
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count())

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "tensornetworks.jl_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

CHI, D, W, NS = 1024, 2, 5, 4
METRIC = "H_eff matvec FP64 TFLOP/s (two-site DMRG, XXZ chain MPO w=5, chi=1024, complex FP64)"


def flops_matvec(chi, d, w):
    return 8.0 * (2.0 * chi ** 3 * d * d * w + 2.0 * chi ** 2 * d ** 3 * w * w)


def make_inputs(chi, seed=0):
    """Random L, R (chi, w, chi), Theta (chi,2,2,chi) and placeholder site tensors; unit-normal re/im."""
    rng = np.random.default_rng(seed)

    def crandn(*shape):
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex128)
    dims = [1, chi, chi, chi, 1]
    sites = [np.asfortranarray(crandn(dims[i], D, dims[i + 1]) / np.sqrt(dims[i] * D)) for i in range(NS)]
    L = np.asfortranarray(crandn(chi, W, chi))
    R = np.asfortranarray(crandn(chi, W, chi))
    theta = np.asfortranarray(crandn(chi, D, D, chi))
    return sites, L, R, theta


def cpu_reference_arm(chi, budget_s=20.0, max_calls=6):
    """The reference's CPU path for this workload: oracle ProjMPS.product (reference contraction order,
    projmps.jl:119-134) on NumPy/OpenBLAS with all host threads.  Returns (tflops, calls, seconds, cores)."""
    import oracle
    from oracle.gmps import GMPS as OG
    import tnb200.models as models
    sites, L, R, theta = make_inputs(chi)
    psi = OG(1, D, sites, 0)
    H = OG(2, D, models.xxz_mpo(NS), 0)
    P = oracle.ProjMPS.__new__(oracle.ProjMPS)
    P.objects = [psi, H, psi]
    P.blocks = [None, None, None, None]
    P.blocks[0], P.blocks[3] = L, R
    P.squared, P.rank, P.center, P.coeff = False, 2, 2, 1.0
    cores = os.cpu_count()
    try:                                   # make sure the BLAS pool really uses every host core (and report what it uses)
        from threadpoolctl import threadpool_limits, threadpool_info
        threadpool_limits(limits=cores)
        used = [p.get("num_threads") for p in threadpool_info() if p.get("user_api") == "blas"]
        if used:
            cores = max(used)
    except Exception:
        pass
    t_tot, calls = 0.0, 0
    out = None
    while calls < max_calls and (calls == 0 or t_tot + t_tot / calls < budget_s):
        t0 = time.perf_counter()
        out = P.product(theta, False, 2)
        t_tot += time.perf_counter() - t0
        calls += 1
    return flops_matvec(chi, D, W) * calls / t_tot / 1e12, calls, t_tot, cores, out


def dmrg_sweep_sample(ctx, N=100, chis=(64, 128, 256)):
    """The other half of BASELINE.json's metric: wall time of full two-site DMRG sweeps (tn_dmrg_sweep: environments,
    Lanczos with <=5 H_eff applications per bond, truncated Jacobi SVD, all on the device) on the C2 chain
    (XXZ N=100, w=5, cutoff=1e-12), ramping maxdim; 2 sweeps per maxdim, the second one is reported."""
    import tnb200
    from tnb200._lib import check, tn_lanczos_t
    rng = np.random.default_rng(1234)
    dims = [1] + [min(2 ** min(i, N - i), 8) for i in range(1, N)] + [1]
    tens = [rng.standard_normal((dims[i], D, dims[i + 1])) for i in range(N)]
    g = tnb200.GMPS(1, D, tens, 0, ctx=ctx)
    gH = tnb200.GMPS(2, D, tnb200.models.xxz_mpo(N, 1.0), ctx=ctx)
    g.movecenter(1)
    Hs = tnb200.ProjMPS(g, gH, g, center=1)
    out, direction = [], False
    for chi in chis:
        for rep in range(2):
            e, mb = C.c_double(), C.c_int64()
            c0 = ctx.counters()
            t0 = time.perf_counter()
            check(ctx.lib.tn_dmrg_sweep(g.h, Hs.h, int(direction), tn_lanczos_t(3, 2, 1e-14), tnb200.Trunc(1e-12, chi, 1), C.byref(e), C.byref(mb)))
            dt = time.perf_counter() - t0
            c1 = ctx.counters()
            direction = not direction
        out.append({"maxdim": chi, "maxbond": mb.value, "seconds_per_sweep": dt, "energy": e.value, "gpu_launches": c1["launches"] - c0["launches"],
                    "heff_applications": c1["matvecs"] - c0["matvecs"], "svds": c1["svds"] - c0["svds"]})
    return {"config": "C2 XXZ chain N=100 w=5 two-site DMRG, cutoff=1e-12, random chi=8 start, 2 sweeps per maxdim (2nd timed)", "sweeps": out}


def qjmc_sample(device, N=64, chi=256, traj=16, workers=16, steps=1):
    """Third part of BASELINE.json's metric: QJMC trajectory throughput at the C4 shapes (dissipative Ising chain N=64, chi=256,
    cutoff=0, seeded random canonical start; tools/bench_qjmc.py is the full tool).  A bounded sample: ``traj`` trajectories x
    ``steps`` steps on ``workers`` worker streams after one warm-up step per worker, wall clock around tn_qjmc_ensemble."""
    import tnb200
    from tnb200 import models
    gamma, dt = 0.1, 5e-3
    onsite = -1j * (1.0 * models.X + 20.0 * models.Z) - 0.5 * gamma * (models.SM.conj().T @ models.SM)
    bond = -1j * 10.0 * np.kron(models.Z, models.Z)
    ss, gg = models.trotter_gates(N, onsite, bond, dt, evol="imag", order=2)
    tens = models.random_canonical_mps(N, D, chi, seed=1)
    args = (tens, 1, ss, gg, list(range(1, N + 1)), [models.SM] * N, [np.sqrt(gamma)] * N)
    kw = dict(workers=workers, device=device, seed=0, obs_op=models.Z, cutoff=0.0, maxdim=chi)
    tnb200.qjmc_ensemble(*args, 1, dt, list(range(10 ** 6, 10 ** 6 + workers)), save_every=1, **kw)      # warm-up
    t0 = time.perf_counter()
    nj, _, _, obs = tnb200.qjmc_ensemble(*args, steps, dt, list(range(traj)), save_every=steps, **kw)
    sec = time.perf_counter() - t0
    return {"config": f"C4 shapes: N={N}, chi={chi}, cutoff=0, {traj} trajectories x {steps} step(s), {workers} worker streams on one GPU",
            "seconds": sec, "traj_steps_per_s": traj * steps / sec, "traj_per_s_at_20_steps": traj * steps / sec / 20.0,
            "jumps": int(nj.sum()), "mean_sum_z": float(np.real(obs[:, -1, :]).sum() / traj)}


class ClockSampler:
    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [l.strip().split(",") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tf, calls, secs, cores, _ = cpu_reference_arm(CHI, budget_s=min(120.0, 8.0 * max(1, args.steps)), max_calls=max(1, args.steps + args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": calls, "warmup": 0,
        "ms_per_step": secs / calls * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128 (complex f64)",
        "data": "synthetic", "config": {"workload": "C2_heff_matvec", "chi": CHI, "d": D, "w": W, "order": "reference (L.M1.M2 first)"},
        "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "port",
                         "sample": f"{calls} matvec(s) at chi={CHI} in the reference contraction order, NumPy/OpenBLAS, {cores} threads"},
        "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--chi", type=int, default=CHI)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the DMRG sweep-time sample")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import tnb200
    from tnb200 import _lib
    from tnb200.api import GMPS, ProjMPS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    chi = args.chi
    W_eff = max(3, args.warmup)

    ctx = tnb200.Context(local)
    lib = ctx.lib
    sites, L, R, theta = make_inputs(chi)
    psi = GMPS(1, D, sites, 0, ctx=ctx)
    psi.center = 2
    H = GMPS(2, D, tnb200.models.xxz_mpo(NS), ctx=ctx)
    env = ProjMPS(psi, H, psi, center=2)
    env.setblock(1, L)
    env.setblock(4, R)
    n = chi * D * D * chi
    th_dev = torch.from_numpy(theta.reshape(-1, order='F').view(np.float64).copy()).cuda()
    out_dev = torch.empty_like(th_dev)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    flops = flops_matvec(chi, D, W)

    def barrier():
        if world > 1:
            dist.barrier()
        ctx.sync()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------
    _lib.check(lib.tn_env_product_dev(env.h, C.c_void_p(th_dev.data_ptr()), 0, C.c_void_p(out_dev.data_ptr()), W_eff))
    barrier()
    c0 = ctx.counters()["launches"]
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    _lib.check(lib.tn_env_product_dev(env.h, C.c_void_p(th_dev.data_ptr()), 0, C.c_void_p(out_dev.data_ptr()), args.steps))
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    launches = ctx.counters()["launches"] - c0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * flops * args.steps / (ms_max * 1e-3) / 1e12

    # ---- roofline of the dominant kernel (live CUDA events around each contraction stage) ---------
    st = (C.c_double * 3)()
    _lib.check(lib.tn_env_product_profile(env.h, C.c_void_p(th_dev.data_ptr()), 0, C.c_void_p(out_dev.data_ptr()), args.steps, st))
    stage_ms = [st[i] / args.steps for i in range(3)]
    big_flops = 8.0 * chi ** 3 * D * D * W            # per chi^3 contraction (2 launches per matvec)
    achieved = 2 * big_flops / ((stage_ms[0] + stage_ms[2]) * 1e-3) / 1e12

    # calibrate the FP64 ceiling live: cuBLAS ZGEMM through torch.matmul (MEASURED_PEAKS.json has no FP64 entry)
    a = torch.randn(4096, 4096, dtype=torch.complex128, device="cuda")
    b = torch.randn(4096, 4096, dtype=torch.complex128, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record(); torch.matmul(a, b); x1.record(); torch.cuda.synchronize()
        best = min(best, x0.elapsed_time(x1))
    zgemm_peak = 8.0 * 4096 ** 3 / (best * 1e-3) / 1e12
    del a, b

    # ---- end to end through the public API with pinned host buffers -------------------------------
    hin = torch.empty(2 * n, dtype=torch.float64).pin_memory()
    hout = torch.empty(2 * n, dtype=torch.float64).pin_memory()
    hin.numpy()[:] = theta.reshape(-1, order='F').view(np.float64)
    A_host = np.ndarray((chi, D, D, chi), dtype=np.complex128, buffer=hin.numpy().data, order='F')
    O_host = np.ndarray((chi, D, D, chi), dtype=np.complex128, buffer=hout.numpy().data, order='F')
    for _ in range(W_eff):
        env.product(A_host, False, out=O_host)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        env.product(A_host, False, out=O_host)
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * flops * args.steps / (float(t.item()) * 1e-3) / 1e12
    # parity spot check of this very run's output against the resident-path output
    dev_out = out_dev.cpu().numpy().view(np.complex128)
    e2e_out = O_host.reshape(-1, order='F')
    same = float(np.linalg.norm(dev_out - e2e_out) / np.linalg.norm(dev_out))

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": W_eff,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c128 (complex f64)", "data": "synthetic",
            "config": {"workload": "C2_heff_matvec", "chi": chi, "d": D, "w": W, "mpo": "XXZ chain (Heisenberg, delta=1)",
                       "order": "flop-optimal (L.Theta).W.R", "multi_gpu": "independent replicas (chi=1024 matvec does not shard)",
                       "cache": "inputs+intermediates (L,R 84 MB each, T1/T2 335 MB each) exceed the 126 MB L2"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": "TFLOP/s", "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 16 * n,
                    "api": "tnb200.ProjMPS.product(A_host) -> tn_env_product", "matches_resident_path_rel": same},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": zgemm_peak, "unit": "TFLOP/s", "frac": achieved / zgemm_peak,
                         "traffic": None, "kernel": "tn::zgemm_sk_kernel<4,1,4,4,true> (persistent stream-K, 128x32 tile, 2 CTAs/SM, DMMA.8x8x4), 2 launches per matvec",
                         "flops_per_launch": big_flops, "ms_per_launch": (stage_ms[0] + stage_ms[2]) / 2,
                         "peak_source": "cuBLAS ZGEMM 4096^3 via torch.matmul measured in this run (MEASURED_PEAKS.json has no FP64 figure; "
                                        "DMMA issue peak measured 37.17 TFLOP/s, profiles/r01_probe_fp64.jsonl)",
                         "stage_ms": {"L.Theta": stage_ms[0], ".W": stage_ms[1], ".R": stage_ms[2]}},
        }
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch, same shapes)
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "zgemm_traffic.json")))
            if tr.get("chi") == chi:
                line["roofline"]["traffic"] = tr["dram_bytes_per_launch"]
                line["roofline"]["traffic_source"] = tr["source"]
                line["roofline"]["algorithmic_bytes_per_launch"] = tr.get("algorithmic_bytes_per_launch")
        except Exception:
            pass
        if world == 1 and not args.no_extras:
            line["extras"] = {"dmrg_sweep": dmrg_sweep_sample(ctx)}
            try:
                line["extras"]["qjmc"] = qjmc_sample(local)
            except Exception as e:          # the sample is a by-product: never lose the bench line over it
                line["extras"]["qjmc"] = {"error": repr(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            tf, calls, secs, cores, ref_out = cpu_reference_arm(chi, budget_s=20.0)
            err = float(np.linalg.norm(ref_out.reshape(-1, order='F') - dev_out) / np.linalg.norm(dev_out))
            line["cpu_baseline"] = {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "port",
                                    "sample": f"{calls} matvec(s) at chi={chi}, reference contraction order, NumPy/OpenBLAS {cores} threads, {secs:.1f} s",
                                    "gpu_vs_oracle_rel_err": err}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
