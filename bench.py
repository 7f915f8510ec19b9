#!/usr/bin/env python
"""bench.py -- headline benchmark of the MPS hot path (BASELINE.json metric).

Workload (config.workload = "C5_heff_matvec"): the two-site effective-Hamiltonian matvec H_eff*Theta of the J1-J2
Heisenberg model on a width-6 cylinder (BASELINE.json configs[4]; bulk MPO bond w = 20, d = 2) at chi = 2048, complex FP64,
random L, R, Theta (unit-normal re/im, seed 0; SURVEY.md 8(d) micro-benchmark).  One "step" = one H_eff application
(reference: ProjMPS product, src/structures/mps/projmps.jl:107-134).  This is the path north_star shards: the same
workload runs at every N,
  N = 1   tn_env_product (3 contraction launches on one GPU),
  N > 1   MPO-bond-sharded over the N ranks (tnb200.sharded.BalancedShardedHeff: local chi^3 stage, NCCL reduce_scatter of
          the partial T2 over (b', w2) pipelined under the next slice's GEMMs, local chi^3 stage, NCCL all_reduce of the
          result), one process per GPU -> "scaling": "strong".

  value      H_eff matvec FP64 TFLOP/s, inputs resident in HBM (algorithmic flops of the flop-optimal order,
             F_mv = 8*(2 chi^3 d^2 w + 2 chi^2 d^3 w^2)); CUDA events, max over ranks.
  e2e        same metric through the public API with HOST buffers (pinned): Theta H2D and result D2H inside the timed
             region, environments resident (they are the state the reference's ProjMPS object carries between calls).
  roofline   dominant kernel tn::zgemm_sk_kernel (the two chi^3 contractions), live CUDA events, against the cuBLAS ZGEMM
             rate measured in the same run.
  extras     (N = 1) the C2 matvec (chi = 1024, w = 5: last round's headline), DMRG sweep wall times on the C2 chain,
             the truncated SVD (F_svd normaliser, next to cuSOLVER on the same box), a bounded QJMC wave at the C4 shapes;
             (every N) extras.qjmc_scaling: one wave of trajectories per GPU of the 8192-trajectory C4 job.
  cpu_baseline / --impl reference: the oracle's restatement of the reference's product() in the REFERENCE contraction
             order, NumPy/OpenBLAS with all host threads, on a bounded sample of the same matvec: the bra-bond rows
             a in [0, ROWS) of the result (every contraction of the reference order is linear in that slice, so the
             sample costs ROWS/chi of a full application and is normalised by ROWS/chi of F_mv).
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

CHI, D, W = 2048, 2, 20
T_START = time.perf_counter()
LX, LY, SITE = 4, 6, 8          # bulk sites 8, 9 of the 4 x 6 cylinder carry the full w = 20 bond on both sides
REF_ROWS = 64                   # bra-bond rows of the bounded CPU sample
METRIC = "H_eff matvec FP64 TFLOP/s (two-site DMRG, J1-J2 width-6 cylinder MPO w=20, chi=2048, complex FP64)"


def flops_matvec(chi, d, w):
    return 8.0 * (2.0 * chi ** 3 * d * d * w + 2.0 * chi ** 2 * d ** 3 * w * w)


def config(chi, w):
    """Identical in both arms (the driver compares the two lines' config)."""
    return {"workload": "C5_heff_matvec", "chi": chi, "d": D, "w": w,
            "mpo": "J1-J2 Heisenberg (J2=0.5) on a width-6 cylinder, two bulk sites",
            "inputs": "random L, R (chi,w,chi) and Theta (chi,2,2,chi), unit-normal re/im, seed 0",
            "cache": "inputs+intermediates (L,R 1.3 GB each, T1/T2 5.4 GB each at chi=2048) exceed the 126 MB L2"}


def make_inputs(chi, w, seed=0, rows=None):
    """Random L, R (chi, w, chi) and Theta (chi,2,2,chi); unit-normal re/im.  ``rows``: keep only L[:rows] (CPU sample)."""
    rng = np.random.default_rng(seed)

    def crandn(*shape):
        x = np.empty(shape, dtype=np.complex128, order='F')
        v = x.reshape(-1, order='F').view(np.float64)
        step = 1 << 24
        for i in range(0, v.size, step):
            v[i:i + step] = rng.standard_normal(min(step, v.size - i))
        return x
    L = crandn(chi, w, chi)
    R = crandn(chi, w, chi)
    theta = crandn(chi, D, D, chi)
    if rows is not None:
        L = np.asfortranarray(L[:rows])
    return L, R, theta


def blas_threads():
    cores = os.cpu_count()
    try:                                   # make sure the BLAS pool really uses every host core (and report what it uses)
        from threadpoolctl import threadpool_limits, threadpool_info
        threadpool_limits(limits=cores)
        used = [p.get("num_threads") for p in threadpool_info() if p.get("user_api") == "blas"]
        if used:
            cores = max(used)
    except Exception:
        pass
    return cores


def oracle_mpo_pair():
    """The reference arm's own MPO: oracle.MPO(spinhalf, H) (restates mpo.jl:323-459) on the 4 x 6 cylinder, bulk sites."""
    import oracle
    from models import j1j2_cylinder
    H = oracle.MPO(oracle.spinhalf(), j1j2_cylinder(LX, LY))
    return np.asfortranarray(H[SITE]), np.asfortranarray(H[SITE + 1])


def julia_reference_sample(chi, w, M1, M2, steps, warmup, rows):
    """The reference itself, when it can run: `julia` on the PATH and TN_REFERENCE_JL = a checkout of lcauser/TensorNetworks.jl with its
    dependencies instantiated (neither exists in the build image or on the GPU boxes of this pool).  Returns the parsed JSON line of
    baseline/run_reference.jl or None."""
    import shutil
    proj = os.environ.get("TN_REFERENCE_JL", "")
    if not shutil.which("julia") or not os.path.isdir(proj):
        return None
    try:
        with tempfile.TemporaryDirectory() as td:
            f1, f2 = os.path.join(td, "M1.bin"), os.path.join(td, "M2.bin")
            open(f1, "wb").write(np.asfortranarray(M1).tobytes(order="F"))
            open(f2, "wb").write(np.asfortranarray(M2).tobytes(order="F"))
            r = subprocess.run(["julia", "-t1", f"--project={proj}", os.path.join(ROOT, "baseline", "run_reference.jl"), str(chi), str(w), str(rows),
                                str(steps), str(warmup), f1, f2], capture_output=True, text=True, timeout=1800)
        return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    except Exception:
        return None


def cpu_reference_sample(chi, w, M1, M2, steps, warmup, rows=REF_ROWS, budget_s=None):
    """The reference's CPU path on the bounded sample: oracle ProjMPS.product (reference contraction order, projmps.jl:119-134)
    with the left block restricted to ``rows`` bra-bond rows.  Returns (tflops, timed steps, seconds, cores, out[:rows])."""
    import oracle
    from oracle.gmps import GMPS as OG
    rows = min(rows, chi)
    L, R, theta = make_inputs(chi, w, rows=rows)
    dims = [1, 1, 1, 1, 1]
    psi = OG(1, D, [np.zeros((dims[i], D, dims[i + 1]), dtype=np.complex128) for i in range(4)], 0)   # placeholders: product() reads blocks, MPO, Theta only
    H = OG(2, D, [M1[:1], M1, M2, M2[..., :1]], 0)
    P = oracle.ProjMPS.__new__(oracle.ProjMPS)
    P.objects = [psi, H, psi]
    P.blocks = [L, None, None, R]
    P.squared, P.rank, P.center, P.coeff = False, 2, 2, 1.0
    cores = blas_threads()
    out = None
    for _ in range(warmup):
        out = P.product(theta, False, 2)
    t_tot, calls = 0.0, 0
    while calls < steps and (budget_s is None or calls == 0 or t_tot + t_tot / calls < budget_s):
        t0 = time.perf_counter()
        out = P.product(theta, False, 2)
        t_tot += time.perf_counter() - t0
        calls += 1
    return flops_matvec(chi, D, w) * (rows / chi) * calls / t_tot / 1e12, calls, t_tot, cores, out


# ------------------------------------------------------------------------------------------------------------------------
# extras (N = 1): the other parts of BASELINE.json's metric
# ------------------------------------------------------------------------------------------------------------------------
def dmrg_sweep_sample(ctx, N=100, chis=(64, 128, 256)):
    """Wall time of full two-site DMRG sweeps (tn_dmrg_sweep: environments, Lanczos with <=5 H_eff applications per bond,
    truncated Jacobi SVD, all on the device) on the C2 chain (XXZ N=100, w=5, cutoff=1e-12), ramping maxdim; 2 sweeps per
    maxdim, the second one is reported."""
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


def c2_matvec_sample(ctx, torch, steps=10):
    """Last round's headline (C2: XXZ chain w = 5, chi = 1024) for continuity."""
    import tnb200
    from tnb200 import _lib
    from tnb200.api import GMPS, ProjMPS
    chi, w = 1024, 5
    L, R, theta = make_inputs(chi, w)
    dims = [1, chi, chi, chi, 1]
    rng = np.random.default_rng(1)
    sites = [np.asfortranarray(rng.standard_normal((dims[i], D, dims[i + 1])) + 0j) for i in range(4)]
    psi = GMPS(1, D, sites, 0, ctx=ctx)
    psi.center = 2
    H = GMPS(2, D, tnb200.models.xxz_mpo(4), ctx=ctx)
    env = ProjMPS(psi, H, psi, center=2)
    env.setblock(1, L)
    env.setblock(4, R)
    th = torch.from_numpy(theta.reshape(-1, order='F').view(np.float64).copy()).cuda()
    out = torch.empty_like(th)
    st = (C.c_double * 3)()
    _lib.check(ctx.lib.tn_env_product_dev(env.h, C.c_void_p(th.data_ptr()), 0, C.c_void_p(out.data_ptr()), 3))
    _lib.check(ctx.lib.tn_env_product_profile(env.h, C.c_void_p(th.data_ptr()), 0, C.c_void_p(out.data_ptr()), steps, st))
    ms = [st[i] / steps for i in range(3)]
    return {"config": "C2_heff_matvec chi=1024 w=5 (XXZ chain)", "ms_per_matvec": sum(ms), "tflops": flops_matvec(chi, D, w) / (sum(ms) * 1e-3) / 1e12,
            "stage_ms": {"L.Theta": ms[0], ".W": ms[1], ".R": ms[2]},
            "chi3_kernel_tflops": 2 * 8.0 * chi ** 3 * D * D * w / ((ms[0] + ms[2]) * 1e-3) / 1e12}


def svd_sample(ctx, torch, zgemm_peak, sizes=(512, 2048)):
    """Truncated SVD (tn_svd_trunc; reference tensors.jl:168-227) on a graded spectrum, host buffers in and out, next to cuSOLVER
    (torch.linalg.svd, driver gesvd) on the same box.  F_svd = 4 (6 m n^2 + 20 n^3) is the LAPACK-style normaliser of SURVEY 8(d)."""
    import tnb200
    out = []
    rng = np.random.default_rng(5)
    for n in sizes:
        u, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        v, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        s = np.exp(-np.arange(n) * 30.0 / n)
        x = np.asfortranarray((u * s) @ v.conj().T)
        tnb200.svd(x, 2, cutoff=0.0, ctx=ctx)
        best = 1e9
        for _ in range(2):
            t0 = time.perf_counter()
            U, S, V = tnb200.svd(x, 2, cutoff=0.0, ctx=ctx)
            best = min(best, time.perf_counter() - t0)
        err = float(np.max(np.abs(np.real(np.diag(S)) - s)))
        fs = 4.0 * (6.0 * n ** 3 + 20.0 * n ** 3)
        rec = {"n": n, "ms": best * 1e3, "F_svd": fs, "tflops_nominal": fs / best / 1e12, "frac_of_zgemm": fs / best / 1e12 / zgemm_peak,
               "max_abs_sigma_err": err, "includes": "H2D of the matrix, D2H of U, S, V^H"}
        # what the MPS drivers call (replacesites! / moveleft! / moveright!): one isometry + the S-weighted other factor, matrix resident
        A, sv, B, sw, ms_dev = tnb200.svd_split(x, 1, cutoff=0.0, ctx=ctx, repeat=3)
        rec["split_device_ms"] = ms_dev
        rec["split_sweeps"] = sw
        rec["split_tflops_nominal"] = fs / ms_dev / 1e9
        rec["split_frac_of_zgemm"] = fs / ms_dev / 1e9 / zgemm_peak
        rec["split_recon_rel"] = float(np.linalg.norm(A @ B - x) / np.linalg.norm(x))
        try:
            xt = torch.from_numpy(np.ascontiguousarray(x)).cuda()
            torch.linalg.svd(xt, full_matrices=False, driver="gesvd")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            torch.linalg.svd(xt, full_matrices=False, driver="gesvd")
            torch.cuda.synchronize()
            rec["cusolver_zgesvd_ms"] = (time.perf_counter() - t0) * 1e3
        except Exception as e:
            rec["cusolver_zgesvd_ms"] = None
            rec["cusolver_error"] = repr(e)[:120]
        out.append(rec)
    return out


def c5_exact_sample(ctx, chi_max=4096):
    """BASELINE config 5 where the answer is known: two-site DMRG (cutoff = 0) of the J1-J2 model on a 4 x 6 cylinder (N = 24, w = 20), maxdim
    ramped to 4096 -- the central bond is then exact (2^12) and the energy must equal exact diagonalisation (tools/ed_j1j2.py, S^z = 0 sector,
    no MPS code).  The central bond of the last passes is this bench's headline workload inside a real sweep (Theta (2048, 2, 2, 2048))."""
    import ctypes as C
    import tnb200
    from tnb200.mpo import MPO
    from tnb200._lib import check, tn_lanczos_t
    e_ed = -47.290880085317
    N = 24
    gH = MPO(N, 2, tnb200.models.j1j2_cylinder_terms(4, 6), ctx=ctx)
    g = tnb200.GMPS(1, 2, tnb200.models.random_canonical_mps(N, 2, 16, seed=7), 1, ctx=ctx)
    g.movecenter(1)
    Hs = tnb200.ProjMPS(g, gH, g, center=1)
    direction, out = False, []
    for chi, passes in ((64, 4), (256, 2), (1024, 2), (2048, 2), (4096, 1)):
        if chi > chi_max:
            break
        for _ in range(passes):
            e, mb = C.c_double(), C.c_int64()
            c0 = ctx.counters()
            t0 = time.perf_counter()
            check(g.lib.tn_dmrg_sweep(g.h, Hs.h, int(direction), tn_lanczos_t(3, 2, 1e-14), tnb200.Trunc(0.0, chi, 1), C.byref(e), C.byref(mb)))
            dt = time.perf_counter() - t0
            c1 = ctx.counters()
            direction = not direction
        out.append({"maxdim": chi, "maxbond": mb.value, "seconds_last_pass": dt, "energy": e.value, "rel_err_vs_ed": abs(e.value - e_ed) / abs(e_ed),
                    "heff_applications": c1["matvecs"] - c0["matvecs"], "svds": c1["svds"] - c0["svds"], "gpu_launches": c1["launches"] - c0["launches"]})
    return {"config": "J1-J2 (J2 = 0.5) 4 x 6 cylinder, N = 24, MPO bond 20, two-site DMRG, cutoff = 0; a pass = one sweep direction over all bonds",
            "ed_energy": e_ed, "ed_source": "tools/ed_j1j2.py (S^z = 0 sector, 2 704 156 states, ARPACK)", "passes": out}


def qjmc_wave(device, N=64, chi=256, traj=64, workers=64, steps=1):
    """QJMC trajectory throughput at the C4 shapes (dissipative Ising chain N=64, chi=256, cutoff=0, seeded random canonical
    start): one wave of ``traj`` trajectories x ``steps`` steps on this GPU after a one-step warm-up, wall clock around
    tn_qjmc_ensemble (SVD batching rounds across the wave's trajectories)."""
    import tnb200
    from tnb200 import models
    os.environ.setdefault("TN_QJMC_BATCH", "1")
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
    return {"seconds": sec, "traj_steps": traj * steps, "jumps": int(nj.sum()), "mean_sum_z": float(np.real(obs[:, -1, :]).sum() / traj)}


def _qjmc_cpu_worker(arg):
    """One trajectory x one step of the oracle's qjmc_simulation (restates qjmc.jl:28-167) at the C4 shapes, `threads` BLAS threads."""
    seed, N, chi, threads = arg
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=threads)
    except Exception:
        pass
    import oracle
    from oracle.gmps import GMPS as OG
    sh = oracle.spinhalf()
    H = oracle.OpList(N)
    J = oracle.OpList(N)
    for i in range(1, N + 1):
        H.add("x", i, 1.0)
        H.add("z", i, 20.0)
        J.add("s-", i, np.sqrt(0.1))
    for i in range(1, N):
        H.add(["z", "z"], [i, i + 1], 10.0)
    rng = np.random.default_rng(1)           # same seeded right-canonical start as the GPU wave (tnb200.models.random_canonical_mps)
    dims = [min(D ** i, D ** (N - i), chi) for i in range(N + 1)]
    tens = []
    for i in range(N):
        l, r = dims[i], dims[i + 1]
        a = rng.standard_normal((l, D * r)) + 1j * rng.standard_normal((l, D * r))
        q, _ = np.linalg.qr(a.T.conj())
        tens.append(np.asfortranarray(np.ascontiguousarray(q.conj().T).reshape(l, D, r, order='F')))
    tens[0] = tens[0] / np.linalg.norm(tens[0])
    psi = OG(1, D, tens, 1)
    dt = 5e-3
    t0 = time.perf_counter()
    oracle.qjmc_simulation(sh, psi, H, J, dt, dt, uniforms=np.random.default_rng(seed).random, cutoff=0.0, maxdim=chi)
    return time.perf_counter() - t0


def qjmc_cpu_sample(N=64, chi=256):
    """CPU baseline of the QJMC leg: ONE trajectory x 1 step of the oracle port with all host threads in its BLAS / LAPACK calls
    (the reference itself is single-process: qjmc.jl:28 runs one trajectory per call) -- about 30 s of CPU work.  Trajectories
    are independent, so a host could also run one per core; that variant was not timed (4 processes x 4 threads exceeded the
    bench budget on the GPU box)."""
    cores = blas_threads()
    sec = _qjmc_cpu_worker((100, N, chi, cores))
    return {"value": 1.0 / sec, "unit": "trajectory-steps/s", "cores": cores, "kind": "port",
            "sample": f"1 trajectory x 1 step at the C4 shapes (N={N}, chi={chi}, cutoff=0), NumPy/SciPy port with {cores} BLAS threads, {sec:.1f} s"}


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
    chi, w = args.chi, W
    M1, M2 = oracle_mpo_pair()
    assert M1.shape == (w, D, D, w) and M2.shape == (w, D, D, w), (M1.shape, M2.shape)
    W_eff = max(3, args.warmup)
    rows = min(REF_ROWS, chi)
    kind = "port"
    jl = julia_reference_sample(chi, w, M1, M2, args.steps, W_eff, rows)
    if jl is not None:
        tf, calls, secs, cores, kind = jl["tflops"], args.steps, jl["seconds_per_step"] * args.steps, jl["threads"], "julia"
        how = f"the reference itself under Julia {jl['julia']} (baseline/run_reference.jl), {cores} BLAS threads"
    else:
        tf, calls, secs, cores, _ = cpu_reference_sample(chi, w, M1, M2, args.steps, W_eff, rows=rows)
        how = f"NumPy/OpenBLAS port of the reference's product(), {cores} threads"
    sample = (f"each step = rows a in [0,{rows}) of one H_eff*Theta at chi={chi}, w={w} ({rows}/{chi} of a full application, normalised by the same "
              f"fraction of F_mv), reference contraction order (L.M1.M2 first), {how}")
    line = {
        "impl": "reference", "metric": METRIC, "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": calls, "warmup": W_eff,
        "ms_per_step": secs / calls * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128 (complex f64)",
        "data": "synthetic", "config": config(chi, w),
        "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    global T_START
    T_START = time.perf_counter()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--chi", type=int, default=CHI)
    ap.add_argument("--slices", type=int, default=0, help="N > 1: slices of Theta's right bond (reduce_scatter of slice j under the GEMMs of slice j+1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extras (C2 matvec, DMRG sweep times, SVD, QJMC)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import tnb200
    from tnb200 import _lib
    from tnb200.api import GMPS, ProjMPS
    from tnb200.mpo import MPO
    from tnb200.sharded import BalancedShardedHeff, GpuContractor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    chi, w = args.chi, W
    W_eff = max(3, args.warmup)
    K = args.steps
    if args.slices <= 0:
        args.slices = max(4, world)

    ctx = tnb200.Context(local)
    lib = ctx.lib
    # the MPO of the workload, built by the product's own builder (host assembly + device compression, tnb200/mpo.py)
    gH = MPO(LX * LY, D, tnb200.models.j1j2_cylinder_terms(LX, LY), ctx=ctx)
    mpo_host = gH.tensors
    M1, M2 = np.asfortranarray(mpo_host[SITE - 1]), np.asfortranarray(mpo_host[SITE])
    assert M1.shape == (w, D, D, w) and M2.shape == (w, D, D, w), (M1.shape, M2.shape)
    L, R, theta = make_inputs(chi, w)
    n = chi * D * D * chi
    flops = flops_matvec(chi, D, w)
    th_dev = torch.from_numpy(theta.reshape(-1, order='F').copy()).cuda()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        ctx.sync()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if world == 1:
        dims = [1, chi, chi, chi, 1]
        rng = np.random.default_rng(1)
        sites = [np.asfortranarray(rng.standard_normal((dims[i], D, dims[i + 1])) + 0j) for i in range(4)]   # placeholders: the matvec reads blocks, MPO, Theta
        psi = GMPS(1, D, sites, 0, ctx=ctx)
        psi.center = 2
        H = GMPS(2, D, [M1[:1], M1, M2, M2[..., :1]], ctx=ctx)
        env = ProjMPS(psi, H, psi, center=2)
        env.setblock(1, L)
        env.setblock(4, R)
        out_dev = torch.empty_like(th_dev)

        def run_steps(k):
            _lib.check(lib.tn_env_product_dev(env.h, C.c_void_p(th_dev.data_ptr()), 0, C.c_void_p(out_dev.data_ptr()), k))
            return out_dev
        api = "tnb200.ProjMPS.product(A_host) -> tn_env_product"
        multi = "single GPU"
    else:
        sh = BalancedShardedHeff(L, R, M1, M2, rank, world, GpuContractor(ctx), "cuda", dist)
        del L, R

        def run_steps(k):
            o = None
            for _ in range(k):
                o = sh.apply_pipelined(th_dev, args.slices, ctx.stream())
            return o
        api = ("tnb200.sharded.BalancedShardedHeff.apply_pipelined(Theta) on every rank; every rank uploads 1/N of the host Theta (NCCL all_gather "
               "completes the device replicas) and downloads 1/N of the result")
        multi = (f"MPO-bond-sharded over {world} ranks: even split of the fused (a,w) rows / (b',w2) contraction index; NCCL reduce_scatter of T2 "
                 f"({16 * n * w / 1e9:.2f} GB per rank before reduction) in {args.slices} slices under the next slice's GEMMs + NCCL all_reduce of the result "
                 f"({16 * n / 1e6:.0f} MB)")

    # ---- device-resident timing ------------------------------------------------------------------
    run_steps(W_eff)
    barrier()
    c0 = ctx.counters()["launches"]
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    out_t = run_steps(K)
    e1.record(stream)
    barrier()
    ms_max = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    launches = ctx.counters()["launches"] - c0
    value = flops * K / (ms_max * 1e-3) / 1e12
    dev_out = out_t.cpu().numpy().view(np.complex128).reshape(-1)

    # ---- roofline of the dominant kernel (live CUDA events) ---------------------------------------
    if world == 1:
        st = (C.c_double * 3)()
        _lib.check(lib.tn_env_product_profile(env.h, C.c_void_p(th_dev.data_ptr()), 0, C.c_void_p(out_dev.data_ptr()), K, st))
        stage_ms = [st[i] / K for i in range(3)]
        big_flops = 8.0 * chi ** 3 * D * D * w            # per chi^3 contraction (2 launches per matvec)
        ms_launch = (stage_ms[0] + stage_ms[2]) / 2
        stage_rec = {"L.Theta": stage_ms[0], ".W": stage_ms[1], ".R": stage_ms[2]}
    else:
        # this rank's share of the first chi^3 contraction, alone on the library stream
        m_loc = sh.mloc
        big_flops = 8.0 * m_loc * chi * (D * D * chi)
        ct = sh.contract
        BIG = 1 << 40

        def stage1():
            ct(m_loc, D * D * chi, chi, sh.Lg, (BIG, 1, 0), (BIG, m_loc, 0), th_dev, (BIG, 1, 0), (BIG, chi, 0), sh.T1[sh.r0:], (BIG, 1, 0), (BIG, chi * sh.nw, 0))
        stage1()
        barrier()
        e0.record(stream)
        for _ in range(K):
            stage1()
        e1.record(stream)
        barrier()
        ms_launch = max_over_ranks(e0.elapsed_time(e1)) / K
        stage_rec = {"L.Theta (this rank's rows)": ms_launch}
    achieved = big_flops / (ms_launch * 1e-3) / 1e12

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
    if world == 1:
        A_host = np.ndarray((chi, D, D, chi), dtype=np.complex128, buffer=hin.numpy().data, order='F')
        O_host = np.ndarray((chi, D, D, chi), dtype=np.complex128, buffer=hout.numpy().data, order='F')

        def e2e_step():
            env.product(A_host, False, out=O_host)
    else:
        th_real = torch.view_as_real(th_dev).reshape(-1)
        # Theta is replicated on the host of every rank: each rank uploads 1/N of it over its own PCIe link and the replicas are
        # completed by an NCCL all_gather over NVLink; the result comes back the same way (each rank downloads its 1/N slice)
        per = (2 * n + world - 1) // world
        lo, hi = min(rank * per, 2 * n), min((rank + 1) * per, 2 * n)
        pad = torch.empty(per * world, dtype=torch.float64, device="cuda") if per * world != 2 * n else None

        def e2e_step():
            if pad is None:
                th_real[lo:hi].copy_(hin[lo:hi], non_blocking=True)
                dist.all_gather_into_tensor(th_real, th_real[lo:hi])
            else:
                pad[rank * per:rank * per + (hi - lo)].copy_(hin[lo:hi], non_blocking=True)
                dist.all_gather_into_tensor(pad, pad[rank * per:(rank + 1) * per])
                th_real.copy_(pad[:2 * n])
            torch.cuda.current_stream().synchronize()
            o = sh.apply_pipelined(th_dev, args.slices, ctx.stream())
            hout[lo:hi].copy_(torch.view_as_real(o).reshape(-1)[lo:hi], non_blocking=True)   # this rank's slice of the result
            torch.cuda.current_stream().synchronize()
    for _ in range(W_eff):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(K):
        e2e_step()
    e1.record(stream)
    barrier()
    # the API call blocks until the result is in the host buffer, so the wall clock and the device events bracket the same work
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), 0.0))
    e2e_wall = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_val = flops * K / (max(e2e_ms, e2e_wall) * 1e-3) / 1e12
    if world == 1:
        e2e_out = hout.numpy().view(np.complex128)
        same = float(np.linalg.norm(dev_out - e2e_out) / np.linalg.norm(dev_out))
    else:                                  # every rank holds its slice of the result on the host
        mine = hout.numpy()[lo:hi]
        ref = dev_out.view(np.float64)[lo:hi]
        same = float(np.linalg.norm(ref - mine) / max(np.linalg.norm(ref), 1e-300))

    line = None
    if rank == 0:
        cfg = config(chi, w)
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": K, "warmup": W_eff,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "c128 (complex f64)", "data": "synthetic", "config": cfg,
            "arm": {"order": "flop-optimal (L.Theta).W.R", "multi_gpu": multi, "mpo_builder": "tnb200.mpo.MPO (host FSM assembly + device SVD compression)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": "TFLOP/s", "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 16 * n,
                    "api": api, "matches_resident_path_rel": same, "ms_per_step_events": e2e_ms / K, "ms_per_step_wall": e2e_wall / K},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": zgemm_peak, "unit": "TFLOP/s", "frac": achieved / zgemm_peak,
                         "traffic": None, "kernel": "tn::zgemm_kernel<4,1,4,4,1> (128x32 CTA tile, 2 CTAs/SM, DMMA.8x8x4, 4-stage cp.async; its persistent stream-K twin zgemm_sk_kernel runs when the tile count leaves a partial last wave), 2 launches per matvec",
                         "flops_per_launch": big_flops, "ms_per_launch": ms_launch,
                         "peak_source": "cuBLAS ZGEMM 4096^3 via torch.matmul measured in this run (MEASURED_PEAKS.json has no FP64 figure; "
                                        "DMMA issue peak measured 37.17 TFLOP/s, profiles/r01_probe_fp64.jsonl)",
                         "stage_ms": stage_rec},
        }
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch, same shapes)
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "zgemm_traffic.json")))
            if tr.get("chi") == chi and tr.get("w") == w and world == 1:
                line["roofline"]["traffic"] = tr["dram_bytes_per_launch"]
                line["roofline"]["traffic_source"] = tr["source"]
                line["roofline"]["algorithmic_bytes_per_launch"] = tr.get("algorithmic_bytes_per_launch")
        except Exception:
            pass
    extras = {}
    if not args.no_extras:
        if world == 1:
            for name, fn in (("c2_matvec", lambda: c2_matvec_sample(ctx, torch)), ("dmrg_sweep", lambda: dmrg_sweep_sample(ctx)),
                             ("svd", lambda: svd_sample(ctx, torch, zgemm_peak)), ("c5_exact", lambda: c5_exact_sample(ctx))):
                if name == "c5_exact" and time.perf_counter() - T_START > 150.0:      # ~45 s: only while the run stays within a few minutes
                    extras[name] = {"skipped": "time budget"}
                    continue
                try:
                    extras[name] = fn()
                except Exception as e:          # the samples are by-products: never lose the bench line over them
                    extras[name] = {"error": repr(e)[:200]}
        # QJMC (every N): one wave of trajectories per GPU of the 8192-trajectory C4 job; trajectories are independent
        # units (no data-path collective), so the job's throughput is N waves' trajectory-steps over the slowest wave
        try:
            del th_dev
            torch.cuda.empty_cache()
            barrier()
            q = qjmc_wave(local)
            sec = max_over_ranks(q["seconds"])
            tot = q["traj_steps"] * world
            extras["qjmc_scaling"] = {"config": "C4 shapes: N=64, chi=256, cutoff=0; one wave of 64 trajectories x 1 step per GPU (the 8192-trajectory job is "
                                                f"8192/(64*{world}) such waves per GPU), SVD batching rounds in two worker groups across the wave",
                                      "n_gpus": world, "seconds_slowest_rank": sec, "traj_steps_per_s": tot / sec, "traj_per_s_at_20_steps": tot / sec / 20.0,
                                      "collective": "none during the evolution (final gather of jump records only)", "jumps_rank0": q["jumps"],
                                      "mean_sum_z_rank0": q["mean_sum_z"]}
            if world == 1 and not args.no_cpu_baseline and time.perf_counter() - T_START < 240.0:     # keep the whole run within a few minutes
                extras["qjmc_scaling"]["cpu_baseline"] = qjmc_cpu_sample()
        except Exception as e:
            extras["qjmc_scaling"] = {"error": repr(e)[:200]}
    if rank == 0:
        if extras:
            line["extras"] = extras
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only (at N > 1 the other ranks' processes share the host cores)
            rows = min(REF_ROWS, chi)
            tf, calls, secs, cores, ref_out = cpu_reference_sample(chi, w, M1, M2, 4, 1, rows=rows, budget_s=25.0)
            got = dev_out.reshape(chi, D, D, chi, order='F')[:rows]
            err = float(np.linalg.norm(ref_out - got) / np.linalg.norm(got))
            line["cpu_baseline"] = {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "port",
                                    "sample": f"{calls} x rows a in [0,{rows}) of one matvec at chi={chi}, w={w} ({rows}/{chi} of an application), reference "
                                              f"contraction order, NumPy/OpenBLAS {cores} threads, {secs:.1f} s",
                                    "gpu_vs_oracle_rel_err": err}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
