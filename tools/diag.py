"""Scratch diagnostics run on the GPU box (not part of the product or the tests)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tensornetworks.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import tnb200
from tnb200 import _lib
import ctypes as C

ctx = tnb200.Context.default()
rng = np.random.default_rng(0)

def crandn(*s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)

what = sys.argv[1:] or ["svd", "matvec"]
if "svdmodes" in what:
    def mk(kind, n):
        if kind == "randn": return crandn(n, n)
        if kind == "hidden":
            u, _ = np.linalg.qr(crandn(n, n)); v, _ = np.linalg.qr(crandn(n, n)); return (u * np.exp(-np.arange(n) * 30.0 / n)) @ v.conj().T
        if kind == "visible":
            lam = np.exp(-np.arange(n // 2) * 28.0 / (n // 2)); return (np.tile(lam, 2)[:, None] * crandn(n, n)) * np.tile(lam, 2)[None, :]
        if kind == "rankdef": return crandn(n, n // 2) @ crandn(n // 2, n)
        if kind == "tall": return crandn(2 * n, n) * np.exp(-np.arange(n) * 20.0 / n)[None, :]
    for n in [int(a) for a in os.environ.get("SVD_N", "256,1024").split(",")]:
        for kind in ("randn", "hidden", "visible", "rankdef", "tall"):
            x = mk(kind, n)
            so = np.linalg.svd(x, compute_uv=False)
            for mode in (1, 0):
                if mode == 0 and kind in ("hidden", "visible") and n > 512: continue
                ctx.lib.tn_svd_set_precond(mode)
                tnb200.svd(x, 2)
                t1 = time.perf_counter()
                U, S, V, sw = tnb200.svd(x, 2, return_sweeps=True)
                t2 = time.perf_counter()
                s = np.real(np.diag(S)); k = int(np.sum(so > 1e-13 * so[0]))
                print(dict(kind=kind, n=n, precond=mode, sweeps=sw, wall_s=round(t2 - t1, 4), sv_err=float(np.max(np.abs(s - so)) / so[0]),
                           recon=float(np.linalg.norm(U @ S @ V - x) / np.linalg.norm(x)),
                           orthU=float(np.linalg.norm(U[:, :k].conj().T @ U[:, :k] - np.eye(k))), orthV=float(np.linalg.norm(V[:k] @ V[:k].conj().T - np.eye(k)))), flush=True)
    ctx.lib.tn_svd_set_precond(1)

if "svdone" in what:
    n = int(os.environ.get("SVD_N", "1024"))
    u, _ = np.linalg.qr(crandn(n, n)); v, _ = np.linalg.qr(crandn(n, n)); x = (u * np.exp(-np.arange(n) * 30.0 / n)) @ v.conj().T
    tnb200.svd(x, 2)
    t1 = time.perf_counter(); U, S, V, sw = tnb200.svd(x, 2, return_sweeps=True); print("svdone", n, sw, time.perf_counter() - t1, flush=True)

if "svd" in what:
    for n in [int(a) for a in os.environ.get("SVD_N", "256,512,1024,2048").split(",")]:
        x = crandn(n, n)
        t0 = time.perf_counter()
        U, S, V, sw = tnb200.svd(x, 2, return_sweeps=True)
        t1 = time.perf_counter()
        U, S, V, sw = tnb200.svd(x, 2, return_sweeps=True)
        t2 = time.perf_counter()
        s = np.real(np.diag(S))
        so = np.linalg.svd(x, compute_uv=False)
        print(dict(n=n, sweeps=sw, wall_s_first=round(t1 - t0, 4), wall_s=round(t2 - t1, 4),
                   sv_err=float(np.max(np.abs(s - so)) / so[0]),
                   recon=float(np.linalg.norm(U @ S @ V - x) / np.linalg.norm(x)),
                   orthU=float(np.linalg.norm(U.conj().T @ U - np.eye(n))), orthV=float(np.linalg.norm(V @ V.conj().T - np.eye(n))),
                   resid=float(np.linalg.norm(x @ V.conj().T - U @ S) / np.linalg.norm(x))), flush=True)

if "matvec" in what:
    from tnb200.api import GMPS, ProjMPS
    d, w, N = 2, 5, 6
    for chi in [int(a) for a in os.environ.get("MV_CHI", "128,256,512,1024").split(",")]:
        # random environment by hand: upload an MPS whose middle bonds are chi (not canonical; fine for timing)
        dims = [1, 2, chi, chi, chi, 2, 1]
        dims = [1] + [min(chi, 2 ** min(i, N - i)) if min(i, N - i) < 12 else chi for i in range(1, N)] + [1]
        dims = [1, chi, chi, chi, chi, chi, 1]
        tens = [crandn(dims[i], d, dims[i + 1]) / np.sqrt(dims[i] * d) for i in range(N)]
        W = tnb200.models.xxz_mpo(N)
        psi = GMPS(1, d, tens, 0)
        psi.center = 3
        H = GMPS(2, d, W)
        env = ProjMPS(psi, H, psi, center=3)
        n = chi * d * d * chi
        th = torch.randn(n, 2, dtype=torch.float64, device="cuda")
        out = torch.empty_like(th)
        torch.cuda.synchronize()
        lib = ctx.lib
        _lib.check(lib.tn_env_product_dev(env.h, C.c_void_p(th.data_ptr()), 0, C.c_void_p(out.data_ptr()), 2))
        ctx.sync()
        reps = 5
        t0 = time.perf_counter()
        _lib.check(lib.tn_env_product_dev(env.h, C.c_void_p(th.data_ptr()), 0, C.c_void_p(out.data_ptr()), reps))
        ctx.sync()
        dt = (time.perf_counter() - t0) / reps
        flops = 8.0 * (2 * chi ** 3 * d * d * w + 2 * chi ** 2 * d ** 3 * w * w)
        print(dict(chi=chi, ms=round(dt * 1e3, 3), tflops=round(flops / dt / 1e12, 2)), flush=True)

if "smallsvd" in what:
    for n in (4, 8, 16, 32, 64, 128):
        x = crandn(n, n)
        tnb200.svd(x, 2)
        t1 = time.perf_counter()
        for _ in range(20):
            U, S, V, sw = tnb200.svd(x, 2, return_sweeps=True)
        dt = (time.perf_counter() - t1) / 20
        so = np.linalg.svd(x, compute_uv=False)
        print(dict(n=n, ms=round(dt * 1e3, 3), sweeps=sw, sv_err=float(np.max(np.abs(np.real(np.diag(S)) - so)) / so[0])), flush=True)

if "qjmcstep" in what:
    import oracle
    N, chi = 32, 64
    X, Z, SM = tnb200.models.X, tnb200.models.Z, tnb200.models.SM
    onsite = -1j * (1.0 * X + 20.0 * Z) - 0.5 * 0.1 * (SM.conj().T @ SM)
    ss, gg = tnb200.models.trotter_gates(N, onsite, -1j * 10.0 * np.kron(Z, Z), 5e-3, evol="imag", order=2)
    psi = tnb200.GMPS(1, 2, tnb200.models.random_canonical_mps(N, 2, chi, seed=1), 1)
    gl = tnb200.GateList(2, ss, gg)
    for label, fn in (("applygates", lambda: tnb200.applygates(psi, gl, cutoff=0.0, maxdim=chi)),
                      ("normalize", lambda: psi.normalize()),
                      ("expect_local", lambda: psi.expect([SM.conj().T @ SM] * N, list(range(1, N + 1))))):
        fn()
        c0 = ctx.counters(); t1 = time.perf_counter(); fn(); dt = time.perf_counter() - t1; c1 = ctx.counters()
        print(dict(op=label, N=N, chi=chi, ms=round(dt * 1e3, 2), launches=c1["launches"] - c0["launches"], svds=c1["svds"] - c0["svds"]), flush=True)
