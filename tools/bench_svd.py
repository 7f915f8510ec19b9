"""Truncated-SVD timing on one GPU (device time of factorisation + gathers, matrix resident; tn_svd_trunc_split):
  python tools/bench_svd.py 512,1024,2048 [graded|randn]
Prints one JSON line per size and mode; TN_SVD_PROFILE=1 adds the library's phase split on stderr."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tensornetworks.jl_b200")):
    sys.path.insert(0, p)
import numpy as np
import tnb200

sizes = [int(a) for a in (sys.argv[1] if len(sys.argv) > 1 else "512,1024,2048").split(",")]
kind = sys.argv[2] if len(sys.argv) > 2 else "graded"
rng = np.random.default_rng(0)
ctx = tnb200.Context.default()
for n in sizes:
    if kind == "randn":
        x = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        so = None
    else:
        u, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        v, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        so = np.exp(-np.arange(n) * 30.0 / n)
        x = (u * so) @ v.conj().T
    fs = 4.0 * 26.0 * n ** 3
    for mode, label in ((1, "split, isometry U (W-only)"), (2, "split, isometry V^H (W-only)"), (0, "full (U, S, V^H)")):
        if mode == 0:
            tnb200.svd(x, 2, ctx=ctx)
            t0 = time.perf_counter(); U, S, V, sw = tnb200.svd(x, 2, ctx=ctx, return_sweeps=True); ms = (time.perf_counter() - t0) * 1e3
            s = np.real(np.diag(S)); note = "wall clock incl. PCIe"
            rec = float(np.linalg.norm(U @ S @ V - x) / np.linalg.norm(x))
        else:
            A, s, B, sw, ms = tnb200.svd_split(x, mode, ctx=ctx, repeat=3)
            note = "device time, matrix resident"
            rec = float(np.linalg.norm(A @ B - x) / np.linalg.norm(x))
        err = float(np.max(np.abs(s - so))) if so is not None else None
        print(json.dumps({"n": n, "kind": kind, "mode": label, "sweeps": sw, "ms": ms, "timing": note, "F_svd": fs, "tflops_nominal": fs / ms / 1e9,
                          "max_abs_sigma_err": err, "recon_rel": rec}), flush=True)
