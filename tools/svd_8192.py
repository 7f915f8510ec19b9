"""One 8192 x 8192 W-only truncated SVD on a graded spectrum (the factorisation size of a chi = 4096 two-site DMRG bond): python tools/svd_8192.py"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tensornetworks.jl_b200")):
    sys.path.insert(0, p)
import numpy as np, scipy.linalg as sla
import tnb200
n = 8192
rng = np.random.default_rng(0)
t0 = time.time()
# graded spectrum without two full QRs: random unitary-ish factors from one QR each of a real Gaussian (cheaper), complex phases on the columns
u, _ = np.linalg.qr(rng.standard_normal((n, n))); v, _ = np.linalg.qr(rng.standard_normal((n, n)))
ph = np.exp(2j * np.pi * rng.random(n))
so = np.exp(-np.arange(n) * 30.0 / n)
x = np.asfortranarray(((u * ph) * so) @ v.T)
print("generated in %.1f s" % (time.time() - t0), flush=True)
ctx = tnb200.Context.default()
A, s, B, sw, ms = tnb200.svd_split(x, 1, ctx=ctx, repeat=2)
print(json.dumps({"n": n, "kind": "graded", "mode": "split, isometry U (W-only)", "sweeps": sw, "ms": ms, "F_svd": 4.0 * 26.0 * n ** 3,
                  "tflops_nominal": 4.0 * 26.0 * n ** 3 / ms / 1e9, "max_abs_sigma_err": float(np.max(np.abs(s - so)))}), flush=True)
