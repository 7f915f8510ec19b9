"""One W-only truncated SVD per size on a graded spectrum (for `ncu` launch lists): python tools/svd_once.py 512,2048"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tensornetworks.jl_b200")):
    sys.path.insert(0, p)
import numpy as np
import tnb200

rng = np.random.default_rng(0)
ctx = tnb200.Context.default()
for n in [int(a) for a in (sys.argv[1] if len(sys.argv) > 1 else "512").split(",")]:
    u, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    v, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    x = (u * np.exp(-np.arange(n) * 30.0 / n)) @ v.conj().T
    A, s, B, sw, ms = tnb200.svd_split(x, 1, ctx=ctx, repeat=1)
    print(n, sw, ms, flush=True)
