"""Batched vs one-at-a-time truncated SVD of same-shape matrices (the QJMC gate step's 512 x 512 problems): wall time per problem
through the host-buffer entry points (tn_svd_trunc_batched vs B calls of tn_svd_trunc; both include the same PCIe copies).
  python tools/bench_svd_batched.py 512 1,4,16,32
Prints one JSON line per batch size."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensornetworks.jl_b200"))
import numpy as np
import tnb200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
batches = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,4,16,32").split(",")]
rng = np.random.default_rng(0)
ctx = tnb200.Context.default()


def graded(n):
    u, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    v, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    return (u * np.exp(-np.arange(n) * (20.0 / n))) @ v.conj().T


pool = [graded(n) for _ in range(max(batches))]
tnb200.svd(pool[0], 2, maxdim=n // 2)
for B in batches:
    mats = np.stack(pool[:B])
    tnb200.svd_batched(mats, maxdim=n // 2)                   # warm-up (workspaces, pair tables)
    t0 = time.perf_counter()
    res = tnb200.svd_batched(mats, maxdim=n // 2)
    tb = time.perf_counter() - t0
    t0 = time.perf_counter()
    for x in pool[:B]:
        U, S, Vh = tnb200.svd(x, 2, maxdim=n // 2)
    ts = time.perf_counter() - t0
    err = max(float(np.max(np.abs(r[1] - np.exp(-np.arange(n // 2) * (20.0 / n))))) for r in res)
    print(json.dumps({"what": "svd_batched", "n": n, "batch": B, "ms_per_problem_batched": 1e3 * tb / B, "ms_per_problem_single": 1e3 * ts / B,
                      "speedup": ts / tb, "max_sv_err": err}), flush=True)
