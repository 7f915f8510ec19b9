"""H_eff matvec TFLOP/s vs chi (C2 micro-benchmark chi in {128..4096} at w=5, and the C5 point
chi=4096, w=20) on ONE GPU, device-resident operands, CUDA events on the library's stream.
  python tools/bench_matvec_chi.py "128:5,256:5,...,4096:20"
Random L, R, Theta and random dense MPO tensors (the matvec cost does not depend on the MPO's sparsity)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tensornetworks.jl_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import tnb200
from tnb200 import _lib
from tnb200.api import GMPS, ProjMPS

D, NS = 2, 4
spec = sys.argv[1] if len(sys.argv) > 1 else "128:5,256:5,512:5,1024:5,2048:5,4096:5,4096:20"
ctx = tnb200.Context(0)
lib = ctx.lib
stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", 0))

# cuBLAS ZGEMM ceiling measured in this run
a = torch.randn(4096, 4096, dtype=torch.complex128, device="cuda"); b = torch.randn(4096, 4096, dtype=torch.complex128, device="cuda")
torch.matmul(a, b); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x0.record(); torch.matmul(a, b); x1.record(); torch.cuda.synchronize(); best = min(best, x0.elapsed_time(x1))
peak = 8.0 * 4096 ** 3 / (best * 1e-3) / 1e12
del a, b
print(json.dumps({"cublas_zgemm_4096_tflops": peak}), flush=True)

for item in spec.split(","):
    chi, w = (int(v) for v in item.split(":"))
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    dims = [1, chi, chi, chi, 1]
    wd = [1, w, w, w, 1]
    rng = np.random.default_rng(0)
    cr = lambda *s: (rng.standard_normal(s) + 1j * rng.standard_normal(s))
    sites = [np.zeros((dims[i], D, dims[i + 1]), dtype=np.complex128) for i in range(NS)]     # placeholders: only the blocks matter
    mpo = [cr(wd[i], D, D, wd[i + 1]) / np.sqrt(w * D) for i in range(NS)]
    psi = GMPS(1, D, sites, 0, ctx=ctx); psi.center = 2
    H = GMPS(2, D, mpo, ctx=ctx)
    env = ProjMPS(psi, H, psi, center=2)
    # blocks generated on the host in slabs to bound host memory / time
    blk = np.empty((chi, w, chi), dtype=np.complex128, order='F')
    r32 = np.random.default_rng(1)
    v = blk.reshape(-1, order='F').view(np.float64)
    step = 1 << 24
    for o in range(0, v.size, step):
        v[o:o + step] = r32.standard_normal(min(step, v.size - o), dtype=np.float32)
    env.setblock(1, blk)
    env.setblock(4, blk)
    del blk, v
    n = chi * D * D * chi
    th = torch.randn(2 * n, dtype=torch.float64, device="cuda", generator=g)
    out = torch.empty_like(th)
    flops = 8.0 * (2.0 * chi ** 3 * D * D * w + 2.0 * chi ** 2 * D ** 3 * w * w)
    reps = max(2, min(50, int(2e13 / flops)))
    _lib.check(lib.tn_env_product_dev(env.h, C.c_void_p(th.data_ptr()), 0, C.c_void_p(out.data_ptr()), 2))
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    _lib.check(lib.tn_env_product_dev(env.h, C.c_void_p(th.data_ptr()), 0, C.c_void_p(out.data_ptr()), reps))
    e1.record(stream)
    ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = (C.c_double * 3)()
    _lib.check(lib.tn_env_product_profile(env.h, C.c_void_p(th.data_ptr()), 0, C.c_void_p(out.data_ptr()), reps, st))
    tf = flops / (ms * 1e-3) / 1e12
    print(json.dumps({"what": "heff_matvec", "chi": chi, "w": w, "d": D, "ms_per_matvec": ms, "tflops": tf, "frac_of_cublas_zgemm": tf / peak,
                      "stage_ms": [st[i] / reps for i in range(3)], "reps": reps,
                      "hbm_gb": {"L,R": 2 * 16 * chi * chi * w / 1e9, "T1,T2": 2 * 16 * chi * chi * w * D * D / 1e9}}), flush=True)
    del env, psi, H, th, out
    torch.cuda.empty_cache()
