"""TEBD step time (BASELINE.json configs[2]): imaginary-time TFIM (h=1, J=1, passed as -H), second-order
Trotter, dt=5e-3, cutoff=0 and maxdim=chi so the shapes are data-independent, seeded random canonical MPS with bonds
min(2^i, 2^(N-i), chi).  One step = applygates! (3 rows) + norm + normalize! (tebd.jl:62-78) through tnb200.tebd.
  python tools/bench_tebd.py "128:256,32:1024,26:2048" [--oracle-max-chi 64]
Prints one JSON line per (N, chi); with --oracle-max-chi the CPU oracle (NumPy/LAPACK zgesdd) runs the same step for chi <= that."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tensornetworks.jl_b200")):
    sys.path.insert(0, p)
import numpy as np
import tnb200
from tnb200 import models

spec = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "128:256"
omax = int(sys.argv[sys.argv.index("--oracle-max-chi") + 1]) if "--oracle-max-chi" in sys.argv else 0
ctx = tnb200.Context(0)
d, dt = 2, 5e-3
for item in spec.split(","):
    N, chi = (int(v) for v in item.split(":"))
    ss, gg = models.trotter_gates(N, -1.0 * models.X, -1.0 * np.kron(models.Z, models.Z), dt, evol="imag", order=2)
    if chi <= 512:
        tens = models.random_canonical_mps(N, d, chi, seed=3)
        psi = tnb200.GMPS(1, d, tens, 1, ctx=ctx)
    else:
        # large bonds: seeded random (non-canonical) tensors, a handful of distinct ones per shape, brought to canonical form on the
        # device (untruncated gauge moves = one QR step each); a host QR per site would take minutes at chi = 2048
        rng = np.random.default_rng(3)
        dims = [min(d ** i, d ** (N - i), chi) for i in range(N + 1)]
        pool = {}
        tens = []
        for i in range(N):
            key = (dims[i], dims[i + 1], i % 4)
            if key not in pool:
                pool[key] = np.asfortranarray((rng.standard_normal((dims[i], d, dims[i + 1])) + 1j * rng.standard_normal((dims[i], d, dims[i + 1]))) / np.sqrt(dims[i] * d))
            tens.append(pool[key])
        psi = tnb200.GMPS(1, d, tens, 0, ctx=ctx)
        del tens, pool
        psi.movecenter(1)
        psi.normalize()
    gl = tnb200.GateList(d, ss, gg, ctx=ctx)
    if "--no-warmup" not in sys.argv:
        tnb200.tebd(psi, gl, 1, cutoff=0.0, maxdim=chi)            # warm-up step (workspaces, bond dimensions settle)
    c0 = ctx.counters()
    t0 = time.perf_counter()
    psi, _, normal = tnb200.tebd(psi, gl, 1, cutoff=0.0, maxdim=chi)
    ctx.sync()
    sec = time.perf_counter() - t0
    c1 = ctx.counters()
    line = {"what": "tebd_step", "sites": N, "chi": chi, "maxbond": psi.maxbonddim(), "gates": sum(len(r) for r in ss), "seconds_per_step": sec,
            "svds": c1["svds"] - c0["svds"], "gpu_launches": c1["launches"] - c0["launches"], "lognorm": normal}
    if chi <= omax and chi <= 512:
        import oracle
        from oracle.gmps import GMPS as OG
        po = OG(1, d, [t.copy() for t in tens], 1)
        glo = type("GL", (), {})()
        glo.sites, glo.gates = ss, gg
        t0 = time.perf_counter()
        oracle.applygates(po, glo, cutoff=0.0, maxdim=chi)
        nrm = po.norm(); po.normalize()
        line["oracle_seconds_per_step"] = time.perf_counter() - t0
        line["oracle_threads"] = os.cpu_count()
        line["oracle_lognorm_first_step"] = float(np.log(np.real(nrm)))
    print(json.dumps(line), flush=True)
    del psi, gl
