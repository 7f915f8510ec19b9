"""T1 / BASELINE config 5 at the size where the answer is known: two-site DMRG of the J1-J2 (J2 = 0.5) Heisenberg model on a 4 x 6 cylinder
(N = 24 sites, compressed MPO bond w = 20) with the bond dimension ramped to chi = 4096 -- the central bond of a 24-site chain is then exact
(2^12), so the energy must reproduce exact diagonalisation (tools/ed_j1j2.py, Sz = 0 sector, independent of any MPS code).  The last passes
run the C5 shapes on ONE GPU: Theta (4096, 2, 2, 4096), H_eff applications of 8.9e13 flop, truncated SVDs of 8192 x 8192.
  python tools/bench_c5_exact.py [chi_max]          (one JSON line per maxdim; a "pass" is one sweep direction over all bonds)"""
import ctypes as C
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tensornetworks.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import tnb200
from tnb200.mpo import MPO
from tnb200._lib import check, tn_lanczos_t

E_ED = float(os.environ.get("E_ED", "nan"))        # tools/ed_j1j2.py 4 6
LX, LY = int(os.environ.get("LX", "4")), 6
N = LX * LY
chi_max = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = tnb200.Context.default()
gH = MPO(N, 2, tnb200.models.j1j2_cylinder_terms(LX, LY), ctx=ctx)
tens = tnb200.models.random_canonical_mps(N, 2, 16, seed=7)
g = tnb200.GMPS(1, 2, tens, 1)
g.movecenter(1)
Hs = tnb200.ProjMPS(g, gH, g, center=1)
direction = False
plan = [(64, 4), (256, 2), (1024, 2)]
if chi_max >= 2048:
    plan.append((2048, int(os.environ.get("PASSES_2048", "2"))))
if chi_max >= 4096:
    plan.append((4096, int(os.environ.get("PASSES_4096", "2"))))
for chi, passes in plan:
    if chi > chi_max:
        break
    res = []
    for _ in range(passes):
        e, mb = C.c_double(), C.c_int64()
        c0 = ctx.counters()
        t0 = time.perf_counter()
        check(g.lib.tn_dmrg_sweep(g.h, Hs.h, int(direction), tn_lanczos_t(3, 2, 1e-14), tnb200.Trunc(0.0, chi, 1), C.byref(e), C.byref(mb)))
        dt = time.perf_counter() - t0
        c1 = ctx.counters()
        res.append(dict(energy=e.value, maxbond=mb.value, seconds=dt, heff_applications=c1["matvecs"] - c0["matvecs"], svds=c1["svds"] - c0["svds"],
                        launches=c1["launches"] - c0["launches"], rel_err_vs_ed=(abs(e.value - E_ED) / abs(E_ED) if E_ED == E_ED else None)))
        direction = not direction
        print(json.dumps(dict(config=f"J1-J2 {LX}x{LY} cylinder N={N} w=20 two-site DMRG cutoff=0", maxdim=chi, **res[-1])), flush=True)
