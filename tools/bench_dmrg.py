"""DMRG sweep-time measurements and parity at the BASELINE.json configs:
  C1  examples/dmrg.jl: TFIM N=100, maxdim=32, cutoff=1e-12, random chi=1 start  (GPU vs oracle, energy + s/sweep)
  C2  XXZ N=100 (w=5), maxdim ramp 64 -> ... -> CHI_MAX, 2 sweeps each; s/sweep at every chi; oracle alongside up to ORACLE_CHI
One JSON line per measurement."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tensornetworks.jl_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import oracle  # noqa: E402
import tnb200  # noqa: E402
from oracle.gmps import GMPS as OG  # noqa: E402
from tnb200._lib import check, tn_lanczos_t  # noqa: E402

what = sys.argv[1:] or ["c1", "c2"]
CHI_MAX = int(os.environ.get("CHI_MAX", "512"))
ORACLE_CHI = int(os.environ.get("ORACLE_CHI", "64"))
CUTOFF = float(os.environ.get("CUTOFF", "1e-12"))
CHI_MIN = int(os.environ.get("CHI_MIN", "64"))
NSWEEPS = int(os.environ.get("NSWEEPS", "2"))      # sweeps per maxdim (the bond dimension at most doubles per sweep)
ctx = tnb200.Context.default()


def gpu_sweeps(psi, Hs, nsweeps, direction, maxdim, cutoff=CUTOFF):
    out = []
    for _ in range(nsweeps):
        e, mb = C.c_double(), C.c_int64()
        c0 = ctx.counters()
        t0 = time.perf_counter()
        check(psi.lib.tn_dmrg_sweep(psi.h, Hs.h, int(direction), tn_lanczos_t(3, 2, 1e-14), tnb200.Trunc(cutoff, maxdim, 1), C.byref(e), C.byref(mb)))
        dt = time.perf_counter() - t0
        c1 = ctx.counters()
        out.append(dict(energy=e.value, maxbond=mb.value, seconds=dt, launches=c1["launches"] - c0["launches"],
                        matvecs=c1["matvecs"] - c0["matvecs"], svds=c1["svds"] - c0["svds"]))
        direction = not direction
    return out, direction


if "c1" in what:
    N = 100
    Ht = tnb200.models.tfim_mpo(N)
    p0 = oracle.randomMPS(2, N, 1, np.random.default_rng(1234))
    hist_g, hist_o = [], []
    g = tnb200.GMPS.from_host(p0)
    gH = tnb200.GMPS(2, 2, Ht)
    t0 = time.perf_counter()
    g, Eg = tnb200.dmrg(g, gH, nsites=2, cutoff=1e-12, maxdim=32, maxsweeps=100, history=hist_g)
    tg = time.perf_counter() - t0
    po = p0.copy()
    oH = OG(2, 2, Ht, 0)
    t0 = time.perf_counter()
    po, Eo = oracle.dmrg(po, oH, nsites=2, cutoff=1e-12, maxdim=32, maxsweeps=100, history=hist_o)
    to = time.perf_counter() - t0
    print(json.dumps(dict(config="C1 examples/dmrg.jl TFIM N=100 maxdim=32 cutoff=1e-12", gpu_energy=Eg, oracle_energy=float(np.real(Eo)),
                          rel_diff=abs(Eg - np.real(Eo)) / abs(np.real(Eo)), gpu_sweeps=len(hist_g), oracle_sweeps=len(hist_o),
                          gpu_s_per_sweep=tg / len(hist_g), oracle_s_per_sweep=to / len(hist_o), gpu_maxbond=hist_g[-1][2],
                          oracle_maxbond=hist_o[-1][2], cpu_threads=os.cpu_count())), flush=True)

if "c2" in what:
    from oracle.dmrg import dmrg_sweeps
    N = 100
    Hx = tnb200.models.xxz_mpo(N, 1.0)
    p0 = oracle.randomMPS(2, N, 8, np.random.default_rng(1234))
    g = tnb200.GMPS.from_host(p0)
    gH = tnb200.GMPS(2, 2, Hx)
    g.movecenter(1)
    Hs = tnb200.ProjMPS(g, gH, g, center=1)
    po = p0.copy()
    oH = OG(2, 2, Hx, 0)
    po.movecenter(1)
    oHs = oracle.ProjMPSSum([oracle.ProjMPS([po, oH, po], rank=2)])
    direction = False
    chi = CHI_MIN
    while chi <= CHI_MAX:
        res, direction2 = gpu_sweeps(g, Hs, NSWEEPS, direction, chi)
        line = dict(config="C2 XXZ N=100 w=5 two-site DMRG", maxdim=chi, cutoff=CUTOFF, gpu=res)
        if chi <= ORACLE_CHI:
            ho = []
            t0 = time.perf_counter()
            dmrg_sweeps(po, oHs, maxdim=chi, cutoff=CUTOFF, minsweeps=2, maxsweeps=2, history=ho)   # 2 sweeps, reference contraction order
            line["oracle"] = [dict(energy=h[1], maxbond=h[2]) for h in ho]
            line["oracle_s_per_sweep"] = (time.perf_counter() - t0) / 2
            line["energy_rel_diff"] = abs(res[-1]["energy"] - ho[-1][1]) / abs(ho[-1][1])
        direction = direction2
        print(json.dumps(line), flush=True)
        chi *= 2
