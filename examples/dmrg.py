"""examples/dmrg.jl of the reference on the B200 path: ground state of the transverse-field Ising chain
H = sum_i (1.0 x_i + 0.05 z_i) + 1.2 sum_i z_i z_{i+1}, N = 100, two-site DMRG with cutoff 1e-12 and maxdim 32.
Run on a machine with a B200:  python examples/dmrg.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensornetworks.jl_b200"))
import tnb200  # noqa: E402
from tnb200 import models  # noqa: E402
from tnb200.mpo import MPO  # noqa: E402

N = 100
terms = [([models.X], [i], 1.0) for i in range(1, N + 1)] + [([models.Z], [i], 0.05) for i in range(1, N + 1)]
terms += [([models.Z, models.Z], [i, i + 1], 1.2) for i in range(1, N)]
H = MPO(N, 2, terms)                                         # MPO(sh, H): host assembly + device compression (w = 3)
rng = np.random.default_rng(1234)
psi = tnb200.GMPS(1, 2, [rng.standard_normal((1, 2, 1)) for _ in range(N)], 0)       # randomMPS(2, N, 1)
psi, energy = tnb200.dmrg(psi, H, nsites=2, cutoff=1e-12, maxdim=32, maxsweeps=100, verbose=True)
print("energy per site", energy / N, "max bond dimension", psi.maxbonddim())
mags = np.real(psi.expect([models.Z] * N, list(range(1, N + 1))))
print("<z> in the bulk", mags[N // 2])
