"""examples/thermal.jl of the reference on the B200 path: the identity MPO evolved in imaginary time to beta / 2 with Trotter gates
(rank-2 applygates!), then energy = trace(H, adjoint(U), U) / trace(adjoint(U), U).   python examples/thermal.py
(tnb200.evolve.thermal_energy evaluates the ratio on doubled sites; it had not been run on a GPU when this file was written.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensornetworks.jl_b200"))
import tnb200  # noqa: E402
from tnb200 import models  # noqa: E402
from tnb200.evolve import trotterize, thermal_energy  # noqa: E402

N, h, J, beta, dt = 40, 1.0, 1.0, 4.0, 5e-3
terms = [([models.X], [i], h) for i in range(1, N + 1)] + [([models.Z, models.Z], [i, i + 1], J) for i in range(1, N)]
rows_sites, rows_gates = trotterize(N, 2, [(o, s, -c) for o, s, c in terms], dt)          # trotterize(sh, -1*H, dt)
gates = tnb200.GateList(2, rows_sites, rows_gates)
U = tnb200.GMPS(2, 2, [np.eye(2).reshape(1, 2, 2, 1) for _ in range(N)], 0)            # productMPO(sh, ["id" ...])
U.movecenter(1)
for step in range(int(round(beta / 2 / dt))):
    tnb200.applygates(U, gates, cutoff=1e-10, maxdim=64)
    if (step + 1) % 50 == 0:
        print("beta = %.3f, maxbonddim = %d" % (2 * (step + 1) * dt, U.maxbonddim()))
print("energy per site at beta = %.1f:" % beta, thermal_energy(U, models.tfim_mpo(N, h, 0.0, J)).real / N)
