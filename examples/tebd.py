"""examples/tebd.jl of the reference on the B200 path: imaginary-time TEBD of the transverse-field Ising chain towards its
ground state, energy and log-norm observers.   python examples/tebd.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensornetworks.jl_b200"))
import tnb200  # noqa: E402
from tnb200 import models  # noqa: E402
from tnb200.evolve import tebd, TEBDEnergy, TEBDNorm  # noqa: E402

N, h, J = 20, 1.0, 1.0
terms = [([models.X], [i], -h) for i in range(1, N + 1)] + [([models.Z, models.Z], [i, i + 1], -J) for i in range(1, N)]
rng = np.random.default_rng(0)
psi = tnb200.GMPS(1, 2, [rng.standard_normal((1, 2, 1)) for _ in range(N)], 0)
psi.movecenter(1)
energy_obs, norm_obs = TEBDEnergy(1e-10), TEBDNorm(1e-10)
# H = -h sum x - J sum zz.  Like the reference, tebd evolves with exp(+dt * (what is passed)), so -H is passed (examples/tebd.jl)
minus_H = [(ops, sites, -c) for ops, sites, c in terms]
psi, energy = tebd(psi, minus_H, dt=0.01, tmax=10.0, save=0.1, observers=[energy_obs, norm_obs], cutoff=1e-12, maxdim=32, verbose=True)
print("final <-H> reported by tebd:", energy, " => ground energy estimate", -energy)
