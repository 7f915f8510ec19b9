"""examples/qjmc.jl of the reference on the B200 path: quantum-jump Monte Carlo of a dissipative Ising chain,
H = sum (x + 20 z) + 10 sum zz with decay sqrt(0.1) s- on every site; one trajectory with observers, then a small ensemble on
worker streams.   python examples/qjmc.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensornetworks.jl_b200"))
import tnb200  # noqa: E402
from tnb200 import models  # noqa: E402

N, gamma, dt, steps, chi = 21, 0.1, 5e-3, 200, 64
onsite = -1j * (1.0 * models.X + 20.0 * models.Z) - 0.5 * gamma * (models.SM.conj().T @ models.SM)     # -iH - 1/2 sum L^dag L (qjmc.jl:9-26)
bond = -1j * 10.0 * np.kron(models.Z, models.Z)
rows_sites, rows_gates = models.trotter_gates(N, onsite, bond, dt, evol="imag", order=2)
up, dn = np.array([1.0, 0.0]), np.array([0.0, 1.0])
tensors = [(up if i % 2 == 0 else dn).reshape(1, 2, 1) for i in range(N)]                                # Z2 initial state
psi = tnb200.GMPS(1, 2, tensors, 0)
psi.movecenter(1)
gates = tnb200.GateList(2, rows_sites, rows_gates)
jump_args = (list(range(1, N + 1)), [models.SM] * N, [np.sqrt(gamma)] * N)
jumps, times, zs = tnb200.qjmc_simulation(psi, gates, *jump_args, steps, dt, seed=0, trajectory=0, obs_op=models.Z, save_every=20,
                                          cutoff=1e-10, maxdim=chi)
print("jumps:", len(jumps), " <z> at the end:", np.round(np.real(zs[-1]), 3))
nj, _, _, obs = tnb200.qjmc_ensemble(tensors, 1, rows_sites, rows_gates, *jump_args, steps, dt, list(range(16)), workers=8, seed=0,
                                     obs_op=models.Z, save_every=steps, cutoff=1e-10, maxdim=chi)
print("ensemble of 16: mean jumps", nj.mean(), " mean magnetisation", float(np.real(obs[:, -1, :]).mean()))
