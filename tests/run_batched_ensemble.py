"""Helper run as a SUBPROCESS by the GPU test of the batching rounds (so that a dead-lock there cannot hang the test session):
runs the same small QJMC ensemble with and without TN_QJMC_BATCH=1 and prints both results as one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tensornetworks.jl_b200")]
import tnb200  # noqa: E402
from tnb200 import models  # noqa: E402

N, d, dt, steps, chi = 8, 2, 0.02, 30, 8
gamma = 0.9
onsite = -1j * (1.0 * models.X + 2.0 * models.Z) - 0.5 * gamma * (models.SM.conj().T @ models.SM)
bond = -1j * 1.0 * np.kron(models.Z, models.Z)
ss, gg = models.trotter_gates(N, onsite, bond, dt, evol="imag", order=2)
tens = models.random_canonical_mps(N, d, chi, seed=5)
ids = [3, 11, 7, 0, 42, 5, 9]
out = {}
for mode in ("0", "1"):
    os.environ["TN_QJMC_BATCH"] = mode
    nj, jumps, times, obs = tnb200.qjmc_ensemble(tens, 1, ss, gg, list(range(1, N + 1)), [models.SM] * N, [np.sqrt(gamma)] * N, steps, dt,
                                                 ids, workers=4, seed=9, obs_op=models.Z, save_every=5, cutoff=1e-12, maxdim=chi)
    out[mode] = dict(nj=nj.tolist(), jumps=[jumps[k, :nj[k]].tolist() for k in range(len(ids))],
                     times=[times[k, :nj[k]].tolist() for k in range(len(ids))], obs=np.real(obs).tolist())
print(json.dumps(out))
