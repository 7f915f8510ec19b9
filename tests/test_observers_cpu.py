"""Host-side observer logic of the tebd / qjmc front-ends (tnb200/evolve.py) against the convergence rules of the reference
(tebd.jl:124-187); no GPU involved."""
import numpy as np


def test_tebd_norm_observer_converges_on_constant_slope():
    from tnb200.evolve import TEBDNorm
    ob = TEBDNorm(tol=1e-6)
    for k in range(2):
        ob.measure(0.1 * k, None, -1.3 * 0.1 * k + 0.01 * np.exp(-k), None)
        assert not ob.checkdone()
    for k in range(2, 40):
        ob.measure(0.1 * k, None, -1.3 * 0.1 * k + 0.01 * np.exp(-k), None)
        if ob.checkdone():
            break
    assert ob.checkdone() and 5 < len(ob.times) < 40
    flat = TEBDNorm(tol=1e-6)                 # |E1 + E2| < 1e-8 branch: absolute difference
    for k in range(3):
        flat.measure(float(k), None, 2.0, None)
    assert flat.checkdone()


def test_tebd_energy_observer():
    from tnb200.evolve import TEBDEnergy
    ob = TEBDEnergy(tol=1e-8)
    vals = [-10.0 - np.exp(-2.0 * k) for k in range(30)]
    n = None
    for k, v in enumerate(vals):
        ob.measure(0.1 * k, None, 0.0, v)
        if ob.checkdone():
            n = k
            break
    assert n is not None and abs(2 * (vals[n - 1] - vals[n]) / (vals[n - 1] + vals[n])) < 1e-8
    assert abs(2 * (vals[n - 2] - vals[n - 1]) / (vals[n - 2] + vals[n - 1])) >= 1e-8 or n == 2
    zero = TEBDEnergy(tol=1e-3)               # near-zero energies: absolute difference
    for k in range(3):
        zero.measure(float(k), None, 0.0, 1e-10 * k)
    assert zero.checkdone()


def test_qjmc_activity_observer():
    from tnb200.evolve import QJMCActivity
    ob = QJMCActivity()
    ob.measure(0.5, None, [1, 3, 2], [0.1, 0.2, 0.4])
    assert ob.time == 0.5 and ob.jumps == 3
