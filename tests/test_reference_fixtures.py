"""Reference-generated fixtures (tests/golden/julia_io/outputs/, written by tests/golden/make_golden_reference.jl where Julia and the
reference are installed) against the oracle-generated golden file.  Skipped while nobody has produced them: the build image has
no Julia, so today parity is pinned by exact-diagonalisation known answers only."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "golden", "julia_io", "outputs")


def _load(name, man):
    dt = {"c128": np.complex128, "f64": np.float64, "i64": np.int64}[man[name]["dtype"]]
    shape = tuple(man[name]["shape"]) or (1,)
    return np.fromfile(os.path.join(OUT, name + ".bin"), dtype=dt).reshape(shape, order="F")


@pytest.mark.skipif(not os.path.exists(os.path.join(OUT, "manifest.json")), reason="no reference-generated fixtures (needs Julia + TensorNetworks.jl)")
def test_oracle_golden_equals_reference_outputs():
    man = json.load(open(os.path.join(OUT, "manifest.json")))
    z = np.load(os.path.join(HERE, "golden", "hotpath_golden.npz"))
    for name in man:
        ref, orc = _load(name, man), np.asarray(z[name])
        ref = ref.reshape(orc.shape, order="F") if ref.size == orc.size else ref
        if name.startswith("dmrg_") and name.endswith("_energy"):
            assert np.max(np.abs(ref - orc) / np.abs(orc)) < 1e-10, name          # north_star: energies to 1e-10 relative
        elif name.endswith("_maxbond"):
            assert np.array_equal(ref.astype(int), orc.astype(int)), name
        elif name.startswith("svd_S_"):
            assert ref.shape == orc.shape and np.max(np.abs(ref - orc)) < 1e-8 * orc.max(), name   # spectra to 1e-8
        else:
            assert np.linalg.norm(ref - orc) <= 1e-12 * np.linalg.norm(orc), name   # contractions: rounding only
