"""Exports the seeded INPUTS of the committed golden file (tests/golden/hotpath_golden.npz) as raw column-major binaries + a JSON
manifest, so that the real reference can be run on them where Julia exists:

    python tests/golden/export_inputs_for_julia.py            # -> tests/golden/julia_io/inputs/
    julia --project=/path/to/TensorNetworks.jl tests/golden/make_golden_reference.jl
    python -m pytest tests/test_reference_fixtures.py         # compares tests/golden/julia_io/outputs/ with the oracle and the npz

(The build image has no Julia, so the outputs directory is not committed: until somebody runs the second line, parity stays
"unpinned by the reference" -- see DESIGN.md section 2.)"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    z = np.load(os.path.join(HERE, "hotpath_golden.npz"))
    out = os.path.join(HERE, "julia_io", "inputs")
    os.makedirs(out, exist_ok=True)
    manifest = {}
    for name in z.files:
        x = np.asarray(z[name])
        kind = "c128" if np.iscomplexobj(x) else ("f64" if x.dtype.kind == "f" else "i64")
        y = x.astype({"c128": np.complex128, "f64": np.float64, "i64": np.int64}[kind])
        with open(os.path.join(out, name + ".bin"), "wb") as f:
            f.write(np.asfortranarray(y).tobytes(order="F"))
        manifest[name] = {"dtype": kind, "shape": list(x.shape)}
    json.dump(manifest, open(os.path.join(out, "manifest.json"), "w"), indent=1)
    print("wrote", len(manifest), "arrays to", out)


if __name__ == "__main__":
    main()
