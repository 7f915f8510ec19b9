"""tnb200 -- host-side mirror of the TensorNetworks.jl API for the MPS hot path,
dispatching into the CUDA shared library ``libtnb200.so`` (sm_100a).

The names follow the reference (``GMPS``, ``ProjMPS``, ``ProjMPSSum``, ``movecenter``, ``applygates``,
``dmrg``, ``vmps``, ``tebd``, ``qjmc_simulation`` ...); site numbers are 1-based; host tensors
are NumPy complex128 in Julia's column-major order.  Tensors live in HBM between
calls; only scalars, observables and explicitly downloaded tensors cross PCIe."""
from ._lib import TNError, load, LIB_PATH  # noqa: F401
from .api import (Context, GMPS, ProjMPS, ProjMPSSum, GateList, applyMPO, svd, svd_split, svd_batched, contract_strided, dmrg, vmps, vmps_sweeps, tebd,  # noqa: F401
                  applygates, qjmc_simulation, qjmc_ensemble, inner, Trunc)
from . import models, mpo, evolve, sharded  # noqa: F401
