"""CPU oracle for the TensorNetworks.jl MPS hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy/SciPy restatement (column-major semantics, complex128,
LAPACK ``zgesdd``) of the reference's algorithm for the path named in
BASELINE.json: ``src/tensors.jl``, ``src/structures/mps/{gmps,mps,mpo,oplist,
projmps,projmpssum,abstractprojmps,gatelist}.jl`` and
``src/algorithms/mps/{dmrg,tebd,qjmc,vmps}.jl`` (and the gate step of ``itebd.jl``).  Every function cites the reference
file:line it follows.

PARITY UNPINNED BY THE REFERENCE: the reference ships no tests, fixtures or
golden vectors (``test/runtests.jl:4-6`` is an empty testset) and Julia is not
available in the build container, so the reference itself cannot be run.  The
oracle is pinned instead by (i) exact-diagonalisation known answers for the
Hamiltonians the reference's examples build (``tests/test_oracle_kat.py``),
(ii) ``einsum`` identities for every contraction, and (iii) analytic
invariants (orthonormality, discarded weight, Lindblad ensemble averages).

Third-party arithmetic the reference reaches but does not vendor (all unpinned
in ``Project.toml``): TensorOperations (``tensorcontract``) -> ``numpy.tensordot``;
LinearAlgebra/LAPACK ``svd(alg=DivideAndConquer())`` -> ``scipy.linalg.svd(
lapack_driver='gesdd')``; KrylovKit ``eigsolve`` (Lanczos, thick restart)
-> ``oracle.lanczos.eigsolve_lowest`` which restates the published algorithm.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product path
(``tensornetworks.jl_b200/``) never does.
"""
from .tensors import contract, moveidx, combineidxs, uncombineidxs, svd, tensor_exp, trace  # noqa: F401
from .sitetypes import Sitetypes, spinhalf  # noqa: F401
from .oplist import OpList  # noqa: F401
from .gmps import (GMPS, randomMPS, randomGMPS, productMPS, productMPO, inner, applyop)  # noqa: F401
from .mpo import MPO, adjoint, trace, applyMPO  # noqa: F401
from .projmps import ProjMPS, ProjMPSSum  # noqa: F401
from .gatelist import GateList, trotterize, applygate, applygates  # noqa: F401
from .lanczos import eigsolve_lowest  # noqa: F401
from .dmrg import dmrg  # noqa: F401
from .vmps import vmps, vmps_sweeps  # noqa: F401
from .itebd import IGMPS, iMPS, itebd_gate, itebd_apply_gates_mps  # noqa: F401
from .tebd import tebd  # noqa: F401
from .qjmc import qjmc_simulation, qjmc_gates, qjmc_emission_rates  # noqa: F401
