"""B200-native evaluator of OpenQuantumBase.jl's master-equation right-hand side.

Hand-written sm_100a CUDA kernels behind the C ABI of ``include/oqb.h`` (``liboqb_b200.so``), with a
host-side mirror of the reference's operator interface in ``host.py``.  No CPU fallback: the
package raises if the CUDA library has not been built or no B200 is present.
"""
from . import _lib  # noqa: F401
from .host import (  # noqa: F401
    ZERO_LAMBSHIFT, AdiabaticFrameHamiltonian, ConstantCouplings, ConstantDenseHamiltonian, ConstantHamiltonian, ConstantSparseHamiltonian, Context, Hamiltonian, DaviesGenerator, DenseHamiltonian, DensityStepper, DeviceBatch,
    DiffEqLiouvillian, GapIndices, Lindblad, LindbladLiouvillian, ODEParams, Ohmic, OhmicBath,
    OneSidedAMELiouvillian, QuadraticBSpline, RedfieldLiouvillian, SparseHamiltonian, TrajectoryEnsemble, TrajectoryStepper, ensemble_population, expect, expect_diag,
    build_lambshift, get_context, haml_eigs, lambshift, unit_scale, update_cache_, update_vectorized_cache_,
)
from .build import build  # noqa: F401
