"""approximategps.jl_b200 -- B200-native SVGP-ELBO / Laplace hot path behind the reference's interface.

The directory name follows the repository convention (``<reference>_b200``); because it contains a
dot it is imported through the ``agp_b200`` loader module at the repository root:

    import agp_b200 as agp
    agp.elbo(agp.SparseVariationalApproximation(fz, q), fx, y)

Contents: ``csrc/`` (CUDA kernels + the C ABI of include/agp.h, built into ``libagp_b200.so``),
``_lib.py`` (ctypes prototypes), ``api.py`` / ``laplace_api.py`` (host-side mirror of the
reference's Julia interface for this path).
"""
from ._lib import AgpError, DomainError, PosDefException, load_library, LIB_PATH, SYMBOLS  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import _prior_kl  # noqa: F401
from .sharding import attach_communicator, shard_range  # noqa: F401
from .laplace_api import (  # noqa: F401
    LaplaceCacheView,
    LaplaceGradient,
    LaplaceObjectiveCache,
    LaplacePosterior,
    LaplaceResult,
    LaplaceStepResult,
    build_laplace_objective,
    build_laplace_objective_,
    laplace_approx_lml_and_gradient,
    laplace_f_and_lml,
    laplace_f_cov,
    laplace_steps,
    newton_inner_loop,
    rrule_newton_inner_loop,
    frule_newton_inner_loop,
    laplace_lml,
    laplace_lml_and_grad_K,
)
