"""CPU oracle for the ApproximateGPs.jl SVGP-ELBO / Laplace hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker (or as the timed CPU baseline), never as the thing shipped.  The
product path (``approximategps.jl_b200``) fails loudly when the CUDA library is
missing and never routes through this package.

What it is: a NumPy float64 restatement, operation by operation, of the
reference's algorithm for the path named by BASELINE.json ``north_star``:

* ``oracle.kernels``      KernelFunctions/Distances semantics reached through
                          ``cov(f.prior, z, x)`` (SVA.jl:216) and ``cov(fx)``
                          (Laplace.jl:174)  [third-party, un-vendored, no pinned
                          version: Project.toml:15 has no compat bound]
* ``oracle.likelihoods``  GPLikelihoods 0.4 ``expected_loglikelihood`` (called at
                          SVA.jl:355) and Distributions ``logpdf`` for
                          Gaussian / Bernoulli-logit / Poisson-exp
* ``oracle.svgp``         src/SparseVariationalApproximationModule.jl:115-136,
                          160-187, 215-253, 307-373 + src/utils.jl:15-20, plus a
                          hand-derived reverse pass (the reference uses Zygote)
* ``oracle.laplace``      src/LaplaceApproximationModule.jl:140-276, 330-369

Pinning status:

* PINNED by reference goldens / known answers (tests/test_oracle_*.py):
  the Laplace L-BFGS optimum ``[7.709076337653239, 1.51820292019697]``
  (test/LaplaceApproximationModule.jl:168) and Nelder-Mead optimum (:159) on the
  fixed 48-point dataset of src/TestUtils.jl:13-37; Laplace AD-vs-FD (:41-54);
  Gaussian-likelihood Laplace == exact GPR (src/TestUtils.jl:99-108); the SVGP
  properties of test/SparseVariationalApproximationModule.jl:61-69, 87-96,
  126-133 (Centered == NonCentered, ELBO <= logpdf, SVGP(Z=X, q*) == exact GPR
  to 1e-10, LatentGP+Gaussian ELBO == FiniteGP ELBO to 1e-10).
* PARITY UNPINNED (no reference test or golden constrains it; the reference
  cannot be executed here because Julia is not installed and its dependencies
  are not vendored): Gauss-Hermite expected log-likelihood for Bernoulli and
  Poisson, the analytic Poisson expectation, ``num_data`` rescaling,
  Matern52/Linear kernels and D>1 inputs in SVGP, SVGP gradients w.r.t. Z and
  kernel parameters, Float32.  For those the oracle follows SURVEY.md
  Appendix A and is cross-checked against torch.float64 autograd of the same
  forward pass, 5-point finite differences and mpmath (tests/).
"""

from . import kernels, likelihoods, svgp, laplace  # noqa: F401
