"""Host-side mirror of src/LaplaceApproximationModule.jl's entry points (Laplace.jl:39-165, :77-132)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _kernel_matrix_host(kernel, x, jitter):
    raise NotImplementedError


class LaplacePosterior:
    """``ApproxPosteriorGP(la, lfx.fx, cache)`` -- Laplace.jl:39-48; prediction :425-463."""

    def __init__(self, approx, fx, cache_fields, ctx):
        self.approx, self.prior, self.data, self.ctx = approx, fx, cache_fields, ctx


def laplace_f_and_lml(K, y, lik, f_init=None, maxiter=100, want_grad=False, ctx=None):
    """``laplace_f_and_lml`` (Laplace.jl:140-145) on a dense ``K = cov(fx)``; returns
    (f_opt, lml, n_newton_steps[, dlml/dK])."""
    from .api import default_context

    ctx = ctx or default_context()
    K = np.asfortranarray(K, dtype=np.float64)
    n = K.shape[0]
    assert K.shape == (n, n)
    y = np.ascontiguousarray(y, dtype=np.float64)
    assert len(y) == n  # Laplace.jl:172
    assert maxiter >= 1  # Laplace.jl:257
    f0 = None if f_init is None else np.ascontiguousarray(f_init, dtype=np.float64)
    f_opt = np.zeros(n)
    lml = C.c_double()
    steps = C.c_int32()
    dK = np.zeros((n, n), order="F") if want_grad else None
    lk = L.AgpLikelihood(lik.kind, float(lik.sigma2))
    L.check(ctx.lib.agp_laplace_f_and_lml(ctx.h, L.dptr(K), n, L.dptr(y), C.byref(lk), L.dptr(f0), int(maxiter), L.dptr(f_opt),
                                          C.byref(lml), C.byref(steps), L.dptr(dK), None))
    if want_grad:
        return f_opt, lml.value, steps.value, dK
    return f_opt, lml.value, steps.value


def laplace_lml_and_grad(K, y, lik, f_init=None, maxiter=100, ctx=None):
    f_opt, lml, steps, dK = laplace_f_and_lml(K, y, lik, f_init, maxiter, True, ctx)
    return lml, dK, f_opt, steps


def laplace_approx_lml(la, lfx, ys, **kwargs):
    raise NotImplementedError("Laplace host API is completed together with the device path")


def laplace_posterior(la, lfx, ys, ctx):
    raise NotImplementedError("Laplace host API is completed together with the device path")


def build_laplace_objective(build_latent_gp, xs, ys, **kwargs):
    raise NotImplementedError("Laplace host API is completed together with the device path")
