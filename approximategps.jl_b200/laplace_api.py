"""Host-side mirror of src/LaplaceApproximationModule.jl's entry points (paths relative to
/root/reference): ``approx_lml`` :58-60, ``posterior`` :39-48, ``build_laplace_objective(!)`` :77-132,
``laplace_f_and_lml`` :140-145, ``laplace_lml`` :152-165, ``_check_laplace_inputs`` :167-179.

Only argument checking and packing happens here; the Newton loop, ``_laplace_lml`` and their reverse
pass run in libagp_b200.so (agp_laplace_f_and_lml).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L

_FIELDS = {"W": 0, "Wsqrt": 1, "d_loglik": 2, "a": 3, "f": 4, "B_ch_L": 5, "fnew": 6, "loglik": 7, "newton_Wsqrt": 8, "newton_d_loglik": 9}


class LaplaceCacheView:
    """``LaplaceCache`` (Laplace.jl:181-199) living on the device; fields are fetched on first access."""

    def __init__(self, lib, handle, n, owning):
        self._lib, self._h, self.n, self._owning, self._memo = lib, handle, n, owning, {}

    def __getattr__(self, name):
        if name.startswith("_") or name not in _FIELDS:
            raise AttributeError(name)
        if name not in self._memo:
            fld = _FIELDS[name]
            out = np.zeros((self.n, self.n), order="F") if fld == 5 else np.zeros(1 if fld == 7 else self.n)
            L.check(self._lib.agp_laplace_cache_fetch(self._h, fld, L.dptr(out)))
            self._memo[name] = float(out[0]) if fld == 7 else out
        return self._memo[name]

    def f_cov(self) -> np.ndarray:
        """``laplace_f_cov(cache)`` (Laplace.jl:376-386), computed on the device (agp_laplace_f_cov)."""
        if "f_cov" not in self._memo:
            out = np.zeros((self.n, self.n), order="F")
            L.check(self._lib.agp_laplace_f_cov(self._h, L.dptr(out)))
            self._memo["f_cov"] = out
        return self._memo["f_cov"]

    def lml_approx(self) -> float:
        """``_laplace_lml(cache.f, cache)`` (Laplace.jl:250-254)."""
        if "lml" not in self._memo:
            out = np.zeros(1)
            L.check(self._lib.agp_laplace_cache_lml(self._h, L.dptr(out)))
            self._memo["lml"] = float(out[0])
        return self._memo["lml"]

    def close(self):
        if self._owning and self._h:
            self._lib.agp_laplace_cache_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


@dataclass
class LaplaceGradient:
    """Structural tangent of ``approx_lml`` w.r.t. the kernel of ``lfx.fx.f`` (and the inputs)."""

    variance: float
    inv_lengthscale: np.ndarray
    linear_c: float
    X: np.ndarray
    comp_variance: np.ndarray = None        # kernel sums / products: per component
    comp_inv_lengthscale: np.ndarray = None


@dataclass
class LaplaceResult:
    f: np.ndarray
    lml: float
    steps: int
    converged: bool
    dK: np.ndarray | None = None
    grad: LaplaceGradient | None = None
    cache: LaplaceCacheView | None = None


def _run(ctx, *, K=None, kernel=None, X=None, jitter=0.0, y, lik, f_init=None, maxiter=100, callback=None, want_dK=False, want_grad=False,
         want_cache=False) -> LaplaceResult:
    from .api import default_context

    ctx = ctx or default_context()
    lib = ctx.lib
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = len(y)
    pr, rs = L.AgpLaplaceProblem(), L.AgpLaplaceResult()
    keep = [y]
    pr.n = n
    if K is not None:
        K = np.asfortranarray(K, dtype=np.float64)
        if K.shape != (n, n):
            raise AssertionError("length(ys) == length(lfx.fx)")  # Laplace.jl:172
        pr.K = L.dptr(K)
        keep.append(K)
    else:
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        if X.shape[0] != n:
            raise AssertionError("length(ys) == length(lfx.fx)")  # Laplace.jl:172
        from .api import agp_kernel_struct

        kk, kkeep = agp_kernel_struct(kernel)
        ils = kkeep[0]
        pr.kernel, pr.X, pr.D, pr.jitter = C.pointer(kk), L.dptr(X), X.shape[1], float(jitter)
        keep += [X, kkeep, kk]
    pr.y = L.dptr(y)
    pr.lik = L.AgpLikelihood(lik.kind, float(lik.sigma2))
    if f_init is not None:
        f0 = np.ascontiguousarray(f_init, dtype=np.float64)
        assert f0.shape == (n,)
        pr.f_init = L.dptr(f0)
        keep.append(f0)
    if maxiter < 1:
        raise AssertionError("maxiter >= 1")  # Laplace.jl:257
    pr.maxiter = int(maxiter)
    cb_error = []
    if callback is not None:
        def _cb(user, it, handle):
            # an exception raised by the user's callback propagates out of the Newton loop in the reference (Laplace.jl:263-265);
            # ctypes would print and swallow it, so it is captured here, the loop is stopped (non-zero status) and it is
            # re-raised once agp_laplace_f_and_lml has returned
            try:
                view = LaplaceCacheView(lib, C.c_void_p(handle), n, owning=False)
                callback(view.fnew, view)
            except BaseException as e:  # noqa: BLE001
                cb_error.append(e)
                return 1
            return 0

        cb = L.NEWTON_CALLBACK(_cb)
        pr.callback = cb
        keep.append(cb)
    f_opt = np.zeros(n)
    rs.f_opt = L.dptr(f_opt)
    dK = np.zeros((n, n), order="F") if want_dK else None
    rs.dK = L.dptr(dK)
    grad = None
    if want_grad:
        if K is not None:
            raise ValueError("kernel-parameter gradients need the (kernel, X) form")
        sc = np.zeros(2)
        dils, dX = np.zeros(ils.size), np.zeros_like(X)
        rs.dvariance, rs.dlinear_c = sc[0:1].ctypes.data_as(L.c_double_p), sc[1:2].ctypes.data_as(L.c_double_p)
        rs.dinv_lengthscale, rs.dX = L.dptr(dils), L.dptr(dX)
        ncomp = len(kernel.components)
        dcv, dcs = np.zeros(ncomp), np.zeros(ncomp)
        if ncomp:
            rs.dcomp_variance, rs.dcomp_inv_lengthscale = L.dptr(dcv), L.dptr(dcs)
    h = C.c_void_p()
    status = lib.agp_laplace_f_and_lml(ctx.h, C.byref(pr), C.byref(rs), C.byref(h) if want_cache else None)
    if cb_error:
        raise cb_error[0]
    L.check(status)
    if want_grad:
        grad = LaplaceGradient(float(sc[0]), dils, float(sc[1]), dX, dcv, dcs)
    cache = LaplaceCacheView(lib, h, n, owning=True) if want_cache else None
    return LaplaceResult(f_opt, rs.lml, rs.steps, bool(rs.converged), dK, grad, cache)


def _check_laplace_inputs(lfx, ys, f_init=None, maxiter=100, callback=None):
    """``_check_laplace_inputs`` (Laplace.jl:167-179): zero prior mean, matching lengths."""
    fx = lfx.fx
    if fx.f.mean_const != 0.0:
        raise AssertionError("mean(lfx.fx) == zero(mean(lfx.fx))")  # Laplace.jl:171
    if len(ys) != len(fx):
        raise AssertionError("length(ys) == length(lfx.fx)")  # Laplace.jl:172
    if not hasattr(lfx.lik, "kind"):
        raise ValueError(f"unsupported likelihood {lfx.lik!r}")
    if np.ndim(fx.Sigma_y) != 0:
        raise ValueError("lfx.fx.Sigma_y must be an isotropic jitter for the device path")
    return dict(kernel=fx.f.kernel, X=fx.x, jitter=float(fx.Sigma_y), y=ys, lik=lfx.lik, f_init=f_init, maxiter=maxiter, callback=callback)


def laplace_f_and_lml(lfx, ys, *, ctx=None, **newton_kwargs):
    """``laplace_f_and_lml(lfx, ys; newton_kwargs...)`` (Laplace.jl:140-145) -> (f_opt, lml)."""
    r = _run(ctx, **_check_laplace_inputs(lfx, ys, **newton_kwargs))
    return r.f, r.lml


def newton_inner_loop(lik, ys, K, *, f_init=None, maxiter=100, callback=None, ctx=None) -> np.ndarray:
    """``newton_inner_loop(dist_y_given_f, ys, K; f_init, maxiter, callback)`` (Laplace.jl:304-307): the mode f_opt."""
    return _run(ctx, K=K, y=ys, lik=lik, f_init=f_init, maxiter=maxiter, callback=callback).f


def rrule_newton_inner_loop(lik, ys, K, *, f_init=None, maxiter=100, ctx=None):
    """``ChainRulesCore.rrule(newton_inner_loop, dist_y_given_f, ys, K; kwargs...)`` (Laplace.jl:330-369): returns
    ``(f_opt, newton_pullback)``; ``newton_pullback(df_opt)`` is the cotangent of K, ``(Wsqrt .* (B_ch \\ (df_opt ./ Wsqrt))) *
    d_loglik'`` (the cotangents of the likelihood and of ys are ``@not_implemented`` in the reference)."""
    r = _run(ctx, K=K, y=ys, lik=lik, f_init=f_init, maxiter=maxiter, want_cache=True)
    n = len(r.f)

    def newton_pullback(df_opt, dense=True):
        df = np.ascontiguousarray(df_opt, dtype=np.float64)
        assert df.shape == (n,)
        u = np.zeros(n)
        dK = np.zeros((n, n), order="F") if dense else None
        L.check(r.cache._lib.agp_laplace_newton_pullback(r.cache._h, L.dptr(df), L.dptr(u), L.dptr(dK)))
        return dK if dense else (u, np.array(r.cache.newton_d_loglik))

    return r.f, newton_pullback


def frule_newton_inner_loop(dK, lik, ys, K, *, f_init=None, maxiter=100, ctx=None):
    """``ChainRulesCore.frule((_, _, _, dK), newton_inner_loop, dist_y_given_f, ys, K; kwargs...)`` (Laplace.jl:309-328):
    returns ``(f_opt, fdot)`` with ``fdot = (B_ch \\ (Wsqrt .* (dK * d_loglik))) ./ Wsqrt``."""
    r = _run(ctx, K=K, y=ys, lik=lik, f_init=f_init, maxiter=maxiter, want_cache=True)
    n = len(r.f)
    dK = np.asfortranarray(dK, dtype=np.float64)
    assert dK.shape == (n, n)
    fdot = np.zeros(n)
    L.check(r.cache._lib.agp_laplace_newton_pushforward(r.cache._h, L.dptr(dK), L.dptr(fdot)))
    return r.f, fdot


def laplace_f_cov(cache: LaplaceCacheView) -> np.ndarray:
    """``laplace_f_cov(cache)`` (Laplace.jl:376-386): ``Wsqrt^-1 (I - B^-1) Wsqrt^-1``."""
    return cache.f_cov()


@dataclass
class LaplaceStepResult:
    """``LaplaceResult(fnew, cache)`` (Laplace.jl:388-395): one record per Newton step of ``laplace_steps``.  ``q`` is
    ``MvNormal(cache.f, _symmetric(f_cov))``, kept as the pair ``(q_mean, q_cov)``."""

    fnew: np.ndarray
    f_cov: np.ndarray
    q_mean: np.ndarray
    q_cov: np.ndarray
    lml_approx: float
    cache: dict


def laplace_steps(lfx, ys, *, ctx=None, **newton_kwargs):
    """``laplace_steps(lfx, ys; newton_kwargs...)`` (Laplace.jl:409-421): the intermediate approximation of every
    Newton step.  The callback's cache view is only valid during the step, so its fields are copied to the host."""
    if "callback" in newton_kwargs:
        raise TypeError("laplace_steps installs its own callback")  # the reference would hit a duplicate keyword
    res_array = []

    def store_result(fnew, cache):
        f_cov = cache.f_cov()
        fields = {k: np.array(getattr(cache, k)) for k in ("W", "Wsqrt", "d_loglik", "a", "f", "B_ch_L")}
        res_array.append(LaplaceStepResult(np.array(fnew), f_cov, fields["f"], 0.5 * (f_cov + f_cov.T), cache.lml_approx(), fields))

    _run(ctx, **_check_laplace_inputs(lfx, ys, **newton_kwargs, callback=store_result))
    return res_array


def laplace_lml(*args, ctx=None, f_init=None, maxiter=100, callback=None):
    """``laplace_lml(lfx, ys; ...)`` (Laplace.jl:152-155) or ``laplace_lml(lik, ys, K; f_init, maxiter)`` (:157-160)."""
    if len(args) == 2:
        return _run(ctx, **_check_laplace_inputs(args[0], args[1], f_init, maxiter, callback)).lml
    lik, ys, K = args
    return _run(ctx, K=K, y=ys, lik=lik, f_init=f_init, maxiter=maxiter, callback=callback).lml


def laplace_lml_and_grad_K(lik, ys, K, *, f_init=None, maxiter=100, ctx=None):
    """Value and d/dK of ``laplace_lml(lik, ys, K)``: what ``Zygote.gradient`` returns through
    ``rrule(newton_inner_loop)`` (Laplace.jl:330-369).  Returns a LaplaceResult."""
    return _run(ctx, K=K, y=ys, lik=lik, f_init=f_init, maxiter=maxiter, want_dK=True)


def laplace_approx_lml(la, lfx, ys, ctx=None):
    """``approx_lml(la::LaplaceApproximation, lfx, ys)`` (Laplace.jl:58-60)."""
    return _run(ctx, **_check_laplace_inputs(lfx, ys, **la.newton_kwargs)).lml


def laplace_approx_lml_and_gradient(la, lfx, ys, ctx=None) -> LaplaceResult:
    """``approx_lml`` and its gradient w.r.t. the kernel parameters of ``lfx.fx.f`` (the tangent Zygote
    would return for ``lfx.fx.f.kernel``) and the inputs."""
    return _run(ctx, want_grad=True, **_check_laplace_inputs(lfx, ys, **la.newton_kwargs))


class LaplacePosterior:
    """``ApproxPosteriorGP(la, lfx.fx, cache)`` (Laplace.jl:39-48); ``data`` is the LaplaceCache at f_opt, which stays
    on the device; the prediction methods (Laplace.jl:425-463) run there (agp_laplace_predict)."""

    def __init__(self, approx, fx, result: LaplaceResult, ctx):
        self.approx, self.prior, self.data, self.ctx = approx, fx, result.cache, ctx
        self.f, self.lml, self.steps = result.f, result.lml, result.steps

    def _predict(self, x, y=None, want_mean=False, want_var=False, want_cov=False):
        from .api import _points

        from .api import agp_kernel_struct

        kk, _kkeep = agp_kernel_struct(self.prior.f.kernel)
        Xt = _points(self.prior.x)
        x = _points(x)
        yy = None if y is None else _points(y)
        n1, n2 = len(x), (len(x) if yy is None else len(yy))
        mu = np.zeros(n1) if want_mean else None
        var = np.zeros(n1) if want_var else None
        cov = np.zeros((n1, n2), order="F") if want_cov else None
        L.check(self.data._lib.agp_laplace_predict(self.data._h, C.byref(kk), L.dptr(Xt), Xt.shape[1], L.dptr(x), n1, L.dptr(yy), n2, L.dptr(mu), L.dptr(var),
                                                   L.dptr(cov)))
        return mu, var, cov

    def mean_and_var(self, x):  # Laplace.jl:433-437
        mu, var, _ = self._predict(x, want_mean=True, want_var=True)
        return mu, var

    def mean_and_cov(self, x):  # Laplace.jl:439-443
        mu, _, cov = self._predict(x, want_mean=True, want_cov=True)
        return mu, cov

    def cov(self, x, y=None):  # Laplace.jl:453-463
        return self._predict(x, y, want_cov=True)[2]


def laplace_posterior(la, lfx, ys, ctx=None):
    """``posterior(la, lfx, ys)`` (Laplace.jl:39-48)."""
    r = _run(ctx, want_cache=True, **_check_laplace_inputs(lfx, ys, **la.newton_kwargs))
    return LaplacePosterior(la, lfx.fx, r, ctx)


class LaplaceObjectiveCache:  # Laplace.jl:91-93
    def __init__(self, f=None):
        self.f = f


class _LaplaceObjective:
    """The closure returned by ``build_laplace_objective`` (Laplace.jl:95-132): ``objective(args...)``
    returns ``-approx_lml``; ``objective.cache.f`` is the warm-start vector."""

    def __init__(self, cache, build_latent_gp, xs, ys, newton_warmstart, newton_callback, newton_maxiter, ctx):
        self.cache, self._build, self._xs, self._ys = cache, build_latent_gp, xs, ys
        self._warm, self._cb, self._maxiter, self._ctx = newton_warmstart, newton_callback, newton_maxiter, ctx
        self._initialize_f = True  # Laplace.jl:104
        self.newton_steps = 0

    def _eval(self, args, want_grad):
        lfx = self._build(*args)(self._xs)  # Laplace.jl:107-108
        n = len(lfx.fx)
        # :109-118 -- `cache.f === nothing` -> mean(lfx.fx) (zeros: _check_laplace_inputs asserts the zero mean); while
        # `initialize_f` is still true (no warm-start store has happened yet) a caller-supplied vector is overwritten IN PLACE
        # with mean(lfx.fx) as well, so build_laplace_objective!(f_init, ...) only fixes the storage, not the first start
        if self.cache.f is None:
            self.cache.f = np.zeros(n)
        elif self._initialize_f:
            self.cache.f[...] = 0.0
        r = _run(self._ctx, want_grad=want_grad, **_check_laplace_inputs(lfx, self._ys, self.cache.f, self._maxiter, self._cb))
        self.newton_steps += r.steps
        if self._warm:
            self.cache.f[...] = r.f  # :122-127 `cache.f .= f_opt` (in place: the caller's vector sees the mode)
            self._initialize_f = False
        return r

    def __call__(self, *args):
        return -self._eval(args, False).lml

    def value_and_gradient(self, *args):
        """(-lml, LaplaceGradient of -lml w.r.t. the kernel built by ``build_latent_gp(args...)``)."""
        r = self._eval(args, True)
        g = r.grad
        return -r.lml, LaplaceGradient(-g.variance, -g.inv_lengthscale, -g.linear_c, -g.X)


def build_laplace_objective(build_latent_gp, xs, ys, *, newton_warmstart=True, newton_callback=None, newton_maxiter=100, ctx=None):
    """``build_laplace_objective(build_latent_gp, xs, ys; kwargs...)`` (Laplace.jl:77-89)."""
    return build_laplace_objective_(LaplaceObjectiveCache(None), build_latent_gp, xs, ys, newton_warmstart=newton_warmstart,
                                    newton_callback=newton_callback, newton_maxiter=newton_maxiter, ctx=ctx)


def build_laplace_objective_(cache, build_latent_gp, xs, ys, *, newton_warmstart=True, newton_callback=None, newton_maxiter=100, ctx=None):
    """``build_laplace_objective!(f_init | cache, ...)`` (Laplace.jl:95-132)."""
    if not isinstance(cache, LaplaceObjectiveCache):
        # build_laplace_objective!(f_init::Vector, ...) (Laplace.jl:85-89): the caller's vector IS the cache storage
        f = cache if (isinstance(cache, np.ndarray) and cache.dtype == np.float64 and cache.flags.c_contiguous) else np.array(cache, dtype=np.float64)
        cache = LaplaceObjectiveCache(f)
    return _LaplaceObjective(cache, build_latent_gp, xs, ys, newton_warmstart, newton_callback, newton_maxiter, ctx)
