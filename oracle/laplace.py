"""Laplace approximation: NumPy restatement of src/LaplaceApproximationModule.jl (oracle; test
infrastructure).

* ``train_intermediates``   Laplace.jl:201-222 (RW Algorithm 3.1 lines 4-7) + ``LaplaceCache`` :181-199
* ``newton_inner_loop``     Laplace.jl:243-248, :256-276 (stop rule ``isapprox(f, fnew)``:
                            ``norm(f - fnew) <= sqrt(eps) * max(norm(f), norm(fnew))``)
* ``laplace_lml``           Laplace.jl:250-254 via :162-165 (re-runs the intermediates at f_opt)
* ``laplace_f_and_lml``     Laplace.jl:140-145
* ``lml_and_grad_K``        reverse pass: Zygote through ``_laplace_train_intermediates`` /
                            ``_laplace_lml`` plus the hand-written ``rrule(newton_inner_loop)``
                            Laplace.jl:330-369 (dK = (Wsqrt .* (B \\ (df ./ Wsqrt))) * d_loglik')
* ``predict_*``             Laplace.jl:425-463
* fixtures                  src/TestUtils.jl:13-37 (48-point dataset, ``build_latent_gp``)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.linalg import cho_solve, solve_triangular

from .kernels import SE, Kernel, kernelmatrix, kernelmatrix_diag, kernelmatrix_pullback
from .likelihoods import BERNOULLI_LOGIT, GAUSSIAN, POISSON_EXP, Likelihood, logistic, loglik_and_derivs

_RTOL = np.sqrt(np.finfo(np.float64).eps)


@dataclass
class LaplaceCache:  # Laplace.jl:181-199
    K: np.ndarray
    f: np.ndarray
    W: np.ndarray
    Wsqrt: np.ndarray
    loglik: float
    d_loglik: np.ndarray
    B_L: np.ndarray  # lower Cholesky factor of B = I + sqrt(W) K sqrt(W)
    a: np.ndarray


def train_intermediates(lik: Likelihood, y, K, f) -> LaplaceCache:
    ll, d_ll, d2_ll = loglik_and_derivs(lik, f, y)
    W = -d2_ll
    if np.any(W < 0):
        raise ValueError("DomainError: sqrt of negative W")  # Laplace.jl:214
    Wsqrt = np.sqrt(W)
    B = np.eye(len(f)) + (Wsqrt[:, None] * K) * Wsqrt[None, :]
    B_L = np.linalg.cholesky(B)
    b = W * f + d_ll
    a = b - Wsqrt * cho_solve((B_L, True), Wsqrt * (K @ b))
    return LaplaceCache(K, np.asarray(f, dtype=np.float64), W, Wsqrt, float(np.sum(ll)), d_ll, B_L, a)


def isapprox(f, fnew) -> bool:
    return bool(np.linalg.norm(f - fnew) <= _RTOL * max(np.linalg.norm(f), np.linalg.norm(fnew)))


def newton_inner_loop(lik: Likelihood, y, K, f_init=None, maxiter=100, callback=None):
    """Returns (f_opt, cache, n_steps); semantics of ``_newton_inner_loop`` (Laplace.jl:256-276):
    on convergence the *previous* iterate is returned together with the cache computed at it."""
    assert maxiter >= 1
    f = np.zeros(len(y)) if f_init is None else np.array(f_init, dtype=np.float64)
    cache = None
    steps = 0
    for _ in range(maxiter):
        cache = train_intermediates(lik, y, K, f)
        fnew = K @ cache.a
        steps += 1
        if callback is not None:
            callback(fnew, cache)
        if isapprox(f, fnew):
            break
        f = fnew
    return f, cache, steps


def _laplace_lml(f, cache: LaplaceCache) -> float:
    return float(-cache.a @ f / 2.0 + cache.loglik - np.sum(np.log(np.diag(cache.B_L))))


def laplace_lml(lik: Likelihood, y, K, f_opt) -> float:
    return _laplace_lml(f_opt, train_intermediates(lik, y, K, f_opt))


def laplace_f_and_lml(lik: Likelihood, y, K, f_init=None, maxiter=100, callback=None):
    f_opt, _, steps = newton_inner_loop(lik, y, K, f_init, maxiter, callback)
    return f_opt, laplace_lml(lik, y, K, f_opt), steps


def newton_pullback(cache: LaplaceCache, df_opt) -> np.ndarray:
    """``newton_pullback`` of ``rrule(newton_inner_loop)`` (Laplace.jl:330-369): the cotangent of K."""
    u = cache.Wsqrt * cho_solve((cache.B_L, True), np.asarray(df_opt, dtype=np.float64) / cache.Wsqrt)
    return np.outer(u, cache.d_loglik)


def newton_pushforward(cache: LaplaceCache, dK) -> np.ndarray:
    """``frule(newton_inner_loop)`` (Laplace.jl:309-328): the tangent of f_opt for a tangent dK."""
    return cho_solve((cache.B_L, True), cache.Wsqrt * (np.asarray(dK, dtype=np.float64) @ cache.d_loglik)) / cache.Wsqrt


def laplace_f_cov(cache: LaplaceCache) -> np.ndarray:
    """Laplace.jl:376-386: (K^-1 + W)^-1 = Wsqrt^-1 (I - B^-1) Wsqrt^-1."""
    n = len(cache.f)
    Binv = cho_solve((cache.B_L, True), np.eye(n))
    wi = 1.0 / cache.Wsqrt
    return wi[:, None] * (np.eye(n) - Binv) * wi[None, :]


def laplace_steps(lik: Likelihood, y, K, f_init=None, maxiter=100):
    """Laplace.jl:409-421 with LaplaceResult :388-395: one dict (fnew, f_cov, q_mean, q_cov, lml_approx, cache) per step."""
    out = []

    def store(fnew, cache):
        fc = laplace_f_cov(cache)
        out.append(dict(fnew=fnew.copy(), f_cov=fc, q_mean=cache.f.copy(), q_cov=0.5 * (fc + fc.T), lml_approx=_laplace_lml(cache.f, cache), cache=cache))

    newton_inner_loop(lik, y, K, f_init, maxiter, callback=store)
    return out


def _d3_loglik(lik: Likelihood, f, y):
    if lik.kind == GAUSSIAN:
        return np.zeros_like(f)
    if lik.kind == BERNOULLI_LOGIT:
        p = logistic(f)
        return -p * (1.0 - p) * (1.0 - 2.0 * p)
    if lik.kind == "bernoulli_probit":
        from scipy.special import erfcx

        sg = np.where(y > 0.5, 1.0, -1.0)
        z = sg * f
        r = np.sqrt(2.0 / np.pi) / erfcx(-z / np.sqrt(2.0))
        return sg * r * ((z + r) * (z + 2.0 * r) - 1.0)
    if lik.kind == POISSON_EXP:
        return -np.exp(f)
    if lik.kind in ("exponential_exp", "gamma_exp"):
        return y * np.exp(-f)
    raise ValueError(lik.kind)


def lml_and_grad_K(lik: Likelihood, y, K, f_init=None, maxiter=100):
    """(lml, dlml/dK, f_opt, n_steps): the total derivative the reference obtains from
    ``Zygote.gradient`` of ``approx_lml`` (explicit part at f_opt + implicit part through the mode)."""
    f_opt, newton_cache, steps = newton_inner_loop(lik, y, K, f_init, maxiter)
    c = train_intermediates(lik, y, K, f_opt)
    lml = _laplace_lml(f_opt, c)
    s, W, g = c.Wsqrt, c.W, c.d_loglik
    n = len(y)
    b = W * f_opt + g
    cvec = K @ b
    u = s * cvec
    v = cho_solve((c.B_L, True), u)
    # reverse
    a_bar = -0.5 * f_opt
    f_bar = -0.5 * c.a + g
    Binv = cho_solve((c.B_L, True), np.eye(n))
    B_bar = -0.5 * Binv
    b_bar = a_bar.copy()
    v_bar = -s * a_bar
    s_bar = -a_bar * v
    u_bar = cho_solve((c.B_L, True), v_bar)
    B_bar = B_bar - np.outer(u_bar, v)
    c_bar = s * u_bar
    s_bar = s_bar + u_bar * cvec
    K_bar = np.outer(c_bar, b)
    b_bar = b_bar + K.T @ c_bar
    K_bar = K_bar + (s[:, None] * B_bar) * s[None, :]
    s_bar = s_bar + (B_bar * K) @ s + (B_bar.T * K.T) @ s
    W_bar = b_bar * f_opt
    f_bar = f_bar + W * b_bar
    g_bar = b_bar
    with np.errstate(divide="ignore", invalid="ignore"):
        W_bar = W_bar + s_bar / (2.0 * s)
    h_bar = -W_bar
    _, _, h = loglik_and_derivs(lik, f_opt, y)
    f_bar = f_bar + g_bar * h + h_bar * _d3_loglik(lik, f_opt, np.asarray(y, dtype=np.float64))
    # rrule(newton_inner_loop), Laplace.jl:361-363, with the cache of the last Newton iteration
    nc = newton_cache
    K_bar = K_bar + np.outer(nc.Wsqrt * cho_solve((nc.B_L, True), f_bar / nc.Wsqrt), nc.d_loglik)
    return lml, K_bar, f_opt, steps


# --- prediction (Laplace.jl:425-463) --------------------------------------------------------


def predict_mean_and_cov(kernel: Kernel, X, cache: LaplaceCache, Xnew):
    k_x_xnew = kernelmatrix(kernel, X, Xnew)
    f_mean = k_x_xnew.T @ cache.d_loglik  # RW (3.21), zero prior mean
    v = solve_triangular(cache.B_L, cache.Wsqrt[:, None] * k_x_xnew, lower=True)  # RW (3.29)
    return f_mean, kernelmatrix(kernel, Xnew) - v.T @ v


def predict_mean_and_var(kernel: Kernel, X, cache: LaplaceCache, Xnew):
    k_x_xnew = kernelmatrix(kernel, X, Xnew)
    f_mean = k_x_xnew.T @ cache.d_loglik
    v = solve_triangular(cache.B_L, cache.Wsqrt[:, None] * k_x_xnew, lower=True)
    return f_mean, kernelmatrix_diag(kernel, Xnew) - np.sum(v * v, axis=0)


def predict_cov_cross(kernel: Kernel, X, cache: LaplaceCache, Xa, Xb):
    """``cov(f::LaplacePosteriorGP, x, y)`` (Laplace.jl:457-463)."""
    vx = solve_triangular(cache.B_L, cache.Wsqrt[:, None] * kernelmatrix(kernel, X, Xa), lower=True)
    vy = solve_triangular(cache.B_L, cache.Wsqrt[:, None] * kernelmatrix(kernel, X, Xb), lower=True)
    return kernelmatrix(kernel, Xa, Xb) - vx.T @ vy


# --- fixtures of the reference's own tests (src/TestUtils.jl:13-37) ---------------------------

#! format: off
_Y48 = [0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0]
#! format: on


def generate_data():
    """``ApproximateGPs.TestUtils.generate_data`` (src/TestUtils.jl:13-28): X = range(0, 23.5; length=48)."""
    return np.linspace(0.0, 23.5, 48), np.array(_Y48, dtype=np.float64)


def softplus(x):
    return np.logaddexp(0.0, x)


def build_latent_gp(theta):
    """``build_latent_gp`` (src/TestUtils.jl:32-37): returns (Kernel, Likelihood, jitter)."""
    variance = float(softplus(theta[0]))
    lengthscale = float(softplus(theta[1]))
    return Kernel(SE, variance, np.array([1.0 / lengthscale])), Likelihood(BERNOULLI_LOGIT), 1e-8


def objective_and_grad(theta, X, y, f_init=None, maxiter=100):
    """``-approx_lml(LaplaceApproximation(), build_latent_gp(theta)(X), y)`` and its theta-gradient
    (the quantity test/LaplaceApproximationModule.jl:41-54 and :167-177 exercise)."""
    theta = np.asarray(theta, dtype=np.float64)
    kernel, lik, jitter = build_latent_gp(theta)
    K = kernelmatrix(kernel, X) + jitter * np.eye(len(y))
    lml, K_bar, f_opt, steps = lml_and_grad_K(lik, y, K, f_init, maxiter)
    _, _, kg = kernelmatrix_pullback(kernel, X, None, K_bar)
    sig = 1.0 / (1.0 + np.exp(-theta))  # d softplus / d theta
    lengthscale = float(softplus(theta[1]))
    dvar = kg.variance * sig[0]
    dls = kg.inv_lengthscale[0] * (-1.0 / lengthscale**2) * sig[1]
    return -lml, -np.array([dvar, dls]), f_opt, steps
