"""GPLikelihoods 0.4 / Distributions semantics (oracle; test infrastructure).

``expected_loglikelihood(quadrature, lik, q_f, y)`` is imported by the reference at
src/SparseVariationalApproximationModule.jl:25 and called at :355; ``logpdf(dist_y_given_f(f), y)``
is called at src/LaplaceApproximationModule.jl:231.  GPLikelihoods / Distributions /
FastGaussQuadrature are not vendored (Project.toml:30-34, compat ranges only), so this restates
their published behaviour:

* ``DefaultExpectationMethod`` -> analytic for ``GaussianLikelihood`` and
  ``PoissonLikelihood{ExpLink}``, else ``GaussHermiteExpectation(20)``.
* Gauss-Hermite: per point ``(1/sqrt(pi)) * sum_k w_k * loglikelihood(lik(mu + sqrt2*sigma*x_k), y)``
  with ``(x_k, w_k) = gausshermite(n)`` == ``numpy.polynomial.hermite.hermgauss(n)``.
* ``BernoulliLikelihood`` -> ``Bernoulli(logistic(f))`` with ``logpdf = y ? log(p) : log(1-p)``;
  ``PoissonLikelihood`` -> ``Poisson(exp(f))`` with ``logpdf = xlogy(y, lam) - lam - loggamma(y+1)``;
  ``GaussianLikelihood(s2)`` -> ``Normal(f, sqrt(s2))``.

Every function returns per-point values and the derivatives of the *finite quadrature sum*
(that is what Zygote differentiates in the reference), so a caller gets
(E_i, dE_i/dmu_i, dE_i/dvar_i, dE_i/dlik_param).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.special import gammaln, xlogy

GAUSSIAN, BERNOULLI_LOGIT, POISSON_EXP = "gaussian", "bernoulli_logit", "poisson_exp"
BERNOULLI_PROBIT = "bernoulli_probit"  # BernoulliLikelihood(ProbitLink()): Bernoulli(normcdf(f))
EXPONENTIAL_EXP, GAMMA_EXP = "exponential_exp", "gamma_exp"  # Exponential / Gamma(alpha) with scale exp(f); alpha rides in sigma2
ANALYTIC, GAUSS_HERMITE, MONTE_CARLO = "analytic", "gauss_hermite", "monte_carlo"

_LOG2PI = np.log(2.0 * np.pi)
_INVSQRTPI = 1.0 / np.sqrt(np.pi)
_SQRT2 = np.sqrt(2.0)


@dataclass
class Likelihood:
    kind: str = GAUSSIAN
    sigma2: float = 1.0  # GaussianLikelihood only


@dataclass
class Expectation:
    """``method == 'default'`` resolves like ``DefaultExpectationMethod()``."""

    method: str = "default"
    n_points: int = 20
    seed: int = 0  # MonteCarloExpectation only (include/agp.h agp_expectation.seed)

    def resolve(self, lik: Likelihood) -> "Expectation":
        if self.method != "default":
            return self
        if lik.kind in (GAUSSIAN, POISSON_EXP, EXPONENTIAL_EXP, GAMMA_EXP):
            return Expectation(ANALYTIC, 0)
        return Expectation(GAUSS_HERMITE, 20)


def philox_normal(seed: int, points, n_samples: int) -> np.ndarray:
    """eps[i, k] for global point index points[i] and sample k: the counter-based N(0,1) stream of the device
    MonteCarloExpectation (Philox4x32-10 keyed by ``seed``, counter (point_lo, point_hi, sample, 0), Box-Muller with
    u1 from output words 0-1 and u2 from words 2-3) restated with NumPy integer arithmetic.  GPLikelihoods'
    MonteCarloExpectation draws ``randn`` from Julia's RNG instead: only the distribution is shared with the reference."""
    pts = np.asarray(points, dtype=np.uint64)
    M32 = np.uint64(0xFFFFFFFF)
    c0 = np.repeat((pts & M32)[:, None], n_samples, axis=1)
    c1 = np.repeat((pts >> np.uint64(32))[:, None], n_samples, axis=1)
    c2 = np.repeat(np.arange(n_samples, dtype=np.uint64)[None, :], len(pts), axis=0)
    c3 = np.zeros_like(c0)
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0
        n1 = p1 & M32
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1
        n3 = p0 & M32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    a = (c0 << np.uint64(32)) | c1
    b = (c2 << np.uint64(32)) | c3
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 1.0) / 9007199254740992.0
    u2 = (b >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def gausshermite(n: int):
    """FastGaussQuadrature.gausshermite(n): physicists' nodes/weights (weight exp(-x^2))."""
    return np.polynomial.hermite.hermgauss(n)


def logistic(f):
    f = np.asarray(f, dtype=np.float64)
    out = np.empty_like(f)
    pos = f >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-f[pos]))
    e = np.exp(f[~pos])
    out[~pos] = e / (1.0 + e)
    return out


def loglik_and_derivs(lik: Likelihood, f, y):
    """Pointwise log p(y|f), d/df, d2/df2 (closed forms of what
    src/LaplaceApproximationModule.jl:230-241 obtains with nested ForwardDiff)."""
    f = np.asarray(f, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if lik.kind == GAUSSIAN:
        r = y - f
        return -0.5 * (_LOG2PI + np.log(lik.sigma2)) - 0.5 * r * r / lik.sigma2, r / lik.sigma2, np.full_like(f, -1.0 / lik.sigma2)
    if lik.kind == BERNOULLI_LOGIT:
        p = logistic(f)
        with np.errstate(divide="ignore"):
            ll = np.where(y > 0.5, np.log(p), np.log(1.0 - p))  # Distributions.logpdf(Bernoulli(p), y)
        return ll, y - p, -p * (1.0 - p)
    if lik.kind == BERNOULLI_PROBIT:
        # logpdf(Bernoulli(normcdf(f)), y) = log Phi(s f), s = 2y - 1, in the overflow-free form (log_ndtr / erfcx); with
        # r = phi(z) / Phi(z): d/df = s r, d2/df2 = -r (z + r)
        from scipy.special import erfcx, log_ndtr

        sg = np.where(y > 0.5, 1.0, -1.0)
        z = sg * f
        r = np.sqrt(2.0 / np.pi) / erfcx(-z / _SQRT2)
        return log_ndtr(z), sg * r, -r * (z + r)
    if lik.kind == POISSON_EXP:
        lam = np.exp(f)
        return xlogy(y, lam) - lam - gammaln(y + 1.0), y - lam, -lam
    if lik.kind in (EXPONENTIAL_EXP, GAMMA_EXP):
        # Distributions.logpdf(Gamma(alpha, theta), y) with theta = exp(f): (alpha-1) log y - y/theta - alpha log theta - loggamma(alpha);
        # Exponential(theta) is alpha = 1
        alpha = lik.sigma2 if lik.kind == GAMMA_EXP else 1.0
        t = y * np.exp(-f)
        cst = (alpha - 1.0) * np.log(y) - gammaln(alpha) if lik.kind == GAMMA_EXP else 0.0
        return cst - t - alpha * f, -alpha + t, -t
    raise ValueError(lik.kind)


def dloglik_dparam(lik: Likelihood, f, y):
    """d log p(y|f) / d(likelihood parameter): sigma2 for the Gaussian, alpha for the Gamma likelihood, else 0."""
    f = np.asarray(f, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if lik.kind == GAUSSIAN:
        r = y - f
        return -0.5 / lik.sigma2 + 0.5 * r * r / lik.sigma2**2
    if lik.kind == GAMMA_EXP:
        from scipy.special import digamma

        return np.log(y) - digamma(lik.sigma2) - f
    return np.zeros(np.broadcast(f, y).shape)


def expected_loglik_terms(exp_: Expectation, lik: Likelihood, mu, var, y, point0: int = 0):
    """Per-point expected log-likelihood under N(mu, var) and its derivatives.

    Returns (E, dE/dmu, dE/dvar, dE/dsigma2_lik); arrays of len(y).  ``var`` is the marginal
    variance *including* the 1e-18 jitter of ``f_post(x)`` (SVA.jl:354): the reference builds
    ``Normal(mu, sqrt(var))`` so GH uses ``std = sqrt(var)`` and the analytic forms use ``std^2``.
    """
    exp_ = exp_.resolve(lik)
    mu = np.asarray(mu, dtype=np.float64)
    var = np.asarray(var, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    std = np.sqrt(var)
    if exp_.method == ANALYTIC:
        v = std * std
        if lik.kind == GAUSSIAN:
            s2 = lik.sigma2
            r = y - mu
            E = -0.5 * (_LOG2PI + np.log(s2) + (r * r + v) / s2)
            return E, r / s2, np.full_like(mu, -0.5 / s2), -0.5 / s2 + 0.5 * (r * r + v) / (s2 * s2)
        if lik.kind == POISSON_EXP:
            e = np.exp(mu + v / 2.0)
            return y * mu - e - gammaln(y + 1.0), y - e, -0.5 * e, np.zeros_like(mu)
        if lik.kind in (EXPONENTIAL_EXP, GAMMA_EXP):
            # GPLikelihoods AnalyticExpectation for the exp link: E[exp(-f)] = exp(-mu + v/2)
            alpha = lik.sigma2 if lik.kind == GAMMA_EXP else 1.0
            t = y * np.exp(-mu + v / 2.0)
            cst = (alpha - 1.0) * np.log(y) - gammaln(alpha) if lik.kind == GAMMA_EXP else 0.0
            dpar = dloglik_dparam(lik, mu, y) if lik.kind == GAMMA_EXP else np.zeros_like(mu)
            return cst - t - alpha * mu, -alpha + t, -0.5 * t, dpar
        raise ValueError(f"no analytic expectation for {lik.kind}")
    if exp_.method == MONTE_CARLO:
        # GPLikelihoods.MonteCarloExpectation(n): mean over n reparameterised samples; derivatives of that finite sum
        eps = philox_normal(exp_.seed, point0 + np.arange(len(mu)), exp_.n_points)
        f = mu[:, None] + std[:, None] * eps
        ll, dll, _ = loglik_and_derivs(lik, f, y[:, None])
        E = ll.mean(axis=1)
        dmu = dll.mean(axis=1)
        dvar = (dll * eps).mean(axis=1) / (2.0 * std)
        ds2 = dloglik_dparam(lik, f, y[:, None]).mean(axis=1)
        return E, dmu, dvar, ds2
    xs, ws = gausshermite(exp_.n_points)
    f = mu[:, None] + (_SQRT2 * std)[:, None] * xs[None, :]
    ll, dll, _ = loglik_and_derivs(lik, f, y[:, None])
    E = _INVSQRTPI * (ll @ ws)
    dmu = _INVSQRTPI * (dll @ ws)
    dstd = _INVSQRTPI * ((dll * (_SQRT2 * xs)[None, :]) @ ws)
    dvar = dstd / (2.0 * std)
    ds2 = _INVSQRTPI * (dloglik_dparam(lik, f, y[:, None]) @ ws)
    return E, dmu, dvar, ds2
