"""SVGP posterior / ELBO / gradient: NumPy restatement of the reference (oracle; test infrastructure).

Follows src/SparseVariationalApproximationModule.jl of /root/reference operation by operation,
with materialised M x n matrices exactly as the reference has them (chunked over the data so the
full-size configs fit in host RAM; the reference itself can only survive them by minibatching):

* ``posterior_data``   SVA.jl:115-136 (Centered), :160-187 (NonCentered); utils.jl:15-20
* ``mean_and_var``     SVA.jl:215-219 (``_A_and_Kuf``), :246-253
* ``prior_kl``         SVA.jl:362 (Centered -> Distributions.kldivergence), :364-373 (NonCentered)
* ``elbo``             SVA.jl:340-360 (and the FiniteGP wrapper :307-317 through ``Likelihood``)
* ``elbo_and_grad``    the reverse pass of the above.  The reference has no hand-written SVGP
                       gradient (Zygote differentiates the forward pass); this is the same
                       derivative written out op by op in the order Zygote's pullbacks run
                       (``\\`` -> two triangular solves and an M x n . n x M product, etc.).

Layout: X is (N, D) row-major (one point per row), Z is (M, D), Lq is the lower Cholesky factor
stored by ``PDMat`` (Distributions / PDMats).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy.linalg import cho_solve, solve_triangular

from .kernels import (
    Kernel,
    KernelGrad,
    kernelmatrix,
    kernelmatrix_diag,
    kernelmatrix_diag_pullback,
    kernelmatrix_pullback,
)
from .likelihoods import Expectation, Likelihood, expected_loglik_terms

POSTERIOR_JITTER = 1e-18  # AbstractGPs default jitter of f_post(x), SVA.jl:354


@dataclass
class SVGP:
    """``SparseVariationalApproximation{Centered|NonCentered}(fz, q)`` (SVA.jl:59-95) with
    ``fz = GP(mean_const, kernel)(Z, jitter)`` and ``q = MvNormal(m, PDMat(Cholesky(Lq)))``."""

    kernel: Kernel
    Z: np.ndarray
    m: np.ndarray
    Lq: np.ndarray
    jitter: float = 1e-18
    centered: bool = False
    mean_const: float = 0.0

    def __post_init__(self):
        self.Z = np.asarray(self.Z, dtype=np.float64)
        if self.Z.ndim == 1:
            self.Z = self.Z[:, None]
        self.m = np.asarray(self.m, dtype=np.float64)
        self.Lq = np.tril(np.asarray(self.Lq, dtype=np.float64))  # LowerTriangular(A) view


@dataclass
class SVGPGrad:
    m: np.ndarray
    Lq: np.ndarray
    Z: np.ndarray
    kernel: KernelGrad
    mean_const: float
    lik_sigma2: float


def _kuu(s: SVGP) -> np.ndarray:
    K = kernelmatrix(s.kernel, s.Z)
    return K + s.jitter * np.eye(K.shape[0])  # cov(fz) = k(Z,Z) + Sigma_y


def posterior_data(s: SVGP):
    """(Lk, B, alpha) == ``posterior(sva).data`` (SVA.jl:131-134 / :181-185)."""
    Lk = np.linalg.cholesky(_kuu(s))  # _chol_cov(fz), utils.jl:17
    if s.centered:
        B = solve_triangular(Lk, s.Lq, lower=True)  # SVA.jl:132
        alpha = cho_solve((Lk, True), s.m - s.mean_const)  # SVA.jl:133
    else:
        alpha = solve_triangular(Lk.T, s.m, lower=False)  # SVA.jl:182
        B = s.Lq  # SVA.jl:183-184
    return Lk, B, alpha


def mean_and_var(s: SVGP, X, data=None):
    """``mean_and_var(posterior(sva), x)`` (SVA.jl:246-253)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    Lk, B, alpha = data if data is not None else posterior_data(s)
    Kuf = kernelmatrix(s.kernel, s.Z, X)  # SVA.jl:216
    A = solve_triangular(Lk, Kuf, lower=True)  # SVA.jl:217
    mu = s.mean_const + Kuf.T @ alpha  # SVA.jl:250
    BtA = B.T @ A
    var = kernelmatrix_diag(s.kernel, X) - np.sum(A * A, axis=0) + np.sum(BtA * BtA, axis=0)  # SVA.jl:251
    return mu, var


def mean_and_cov(s: SVGP, X, data=None):
    """``mean_and_cov`` (SVA.jl:237-244)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    Lk, B, alpha = data if data is not None else posterior_data(s)
    Kuf = kernelmatrix(s.kernel, s.Z, X)
    A = solve_triangular(Lk, Kuf, lower=True)
    BtA = B.T @ A
    return s.mean_const + Kuf.T @ alpha, kernelmatrix(s.kernel, X) - A.T @ A + BtA.T @ BtA


def cov_cross(s: SVGP, X, Y, data=None):
    """``cov(f_post, x, y)`` (SVA.jl:255-264)."""
    X, Y = np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64)
    X = X[:, None] if X.ndim == 1 else X
    Y = Y[:, None] if Y.ndim == 1 else Y
    Lk, B, _ = data if data is not None else posterior_data(s)
    Ax = solve_triangular(Lk, kernelmatrix(s.kernel, s.Z, X), lower=True)
    Ay = solve_triangular(Lk, kernelmatrix(s.kernel, s.Z, Y), lower=True)
    return kernelmatrix(s.kernel, X, Y) - Ax.T @ Ay + Ax.T @ B @ B.T @ Ay


def prior_kl(s: SVGP) -> float:
    M = s.m.size
    if not s.centered:  # SVA.jl:364-373
        trace_term = np.sum(s.Lq**2)
        logdet_S = 2.0 * np.sum(np.log(np.diag(s.Lq)))
        return float((trace_term + s.m @ s.m - M - logdet_S) / 2.0)
    # SVA.jl:362 -> Distributions.kldivergence(q, fz) with dense `\` and logdet (LU) on cov(fz)
    Kuu = _kuu(s)
    S = s.Lq @ s.Lq.T
    r = s.mean_const - s.m
    sign, logdet_K = np.linalg.slogdet(Kuu)
    assert sign > 0
    logdet_S = 2.0 * np.sum(np.log(np.diag(s.Lq)))
    return float((np.trace(np.linalg.solve(Kuu, S)) + r @ np.linalg.solve(Kuu, r) - M + logdet_K - logdet_S) / 2.0)


def _chunks(N: int, chunk: int | None):
    if not chunk or chunk >= N:
        yield 0, N
        return
    for lo in range(0, N, chunk):
        yield lo, min(N, lo + chunk)


def elbo(s: SVGP, X, y, lik: Likelihood, exp_: Expectation | None = None, num_data=None, chunk=None) -> float:
    """``elbo(sva, lfx, y; num_data, quadrature)`` (SVA.jl:340-360)."""
    exp_ = exp_ or Expectation()
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    N = X.shape[0]
    data = posterior_data(s)
    total = 0.0
    for lo, hi in _chunks(N, chunk):
        mu, var = mean_and_var(s, X[lo:hi], data)
        E, _, _, _ = expected_loglik_terms(exp_, lik, mu, var + POSTERIOR_JITTER, y[lo:hi], point0=lo)
        total += float(np.sum(E))
    scale = (N if num_data is None else num_data) / N  # SVA.jl:357-358
    return total * scale - prior_kl(s)


def _chol_pullback(L: np.ndarray, Lbar: np.ndarray) -> np.ndarray:
    """Cotangent of Sigma for Sigma = L L^T given the (lower-triangular) cotangent of L;
    symmetrised (both triangles of Sigma free)."""
    P = np.tril(L.T @ Lbar)
    P[np.diag_indices_from(P)] *= 0.5
    S = solve_triangular(L.T, P, lower=False)  # L^-T P
    S = solve_triangular(L.T, S.T, lower=False).T  # (L^-T P) L^-1
    return 0.5 * (S + S.T)


def elbo_and_grad(s: SVGP, X, y, lik: Likelihood, exp_: Expectation | None = None, num_data=None, chunk=None):
    """ELBO and its gradient w.r.t. (m, Lq lower triangle, Z, kernel params, mean const, lik sigma2)."""
    exp_ = exp_ or Expectation()
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    y = np.asarray(y, dtype=np.float64)
    N = X.shape[0]
    M = s.m.size
    k = s.kernel
    scale = (N if num_data is None else num_data) / N
    Lk, B, alpha = posterior_data(s)

    total = 0.0
    alpha_bar = np.zeros(M)
    B_bar = np.zeros((M, M))
    Lacc = np.zeros((M, M))
    Z_bar = np.zeros_like(s.Z)
    kg = KernelGrad(0.0, np.zeros_like(k.inv_lengthscale), 0.0)
    c_bar = 0.0
    s2_bar = 0.0
    for lo, hi in _chunks(N, chunk):
        Xc, yc = X[lo:hi], y[lo:hi]
        Kuf = kernelmatrix(k, s.Z, Xc)
        A = solve_triangular(Lk, Kuf, lower=True)
        mu = s.mean_const + Kuf.T @ alpha
        BtA = B.T @ A
        var = kernelmatrix_diag(k, Xc) - np.sum(A * A, axis=0) + np.sum(BtA * BtA, axis=0)
        E, dmu, dvar, ds2 = expected_loglik_terms(exp_, lik, mu, var + POSTERIOR_JITTER, yc, point0=lo)
        total += float(np.sum(E))
        dmu = dmu * scale
        dvar = dvar * scale
        s2_bar += float(np.sum(ds2)) * scale
        # reverse of the chunk
        _, g = kernelmatrix_diag_pullback(k, Xc, dvar)
        kg = kg.add(g)
        c_bar += float(np.sum(dmu))
        alpha_bar += Kuf @ dmu
        BtAdv = BtA * dvar[None, :]
        A_bar = -2.0 * A * dvar[None, :] + 2.0 * (B @ BtAdv)
        B_bar += 2.0 * (A @ BtAdv.T)
        T = solve_triangular(Lk.T, A_bar, lower=False)  # pullback of Lk \ Kuf w.r.t. Kuf
        Lacc += T @ A.T  # ... and w.r.t. Lk (negated, lower triangle, below)
        Kuf_bar = np.outer(alpha, dmu) + T
        Zb, _, g = kernelmatrix_pullback(k, s.Z, Xc, Kuf_bar)
        Z_bar += Zb
        kg = kg.add(g)

    Lk_bar = -np.tril(Lacc)
    B_bar = np.tril(B_bar)
    if not s.centered:
        m_data = solve_triangular(Lk, alpha_bar, lower=True)  # alpha = Lk^-T m
        Lk_bar -= np.tril(np.outer(alpha, m_data))
        m_bar = m_data - s.m
        Lq_bar = B_bar - s.Lq + np.diag(1.0 / np.diag(s.Lq))
        Kuu_bar = _chol_pullback(Lk, Lk_bar)
        kl = prior_kl(s)
    else:
        Kuu = _kuu(s)
        S = s.Lq @ s.Lq.T
        # B = Lk \ Lq
        TB = solve_triangular(Lk.T, B_bar, lower=False)
        Lq_bar = np.tril(TB)
        Lk_bar -= np.tril(TB @ B.T)
        # alpha = Kuu \ (m - mean(fz))
        r_bar = cho_solve((Lk, True), alpha_bar)
        m_bar = r_bar.copy()
        c_bar -= float(np.sum(r_bar))
        Kuu_dir = -0.5 * (np.outer(r_bar, alpha) + np.outer(alpha, r_bar))
        # -KL(q || p(u))
        m_bar -= alpha
        c_bar += float(np.sum(alpha))
        Kinv_Lq = cho_solve((Lk, True), s.Lq)
        Lq_bar += -np.tril(Kinv_Lq) + np.diag(1.0 / np.diag(s.Lq))
        Kinv = cho_solve((Lk, True), np.eye(M))
        Kuu_dir += 0.5 * (Kinv @ S @ Kinv + np.outer(alpha, alpha) - Kinv)
        Kuu_bar = Kuu_dir + _chol_pullback(Lk, Lk_bar)
        kl = prior_kl(s)
    Zb, _, g = kernelmatrix_pullback(k, s.Z, None, Kuu_bar)
    Z_bar += Zb
    kg = kg.add(g)
    value = total * scale - kl
    return value, SVGPGrad(m_bar, Lq_bar, Z_bar, kg, c_bar, s2_bar)


# --- helpers used by the reference's own tests (test/test_utils.jl:7-17) -------------------


def optimal_variational_posterior(kernel: Kernel, Z, jitter, X, y, sigma2):
    """Closed-form optimal q(u) = N(m, S) (Titsias), zero prior mean; returns (m, S)."""
    Z = np.asarray(Z, dtype=np.float64)
    Kuf = kernelmatrix(kernel, Z, X)
    Kuu = kernelmatrix(kernel, Z) + jitter * np.eye(Kuf.shape[0])
    Sig = Kuu + Kuf @ Kuf.T / sigma2
    Sig = 0.5 * (Sig + Sig.T)
    m = (Kuu @ np.linalg.solve(Sig, Kuf)) @ y / sigma2
    S = Kuu @ np.linalg.solve(Sig, Kuu)
    return m, 0.5 * (S + S.T)


def exact_gpr_posterior(kernel: Kernel, X, y, sigma2, Xnew):
    """Exact GP regression mean/cov at Xnew (AbstractGPs ``posterior(fx, y)``), zero mean."""
    K = kernelmatrix(kernel, X) + sigma2 * np.eye(len(y))
    L = np.linalg.cholesky(K)
    Ks = kernelmatrix(kernel, X, Xnew)
    a = cho_solve((L, True), y)
    V = solve_triangular(L, Ks, lower=True)
    return Ks.T @ a, kernelmatrix(kernel, Xnew) - V.T @ V


def exact_gpr_logpdf(kernel: Kernel, X, y, sigma2) -> float:
    K = kernelmatrix(kernel, X) + sigma2 * np.eye(len(y))
    L = np.linalg.cholesky(K)
    a = solve_triangular(L, y, lower=True)
    return float(-0.5 * a @ a - np.sum(np.log(np.diag(L))) - 0.5 * len(y) * np.log(2 * np.pi))
