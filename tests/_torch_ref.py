"""torch.float64 forward pass of the reference's SVGP ELBO (mirrors SVA.jl:340-373 op by op) so that
torch.autograd plays the role Zygote plays in the reference.  Used only to cross-check the
hand-derived reverse passes (oracle and CUDA); CPU only."""
import math

import numpy as np
import torch

from oracle.likelihoods import ANALYTIC, BERNOULLI_LOGIT, GAUSSIAN, POISSON_EXP, gausshermite


def _scale_vec(inv_ls, D):
    return inv_ls.expand(D) if inv_ls.numel() == 1 else inv_ls


def t_kernelmatrix(kind, variance, inv_ls, c, X, Y=None):
    D = X.shape[1]
    s = _scale_vec(inv_ls, D)
    Xs = X * s
    sym = Y is None
    Ys = Xs if sym else Y * s
    if kind == "linear":
        return variance * (Xs @ Ys.T + c)
    if D == 1:
        u = (Xs[:, 0][:, None] - Ys[:, 0][None, :]) ** 2
    else:
        u = torch.clamp((Xs * Xs).sum(1)[:, None] + (Ys * Ys).sum(1)[None, :] - 2.0 * Xs @ Ys.T, min=0.0)
    if sym:
        u = torch.triu(u, 1)
        u = u + u.T
    if kind == "se":
        return variance * torch.exp(-u / 2)
    # differentiate through u (finite at 0): use a safe sqrt whose gradient is supplied analytically
    d = _SafeSqrt.apply(u)
    if kind == "matern32":
        return variance * (1 + math.sqrt(3) * d) * torch.exp(-math.sqrt(3) * d)
    return variance * (1 + math.sqrt(5) * d + 5 * u / 3) * torch.exp(-math.sqrt(5) * d)


class _SafeSqrt(torch.autograd.Function):
    """sqrt with the 1/(2 max(d, eps)) pullback KernelFunctions uses for Euclidean distances."""

    @staticmethod
    def forward(ctx, u):
        d = torch.sqrt(u)
        ctx.save_for_backward(d)
        return d

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return g / (2 * torch.clamp(d, min=1e-300))


def t_kernel_diag(kind, variance, inv_ls, c, X):
    if kind == "linear":
        Xs = X * _scale_vec(inv_ls, X.shape[1])
        return variance * ((Xs * Xs).sum(1) + c)
    return variance * torch.ones(X.shape[0], dtype=X.dtype)


def t_expected_loglik(lik_kind, sigma2, method, n_points, mu, var, y):
    std = torch.sqrt(var)
    if method == ANALYTIC:
        v = std * std
        if lik_kind == GAUSSIAN:
            return torch.sum(-0.5 * (math.log(2 * math.pi) + torch.log(sigma2) + ((y - mu) ** 2 + v) / sigma2))
        if lik_kind == POISSON_EXP:
            return torch.sum(y * mu - torch.exp(mu + v / 2) - torch.lgamma(y + 1))
        if lik_kind in ("exponential_exp", "gamma_exp"):
            alpha = sigma2 if lik_kind == "gamma_exp" else torch.tensor(1.0, dtype=mu.dtype)
            cst = (alpha - 1) * torch.log(y) - torch.lgamma(alpha) if lik_kind == "gamma_exp" else 0.0
            return torch.sum(cst - y * torch.exp(-mu + v / 2) - alpha * mu)
        raise ValueError
    xs, ws = gausshermite(n_points)
    xs = torch.tensor(xs)
    ws = torch.tensor(ws)
    f = mu[:, None] + (math.sqrt(2) * std)[:, None] * xs[None, :]
    yy = y[:, None]
    if lik_kind == GAUSSIAN:
        ll = -0.5 * (math.log(2 * math.pi) + torch.log(sigma2)) - 0.5 * (yy - f) ** 2 / sigma2
    elif lik_kind == BERNOULLI_LOGIT:
        p = torch.sigmoid(f)
        ll = torch.where(yy > 0.5, torch.log(p), torch.log(1 - p))
    elif lik_kind == "bernoulli_probit":
        # p = normcdf(f); logpdf(Bernoulli(p), y) = y ? log(p) : log(1 - p) = log Phi(+-f).  (torch.where over the literal
        # log(p) / log(1 - p) pair back-propagates 0 * inf = NaN from the unselected branch once 1 - p rounds to 0.)
        ll = torch.special.log_ndtr(torch.where(yy > 0.5, f, -f))
    elif lik_kind in ("exponential_exp", "gamma_exp"):
        alpha = sigma2 if lik_kind == "gamma_exp" else torch.tensor(1.0, dtype=mu.dtype)
        cst = (alpha - 1) * torch.log(yy) - torch.lgamma(alpha) if lik_kind == "gamma_exp" else 0.0
        ll = cst - yy * torch.exp(-f) - alpha * f
    else:
        ll = yy * f - torch.exp(f) - torch.lgamma(yy + 1)
    return torch.sum((ll @ ws) / math.sqrt(math.pi))


def t_elbo(kind, variance, inv_ls, c, Z, jitter, m, Lq_param, centered, mean_const, X, y, lik_kind, sigma2, method, n_points, num_data):
    """All arguments torch.float64 tensors where differentiable; returns the scalar ELBO."""
    M = Z.shape[0]
    N = X.shape[0]
    Lq = torch.tril(Lq_param)
    Kuu = t_kernelmatrix(kind, variance, inv_ls, c, Z) + jitter * torch.eye(M, dtype=Z.dtype)
    Lk = torch.linalg.cholesky(Kuu)
    if centered:
        B = torch.linalg.solve_triangular(Lk, Lq, upper=False)
        alpha = torch.cholesky_solve((m - mean_const)[:, None], Lk)[:, 0]
    else:
        alpha = torch.linalg.solve_triangular(Lk.T, m[:, None], upper=True)[:, 0]
        B = Lq
    Kuf = t_kernelmatrix(kind, variance, inv_ls, c, Z, X)
    A = torch.linalg.solve_triangular(Lk, Kuf, upper=False)
    mu = mean_const + Kuf.T @ alpha
    BtA = B.T @ A
    var = t_kernel_diag(kind, variance, inv_ls, c, X) - (A * A).sum(0) + (BtA * BtA).sum(0) + 1e-18
    E = t_expected_loglik(lik_kind, sigma2, method, n_points, mu, var, y)
    if centered:
        S = Lq @ Lq.T
        r = mean_const - m
        kl = 0.5 * (
            torch.trace(torch.linalg.solve(Kuu, S)) + r @ torch.linalg.solve(Kuu, r) - M + torch.logdet(Kuu) - 2 * torch.log(torch.diagonal(Lq)).sum()
        )
    else:
        kl = 0.5 * ((Lq**2).sum() + m @ m - M - 2 * torch.log(torch.diagonal(Lq)).sum())
    return E * (num_data / N) - kl


def torch_elbo_and_grad(s, X, y, lik, exp_, num_data=None):
    """Takes oracle-side objects (oracle.svgp.SVGP, Likelihood, Expectation) and returns
    (elbo, dict of gradients) from torch autograd."""
    exp_ = exp_.resolve(lik)
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True)
    variance, inv_ls, c = t(s.kernel.variance), t(s.kernel.inv_lengthscale), t(s.kernel.c)
    Z, m, Lq, mc, s2 = t(s.Z), t(s.m), t(s.Lq), t(s.mean_const), t(lik.sigma2)
    val = t_elbo(
        s.kernel.kind, variance, inv_ls, c, Z, s.jitter, m, Lq, s.centered, mc,
        torch.tensor(X), torch.tensor(np.asarray(y, dtype=np.float64)), lik.kind, s2, exp_.method, exp_.n_points,
        float(X.shape[0] if num_data is None else num_data),
    )
    val.backward()
    g = lambda a: None if a.grad is None else a.grad.numpy().copy()
    return float(val), dict(variance=g(variance), inv_lengthscale=g(inv_ls), c=g(c), Z=g(Z), m=g(m), Lq=g(Lq), mean_const=g(mc), lik_sigma2=g(s2))
