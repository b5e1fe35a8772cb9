"""The reference's "optimised posterior" test (test/SparseVariationalApproximationModule.jl:136-186): train (m, A) of a
NonCentered SVGP with Z = X by 20 000 full-batch Flux.Adam(1e-3) steps from m = 0, A = I and compare with exact GPR."""
import numpy as np

from oracle import kernels as ok, svgp as osv


def problem():
    rng = np.random.default_rng(654321)
    N = 20
    x = rng.random(N) * 10
    y = np.sin(x) + 0.9 * np.cos(x * 1.6) + 0.4 * rng.random(N)
    sp = lambda v: np.logaddexp(0.0, v)
    return x, y, sp(0.2), sp(0.6), 0.1, 1e-5  # variance, inverse length scale (ScaleTransform), noise, jitter


def adam_train(neg_elbo_and_grad, N, steps=20000, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """Flux.Adam on (m, A); neg_elbo_and_grad(m, A) -> (loss, dm, dA)."""
    m, A = np.zeros(N), np.eye(N)
    mom = [np.zeros(N), np.zeros((N, N))]
    vel = [np.zeros(N), np.zeros((N, N))]
    for t in range(1, steps + 1):
        _, gm, gA = neg_elbo_and_grad(m, A)
        for i, (p, g) in enumerate(((m, gm), (A, gA))):
            mom[i] = b1 * mom[i] + (1 - b1) * g
            vel[i] = b2 * vel[i] + (1 - b2) * g * g
            p -= lr * (mom[i] / (1 - b1**t)) / (np.sqrt(vel[i] / (1 - b2**t)) + eps)
    return m, A


def exact_gpr(x, y, variance, inv_ls, noise):
    k = ok.Kernel("se", variance, np.array([inv_ls]))
    return osv.exact_gpr_posterior(k, x[:, None], y, noise, x[:, None])
