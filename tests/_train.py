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


def adam_train(neg_elbo_and_grad, N, steps=20000, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, tail=200):
    """Flux.Adam on (m, A); neg_elbo_and_grad(m, A) -> (loss, dm, dA).  Returns the mean of the last `tail` iterates.

    Why not the 20 000th iterate itself, as the reference's test does: at the flat optimum Adam's normalised steps keep the iterate
    hopping, and a single iterate's posterior mean sits at 1.4e-5 from exact GPR most of the time but spikes to 2e-4 .. 4e-4 every few
    hundred steps.  Which side of the reference's atol = 1e-4 step 20 000 lands on is decided by rounding-level differences in the
    gradient: the oracle itself gives 1.2e-5 .. 1.4e-5 for five of six runs with the gradient perturbed by 1e-14 .. 1e-10 relative,
    and 4.3e-4 for the sixth (measured in round 2 when a change of summation order inside the Cholesky kernel flipped the CUDA run
    from 3.9e-5 to 2.8e-4).  The mean of the last 200 iterates is free of the hopping (1.4e-5 / 1.0e-5 on mean / covariance) and
    keeps the reference's protocol, optimiser and tolerance."""
    m, A = np.zeros(N), np.eye(N)
    mom = [np.zeros(N), np.zeros((N, N))]
    vel = [np.zeros(N), np.zeros((N, N))]
    m_sum, A_sum = np.zeros(N), np.zeros((N, N))
    for t in range(1, steps + 1):
        _, gm, gA = neg_elbo_and_grad(m, A)
        for i, (p, g) in enumerate(((m, gm), (A, gA))):
            mom[i] = b1 * mom[i] + (1 - b1) * g
            vel[i] = b2 * vel[i] + (1 - b2) * g * g
            p -= lr * (mom[i] / (1 - b1**t)) / (np.sqrt(vel[i] / (1 - b2**t)) + eps)
        if t > steps - tail:
            m_sum += m
            A_sum += A
    return m_sum / tail, A_sum / tail


def exact_gpr(x, y, variance, inv_ls, noise):
    k = ok.Kernel("se", variance, np.array([inv_ls]))
    return osv.exact_gpr_posterior(k, x[:, None], y, noise, x[:, None])
