"""CPU: pins the Laplace oracle (oracle/laplace.py) against the reference's known answers
(test/LaplaceApproximationModule.jl) on the deterministic 48-point data set of src/TestUtils.jl:13-37."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import kernels as ok, laplace as olap, likelihoods as ol  # noqa: E402

GOLDEN_LBFGS = np.array([7.709076337653239, 1.51820292019697])  # test/LaplaceApproximationModule.jl:168
GOLDEN_NELDER_MEAD = np.array([7.708967951453345, 1.5182348363613536])  # :159 (rtol 1e-4)


def test_fixture_matches_reference():
    X, y = olap.generate_data()
    assert len(X) == 48 and X[0] == 0.0 and X[-1] == 23.5 and np.isclose(X[1] - X[0], 0.5)
    ref = "000100000011111111101000000010111011111110000000"  # src/TestUtils.jl:19, transcribed digit by digit
    assert "".join(str(int(v)) for v in y) == ref


def test_golden_lbfgs_optimum():
    """:167-177 -- L-BFGS from theta0 = [5, 1] reaches the pinned optimum (true known-answer test)."""
    from scipy.optimize import minimize

    X, y = olap.generate_data()
    res = minimize(lambda t: olap.objective_and_grad(t, X, y)[:2], np.array([5.0, 1.0]), jac=True, method="L-BFGS-B",
                   options=dict(gtol=1e-10, ftol=1e-15, maxiter=500))
    assert np.allclose(res.x, GOLDEN_LBFGS, rtol=1e-6)
    assert np.allclose(res.x, GOLDEN_NELDER_MEAD, rtol=1e-4)
    val, grad, _, _ = olap.objective_and_grad(GOLDEN_LBFGS, X, y)
    assert np.max(np.abs(grad)) < 1e-7  # the reference's optimum is a stationary point of the oracle objective
    assert abs(val - 25.661864672178) < 1e-9


def test_gradient_matches_finite_differences():
    """:41-54 -- approx_lml gradient vs central_fdm(5, 1), rtol 1e-6."""
    X, y = olap.generate_data()
    theta = np.array([1.234, 0.789])
    _, grad, _, _ = olap.objective_and_grad(theta, X, y)
    h = 1e-3
    c = np.array([1.0, -8.0, 8.0, -1.0]) / (12.0 * h)
    for i in range(2):
        vals = []
        for d in (-2, -1, 1, 2):
            t = theta.copy()
            t[i] += d * h
            vals.append(olap.objective_and_grad(t, X, y)[0])
        assert abs(np.dot(c, vals) - grad[i]) < 1e-6 * max(1.0, abs(grad[i]))


def test_dK_matches_finite_differences_through_K():
    """:78-145 -- the rrule of newton_inner_loop, checked through a symmetric perturbation of K."""
    rng = np.random.default_rng(0)
    X, y = olap.generate_data()
    k = ok.Kernel("se", 2.0, np.array([0.5]))
    K = ok.kernelmatrix(k, X[:, None]) + 1e-8 * np.eye(48)
    lik = ol.Likelihood("bernoulli_logit")
    lml, Kbar, _, _ = olap.lml_and_grad_K(lik, y, K)
    V = rng.normal(size=(48, 48))
    V = 1e-3 * (V + V.T)
    h = 1e-3
    vals = [olap.laplace_f_and_lml(lik, y, K + d * h * V)[1] for d in (-2, -1, 1, 2)]
    fd = np.dot(np.array([1.0, -8.0, 8.0, -1.0]) / (12.0 * h), vals)
    assert abs(fd - np.sum(Kbar * V)) < 1e-6 * max(1.0, abs(fd))


def test_gaussian_likelihood_is_exact_gpr():
    """src/TestUtils.jl:99-108 -- `f -> Normal(f, 0.1)`, maxiter = 2."""
    rng = np.random.default_rng(1)
    X = np.sort(rng.uniform(0, 5, 30))
    y = np.sin(X) + 0.1 * rng.normal(size=30)
    k = ok.Kernel("se", 1.0, np.array([1.0]))
    K = ok.kernelmatrix(k, X[:, None]) + 1e-8 * np.eye(30)
    lik = ol.Likelihood("gaussian", 0.01)
    f, cache, steps = olap.newton_inner_loop(lik, y, K, maxiter=2)
    assert np.allclose(f, K @ np.linalg.solve(K + 0.01 * np.eye(30), y), atol=1e-9)
    c = olap.train_intermediates(lik, y, K, f)
    mu, cov = olap.predict_mean_and_cov(k, X[:, None], c, X[:, None])
    cov_e = K - K @ np.linalg.solve(K + 0.01 * np.eye(30), K)
    assert np.allclose(mu, f, atol=1e-6) and np.allclose(cov, cov_e, atol=1e-6)


def test_warm_start_saves_newton_steps():
    """:180-204"""
    X, y = olap.generate_data()
    thetas = [np.array([5.0, 1.0]) + 0.05 * i for i in range(6)]
    cold = sum(olap.objective_and_grad(t, X, y)[3] for t in thetas)
    warm, f = 0, None
    for t in thetas:
        _, _, f, s = olap.objective_and_grad(t, X, y, f_init=f)
        warm += s
    assert warm < cold
