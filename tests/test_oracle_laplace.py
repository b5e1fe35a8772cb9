"""CPU: pins the Laplace oracle (oracle/laplace.py) against the reference's known answers
(test/LaplaceApproximationModule.jl) on the deterministic 48-point data set of src/TestUtils.jl:13-37."""
import os
import sys

import numpy as np
import pytest

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


def test_laplace_f_cov_is_posterior_covariance_and_steps_record_every_iteration():
    """laplace_f_cov (Laplace.jl:376-386) is (K^-1 + W)^-1 (checked against the direct inverse); laplace_steps
    (:409-421, test :207-217) returns one LaplaceResult per Newton step, the last one at the mode."""
    X, y = olap.generate_data()
    K = ok.kernelmatrix(ok.Kernel(ok.SE, 1.3, np.array([1.0 / 0.9])), X) + 1e-8 * np.eye(len(y))
    lik = ol.Likelihood(ol.BERNOULLI_LOGIT)
    res = olap.laplace_steps(lik, y, K)
    f_opt, lml, steps = olap.laplace_f_and_lml(lik, y, K)
    assert len(res) == steps
    last = res[-1]
    assert np.allclose(last["q_mean"], f_opt, rtol=0, atol=1e-12)
    assert abs(last["lml_approx"] - lml) < 1e-10 * abs(lml)
    direct = np.linalg.inv(np.linalg.inv(K) + np.diag(last["cache"].W))
    assert np.allclose(last["f_cov"], direct, rtol=0, atol=1e-7 * np.abs(direct).max())  # K^-1 with jitter 1e-8 is ill-conditioned
    assert np.allclose(last["q_cov"], last["q_cov"].T, rtol=0, atol=0)
    assert np.all(np.linalg.eigvalsh(last["q_cov"]) > -1e-10)
    # Gaussian 'likelihood': q(f) is the exact GP posterior over f, K - K (K + s2 I)^-1 K
    s2 = 0.3
    yg = np.sin(X) + 0.1
    resg = olap.laplace_steps(ol.Likelihood(ol.GAUSSIAN, s2), yg, K)
    exact = K - K @ np.linalg.solve(K + s2 * np.eye(len(yg)), K)
    assert np.allclose(resg[-1]["f_cov"], exact, rtol=0, atol=1e-10)


def test_probit_link_derivatives():
    """BernoulliLikelihood(ProbitLink()): the oracle's overflow-free log Phi(s f) and its closed-form derivatives against
    (i) mpmath log(ncdf) at 50 digits, (ii) torch autograd through the reference's op sequence log(normcdf(f)) /
    log(1 - normcdf(f)) where that is well conditioned, (iii) the third derivative by finite differences of the second,
    and the Laplace gradient w.r.t. K by finite differences."""
    import pytest
    import torch

    mp = pytest.importorskip("mpmath")
    lik = ol.Likelihood("bernoulli_probit")
    f = np.array([-9.0, -3.0, -0.7, 0.0, 0.4, 2.5, 6.0])
    for yv in (0.0, 1.0):
        y = np.full_like(f, yv)
        ll, d1, d2 = ol.loglik_and_derivs(lik, f, y)
        sg = 1.0 if yv > 0.5 else -1.0
        with mp.workdps(50):
            ref = np.array([float(mp.log(mp.ncdf(sg * v))) for v in f])
        assert np.allclose(ll, ref, rtol=1e-14, atol=0)
        sel = np.abs(f) < 5  # log(1 - normcdf(f)) is only accurate away from the saturated tail
        ft = torch.tensor(f[sel], requires_grad=True)
        p = torch.special.ndtr(ft)
        lt = (torch.log(p) if yv > 0.5 else torch.log(1 - p)).sum()
        (g1,) = torch.autograd.grad(lt, ft, create_graph=True)
        (g2,) = torch.autograd.grad(g1.sum(), ft)
        assert np.allclose(d1[sel], g1.detach().numpy(), rtol=1e-9) and np.allclose(d2[sel], g2.numpy(), rtol=1e-8)
        h = 1e-4
        d3 = olap._d3_loglik(lik, f, y)
        fd = (ol.loglik_and_derivs(lik, f + h, y)[2] - ol.loglik_and_derivs(lik, f - h, y)[2]) / (2 * h)
        assert np.allclose(d3, fd, rtol=1e-6, atol=1e-9)
    X, y = olap.generate_data()
    K = ok.kernelmatrix(ok.Kernel(ok.SE, 1.3, np.array([1.0 / 0.9])), X) + 1e-6 * np.eye(len(y))
    lml, Kbar, _, _ = olap.lml_and_grad_K(lik, y, K)
    rng = np.random.default_rng(0)
    Dm = rng.normal(size=K.shape)
    Dm = 0.5 * (Dm + Dm.T) * 1e-2
    h = 1e-5
    fd = (olap.laplace_f_and_lml(lik, y, K + h * Dm)[1] - olap.laplace_f_and_lml(lik, y, K - h * Dm)[1]) / (2 * h)
    assert abs(fd - np.sum(Kbar * Dm)) < 1e-6 * max(1.0, abs(fd))


def test_newton_inner_loop_chain_rules():
    """test/LaplaceApproximationModule.jl:78-145 (test_frule / test_rrule of newton_inner_loop through K = L'L): the frule is
    the finite-difference derivative of f_opt, and the rrule is its adjoint."""
    xs = np.array([0.2, 0.3, 0.7])
    ys = np.array([1.0, 1.0, 0.0])
    rng = np.random.default_rng(54321)
    Lm = rng.normal(size=(3, 3))
    dL = rng.normal(size=(3, 3))
    lik = ol.Likelihood(ol.BERNOULLI_LOGIT)
    K = Lm.T @ Lm
    dK = dL.T @ Lm + Lm.T @ dL
    f_opt, cache, _ = olap.newton_inner_loop(lik, ys, K)
    fdot = olap.newton_pushforward(cache, dK)
    h = 1e-5
    fp = olap.newton_inner_loop(lik, ys, (Lm + h * dL).T @ (Lm + h * dL))[0]
    fm = olap.newton_inner_loop(lik, ys, (Lm - h * dL).T @ (Lm - h * dL))[0]
    assert np.allclose(fdot, (fp - fm) / (2 * h), rtol=1e-6, atol=1e-9)
    df = rng.normal(size=3)
    Kbar = olap.newton_pullback(cache, df)
    assert abs(np.sum(Kbar * dK) - df @ fdot) < 1e-12 * max(1.0, abs(df @ fdot))
    # through L: Lbar = L (Kbar' + Kbar)  (:128)
    Lbar = Lm @ (Kbar.T + Kbar)
    assert abs(np.sum(Lbar * dL) - df @ fdot) < 1e-12 * max(1.0, abs(df @ fdot))
    assert xs.shape == ys.shape


def test_third_party_pin_scikit_learn_gpc():
    """Independent implementation of the same algorithm: scikit-learn's GaussianProcessClassifier is Rasmussen & Williams
    Alg. 3.1 (Newton mode finding) and Alg. 5.1 (log marginal likelihood and its total derivative, implicit part through f-hat
    included) with the logistic link -- what Laplace.jl:201-276 / :330-369 implement.  On the reference's 48-point fixture
    (src/TestUtils.jl:13-28) the oracle's lml and d lml / d log(theta) agree with it to 1e-10 (measured 1e-12 .. 1e-14): the oracle's
    Laplace value AND gradient are pinned by a second, unrelated code base, not only by the reference's golden optimum."""
    skl = pytest.importorskip("sklearn.gaussian_process")
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, WhiteKernel

    X, y = olap.generate_data()
    for var, ls in [(1.7, 2.3), (5.0067, 1.3133), (0.6, 0.9)]:
        kern = ConstantKernel(var) * RBF(ls) + WhiteKernel(1e-8, noise_level_bounds="fixed")
        gpc = skl.GaussianProcessClassifier(kernel=kern, optimizer=None, max_iter_predict=1000).fit(X[:, None], y)
        lml_s, g_s = gpc.log_marginal_likelihood(gpc.kernel_.theta, eval_gradient=True)  # gradient w.r.t. log(variance), log(lengthscale)
        k = ok.Kernel(ok.SE, var, np.array([1.0 / ls]))
        K = ok.kernelmatrix(k, X) + 1e-8 * np.eye(len(y))
        lml, Kbar, _, _ = olap.lml_and_grad_K(ol.Likelihood("bernoulli_logit"), y, K)
        _, _, kg = ok.kernelmatrix_pullback(k, X, None, Kbar)
        g = np.array([kg.variance * var, kg.inv_lengthscale[0] * (-1.0 / ls**2) * ls])
        assert abs(lml - lml_s) < 1e-10 * abs(lml_s)
        assert np.max(np.abs(g - g_s) / np.abs(g_s)) < 1e-10
