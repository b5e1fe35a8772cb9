"""CPU: pins the SVGP oracle (oracle/svgp.py) with the property tests of the reference's own suite
(test/SparseVariationalApproximationModule.jl) and cross-checks its hand-derived reverse pass against
torch.float64 autograd of the same forward pass (the role Zygote plays in the reference), finite
differences and mpmath.  No GPU, no product code."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _cases import make_problem, oracle_objects, rel_err  # noqa: E402
from _torch_ref import torch_elbo_and_grad  # noqa: E402

from oracle import kernels as ok, likelihoods as ol, svgp as osv  # noqa: E402


def _small(seed=123456, N=5, M=4, kind="matern32"):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, N)
    z = x[:M].copy()
    k = ok.Kernel(kind, 1.0, np.array([1.0]))
    return rng, x, z, k


def test_centered_equals_noncentered():
    """test/SVA...:45-69: whitening round trip; KL rtol 1e-5; mean / cov / elbo agree."""
    rng, x, z, k = _small()
    M = len(z)
    jitter = 1e-12
    Kuu = ok.kernelmatrix(k, z[:, None]) + jitter * np.eye(M)
    Lk = np.linalg.cholesky(Kuu)
    m_w = rng.normal(size=M)
    A = np.tril(rng.normal(size=(M, M)))
    A[np.diag_indices(M)] = np.abs(np.diag(A)) + 0.5
    nc = osv.SVGP(k, z, m_w, A, jitter=jitter, centered=False)
    ce = osv.SVGP(k, z, Lk @ m_w, Lk @ A, jitter=jitter, centered=True)  # q_ex = (Lk m, Lk S Lk')
    assert abs(osv.prior_kl(nc) - osv.prior_kl(ce)) <= 1e-5 * abs(osv.prior_kl(ce))
    mu1, c1 = osv.mean_and_cov(nc, x)
    mu2, c2 = osv.mean_and_cov(ce, x)
    assert np.allclose(mu1, mu2) and np.allclose(c1, c2)
    y = rng.normal(size=len(x))
    lik = ol.Likelihood("gaussian", 0.1)
    assert np.isclose(osv.elbo(nc, x, y, lik), osv.elbo(ce, x, y, lik))


def test_elbo_below_logpdf_and_gpr_equivalence():
    """test/SVA...:99-134 with the reference's recipe (x = 10 rand(20), z = x, fz = f(z) i.e. jitter 1e-18,
    kernel = softplus(0.2) * SE o ScaleTransform(softplus(0.6)), noise 0.1):
    SVGP(Centered, q*, Z = X) == exact GPR at atol 1e-10 (:126-127); elbo <= logpdf + 1e-5 (:132-133)."""
    sp = lambda v: np.logaddexp(0.0, v)
    k = ok.Kernel("se", sp(0.2), np.array([sp(0.6)]))
    worst, used = 0.0, 0
    for seed in range(8):
        rng = np.random.default_rng(654321 + seed)
        N = 20
        x = rng.random(N) * 10
        if np.min(np.diff(np.sort(x))) < 0.05:
            continue  # near-duplicate inputs: cholesky(Kuu + 1e-18 I) throws PosDefException in the reference too
        used += 1
        y = np.sin(x) + 0.9 * np.cos(x * 1.6) + 0.4 * rng.random(N)
        s2, jitter = 0.1, 1e-18
        m, S = osv.optimal_variational_posterior(k, x[:, None], jitter, x[:, None], y, s2)
        s = osv.SVGP(k, x, m, np.linalg.cholesky(S), jitter=jitter, centered=True)
        mu, cov = osv.mean_and_cov(s, x)
        mu_e, cov_e = osv.exact_gpr_posterior(k, x[:, None], y, s2, x[:, None])
        worst = max(worst, np.max(np.abs(mu - mu_e)), np.max(np.abs(cov - cov_e)))
        lik = ol.Likelihood("gaussian", s2)
        lp = osv.exact_gpr_logpdf(k, x[:, None], y, s2)
        e = osv.elbo(s, x, y, lik)
        assert e <= lp + 1e-5
        assert lp - e < 1e-3  # and the bound is tight at the optimum (not asserted by the reference)
    assert used >= 3 and worst < 1e-10, (used, worst)


def test_mean_and_var_is_diag_of_cov():
    """AbstractGPs internal-interface consistency (test/SVA...:30-34, :54-58)."""
    for centered in (False, True):
        p = make_problem(seed=3, kind="matern32", N=40, M=7, D=2, centered=centered)
        s, _, _ = oracle_objects(p)
        mu, var = osv.mean_and_var(s, p["X"])
        mu2, cov = osv.mean_and_cov(s, p["X"])
        assert np.allclose(mu, mu2) and np.allclose(var, np.diag(cov), atol=1e-12)


CASES = [
    dict(kind="se", D=2, centered=False, lik="gaussian", method="default"),
    dict(kind="matern52", D=3, centered=False, lik="bernoulli_logit", method="default"),
    dict(kind="matern32", D=1, centered=True, lik="poisson_exp", method="default"),
    dict(kind="se", D=4, centered=True, lik="poisson_exp", method="gauss_hermite", ard=True, mean_const=0.3),
    dict(kind="linear", D=3, centered=False, lik="gaussian", method="gauss_hermite", jitter=1e-3, zdist="random", lengthscale=1.5, M=3),
    dict(kind="matern52", D=2, centered=False, lik="gamma_exp", method="default"),
    dict(kind="se", D=2, centered=False, lik="gamma_exp", method="gauss_hermite"),
    dict(kind="matern32", D=3, centered=False, lik="exponential_exp", method="default"),
    dict(kind="matern52", D=2, centered=False, lik="bernoulli_probit", method="default"),
    dict(kind="se", D=3, centered=True, lik="bernoulli_probit", method="gauss_hermite", n_gh=32),
]


@pytest.mark.parametrize("case", CASES)
def test_reverse_pass_matches_autograd(case):
    """The oracle's gradient == torch autograd through the reference's forward op sequence (rtol 1e-9)."""
    kw = dict(seed=21, N=60, M=8)
    kw.update(case)
    p = make_problem(**kw)
    s, lik, ex = oracle_objects(p)
    val, g = osv.elbo_and_grad(s, p["X"], p["y"], lik, ex, num_data=777.0)
    tval, tg = torch_elbo_and_grad(s, p["X"], p["y"], lik, ex, num_data=777.0)
    assert abs(val - tval) <= 1e-12 * abs(tval)
    assert rel_err(g.m, tg["m"]) < 1e-9 and rel_err(g.Lq, np.tril(tg["Lq"])) < 1e-9 and rel_err(g.Z, tg["Z"]) < 1e-8
    assert rel_err(g.kernel.variance, tg["variance"]) < 1e-9 and rel_err(g.kernel.inv_lengthscale, tg["inv_lengthscale"]) < 1e-8
    if p["kind"] == "linear":
        assert rel_err(g.kernel.c, tg["c"]) < 1e-9
    if p["mean_const"] != 0.0:
        assert rel_err(g.mean_const, tg["mean_const"]) < 1e-9
    if p["lik"] in ("gaussian", "gamma_exp"):
        assert rel_err(g.lik_sigma2, tg["lik_sigma2"]) < 1e-9


def test_gradient_vs_finite_differences():
    """central_fdm(5, 1)-style check (rtol 1e-6 as in test/Laplace...:51-53) of d elbo / d m, d variance."""
    p = make_problem(seed=5, kind="matern52", N=50, M=6, D=2, lik="bernoulli_logit")
    s, lik, ex = oracle_objects(p)
    _, g = osv.elbo_and_grad(s, p["X"], p["y"], lik, ex)
    h = 1e-3
    c = np.array([1.0, -8.0, 8.0, -1.0]) / (12.0 * h)
    for i in range(3):
        vals = []
        for d in (-2, -1, 1, 2):
            m2 = p["m"].copy()
            m2[i] += d * h
            s2 = osv.SVGP(s.kernel, s.Z, m2, s.Lq, jitter=s.jitter)
            vals.append(osv.elbo(s2, p["X"], p["y"], lik, ex))
        assert abs(np.dot(c, vals) - g.m[i]) < 1e-6 * max(1.0, abs(g.m[i]))
    vals = []
    for d in (-2, -1, 1, 2):
        k2 = ok.Kernel(s.kernel.kind, s.kernel.variance + d * h, s.kernel.inv_lengthscale, s.kernel.c)
        vals.append(osv.elbo(osv.SVGP(k2, s.Z, s.m, s.Lq, jitter=s.jitter), p["X"], p["y"], lik, ex))
    assert abs(np.dot(c, vals) - g.kernel.variance) < 1e-6 * max(1.0, abs(g.kernel.variance))


def test_gauss_hermite_against_mpmath():
    """GH-20 expected log-likelihood (Appendix A) vs 40-digit quadrature of the same integral."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    mu, var = 0.3, 0.49
    for kind, y in (("bernoulli_logit", 1.0), ("poisson_exp", 3.0)):
        lik = ol.Likelihood(kind)
        E, _, _, _ = ol.expected_loglik_terms(ol.Expectation("gauss_hermite", 20), lik, np.array([mu]), np.array([var]), np.array([y]))

        def integrand(f):
            if kind == "bernoulli_logit":
                ll = -mp.log(1 + mp.exp(-f))
            else:
                ll = y * f - mp.exp(f) - mp.loggamma(y + 1)
            return ll * mp.npdf(f, mu, mp.sqrt(var))

        exact = mp.quad(integrand, [-mp.inf, mu, mp.inf])
        assert abs(E[0] - float(exact)) < 1e-7 * abs(float(exact))  # truncation error of 20 nodes, not rounding
    # analytic Poisson expectation is exact
    E, _, _, _ = ol.expected_loglik_terms(ol.Expectation("analytic"), ol.Likelihood("poisson_exp"), np.array([mu]), np.array([var]), np.array([3.0]))
    exact = mp.quad(lambda f: (3.0 * f - mp.exp(f) - mp.loggamma(4.0)) * mp.npdf(f, mu, mp.sqrt(var)), [-mp.inf, mu, mp.inf])
    assert abs(E[0] - float(exact)) < 1e-12 * abs(float(exact))


def test_num_data_rescaling_and_chunking():
    """SVA.jl:357-359: sum * num_data / length(y); chunked evaluation (bench CPU baseline) is identical."""
    p = make_problem(seed=9, kind="se", N=300, M=10, D=2, lik="poisson_exp")
    s, lik, ex = oracle_objects(p)
    e1 = osv.elbo(s, p["X"], p["y"], lik, ex)
    e2 = osv.elbo(s, p["X"], p["y"], lik, ex, num_data=3000)
    kl = osv.prior_kl(s)
    assert np.isclose(e2 + kl, 10.0 * (e1 + kl), rtol=1e-12)
    v1, g1 = osv.elbo_and_grad(s, p["X"], p["y"], lik, ex, num_data=3000)
    v2, g2 = osv.elbo_and_grad(s, p["X"], p["y"], lik, ex, num_data=3000, chunk=64)
    assert np.isclose(v1, v2, rtol=1e-13) and rel_err(g1.Z, g2.Z) < 1e-11 and rel_err(g1.Lq, g2.Lq) < 1e-11


def test_optimised_posterior_matches_gpr():
    """test/SVA...:136-186 with the oracle's ELBO and gradient driving Flux.Adam(1e-3) for 20 000 steps: atol 1e-4 (:184-185)."""
    from _train import adam_train, exact_gpr, problem

    x, y, variance, inv_ls, noise, jitter = problem()
    k = ok.Kernel("se", variance, np.array([inv_ls]))
    lik = ol.Likelihood("gaussian", noise)

    def f(m, A):
        v, g = osv.elbo_and_grad(osv.SVGP(k, x, m, A, jitter=jitter), x, y, lik)
        return -v, -g.m, -g.Lq

    from threadpoolctl import threadpool_limits

    with threadpool_limits(limits=1):  # 20 x 20 matrices: BLAS threading only adds overhead
        m, A = adam_train(f, len(x))
    mu, cov = osv.mean_and_cov(osv.SVGP(k, x, m, A, jitter=jitter), x)
    mu_e, cov_e = exact_gpr(x, y, variance, inv_ls, noise)
    assert np.max(np.abs(mu - mu_e)) < 1e-4 and np.max(np.abs(cov - cov_e)) < 1e-4


def test_monte_carlo_stream_and_expectation():
    """The counter-based normal stream of MonteCarloExpectation: Philox4x32-10 known answers (Random123 kat_vectors), N(0,1)
    moments, convergence of the Monte-Carlo expectation to Gauss-Hermite, and its gradient against finite differences."""
    import oracle.likelihoods as L

    # Random123 known-answer vectors for philox4x32-10 (counter, key -> output)
    def raw(counter, key):
        c = [np.uint64(v) for v in counter]
        k = [np.uint64(v) for v in key]
        M = np.uint64(0xFFFFFFFF)
        for _ in range(10):
            p0, p1 = np.uint64(0xD2511F53) * c[0], np.uint64(0xCD9E8D57) * c[2]
            c = [(p1 >> np.uint64(32)) ^ c[1] ^ k[0], p1 & M, (p0 >> np.uint64(32)) ^ c[3] ^ k[1], p0 & M]
            k = [(k[0] + np.uint64(0x9E3779B9)) & M, (k[1] + np.uint64(0xBB67AE85)) & M]
        return [int(v) for v in c]

    assert raw([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert raw([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    eps = L.philox_normal(7, np.arange(50000), 4)
    assert abs(eps.mean()) < 0.01 and abs(eps.std() - 1.0) < 0.01 and np.all(np.isfinite(eps))
    assert np.array_equal(L.philox_normal(7, np.array([123]), 4), eps[123:124])  # counter-based: independent of batching
    mu, var, y = np.full(1, 0.3), np.full(1, 0.49), np.ones(1)
    lik = L.Likelihood("bernoulli_logit")
    gh = L.expected_loglik_terms(L.Expectation("gauss_hermite", 20), lik, mu, var, y)
    mc = L.expected_loglik_terms(L.Expectation("monte_carlo", 200000, 3), lik, mu, var, y)
    assert abs(mc[0][0] - gh[0][0]) < 5e-3 and abs(mc[1][0] - gh[1][0]) < 5e-3 and abs(mc[2][0] - gh[2][0]) < 5e-3
    ex = L.Expectation("monte_carlo", 16, 11)
    E, dmu, dvar, _ = L.expected_loglik_terms(ex, lik, mu, var, y)
    h = 1e-6
    fd_mu = (L.expected_loglik_terms(ex, lik, mu + h, var, y)[0] - L.expected_loglik_terms(ex, lik, mu - h, var, y)[0]) / (2 * h)
    fd_var = (L.expected_loglik_terms(ex, lik, mu, var + h, y)[0] - L.expected_loglik_terms(ex, lik, mu, var - h, y)[0]) / (2 * h)
    assert abs(fd_mu[0] - dmu[0]) < 1e-8 and abs(fd_var[0] - dvar[0]) < 1e-8


def test_third_party_pin_scikit_learn_kernels_and_gpr():
    """Third-party pins for what no reference test constrains (SURVEY.md section 8c "parity unpinned"): the covariance functions on
    D > 1 inputs and the Gaussian SVGP bound.  scikit-learn's kernels are an unrelated implementation of the same definitions
    (KernelFunctions: SqExponential exp(-d^2/2), Matern32 (1+sqrt3 d) exp(-sqrt3 d), Matern52 (1+sqrt5 d+5d^2/3) exp(-sqrt5 d),
    Linear x.y + c; scale transforms multiply the inputs by 1/lengthscale): values AND hyper-parameter gradients agree to 1e-12,
    for isotropic and ARD length scales, D = 8.  And with Z = X and the optimal q the ELBO equals scikit-learn's exact GPR log
    marginal likelihood up to the jitter (test/SparseVariationalApproximationModule.jl:126-133 states the same about AbstractGPs)."""
    skl = pytest.importorskip("sklearn.gaussian_process")
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, DotProduct, Matern, WhiteKernel

    rng = np.random.default_rng(11)
    X, Z = rng.normal(size=(40, 8)), rng.normal(size=(13, 8))
    for ard in (False, True):
        ls = rng.uniform(1.5, 3.0, size=8) if ard else np.array([2.2])
        for kind, sk in ((ok.SE, RBF(ls if ard else ls[0])), (ok.MATERN32, Matern(ls if ard else ls[0], nu=1.5)), (ok.MATERN52, Matern(ls if ard else ls[0], nu=2.5))):
            k = ok.Kernel(kind, 1.7, 1.0 / ls)
            full = ConstantKernel(1.7) * sk
            assert np.max(np.abs(ok.kernelmatrix(k, X, Z) - full(X, Z))) < 1e-13
            Kx, dK = full(X, eval_gradient=True)  # dK[..., j]: derivative w.r.t. log(theta_j), theta = (variance, lengthscales)
            assert np.max(np.abs(ok.kernelmatrix(k, X) - Kx)) < 1e-13
            W = rng.normal(size=Kx.shape)
            W = W + W.T
            _, _, kg = ok.kernelmatrix_pullback(k, X, None, W)
            g_sk = np.einsum("ij,ijp->p", W, dK)
            g_or = np.concatenate([[kg.variance * 1.7], kg.inv_lengthscale * (-1.0 / ls**2) * ls])
            assert np.max(np.abs(g_or - g_sk)) < 1e-11 * np.max(np.abs(g_sk))
    klin = ok.Kernel(ok.LINEAR, 1.0, np.array([1.0]), 0.4)
    assert np.max(np.abs(ok.kernelmatrix(klin, X, Z) - DotProduct(sigma_0=np.sqrt(0.4))(X, Z))) < 1e-13
    # Gaussian SVGP bound at Z = X with the optimal q  ==  exact GPR log marginal likelihood (scikit-learn), D = 3, Matern52
    n, s2 = 30, 0.3
    X = rng.normal(size=(n, 3))
    y = np.sin(X @ rng.normal(size=3)) + 0.3 * rng.normal(size=n)
    k = ok.Kernel(ok.MATERN52, 1.3, np.array([1.0 / 1.4]))
    gpr = skl.GaussianProcessRegressor(kernel=ConstantKernel(1.3) * Matern(1.4, nu=2.5) + WhiteKernel(s2), optimizer=None).fit(X, y)
    K = ok.kernelmatrix(k, X)
    S = K - K @ np.linalg.solve(K + s2 * np.eye(n), K)
    mq = K @ np.linalg.solve(K + s2 * np.eye(n), y)
    S = 0.5 * (S + S.T) + 1e-12 * np.eye(n)
    s = osv.SVGP(k, X, mq, np.linalg.cholesky(S), jitter=1e-10, centered=True)
    val = osv.elbo(s, X, y, ol.Likelihood("gaussian", s2), ol.Expectation())
    assert abs(val - gpr.log_marginal_likelihood_value_) < 1e-5


@pytest.mark.parametrize("op", ["sum", "product"])
@pytest.mark.parametrize("centered", [False, True])
def test_kernel_sum_product_gradient_vs_finite_differences(op, centered):
    """KernelSum / KernelProduct of stationary components under a shared ARD transform (KernelFunctions `k1 + k2`, `k1 * k2`, public through
    `@reexport using AbstractGPs`, src/ApproximateGPs.jl:5): the oracle's hand-derived reverse pass against 5-point finite differences
    of its own forward pass, for every kernel parameter (outer variance, outer ARD scales, component variances and lengthscales) and Z."""
    p = make_problem(seed=11, kind="se", N=60, M=7, D=3, lik="poisson_exp", centered=centered, ard=True)
    comps = (("se", 0.7, 1.3), ("matern32", 1.1, 0.6), ("matern52", 0.4, 2.0))
    _, lik, ex = oracle_objects(p)

    def build(variance, inv, comps, Z):
        k = ok.Kernel(op, variance, inv, 0.0, comps)
        return osv.SVGP(k, Z, p["m"], p["A"], jitter=1e-6, centered=centered)

    def val(variance=p["variance"], inv=p["inv"], comps=comps, Z=p["Z"]):
        return osv.elbo(build(variance, inv, comps, Z), p["X"], p["y"], lik, ex, num_data=200)

    _, g = osv.elbo_and_grad(build(p["variance"], p["inv"], comps, p["Z"]), p["X"], p["y"], lik, ex, num_data=200)
    h = 1e-3
    cf = np.array([1.0, -8.0, 8.0, -1.0]) / (12.0 * h)
    steps = (-2, -1, 1, 2)

    def fd(fn):
        return float(np.dot(cf, [fn(d * h) for d in steps]))

    def close(a, b):
        assert abs(a - b) < 2e-6 * max(1.0, abs(b)), (a, b)

    close(fd(lambda e: val(variance=p["variance"] + e)), g.kernel.variance)
    for d in range(3):
        close(fd(lambda e: val(inv=p["inv"] + e * np.eye(3)[d])), g.kernel.inv_lengthscale[d])
    for i in range(3):
        def with_v(e, i=i):
            c2 = list(comps)
            c2[i] = (comps[i][0], comps[i][1] + e, comps[i][2])
            return val(comps=tuple(c2))

        def with_s(e, i=i):
            c2 = list(comps)
            c2[i] = (comps[i][0], comps[i][1], comps[i][2] + e)
            return val(comps=tuple(c2))

        close(fd(with_v), g.kernel.comp_variance[i])
        close(fd(with_s), g.kernel.comp_inv_lengthscale[i])
    E = np.zeros_like(p["Z"])
    E[2, 1] = 1.0
    close(fd(lambda e: val(Z=p["Z"] + e * E)), g.Z[2, 1])
    # a one-component sum is the plain kernel
    k1 = ok.Kernel("sum", 1.3, p["inv"], 0.0, (("matern52", 1.0, 1.0),))
    k0 = ok.Kernel("matern52", 1.3, p["inv"])
    assert np.allclose(ok.kernelmatrix(k1, p["X"], p["Z"]), ok.kernelmatrix(k0, p["X"], p["Z"]), rtol=0, atol=1e-15)
