"""Shared builders: one seeded problem expressed both as oracle objects and as agp (device) objects."""
import numpy as np

from oracle import kernels as ok, likelihoods as ol, svgp as osv

KIND_NAMES = {"se": 0, "matern32": 1, "matern52": 2, "linear": 3}

# measured parity errors of the GPU tests (case label -> {quantity: relative error, "tol": asserted tolerance}); conftest.py writes
# them to gpurun_out/parity_errors.json at the end of a GPU session, the tracked copy is profiles/parity_errors.json
PARITY_ERRORS = {}


def record_parity(label, errs, **extra):
    PARITY_ERRORS[label] = {**{k: float(v) for k, v in errs.items()}, **extra}


def make_problem(seed=0, kind="se", N=300, M=20, D=2, centered=False, lik="gaussian", method="default", n_gh=20, ard=False,
                 mean_const=0.0, jitter=1e-6, lengthscale=None, variance=1.3, zdist="data"):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, D))
    if zdist == "data":
        Z = X[rng.choice(N, size=M, replace=False)] + 1e-2 * rng.normal(size=(M, D))
    else:
        Z = rng.normal(size=(M, D))
    if lengthscale is None:
        lengthscale = 0.5 if D == 1 else np.sqrt(D)
    inv = (1.0 / lengthscale) * (rng.uniform(0.8, 1.2, size=D) if ard else np.ones(1))
    w = rng.normal(size=D)
    g = np.sin(X @ w)
    if lik == "gaussian":
        y = g + 0.3 * rng.normal(size=N)
    elif lik in ("bernoulli_logit", "bernoulli_probit"):
        y = (rng.random(N) < 1 / (1 + np.exp(-2 * g))).astype(np.float64)
    elif lik == "exponential_exp":
        y = rng.exponential(np.exp(0.5 * g))
    elif lik == "gamma_exp":
        y = rng.gamma(2.5, np.exp(0.5 * g))
    else:
        y = rng.poisson(np.exp(0.5 * g)).astype(np.float64)
    m = 0.1 * rng.normal(size=M)
    A = 0.5 * np.eye(M) + 0.01 * np.tril(rng.normal(size=(M, M)))
    A[np.diag_indices(M)] = np.abs(np.diag(A))
    return dict(X=X, y=y, Z=Z, m=m, A=A, kind=kind, variance=variance, inv=inv, c=0.4 if kind == "linear" else 0.0, centered=centered,
                lik=lik, method=method, n_gh=n_gh, mean_const=mean_const, jitter=jitter, sigma2=2.5 if lik == "gamma_exp" else 0.3)


def oracle_objects(p):
    k = ok.Kernel(p["kind"], p["variance"], p["inv"], p["c"])
    s = osv.SVGP(k, p["Z"], p["m"], p["A"], jitter=p["jitter"], centered=p["centered"], mean_const=p["mean_const"])
    lik = ol.Likelihood(p["lik"], p["sigma2"])
    ex = ol.Expectation(p["method"], p["n_gh"], p.get("mc_seed", 0))
    return s, lik, ex


def agp_objects(agp, p, x=None):
    base = {"se": agp.SqExponentialKernel, "matern32": agp.Matern32Kernel, "matern52": agp.Matern52Kernel}.get(p["kind"])
    kb = base() if base else agp.LinearKernel(p["c"])
    inv = p["inv"]
    kern = p["variance"] * (agp.ScaleTransform(kb, inv[0]) if inv.size == 1 else agp.ARDTransform(kb, inv))
    f = agp.GP(p["mean_const"], kern) if p["mean_const"] != 0.0 else agp.GP(kern)
    fz = f(p["Z"], p["jitter"])
    q = agp.MvNormal(p["m"], chol_lower=p["A"])
    sva = agp.SparseVariationalApproximation(agp.Centered() if p["centered"] else agp.NonCentered(), fz, q)
    lik = {"gaussian": agp.GaussianLikelihood(p["sigma2"]), "bernoulli_logit": agp.BernoulliLikelihood(), "bernoulli_probit": agp.BernoulliLikelihood(agp.ProbitLink()), "poisson_exp": agp.PoissonLikelihood(),
           "exponential_exp": agp.ExponentialLikelihood(), "gamma_exp": agp.GammaLikelihood(p["sigma2"])}[p["lik"]]
    quad = {"default": agp.DefaultExpectationMethod(), "analytic": agp.AnalyticExpectation(), "gauss_hermite": agp.GaussHermiteExpectation(p["n_gh"]),
            "monte_carlo": agp.MonteCarloExpectation(p["n_gh"], p.get("mc_seed", 0))}[p["method"]]
    lfx = agp.LatentGP(f, lik, 1e-18)(p["X"] if x is None else x)
    return sva, lfx, quad, f


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def compare_grads(g, rg, p):
    """dict of relative errors (scaled by the max-abs of each oracle array)."""
    out = dict(m=rel_err(g.m, rg.m), Lq=rel_err(g.Lq, rg.Lq), Z=rel_err(g.Z, rg.Z), variance=rel_err(g.variance, rg.kernel.variance),
               inv_lengthscale=rel_err(g.inv_lengthscale, rg.kernel.inv_lengthscale))
    if p["kind"] == "linear":
        out["linear_c"] = rel_err(g.linear_c, rg.kernel.c)
    if p["mean_const"] != 0.0:
        out["mean_const"] = rel_err(g.mean_const, rg.mean_const)
    if p["lik"] in ("gaussian", "gamma_exp"):
        out["lik_sigma2"] = rel_err(g.lik_sigma2, rg.lik_sigma2)
    return out
