"""Committed golden fixtures (tests/golden/, see make_golden.py): the oracle must reproduce them on the CPU
and the CUDA path must match them on the GPU (through the C ABI), independent of any random stream."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _cases import agp_objects, oracle_objects, rel_err  # noqa: E402

from oracle import laplace as olap, svgp as osv  # noqa: E402

with open(os.path.join(HERE, "golden", "svgp_golden.json")) as f:
    SVGP = json.load(f)["cases"]
with open(os.path.join(HERE, "golden", "laplace_golden.json")) as f:
    LAPLACE = json.load(f)


def _problem(case):
    p = dict(case["inputs"])
    for k in ("X", "y", "Z", "m", "A", "inv"):
        p[k] = np.array(p[k], dtype=np.float64)
    return p


@pytest.mark.parametrize("case", SVGP, ids=[c["name"] for c in SVGP])
def test_oracle_reproduces_svgp_fixture(case):
    p, o = _problem(case), case["outputs"]
    s, lik, ex = oracle_objects(p)
    val, g = osv.elbo_and_grad(s, p["X"], p["y"], lik, ex, num_data=case["num_data"])
    assert abs(val - o["elbo"]) <= 1e-13 * abs(o["elbo"])
    assert rel_err(g.m, o["dm"]) < 1e-12 and rel_err(g.Lq, o["dLq"]) < 1e-12 and rel_err(g.Z, o["dZ"]) < 1e-12
    assert abs(osv.prior_kl(s) - o["kl"]) <= 1e-13 * abs(o["kl"])


def test_oracle_reproduces_laplace_fixture():
    X, y = np.array(LAPLACE["X"]), np.array(LAPLACE["y"])
    assert np.array_equal(X, olap.generate_data()[0]) and np.array_equal(y, olap.generate_data()[1])
    for pt in LAPLACE["points"]:
        val, grad, f_opt, steps = olap.objective_and_grad(np.array(pt["theta"]), X, y)
        assert abs(val - pt["objective"]) <= 1e-12 * abs(val) and steps == pt["newton_steps"]
        assert np.max(np.abs(grad - np.array(pt["gradient"]))) < 1e-10
    assert np.allclose(LAPLACE["lbfgs_optimum"], LAPLACE["nelder_mead_optimum"], rtol=1e-4)  # test/Laplace...:159-164


@pytest.mark.gpu
@pytest.mark.parametrize("case", SVGP, ids=[c["name"] for c in SVGP])
def test_cuda_matches_svgp_fixture(case):
    import agp_b200 as agp

    p, o = _problem(case), case["outputs"]
    sva, lfx, quad, _ = agp_objects(agp, p)
    val, g = agp.elbo_and_gradient(sva, lfx, p["y"], num_data=case["num_data"], quadrature=quad)
    assert abs(val - o["elbo"]) < 1e-10 * abs(o["elbo"])
    assert rel_err(g.m, o["dm"]) < 1e-9 and rel_err(g.Lq, o["dLq"]) < 1e-9 and rel_err(g.Z, o["dZ"]) < 1e-9
    assert rel_err(g.variance, o["dvariance"]) < 1e-9 and rel_err(g.inv_lengthscale, o["dinv_lengthscale"]) < 1e-9
    if p["kind"] == "linear":
        assert rel_err(g.linear_c, o["dlinear_c"]) < 1e-9
    if p["mean_const"] != 0.0:
        assert rel_err(g.mean_const, o["dmean_const"]) < 1e-9
    if p["lik"] == "gaussian":
        assert rel_err(g.lik_sigma2, o["dlik_sigma2"]) < 1e-9
    assert abs(agp._prior_kl(sva) - o["kl"]) < 1e-10 * abs(o["kl"])
    mu, var = agp.mean_and_var(agp.posterior(sva), p["X"][:16])
    assert rel_err(mu, o["mu16"]) < 1e-10 and rel_err(var, o["var16"]) < 1e-10


@pytest.mark.gpu
def test_cuda_matches_laplace_fixture():
    import agp_b200 as agp

    X, y = np.array(LAPLACE["X"]), np.array(LAPLACE["y"])
    sp = lambda v: np.logaddexp(0.0, v)
    for pt in LAPLACE["points"]:
        th = np.array(pt["theta"])
        kernel = sp(th[0]) * agp.with_lengthscale(agp.SqExponentialKernel(), sp(th[1]))
        lfx = agp.LatentGP(agp.GP(kernel), agp.BernoulliLikelihood(), 1e-8)(X)
        r = agp.laplace_approx_lml_and_gradient(agp.LaplaceApproximation(), lfx, y)
        sig = 1.0 / (1.0 + np.exp(-th))
        grad = -np.array([r.grad.variance * sig[0], r.grad.inv_lengthscale[0] * (-1.0 / sp(th[1]) ** 2) * sig[1]])
        assert abs(-r.lml - pt["objective"]) < 1e-10 * abs(pt["objective"]) and r.steps == pt["newton_steps"]
        assert np.max(np.abs(grad - np.array(pt["gradient"]))) < 1e-8
        assert rel_err(r.f, pt["f_opt"]) < 1e-10
