"""Generates tests/golden/*.json.

The reference is Julia and cannot run in the build image, so these vectors come from two sources:
 * reference-derived known answers copied from the reference's own test-suite (the Laplace optima of
   test/LaplaceApproximationModule.jl:159,168 and the fixed 48-point data set of src/TestUtils.jl:13-28);
 * oracle-generated regression vectors: inputs and outputs of oracle/ (the NumPy restatement) on small
   seeded problems, stored with full precision so that the CPU suite (oracle vs fixture) and the GPU suite
   (CUDA path vs fixture) check the same numbers without depending on a random-number stream.

Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from _cases import make_problem, oracle_objects  # noqa: E402

from oracle import laplace as olap, svgp as osv  # noqa: E402

SVGP_CASES = [
    dict(name="se_gaussian_nc", seed=101, kind="se", N=64, M=8, D=2, lik="gaussian", num_data=640.0),
    dict(name="matern52_bernoulli_gh20_nc", seed=102, kind="matern52", N=96, M=10, D=3, lik="bernoulli_logit", num_data=9600.0),
    dict(name="matern32_poisson_analytic_centered", seed=103, kind="matern32", N=80, M=9, D=1, lik="poisson_exp", centered=True, num_data=None),
    dict(name="se_ard_poisson_gh_mean", seed=104, kind="se", N=72, M=12, D=4, lik="poisson_exp", method="gauss_hermite", n_gh=20, ard=True,
         mean_const=0.25, num_data=7200.0),
    dict(name="linear_gaussian", seed=105, kind="linear", N=50, M=3, D=3, lik="gaussian", jitter=1e-3, zdist="random", lengthscale=1.5, num_data=None),
    dict(name="se_bernoulli_probit_gh20_centered", seed=106, kind="se", N=88, M=9, D=2, lik="bernoulli_probit", centered=True, num_data=880.0),
    dict(name="matern52_gamma_analytic", seed=107, kind="matern52", N=70, M=8, D=3, lik="gamma_exp", num_data=None),
    dict(name="matern32_exponential_mc16", seed=108, kind="matern32", N=60, M=7, D=2, lik="exponential_exp", method="monte_carlo", n_gh=16, num_data=600.0),
]


def tolist(a):
    return np.asarray(a, dtype=np.float64).tolist()


def main():
    out = []
    path = os.path.join(HERE, "svgp_golden.json")
    existing = {}
    if os.path.exists(path):  # committed vectors are kept verbatim; only cases that are not in the file yet are generated
        with open(path) as f:
            existing = {c["name"]: c for c in json.load(f)["cases"]}
    for c in SVGP_CASES:
        c = dict(c)
        name, num_data = c.pop("name"), c.pop("num_data")
        if name in existing:
            out.append(existing[name])
            continue
        p = make_problem(**c)
        s, lik, ex = oracle_objects(p)
        val, g = osv.elbo_and_grad(s, p["X"], p["y"], lik, ex, num_data=num_data)
        mu, var = osv.mean_and_var(s, p["X"][:16])
        out.append(dict(name=name, num_data=num_data,
                        inputs={k: (tolist(v) if isinstance(v, np.ndarray) else v) for k, v in p.items()},
                        outputs=dict(elbo=val, kl=osv.prior_kl(s), dm=tolist(g.m), dLq=tolist(g.Lq), dZ=tolist(g.Z), dvariance=g.kernel.variance,
                                     dinv_lengthscale=tolist(g.kernel.inv_lengthscale), dlinear_c=g.kernel.c, dmean_const=g.mean_const,
                                     dlik_sigma2=g.lik_sigma2, mu16=tolist(mu), var16=tolist(var))))
    with open(os.path.join(HERE, "svgp_golden.json"), "w") as f:
        json.dump(dict(source="oracle-generated regression vectors (oracle/svgp.py); see make_golden.py", cases=out), f)

    X, y = olap.generate_data()
    lap = dict(source="reference known answers (test/LaplaceApproximationModule.jl:159,168; src/TestUtils.jl:13-28) + oracle values at them",
               X=tolist(X), y=tolist(y), theta0=[5.0, 1.0], lbfgs_optimum=[7.709076337653239, 1.51820292019697],
               nelder_mead_optimum=[7.708967951453345, 1.5182348363613536], points=[])
    for theta in ([5.0, 1.0], [1.0, 2.0], [7.709076337653239, 1.51820292019697]):
        val, grad, f_opt, steps = olap.objective_and_grad(np.array(theta), X, y)
        lap["points"].append(dict(theta=theta, objective=val, gradient=tolist(grad), f_opt=tolist(f_opt), newton_steps=steps))
    with open(os.path.join(HERE, "laplace_golden.json"), "w") as f:
        json.dump(lap, f)
    print("wrote", len(out), "SVGP cases and", len(lap["points"]), "Laplace points")


if __name__ == "__main__":
    main()
