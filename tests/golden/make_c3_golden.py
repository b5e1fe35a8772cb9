"""Generates tests/golden/laplace_c3_golden.npz: the oracle (oracle/laplace.py, the NumPy restatement of
Laplace.jl:201-276 and :330-369) on BASELINE.json configs[2] / SURVEY.md section 8(d) "C3" at its FULL size --
N = 8192 points of U(0,10)^2 (seed 3), SqExponential variance 1 / lengthscale 1, K = k(X,X) + 1e-8 I,
y ~ Bernoulli(logistic(3 sin x1)), f_init = 0, maxiter 100.  About 70 s and 4 GB on 8 cores, which is why the GPU
test (tests/test_gpu_laplace.py::test_c3_full_size) compares against this fixture instead of re-running the oracle on the GPU box
(AGP_C3_ORACLE=1 makes it re-run the oracle there as well).

Run from the repository root:  python tests/golden/make_c3_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import kernels as ok, laplace as olap, likelihoods as ol  # noqa: E402


def c3_problem(n=8192):
    rng = np.random.default_rng(3)
    X = rng.uniform(0, 10, size=(n, 2))
    y = (rng.random(n) < 1 / (1 + np.exp(-3 * np.sin(X[:, 0])))).astype(np.float64)
    return X, y


def c3_oracle(X, y):
    k = ok.Kernel(ok.SE, 1.0, np.array([1.0]))
    K = ok.kernelmatrix(k, X) + 1e-8 * np.eye(len(y))
    lml, Kbar, f_opt, steps = olap.lml_and_grad_K(ol.Likelihood("bernoulli_logit"), y, K)
    dX, _, kg = ok.kernelmatrix_pullback(k, X, None, Kbar)
    return dict(lml=lml, steps=steps, f_opt=f_opt, dvariance=kg.variance, dinv_lengthscale=kg.inv_lengthscale, dX=dX)


if __name__ == "__main__":
    X, y = c3_problem()
    r = c3_oracle(X, y)
    print("lml", r["lml"], "steps", r["steps"], "dvariance", r["dvariance"], "dinv_lengthscale", r["dinv_lengthscale"])
    np.savez_compressed(os.path.join(HERE, "laplace_c3_golden.npz"), x_checksum=np.array([X.sum(), y.sum()]), **r)
