#!/usr/bin/env python
"""C1 (BASELINE.json configs[0], examples/a-regression/script.jl): SVGP GaussianLikelihood, SqExponential, N = 10 000 1-D
points, M = 20 / 50 inducing points, minibatches of 100 with num_data = N.  This case is latency-bound: the figure of merit
is the time of one ELBO+gradient evaluation (one optimiser step of the example's 30 000).  Writes one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import agp_b200 as agp
    from oracle import kernels as ok, likelihoods as ol, svgp as osv

    rng = np.random.default_rng(1234)
    N = 10_000
    x = rng.uniform(-1, 1, N)
    y = np.sin(3 * np.pi * x) + 0.3 * np.cos(9 * np.pi * x) + 0.5 * np.sin(7 * np.pi * x) + 0.3 * rng.normal(size=N)
    ctx = agp.default_context()
    ds = agp.DeviceData(x, y, ctx=ctx)
    out = {"workload": "C1 a-regression: N=1e4, D=1, batch=100, num_data=N, SE(1.3, 0.3), sigma2=0.3, jitter=1e-5"}
    for M in (20, 50):
        z = x[:M].copy()
        f = agp.GP(1.3 * agp.with_lengthscale(agp.SqExponentialKernel(), 0.3))
        sva = agp.SparseVariationalApproximation(f(z, 1e-5), agp.MvNormal(np.zeros(M), chol_lower=np.eye(M)))
        fx = agp.FiniteGP(f, ds, 0.3)
        steps = 300
        for w in range(20):
            agp.elbo_and_gradient(sva, fx, None, num_data=N, offset=100 * w, count=100)
        t0 = time.perf_counter()
        for s in range(steps):
            val, g = agp.elbo_and_gradient(sva, fx, None, num_data=N, offset=100 * (s % 100), count=100)
        t = (time.perf_counter() - t0) / steps
        l0 = ctx.launch_count()
        agp.elbo_and_gradient(sva, fx, None, num_data=N, offset=0, count=100)
        launches = ctx.launch_count() - l0
        # full-batch evaluation for reference
        t1 = time.perf_counter()
        for _ in range(20):
            agp.elbo_and_gradient(sva, fx, None, num_data=N)
        tf = (time.perf_counter() - t1) / 20
        s_or = osv.SVGP(ok.Kernel("se", 1.3, np.array([1 / 0.3])), z, np.zeros(M), np.eye(M), jitter=1e-5)
        lik = ol.Likelihood("gaussian", 0.3)
        tc0 = time.perf_counter()
        for s in range(100):
            lo = 100 * (s % 100)
            osv.elbo_and_grad(s_or, x[lo:lo + 100], y[lo:lo + 100], lik, ol.Expectation(), num_data=N)
        tc = (time.perf_counter() - tc0) / 100
        out[f"M{M}"] = {"gpu_us_per_minibatch_step": 1e6 * t, "gpu_kernel_launches_per_step": launches, "gpu_us_full_batch_N1e4": 1e6 * tf,
                        "cpu_port_us_per_minibatch_step": 1e6 * tc, "elbo": val}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
