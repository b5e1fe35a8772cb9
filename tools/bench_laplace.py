#!/usr/bin/env python
"""C3 (BASELINE.json configs[2]): LaplaceApproximation Bernoulli-logit on a dense N = 8192 latent GP.
Reports Newton-loop time, time per Newton iteration and FP64 TFLOP/s against N^3/3 + 6 N^2 flop per iteration
(SURVEY.md section 8d), plus the CPU restatement on a reduced N.  Writes one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-n", type=int, default=2048)
    args = ap.parse_args()
    import agp_b200 as agp

    rng = np.random.default_rng(3)
    n = args.n
    X = rng.uniform(0, 10, size=(n, 2))
    y = (rng.random(n) < 1 / (1 + np.exp(-3 * np.sin(X[:, 0])))).astype(np.float64)
    f = agp.GP(1.0 * agp.with_lengthscale(agp.SqExponentialKernel(), 1.0))
    lfx = agp.LatentGP(f, agp.BernoulliLikelihood(), 1e-8)(X)
    la = agp.LaplaceApproximation(maxiter=100)
    ctx = agp.default_context()
    out = {"workload": f"C3 Laplace Bernoulli-logit, SE, N={n}, D=2, FP64", "n": n}
    for name, fn in (("lml", lambda: agp.laplace_api._run(ctx, **agp.laplace_api._check_laplace_inputs(lfx, y, **la.newton_kwargs))),
                     ("lml_and_grad", lambda: agp.laplace_approx_lml_and_gradient(la, lfx, y))):
        fn()
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            r = fn()
            ts.append(time.perf_counter() - t0)
        t = float(np.median(ts))
        iters = r.steps + (0 if name == "lml" else 1)  # the gradient path re-runs the intermediates once
        flop_iter = n**3 / 3 + 6.0 * n * n
        out[name] = {"seconds": t, "newton_steps": r.steps, "lml": r.lml, "ms_per_newton_iteration": 1e3 * t / iters if name == "lml" else None,
                     "tflops_newton": flop_iter * iters / t / 1e12 if name == "lml" else None}
    # CPU restatement (oracle) at a reduced size, all host threads
    from oracle import kernels as ok, laplace as olap, likelihoods as ol

    m = args.cpu_n
    K = ok.kernelmatrix(ok.Kernel("se", 1.0, np.array([1.0])), X[:m]) + 1e-8 * np.eye(m)
    t0 = time.perf_counter()
    _, _, steps = olap.laplace_f_and_lml(ol.Likelihood("bernoulli_logit"), y[:m], K)
    tc = time.perf_counter() - t0
    out["cpu_port"] = {"n": m, "seconds": tc, "newton_steps": steps, "gflops": (m**3 / 3 + 6.0 * m * m) * (steps + 1) / tc / 1e9,
                       "cores": len(os.sched_getaffinity(0))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
