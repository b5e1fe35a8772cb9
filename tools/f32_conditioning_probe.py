import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import agp_b200 as agp
from _cases import agp_objects, compare_grads, make_problem, oracle_objects
from oracle import svgp as osv
cases = {"m4096": make_problem(seed=55, kind="matern52", N=3000, M=4096, D=4, lik="poisson_exp", zdist="random", lengthscale=1.0),
         "c4twin": make_problem(seed=4, kind="se", N=4096, M=1024, D=8, lik="poisson_exp", lengthscale=np.sqrt(8.0)),
         "m1024_ls1": make_problem(seed=56, kind="se", N=3000, M=1024, D=4, lik="poisson_exp", zdist="random", lengthscale=1.5)}
for name, p in cases.items():
    s, lik, ex = oracle_objects(p)
    ref, rg = osv.elbo_and_grad(s, p["X"], p["y"], lik, ex, num_data=1e6)
    sva, lfx, quad, _ = agp_objects(agp, p)
    val, g = agp.elbo_and_gradient(sva, lfx, p["y"], num_data=1e6, quadrature=quad, dtype="f32")
    errs = compare_grads(g, rg, p)
    print(name, os.environ.get("AGP_F32_S1", "i8"), "elbo_rel=%.2e" % (abs(val - ref) / abs(ref)), {k: "%.1e" % v for k, v in errs.items()}, flush=True)
