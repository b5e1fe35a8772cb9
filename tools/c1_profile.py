import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import agp_b200 as agp
rng = np.random.default_rng(1234); N = 10000
x = rng.uniform(-1, 1, N); y = np.sin(3*np.pi*x) + 0.3*rng.normal(size=N)
ctx = agp.default_context(); ds = agp.DeviceData(x, y, ctx=ctx)
M = 20; z = x[:M].copy()
f = agp.GP(1.3 * agp.with_lengthscale(agp.SqExponentialKernel(), 0.3))
sva = agp.SparseVariationalApproximation(f(z, 1e-5), agp.MvNormal(np.zeros(M), chol_lower=np.eye(M)))
fx = agp.FiniteGP(f, ds, 0.3)
for w in range(20): agp.elbo_and_gradient(sva, fx, None, num_data=N, offset=100*w, count=100)
ctx.profile_read(); ctx.profile(True)
steps = 200
t0 = time.perf_counter()
for s in range(steps): agp.elbo_and_gradient(sva, fx, None, num_data=N, offset=100*(s % 100), count=100)
wall = (time.perf_counter() - t0) / steps
pr = ctx.profile_read(); ctx.profile(False)
print("wall us/step", 1e6*wall)
tot = 0
for k, (ms, cnt) in pr.items():
    if cnt: print(f"  {k:18s} {1e3*ms/steps:8.1f} us/step"); tot += ms
print("  sum of classes", 1e3*tot/steps)
# python-side overhead: time packing alone
from agp_b200.api import _Packed
t0 = time.perf_counter()
for s in range(steps): _Packed(sva, agp.GaussianLikelihood(0.3), None)
print("python _Packed us", 1e6*(time.perf_counter()-t0)/steps)
