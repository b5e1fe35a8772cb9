"""GPU parity of the Float32 fast mode (AGP_COMPUTE_F32: S2 / S4 / S5 / S6 as 3xTF32 split products on the tcgen05 tensor cores,
csrc/f32sweep.cuh) against the FLOAT64 NumPy oracle: north_star's tolerance for Float32 is relative 1e-4 on the ELBO and on every
gradient buffer (relative to its max-abs entry).  The reference is type-generic (SparseVariationalApproximation{P,Tfz,Tq},
SVA.jl:59-62, accepts Float32 GPs), so its own Float32 result differs from Float64 by the same order."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _cases import agp_objects, compare_grads, make_problem, oracle_objects, record_parity  # noqa: E402

from oracle import svgp as osv  # noqa: E402

pytestmark = pytest.mark.gpu
TOL32 = 1e-4


@pytest.fixture(scope="module")
def agp():
    import agp_b200

    return agp_b200


def _run(agp, p, num_data=None, tol=TOL32, dtype="f32"):
    s, lik, ex = oracle_objects(p)
    ref, rg = osv.elbo_and_grad(s, p["X"], p["y"], lik, ex, num_data=num_data)
    sva, lfx, quad, _ = agp_objects(agp, p)
    val, g = agp.elbo_and_gradient(sva, lfx, p["y"], num_data=num_data, quadrature=quad, dtype=dtype)
    fwd = agp.elbo(sva, lfx, p["y"], num_data=num_data, quadrature=quad, dtype=dtype)
    v64, g64 = agp.elbo_and_gradient(sva, lfx, p["y"], num_data=num_data, quadrature=quad)
    errs = {"elbo": abs(val - ref) / abs(ref), **compare_grads(g, rg, p)}
    label = f"{dtype} {p['kind']} D={p['X'].shape[1]} M={len(p['m'])} N={len(p['y'])} cent={p['centered']} {p['lik']}/{p['method']}"
    print(f"\n[{label}] elbo={val:.8f} (f64 {v64:.8f}) " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()))
    record_parity(label, errs, tol=tol)
    assert abs(fwd - val) <= 1e-6 * abs(val)
    assert np.all(np.triu(g.Lq, 1) == 0.0)
    assert val != v64 or len(p["y"]) < 64  # the Float32 path really ran (it cannot be bit-identical to the Float64 one)
    for k, v in errs.items():
        assert v < tol, (k, v)


def test_f32_c2_twin(agp):
    _run(agp, make_problem(seed=2, kind="matern52", N=4096, M=512, D=8, lik="bernoulli_logit", lengthscale=np.sqrt(8.0), variance=1.0), num_data=1e6)


def test_f32_c4_twin(agp):
    _run(agp, make_problem(seed=4, kind="se", N=4096, M=1024, D=8, lik="poisson_exp", lengthscale=np.sqrt(8.0), variance=1.0), num_data=1e7)


@pytest.mark.parametrize("centered", [False, True])
def test_f32_small_and_ragged(agp, centered):
    # M = 50 (one padded 128-block), N = 777 (not a multiple of the 128-point tile), ARD, constant mean
    _run(agp, make_problem(seed=21, kind="matern32", N=777, M=50, D=3, ard=True, centered=centered, mean_const=0.4, lik="gaussian"), num_data=5000.0)


def test_f32_multi_chunk(agp, monkeypatch):
    monkeypatch.setenv("AGP_CHUNK_COLS", "384")
    _run(agp, make_problem(seed=8, kind="matern52", N=1000, M=140, D=3, lik="bernoulli_logit"), num_data=12345)


def test_f32_tensor_core_solve_variant(agp):
    """AGP_COMPUTE_F32_TC_SOLVE: the reverse-pass solve on the tensor cores as well (product with the explicit inverse).  Its error
    on dZ / d theta carries a factor cond(Lk): within 1e-4 on the Matern52 twin of config 2, 2e-4 (measured 1.4e-4) on the
    SqExponential M = 1024 twin of config 4 -- which is why it is not the default Float32 mode."""
    _run(agp, make_problem(seed=2, kind="matern52", N=4096, M=512, D=8, lik="bernoulli_logit", lengthscale=np.sqrt(8.0), variance=1.0), num_data=1e6,
         dtype="f32_tc_solve")
    _run(agp, make_problem(seed=4, kind="se", N=4096, M=1024, D=8, lik="poisson_exp", lengthscale=np.sqrt(8.0), variance=1.0), num_data=1e7, tol=3e-4,
         dtype="f32_tc_solve")
