"""GPU parity of AGP_COMPUTE_F64_EMU against the FLOAT64 NumPy oracle at the FLOAT64 tolerance (1e-10): the reverse pass's point-sum product
G += (dv A) A^T (S6, the pullback of the two M x N . N x M products behind SVA.jl:251) and the forward product C = Bt^T A (S2, `B' * A` of SVA.jl:251)
run as FP64-accurate INT8-slice products on the tcgen05 tensor path (csrc/i8emu.cuh: seven round-to-nearest 7-bit slices per element under one power-of-two scale per inducing row, 28 exact
slice products in INT32 tensor memory, Float64 recombination); every other stage is the Float64 mode's.  The mode is opt-in."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _cases import agp_objects, compare_grads, make_problem, oracle_objects, record_parity  # noqa: E402

from oracle import svgp as osv  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def agp():
    import agp_b200

    return agp_b200


def _run(agp, p, num_data=None, expect_engine=True):
    s, lik, ex = oracle_objects(p)
    ref, rg = osv.elbo_and_grad(s, p["X"], p["y"], lik, ex, num_data=num_data)
    sva, lfx, quad, _ = agp_objects(agp, p)
    val, g = agp.elbo_and_gradient(sva, lfx, p["y"], num_data=num_data, quadrature=quad, dtype="f64emu")
    v64, g64 = agp.elbo_and_gradient(sva, lfx, p["y"], num_data=num_data, quadrature=quad)
    errs = {"elbo": abs(val - ref) / abs(ref), **compare_grads(g, rg, p)}
    label = f"f64emu {p['kind']} D={p['X'].shape[1]} M={len(p['m'])} N={len(p['y'])} cent={p['centered']} {p['lik']}/{p['method']}"
    print(f"\n[{label}] elbo={val:.10f} " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()))
    record_parity(label, errs, tol=TOL)
    if expect_engine:  # S2 (forward pass) runs on the engine too: the ELBO agrees with the Float64 mode's far inside the tolerance, not to the last bit
        assert abs(val - v64) <= 1e-12 * abs(v64)
    else:
        assert val == v64
    assert np.all(np.triu(g.Lq, 1) == 0.0)
    if expect_engine:  # the INT8 product really ran: G, hence dLq, cannot agree with the DMMA result to the last bit
        assert not np.array_equal(g.Lq, g64.Lq)
    else:  # below the engine's size threshold the mode is the Float64 mode
        assert np.array_equal(g.Lq, g64.Lq)
    for k, v in errs.items():
        assert v < TOL, (k, v)


def test_f64emu_c2_twin_is_plain_f64(agp):
    # M = 512: below the engine's threshold (the DMMA SYRK is as fast there), the mode falls back to Float64 arithmetic
    _run(agp, make_problem(seed=2, kind="matern52", N=4096, M=512, D=8, lik="bernoulli_logit", lengthscale=np.sqrt(8.0), variance=1.0), num_data=1e6, expect_engine=False)


def test_f64emu_matern52_m768(agp):
    _run(agp, make_problem(seed=12, kind="matern52", N=4096, M=768, D=8, lik="bernoulli_logit", lengthscale=np.sqrt(8.0), variance=1.0), num_data=1e6)


def test_f64emu_c4_twin(agp):
    _run(agp, make_problem(seed=4, kind="se", N=4096, M=1024, D=8, lik="poisson_exp", lengthscale=np.sqrt(8.0), variance=1.0), num_data=1e7)


def test_f64emu_c5_shape_ragged(agp):
    # M = 2048, D = 16 (BASELINE config 5's shape), N not a multiple of the engine's 128-point k-block
    _run(agp, make_problem(seed=54, kind="se", N=2500, M=2048, D=16, lik="gaussian", lengthscale=4.0, variance=1.0, zdist="random"), num_data=1e8)


def test_f64emu_multi_chunk_and_wide_rows(agp, monkeypatch):
    # several launch groups (G accumulates across them) and inducing rows whose entries span many decades (short length scale)
    monkeypatch.setenv("AGP_CHUNK_COLS", "2560")
    _run(agp, make_problem(seed=9, kind="matern32", N=6000, M=800, D=4, lik="poisson_exp", lengthscale=0.8, zdist="random"), num_data=1e5)


def test_f64emu_small_problem_is_plain_f64(agp):
    _run(agp, make_problem(seed=21, kind="matern32", N=777, M=50, D=3, ard=True, mean_const=0.4, lik="gaussian"), num_data=5000.0, expect_engine=False)
