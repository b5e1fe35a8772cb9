"""CPU: host-side logic of the data-parallel path.  The N > 1 protocol (contiguous shards, partial sums
scaled with the GLOBAL batch size, one sum-all-reduce of the packed buffer, replicated KL epilogue) is run
with world_size = 2 over gloo, the oracle standing in for the device sweep, and must equal the
single-process evaluation."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def test_shard_range_partitions():
    import agp_b200 as agp

    for n in (0, 1, 7, 10_000_000, 10_000_003):
        for world in (1, 2, 3, 4, 8):
            edges = [agp.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        agp.shard_range(10, 2, 2)


def _pack(val, g):
    return np.concatenate([[val], g.m, g.Lq.ravel(), g.Z.ravel(), [g.kernel.variance], g.kernel.inv_lengthscale, [g.mean_const, g.lik_sigma2]])


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    from _cases import make_problem, oracle_objects

    import agp_b200 as agp
    from oracle import svgp as osv

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = make_problem(seed=77, kind="matern52", N=301, M=9, D=3, lik="poisson_exp")
    s, lik, ex = oracle_objects(p)
    N, num_data = len(p["y"]), 5000.0
    lo, hi = agp.shard_range(N, rank, world)
    # "sweep": this rank's share of the data term, scaled with the global batch (num_data / N, SVA.jl:357-358)
    nd_local = num_data * (hi - lo) / N
    v1, g1 = osv.elbo_and_grad(s, p["X"][lo:hi], p["y"][lo:hi], lik, ex, num_data=nd_local)
    v0, g0 = osv.elbo_and_grad(s, p["X"][lo:hi], p["y"][lo:hi], lik, ex, num_data=0.0)  # = -KL and its gradient
    buf = torch.from_numpy(_pack(v1, g1) - _pack(v0, g0))
    dist.all_reduce(buf)  # the one collective of the step
    total = buf.numpy() + _pack(v0, g0)  # replicated epilogue: subtract KL once
    if rank == 0:
        vf, gf = osv.elbo_and_grad(s, p["X"], p["y"], lik, ex, num_data=num_data)
        np.save(out, np.stack([total, _pack(vf, gf)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_single_process(tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got, ref = np.load(out)
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6)) < 1e-9


def test_reference_interface_objects_without_a_device():
    """Host mirror of the reference's constructors (no device call): likelihood / link resolution, kernel composition,
    SparseVariationalApproximation defaults, argument errors that must be raised before anything crosses the C ABI."""
    import agp_b200 as agp
    from agp_b200 import _lib as L

    assert agp.BernoulliLikelihood().kind == L.LIK_BERNOULLI_LOGIT
    assert agp.BernoulliLikelihood(agp.LogisticLink()).kind == L.LIK_BERNOULLI_LOGIT
    assert agp.BernoulliLikelihood(agp.ProbitLink()).kind == agp.BernoulliLikelihood("probit").kind == L.LIK_BERNOULLI_PROBIT
    with pytest.raises(ValueError):
        agp.BernoulliLikelihood("cloglog")  # unsupported link: ArgumentError, there is no CPU fallback
    assert agp.GammaLikelihood(2.5).sigma2 == 2.5 and agp.ExponentialLikelihood().kind == L.LIK_EXPONENTIAL_EXP
    # variance * (k o ScaleTransform(1 / l)): KernelFunctions composition
    k = 1.3 * agp.with_lengthscale(agp.Matern52Kernel(), 0.5)
    assert k.kind == L.KERNEL_MATERN52 and k.variance == 1.3 and np.allclose(k.inv_lengthscale, [2.0])
    k2 = agp.ARDTransform(agp.SqExponentialKernel(), [1.0, 2.0, 4.0])
    assert np.allclose(k2.inv_lengthscale, [1.0, 2.0, 4.0])
    f = agp.GP(k)
    z = np.linspace(-1, 1, 5)
    sva = agp.SparseVariationalApproximation(f(z, 1e-6), agp.MvNormal(np.zeros(5), chol_lower=np.eye(5)))
    assert not sva.centered  # NonCentered is the default (SVA.jl:93)
    assert agp.SparseVariationalApproximation(agp.Centered(), f(z, 1e-6), agp.MvNormal(np.zeros(5), chol_lower=np.eye(5))).centered
    # laplace_steps installs its own Newton callback (Laplace.jl:409-421)
    lfx = agp.LatentGP(f, agp.BernoulliLikelihood(), 1e-8)(z)
    with pytest.raises(TypeError):
        agp.laplace_steps(lfx, np.ones(5), callback=lambda *a: None)
    # _check_laplace_inputs (Laplace.jl:167-179): zero prior mean, matching lengths
    from agp_b200.laplace_api import _check_laplace_inputs

    with pytest.raises(AssertionError):
        _check_laplace_inputs(lfx, np.ones(4))
    with pytest.raises(AssertionError):
        _check_laplace_inputs(agp.LatentGP(agp.GP(0.5, k), agp.BernoulliLikelihood(), 1e-8)(z), np.ones(5))


def test_compute_dtype_selection_matches_the_header():
    """The host mirror's dtype names map onto the header's AGP_COMPUTE_* values (include/agp.h); unknown names are an ArgumentError."""
    import re

    import numpy as np
    import pytest

    import agp_b200  # noqa: F401  (loads the package)
    from agp_b200 import _lib as L
    from agp_b200 import api

    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "agp.h")).read()
    vals = {name: int(v) for name, v in re.findall(r"#define (AGP_COMPUTE_[A-Z0-9_]+) (\d+)", hdr)}
    assert vals == {"AGP_COMPUTE_F64": L.COMPUTE_F64, "AGP_COMPUTE_F32": L.COMPUTE_F32, "AGP_COMPUTE_F32_TC_SOLVE": L.COMPUTE_F32_TC_SOLVE,
                    "AGP_COMPUTE_F64_EMU": L.COMPUTE_F64_EMU}
    assert api._compute_dtype(None) == api._compute_dtype("f64") == api._compute_dtype(np.float64) == vals["AGP_COMPUTE_F64"]
    assert api._compute_dtype("f32") == api._compute_dtype(np.float32) == vals["AGP_COMPUTE_F32"]
    assert api._compute_dtype("f64emu") == vals["AGP_COMPUTE_F64_EMU"]
    with pytest.raises(ValueError, match="ArgumentError"):
        api._compute_dtype("bf16")
