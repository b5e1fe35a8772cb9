"""Worker of tests/test_gpu_multi.py: one rank of a data-parallel ELBO + gradient evaluation through the Python mirror
(launched with torchrun, one process per GPU).  Rank 0 compares with the oracle on the full data and prints DP_OK."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from _cases import agp_objects, compare_grads, make_problem, oracle_objects  # noqa: E402

import agp_b200 as agp  # noqa: E402
from oracle import svgp as osv  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = agp.Context(local)
    agp.attach_communicator(ctx, dist)
    p = make_problem(seed=81, kind="matern52", N=5003, M=150, D=3, lik="bernoulli_logit")
    sva, lfx, quad, f = agp_objects(agp, p)
    lo, hi = agp.shard_range(len(p["y"]), rank, world)
    ds = agp.DeviceData(p["X"][lo:hi], p["y"][lo:hi], ctx=ctx)
    lds = agp.LatentGP(f, agp.BernoulliLikelihood(), 1e-18)(ds)
    val, g = agp.elbo_and_gradient(sva, lds, None, num_data=1e5, quadrature=quad, ctx=ctx, global_batch=len(p["y"]))
    fwd = agp.elbo(sva, lds, None, num_data=1e5, quadrature=quad, ctx=ctx, global_batch=len(p["y"]))
    vals = [None] * world
    dist.all_gather_object(vals, (val, fwd, float(np.abs(g.Z).sum())))
    if rank == 0:
        s, lik, ex = oracle_objects(p)
        ref, rg = osv.elbo_and_grad(s, p["X"], p["y"], lik, ex, num_data=1e5)
        errs = compare_grads(g, rg, p)
        assert abs(val - ref) < 1e-10 * abs(ref), (val, ref)
        assert all(v < 1e-9 for v in errs.values()), errs
        assert all(v == vals[0] for v in vals), vals  # every rank returns the same numbers
        assert abs(fwd - val) <= 1e-12 * abs(val)
    # a communicator without global_batch would silently scale the ELBO by the number of ranks: an argument error on every rank
    try:
        agp.elbo(sva, lds, None, num_data=1e5, quadrature=quad, ctx=ctx)
        raise SystemExit("global_batch = 0 with a communicator must be rejected")
    except ValueError:
        pass
    # fewer points than ranks: the ranks with an empty shard contribute zeros and still take part in the collective
    n_tiny = world - 1
    tlo, thi = agp.shard_range(n_tiny, rank, world)
    dst = agp.DeviceData(capacity=1, D=3, ctx=ctx)
    if thi > tlo:
        dst.upload(p["X"][tlo:thi], p["y"][tlo:thi])
    ldt = agp.LatentGP(f, agp.BernoulliLikelihood(), 1e-18)(dst)
    vt, gt = agp.elbo_and_gradient(sva, ldt, None, num_data=1e5, quadrature=quad, ctx=ctx, global_batch=n_tiny, count=thi - tlo)
    if rank == 0:
        reft, rgt = osv.elbo_and_grad(s, p["X"][:n_tiny], p["y"][:n_tiny], lik, ex, num_data=1e5)
        assert abs(vt - reft) < 1e-10 * abs(reft), (vt, reft)
        assert all(v < 1e-9 for v in compare_grads(gt, rgt, p).values())
        print("DP_OK", world, val, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
