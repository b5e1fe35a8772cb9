"""Data-parallel plumbing of the SVGP sweep: the N points are partitioned
contiguously over the ranks, every rank holds the (tiny) replicated parameters, and one all-reduce per
step combines the packed partial sums.  ``torch.distributed`` is only used to hand the NCCL unique id to
the library (the collective itself is the library's own ncclAllReduce on its stream)."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Rows ``[lo, hi)`` of an ``n``-point data set owned by ``rank`` (block partition, remainder spread
    over the first ranks)."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError(f"bad shard request n={n} rank={rank} world={world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def attach_communicator(ctx, dist, device=None) -> None:
    """Create the library's NCCL communicator for ``ctx`` on every rank of an initialised
    ``torch.distributed`` process group (rank 0 creates the unique id, everyone receives it)."""
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(ctx.unique_id()), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        uid = uid.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.broadcast(uid, 0)
    ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
