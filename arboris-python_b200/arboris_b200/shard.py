"""Sharding of a batch of independent worlds over the GPUs of one box.

Worlds never interact (the reference ``World`` owns all of its state,
core.py:341-436), so rank r of R steps the contiguous block
``shard_range(total, r, R)`` and the step path needs **no collective**.
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) only carries the
few scalars of a report: max of the timed duration, sums of counters.
"""
import os


def env_rank():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(total, rank, world_size):
    """Worlds [w0, w1) of rank ``rank``: contiguous blocks, the first
    ``total % world_size`` ranks take one world more."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world size %d" % (rank, world_size))
    if total < 0:
        raise ValueError("negative number of worlds")
    q, r = divmod(int(total), int(world_size))
    w0 = rank*q + min(rank, r)
    return w0, w0 + q + (1 if rank < r else 0)


def block_ranges(total, chunks, align=32):
    """Column blocks [w0, w1) of a host-resident batch (``batch.HostPipeline``): ``chunks`` is a
    number of equal blocks, or the relative sizes of the blocks (e.g. ``(1, 2, 2, 2, 1)``: small
    first and last blocks shorten the exposed first copy-in / last copy-out, large middle blocks
    keep the kernels efficient).  Relative blocks start on multiples of ``align`` worlds (a tile of
    the scratch); empty blocks are dropped; the blocks cover [0, total) exactly, in order."""
    total = int(total)
    if total < 0:
        raise ValueError("negative number of worlds")
    if isinstance(chunks, (list, tuple)):
        if not chunks or any(c < 0 for c in chunks) or sum(chunks) <= 0:
            raise ValueError("relative block sizes must be non-negative with a positive sum")
        tot = float(sum(chunks))
        edges = [0]
        for k in range(len(chunks)):
            e = int(round(total*sum(chunks[:k + 1])/tot/float(align)))*align
            edges.append(min(total, max(edges[-1], e)))
        edges[-1] = total
        ranges = list(zip(edges[:-1], edges[1:]))
    else:
        n = max(1, min(int(chunks), (total + align - 1)//align))
        ranges = [shard_range(total, k, n) for k in range(n)]
    return [r for r in ranges if r[1] > r[0]]


def reduce_report(maxes, sums, device=None, group=None):
    """All-reduce a report: ``maxes`` (e.g. timed milliseconds) with MAX over
    ranks, ``sums`` (world counts, launch counts, non-finite counts) with SUM.
    Returns (list, list) of floats, identical on every rank.  Without an
    initialised process group the inputs come back unchanged."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [float(x) for x in maxes], [float(x) for x in sums]
    tm = torch.tensor([float(x) for x in maxes], dtype=torch.float64, device=device)
    ts = torch.tensor([float(x) for x in sums], dtype=torch.float64, device=device)
    if tm.numel():
        dist.all_reduce(tm, op=dist.ReduceOp.MAX, group=group)
    if ts.numel():
        dist.all_reduce(ts, op=dist.ReduceOp.SUM, group=group)
    return tm.tolist(), ts.tolist()
