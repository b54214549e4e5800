"""Multi-GPU host logic on CPU: world sharding and the diagnostics reduction,
world_size 2 over gloo (the step path itself has no collective)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from arboris_b200 import scenarios
from arboris_b200.flatten import flatten
from arboris_b200.shard import reduce_report, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 64, 262144, 262145):
        for R in (1, 2, 3, 4, 8):
            blocks = [shard_range(total, r, R) for r in range(R)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            for a, b in zip(blocks, blocks[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_reduce_report_without_group_is_identity():
    assert reduce_report([1.5], [2, 3]) == ([1.5], [2.0, 3.0])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        w0, w1 = shard_range(total, rank, world_size)
        model = flatten(scenarios.BUILDERS["human36_contact"]())
        gp, gv = scenarios.initial_states(model, "human36_contact", w0, w1)
        # every rank sees only its block; the checksum of the blocks must equal the whole
        ms = 10. + rank            # pretend timings: the report keeps the slowest rank
        maxes, sums = reduce_report([ms], [w1 - w0, float(gp.sum()), float(gv.sum())])
        if rank == 0:
            torch.save({"maxes": maxes, "sums": sums}, out)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_cover_the_batch(tmp_path):
    total = 10
    out = str(tmp_path / "report.pt")
    mp.spawn(_worker, args=(2, _free_port(), total, out), nprocs=2, join=True)
    rep = torch.load(out)
    model = flatten(scenarios.BUILDERS["human36_contact"]())
    gp, gv = scenarios.initial_states(model, "human36_contact", 0, total)
    assert rep["maxes"] == [11.]
    assert rep["sums"][0] == total
    np.testing.assert_allclose(rep["sums"][1], gp.sum(), rtol=1e-12)
    np.testing.assert_allclose(rep["sums"][2], gv.sum(), rtol=1e-12, atol=1e-12)


def test_block_ranges_cover_the_batch_in_order():
    """HostPipeline's column blocks: equal or relative sizes, aligned, exact cover, no empty block."""
    from arboris_b200.shard import block_ranges
    for total in (0, 1, 31, 32, 1000, 4096, 262144):
        for chunks in (1, 3, 7, 8, 64, (1, 2, 1), (1, 2, 2, 2, 1), [1, 3, 4, 4, 3, 1], (0, 1, 0)):
            r = block_ranges(total, chunks)
            assert all(w1 > w0 for w0, w1 in r)
            if total == 0:
                assert r == []
                continue
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            if isinstance(chunks, (list, tuple)):
                assert all(w0 % 32 == 0 for w0, _ in r)
    assert block_ranges(262144, (1, 2, 2, 2, 1)) == [(0, 32768), (32768, 98304), (98304, 163840),
                                                     (163840, 229376), (229376, 262144)]
    assert block_ranges(1000, 7) == [(0, 143), (143, 286), (286, 429), (429, 572), (572, 715), (715, 858), (858, 1000)]
    import pytest
    with pytest.raises(ValueError):
        block_ranges(10, ())
    with pytest.raises(ValueError):
        block_ranges(10, (1, -1))
    with pytest.raises(ValueError):
        block_ranges(-1, 2)
