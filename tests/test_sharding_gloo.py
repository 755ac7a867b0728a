"""The N>1 path on CPU: two processes over the gloo backend run the same host-side sharding / reduction /
ordered-gather code that bench.py and the drop-in binary use (16gt_b200/sharding.py).  The data path has no
collective; what is exercised is the batch-aligned partition, the max-over-ranks timing reduce and the final
ordered gather, with the oracle standing in for the per-shard caller (there is no GPU here)."""
import importlib
import os
import socket
import sys

import pytest

import gt16_testlib as T

sharding = importlib.import_module("16gt_b200.sharding")


def test_cuts_are_batch_aligned_and_tile_the_range():
    B = sharding.BATCH_SITES
    for total in (1, B - 1, B, B + 1, 64444167, 3100000000):
        for n in (1, 2, 3, 4, 8):
            cuts = sharding.shard_cuts(total, n)
            assert cuts[0] == 0 and cuts[-1] == total and cuts == sorted(cuts)
            assert all(c % B == 0 for c in cuts[1:-1])
            assert len(cuts) - 1 <= max(1, (total + B - 1) // B)
    assert sharding.my_shard(100, 3, 8) == (100, 100)       # more ranks than batches: empty shard


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = 3 * sharding.BATCH_SITES + 12345
        begin, end = sharding.my_shard(total, rank, world)
        # stand-in for the per-shard call list: the shard's first and last site index
        calls = [begin, end - 1] if end > begin else []
        gathered = sharding.gather_in_rank_order(calls, dist)
        t = sharding.max_over_ranks(1.0 + rank, dist)
        q.put((rank, begin, end, gathered, t))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = 3 * sharding.BATCH_SITES + 12345
    (_, b0, e0, g0, t0), (_, b1, e1, g1, t1) = res
    assert b0 == 0 and e0 == b1 and e1 == total and e0 % sharding.BATCH_SITES == 0
    assert g0 == g1 == [[0, e0 - 1], [b1, total - 1]]         # same ordered gather on every rank
    assert t0 == t1 == 2.0                                    # max over ranks
    flat = [x for part in g0 for x in part]
    assert flat == sorted(flat)
