"""Region sharding of the calling path across GPUs: contiguous ranges of the processed-site
sequence, cut on multiples of the reference's batch size (8,388,608 sites, VariantCaller.cpp:526-528) so that
the per-batch window semantics are the same whatever the number of shards.  No collective on the data path;
the only communication is the final ordered gather of the (small) call lists and, in bench.py, a max-reduce
of the measured time."""
BATCH_SITES = 8388608


def shard_cuts(total_sites, n_shards):
    """cuts[d] .. cuts[d+1] is shard d's range of processed-site indices; every interior cut is batch-aligned"""
    nb = (total_sites + BATCH_SITES - 1) // BATCH_SITES
    n = max(1, min(n_shards, nb if nb else 1))
    cuts = [min(total_sites, (nb * d // n) * BATCH_SITES) for d in range(n)] + [total_sites]
    return cuts


def my_shard(total_sites, rank, world):
    """(begin, end) of `rank`; ranks beyond the number of batches get an empty range"""
    cuts = shard_cuts(total_sites, world)
    if rank >= len(cuts) - 1:
        return total_sites, total_sites
    return cuts[rank], cuts[rank + 1]


def max_over_ranks(value, dist=None, device=None):
    """max-reduce of a python float over the process group (identity when not distributed)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_in_rank_order(obj, dist=None):
    """the "final ordered gather": every rank's call list, concatenated in shard order on every rank"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
