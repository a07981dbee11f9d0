"""Block sharding across ranks (one process per GPU) and the dense merge it relies on.

Blocks are independent units (table_query.go:110-111, aggregate.go:326-332) and results
merge by one associative combine (CombineResults, aggregate.go:414-467).  Ranks take
contiguous block ranges; because every rank numbers group keys identically the partial
results are element-wise summable — on GPUs by sg_query_allreduce (NCCL), in the CPU test by
a gloo all_reduce over the same dense layout.
"""


def shard_range(nblocks, rank, world):
    """[first, first+count) of the blocks rank scans: contiguous, sizes differ by at most one block."""
    base, extra = divmod(nblocks, world)  # the first `extra` ranks take one block more
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def dense_layout(all_keys, naggs, nvalues):
    """Shared numbering of group keys -> rows of a dense partial-result matrix.
    Row layout: [Count, then per aggregation: hist Count, exact sum, nvalues bucket counters]."""
    keys = sorted(set(all_keys))
    return {k: i for i, k in enumerate(keys)}, 1 + naggs * (2 + nvalues)
