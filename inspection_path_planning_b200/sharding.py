"""Host-side partition of a population over ranks: contiguous shards of ceil(pop / world) plans -- the same split
csrc/ipp_api.cu (ipp_evaluate_population_sharded) applies before the NCCL all-gather of the fitness rows."""
import numpy as np


def shard_bounds(pop, world, rank):
    """-> (rows per rank slot, first plan, one past the last plan) of `rank`."""
    per = (pop + world - 1) // world if pop > 0 else 0
    lo = min(pop, per * rank)
    hi = min(pop, lo + per)
    return per, lo, hi


def shard_csr(ids, offsets, lo, hi):
    """CSR slice of plans [lo, hi) with offsets rebased to 0."""
    g0, g1 = int(offsets[lo]), int(offsets[hi])
    return np.ascontiguousarray(ids[g0:g1]), np.ascontiguousarray(offsets[lo:hi + 1] - g0)
