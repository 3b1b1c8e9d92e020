"""Host-side sharding helpers for the multi-GPU path (one process per GPU).

A batched MSM / NTT shards by batch index with no collective; a single large MSM shards by POINT RANGE: each rank computes
the partial sum over its contiguous range and the partial results (96 B each) are exchanged with one all-gather and summed
with the ec_sum kernel (include/icicle_b200.h b200_ec_sum).  The reference has no multi-device reduction
(docs/docs/start/architecture/multi-device.md:32-36): one host thread per device, the user shards the work."""


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of n items for `rank` of `world`; sizes differ by at most one, earlier ranks take the extras."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch, rank, world):
    """Batch indices [lo, hi) handled by `rank` when a batch of independent MSMs / NTTs is split across GPUs."""
    return shard_range(batch, rank, world)
