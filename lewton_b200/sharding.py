"""Stream sharding for multi-GPU runs.

The synthesis path has no cross-stream dependency: the only couplings are packet k <-> k+1 of one
stream (overlap-add) and the channels of one packet (inverse coupling), so streams are partitioned
by contiguous ranges over the ranks and NO data-path collective exists.  torch.distributed is used
by callers only to agree on timings (max over ranks)."""


def stream_range(n_streams, world_size, rank):
    """Contiguous, balanced [lo, hi) of the streams rank `rank` owns (sizes differ by at most 1)."""
    if not (0 <= rank < world_size) or n_streams < 0:
        raise ValueError("bad rank / world size / stream count")
    base, extra = divmod(n_streams, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of(stream, n_streams, world_size):
    """Rank that owns `stream` under stream_range."""
    base, extra = divmod(n_streams, world_size)
    cut = extra * (base + 1)
    if stream < cut:
        return stream // (base + 1)
    return extra + (stream - cut) // base if base else world_size - 1
