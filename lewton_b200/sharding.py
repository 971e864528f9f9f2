"""Stream sharding for multi-GPU runs.

The synthesis path has no cross-stream dependency: the only couplings are packet k <-> k+1 of one
stream (overlap-add) and the channels of one packet (inverse coupling), so streams are partitioned
by contiguous ranges over the ranks and the COMPUTE needs no collective.  When a batch originates and
terminates on one rank (a front end that entropy-decodes on one socket, a sink that wants all PCM in one
place), the batch is scattered and the PCM gathered over NVLink with grouped NCCL send / receive
(`scatter_streams` / `gather_streams`: torch.distributed's batch_isend_irecv = ncclGroupStart ..
ncclSend / ncclRecv .. ncclGroupEnd) -- the only NCCL traffic of the path, timed separately by bench.py."""


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


def _p2p(ops):
    import torch.distributed as dist
    if not ops:
        return
    if dist.get_backend() == "nccl":
        for w in dist.batch_isend_irecv(ops):      # one ncclGroupStart/End around all sends and receives
            w.wait()
    else:                                          # gloo (CPU tests): the same transfers one by one
        for w in [op.op(op.tensor, op.peer) for op in ops]:
            w.wait()


def scatter_streams(root_tensor, local_tensor, n_streams, root=0):
    """Rows [lo, hi) = stream_range(n_streams, world, rank) of `root_tensor` (first dimension = streams, present on
    `root` only) land in `local_tensor` on every rank.  Grouped point-to-point: the root sends each peer its slice."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == root:
        ops = []
        for r in range(world):
            lo, hi = stream_range(n_streams, world, r)
            if r == root:
                local_tensor.copy_(root_tensor[lo:hi])
            elif hi > lo:
                ops.append(dist.P2POp(dist.isend, root_tensor[lo:hi], r))
        _p2p(ops)
    else:
        lo, hi = stream_range(n_streams, world, rank)
        if hi > lo:
            _p2p([dist.P2POp(dist.irecv, local_tensor, root)])


def gather_streams(local_tensor, root_tensor, n_streams, root=0):
    """Inverse of scatter_streams: every rank's rows return to their place in `root_tensor` on `root`."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == root:
        ops = []
        for r in range(world):
            lo, hi = stream_range(n_streams, world, r)
            if r == root:
                root_tensor[lo:hi].copy_(local_tensor)
            elif hi > lo:
                ops.append(dist.P2POp(dist.irecv, root_tensor[lo:hi], r))
        _p2p(ops)
    else:
        lo, hi = stream_range(n_streams, world, rank)
        if hi > lo:
            _p2p([dist.P2POp(dist.isend, local_tensor, root)])
