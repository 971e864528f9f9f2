"""Worker for tests/test_multirank_cpu.py: run under torch.distributed.run with the gloo backend."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lewton_b200.sharding import gather_streams, owner_of, scatter_streams, stream_range  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_streams = int(sys.argv[1])
    lo, hi = stream_range(n_streams, world, rank)
    mine = torch.zeros(n_streams, dtype=torch.int64)
    mine[lo:hi] = 1
    dist.all_reduce(mine)                                   # every stream owned exactly once
    assert bool((mine == 1).all()), mine
    for s in range(lo, hi):
        assert owner_of(s, n_streams, world) == rank
    # the bench's timing protocol: value = units of all ranks / max over ranks of the device time
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    units = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(units)
    assert units.item() == n_streams
    # batch scatter / PCM gather (grouped send / receive): every rank gets exactly its streams' rows and the
    # root gets every rank's results back in stream order
    full = torch.arange(n_streams * 3, dtype=torch.float32).reshape(n_streams, 3) if rank == 0 else None
    local = torch.full((hi - lo, 3), -1.0)
    scatter_streams(full, local, n_streams, root=0)
    assert torch.equal(local, torch.arange(n_streams * 3, dtype=torch.float32).reshape(n_streams, 3)[lo:hi])
    back = torch.zeros(n_streams, 3) if rank == 0 else None
    gather_streams(local * 2 + rank, back, n_streams, root=0)
    if rank == 0:
        for r in range(world):
            a, b = stream_range(n_streams, world, r)
            assert torch.equal(back[a:b], full[a:b] * 2 + r)
    dist.barrier()
    if rank == 0:
        print(f"RANKS_OK {world} {n_streams}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
