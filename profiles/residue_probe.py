#!/usr/bin/env python3
"""Small driver for profiling the residue entry's front stages under ncu: S stereo streams x P long packets,
device-resident residues and floor arrays, prepared plan, a few replays."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch

    import lewton_b200 as L
    from lewton_b200 import _cabi as cabi
    from helpers import make_setup, random_floor1_y

    S, P, C = int(os.environ.get("S", 4096)), 16, 2
    BS = int(os.environ.get("BS", 11))                  # 11: k_long; 10 / 9: k_mid (LWB_NO_MID=1: the chain kernel)
    N2 = 1 << (BS - 1)
    ctx = L.Context(0)
    rng = np.random.default_rng(1234)
    floors = [(1, [0, N2] + [int(v) for v in rng.permutation(np.arange(1, N2))[:30]])]
    su = make_setup(ctx, C, 8 if BS == 11 else BS, BS, mappings=[{"coupling": [(0, 1)], "floor_of_channel": [0, 0]}], floors=floors)
    res = torch.randn(S * P * C * N2, device="cuda") * 1e-2
    pcm = torch.empty(S * C * P * N2, device="cuda")
    pool = np.zeros((64, cabi.MAX_POSTS), np.uint32)
    for i in range(64):
        pool[i, :len(floors[0][1])] = random_floor1_y(rng, 1, len(floors[0][1]))
    ys = torch.from_numpy(pool[rng.integers(0, 64, S * P * C)].view(np.int32)).cuda()
    kinds = torch.full((S * P * C,), cabi.FLOOR_ONE, dtype=torch.uint8, device="cuda")
    pw = [L.PreviousWindowRight(su) for _ in range(S)]
    modes = np.ones(P, np.uint8)
    chains = [L.ChainSpec(pw[s], modes, coeff_offset=s * P * C * N2, packet_index=s * P, out_offset=s * C * P * N2,
                          out_stride=P * N2) for s in range(S)]
    batch = L.Batch(ctx, chains, cabi.ENTRY_RESIDUE, cabi.MEM_DEVICE, res.data_ptr(), pcm.data_ptr(), cabi.OUT_F32_PLANAR,
                    floor_kind=kinds.data_ptr(), floor1_y=ys.data_ptr(), floor_memory=cabi.MEM_DEVICE)
    for _ in range(int(os.environ.get("REPS", 4))):
        batch.run()
    ctx.synchronize()
    stream = torch.cuda.ExternalStream(ctx.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        batch.run()
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print({"bs": BS, "ms": ms, "msamples_per_s": S * P * C * N2 / ms / 1e3})


if __name__ == "__main__":
    main()
