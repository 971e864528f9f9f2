#!/usr/bin/env python3
"""Driver for profiling the BASELINE configs[2] shape under ncu: 6 channels, 25 % short blocks, coupling chain, residue entry,
device floor arrays, prepared plan (the same batch configs_bench.py times)."""
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

    ctx = L.Context(0)
    rng = np.random.default_rng(1234)
    floors = [(1, [0, 1024] + [int(v) for v in rng.permutation(np.arange(1, 1024))[:30]])]
    S3, P3, C3 = 1024, 64, 6
    su3 = make_setup(ctx, C3, 8, 11, mappings=[{"coupling": [(0, 1), (2, 3), (0, 4)], "floor_of_channel": [0] * C3}], floors=floors)
    seqs, coeff_off, offs = [], 0, []
    for s in range(S3):
        bf = (rng.random(P3) >= float(os.environ.get("P_SHORT", 0.25))).astype(np.uint8)
        bf[0] = bf[-1] = 1
        prev, nxt = np.ones(P3, np.uint8), np.ones(P3, np.uint8)
        for i in range(P3):
            if bf[i]:
                prev[i] = bf[i - 1] if i else 1
                nxt[i] = bf[i + 1] if i + 1 < P3 else 1
        seqs.append((bf, prev, nxt))
        offs.append(coeff_off)
        coeff_off += int(sum(C3 * (1024 if b else 128) for b in bf))
    res = torch.randn(coeff_off, device="cuda") * 1e-2
    pcm = torch.empty(S3 * C3 * P3 * 1024, device="cuda")
    pw = [L.PreviousWindowRight(su3) for _ in range(S3)]
    chains = [L.ChainSpec(pw[s], seqs[s][0], seqs[s][1], seqs[s][2], coeff_offset=offs[s], packet_index=s * P3,
                          out_offset=s * C3 * P3 * 1024, out_stride=P3 * 1024) for s in range(S3)]
    rows = S3 * P3 * C3
    pool = np.zeros((64, cabi.MAX_POSTS), np.uint32)
    for i in range(64):
        pool[i, :len(floors[0][1])] = random_floor1_y(rng, 1, len(floors[0][1]))
    d_ys = torch.from_numpy(pool[rng.integers(0, 64, rows)].view(np.int32)).cuda()
    d_kinds = torch.full((rows,), cabi.FLOOR_ONE, dtype=torch.uint8, device="cuda")
    batch = L.Batch(ctx, chains, cabi.ENTRY_RESIDUE, cabi.MEM_DEVICE, res.data_ptr(), pcm.data_ptr(), cabi.OUT_F32_PLANAR,
                    floor_kind=d_kinds.data_ptr(), floor1_y=d_ys.data_ptr(), floor_memory=cabi.MEM_DEVICE)
    for _ in range(int(os.environ.get("REPS", 4))):
        batch.run()
    ctx.synchronize()
    print("launches", ctx.launch_count)


if __name__ == "__main__":
    main()
