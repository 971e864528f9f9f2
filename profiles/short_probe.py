#!/usr/bin/env python3
"""Driver for profiling k_short under ncu: mono chains of 256-point blocks (the n = 256 sweep shape)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import lewton_b200 as L
    from lewton_b200 import _cabi as cabi

    chains_n, pk = int(os.environ.get("CHAINS", 131072)), int(os.environ.get("PK", 8))
    ctx = L.Context(0)
    su = L.Setup(ctx, 1, 8, 8, [L.FloorTypeOne(1, [0, 128])], [L.Mapping(1)], [L.ModeInfo(True)])
    total = chains_n * pk * 128
    spec = torch.randn(total, device="cuda") * 1e-2
    pcm = torch.empty(total, device="cuda")
    pw = [L.PreviousWindowRight(su) for _ in range(chains_n)]
    modes = np.zeros(pk, np.uint8)
    chains = [L.ChainSpec(pw[i], modes, coeff_offset=i * pk * 128, out_offset=i * pk * 128, out_stride=pk * 128) for i in range(chains_n)]
    batch = L.Batch(ctx, chains, cabi.ENTRY_SPECTRUM, cabi.MEM_DEVICE, spec.data_ptr(), pcm.data_ptr(), cabi.OUT_F32_PLANAR)
    for _ in range(int(os.environ.get("REPS", 4))):
        batch.run()
    ctx.synchronize()
    print("launches", ctx.launch_count)


if __name__ == "__main__":
    main()
