#!/usr/bin/env python3
"""Throughput of the other shapes SURVEY.md section 8(d) asks to be reported beside the headline
(device-resident, one B200, CUDA events on the library's stream, one JSON line per case):

  long_f32        S x P stereo long packets, spectrum entry, f32 planar           8 B / sample (the headline)
  long_i16        same, i16 planar output                                          6 B / sample
  long_residue    same, residue entry: coupling (0,1) + floor-1 + multiply         16 B / sample (k_prologue + k_long)
  streaming_p1    one packet per stream per call (state round-trips HBM)           16 B / sample
  config3_6ch     BASELINE.json configs[2] shape: 6 channels, Bernoulli(0.25) short blocks, coupling chain
                  (0,1),(2,3),(0,4), floor-1, residue entry; segmented path (k_prologue + k_long, k_chain)
Inputs are synthetic (N(0,1)*1e-2 residues, random valid floor-1 posts drawn from a pool of 64 rows)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch

    import lewton_b200 as L
    from lewton_b200 import _cabi as cabi
    from helpers import make_setup, random_floor1_y

    peak = 6650.0
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        peak = float(json.load(open(pth))["hbm_gbs"])
    ctx = L.Context(0)
    stream = torch.cuda.ExternalStream(ctx.cuda_stream)
    rng = np.random.default_rng(1234)
    floors = [(1, [0, 1024] + [int(v) for v in rng.permutation(np.arange(1, 1024))[:30]])]

    def timed(batch, reps=10, warm=3):
        l0 = ctx.launch_count
        batch.run()
        launches = ctx.launch_count - l0
        for _ in range(warm - 1):
            batch.run()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(reps):
            batch.run()
        host_ms = (time.perf_counter() - t0) / reps * 1e3
        e1.record(stream)
        ctx.synchronize()
        return e0.elapsed_time(e1) / reps, host_ms, launches

    def report(case, samples, ms, host_ms, launches, bytes_per_sample, note):
        print(json.dumps({"case": case, "ms": ms, "host_enqueue_ms": host_ms, "launches": launches,
                          "msamples_per_s": samples / ms / 1e3, "algorithmic_bytes_per_sample": bytes_per_sample,
                          "achieved_gbs": samples * bytes_per_sample / ms / 1e6,
                          "frac_of_hbm_peak": samples * bytes_per_sample / ms / 1e6 / peak, "note": note}), flush=True)

    def floor_rows(rows, nposts, mult):
        pool = np.zeros((64, cabi.MAX_POSTS), np.uint32)
        for i in range(64):
            pool[i, :nposts] = random_floor1_y(rng, mult, nposts)
        return pool[rng.integers(0, 64, rows)]

    # ---- uniform long batches ------------------------------------------------------------------
    S, P, C = 4096, 16, 2
    su = make_setup(ctx, C, 8, 11, mappings=[{"coupling": [(0, 1)], "floor_of_channel": [0, 0]}], floors=floors)
    spec = torch.randn(S * P * C * 1024, device="cuda") * 1e-2
    modes = np.ones(P, np.uint8)
    for case, fmt, entry, bps in (("long_f32", cabi.OUT_F32_PLANAR, cabi.ENTRY_SPECTRUM, 8),
                                  ("long_i16", cabi.OUT_I16_PLANAR, cabi.ENTRY_SPECTRUM, 6),
                                  ("long_residue", cabi.OUT_F32_PLANAR, cabi.ENTRY_RESIDUE, 16)):
        pcm = torch.empty(S * C * P * 1024, device="cuda", dtype=torch.float32 if fmt == cabi.OUT_F32_PLANAR else torch.int16)
        pw = [L.PreviousWindowRight(su) for _ in range(S)]
        chains = [L.ChainSpec(pw[s], modes, coeff_offset=s * P * C * 1024, packet_index=s * P, out_offset=s * C * P * 1024,
                              out_stride=P * 1024) for s in range(S)]
        kw = {}
        if entry == cabi.ENTRY_RESIDUE:
            kw = dict(floor_kind=np.full(S * P * C, cabi.FLOOR_ONE, np.uint8), floor1_y=floor_rows(S * P * C, len(floors[0][1]), 1))
        batch = L.Batch(ctx, chains, entry, cabi.MEM_DEVICE, spec.data_ptr(), pcm.data_ptr(), fmt, **kw)
        ms, host_ms, launches = timed(batch)
        report(case, S * P * C * 1024, ms, host_ms, launches, bps,
               f"{S} stereo streams x {P} long packets" + (", per-packet floor posts uploaded every step (host arrays)" if kw else ""))
        batch.close()
        for p in pw:
            p.close()
        del pcm
    del spec

    # ---- one packet per stream per call ----------------------------------------------------------
    S1 = 65536
    spec = torch.randn(S1 * C * 1024, device="cuda") * 1e-2
    pcm = torch.empty(S1 * C * 1024, device="cuda")
    pw = [L.PreviousWindowRight(su) for _ in range(S1)]
    one = np.ones(1, np.uint8)
    chains = [L.ChainSpec(pw[s], one, coeff_offset=s * C * 1024, out_offset=s * C * 1024, out_stride=1024) for s in range(S1)]
    batch = L.Batch(ctx, chains, cabi.ENTRY_SPECTRUM, cabi.MEM_DEVICE, spec.data_ptr(), pcm.data_ptr(), cabi.OUT_F32_PLANAR)
    ms, host_ms, launches = timed(batch)
    report("streaming_p1", S1 * C * 1024, ms, host_ms, launches, 16, f"{S1} stereo streams x 1 long packet per call")
    batch.close()
    for p in pw:
        p.close()
    del spec, pcm
    su.close()

    # ---- config 3: 6 channels, mixed blocks, coupling chain, floor-1, residue entry ----------------
    S3, P3, C3 = 1024, 64, 6
    su3 = make_setup(ctx, C3, 8, 11, mappings=[{"coupling": [(0, 1), (2, 3), (0, 4)], "floor_of_channel": [0] * C3}], floors=floors)
    seqs, coeff_off, offs, total_rows = [], 0, [], 0
    for s in range(S3):
        bf = (rng.random(P3) >= 0.25).astype(np.uint8)
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
    batch = L.Batch(ctx, chains, cabi.ENTRY_RESIDUE, cabi.MEM_DEVICE, res.data_ptr(), pcm.data_ptr(), cabi.OUT_F32_PLANAR,
                    floor_kind=np.full(rows, cabi.FLOOR_ONE, np.uint8), floor1_y=floor_rows(rows, len(floors[0][1]), 1))
    ms, host_ms, launches = timed(batch, reps=5)
    batch.collect()
    samples = sum(ch.n_samples for ch in chains) * C3
    report("config3_6ch", samples, ms, host_ms, launches, 8,
           f"{S3} streams x {P3} packets x 6 ch, 25 % short blocks, residue entry; host plans every step (no capture for residue entry)")
    batch.close()
    ctx.close()


if __name__ == "__main__":
    main()
