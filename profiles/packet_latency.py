#!/usr/bin/env python3
"""Per-call latency of the literal drop-in, lwb_decode_packet (one stream, host buffers, synchronous):
median / p10 / p90 microseconds per packet over a stereo 256/2048 stream, residue entry (coupling + floor-1),
f32 planar and i16 interleaved output; and of lwb_decode_spectrum (entry at audio.rs:1041).  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import lewton_b200 as L
    from helpers import make_setup, mode_sequence, random_floor1_y

    ctx = L.Context(0)
    rng = np.random.default_rng(7)
    floors = [(1, [0, 1024] + [int(v) for v in rng.permutation(np.arange(1, 1024))[:30]])]
    su = make_setup(ctx, 2, 8, 11, mappings=[{"coupling": [(0, 1)], "floor_of_channel": [0, 0]}], floors=floors)
    n = 400
    bf, prev, nxt = mode_sequence(rng, n, p_short=0.1)
    out = {}
    for name, kw in (("residue_f32_planar", dict(sample="f32", interleaved=False)), ("residue_i16_interleaved", dict(sample="i16", interleaved=True))):
        pwr = L.PreviousWindowRight(su)
        ts = {0: [], 1: []}
        for i in range(n):
            n2 = 1024 if bf[i] else 128
            res = (rng.standard_normal((2, n2)) * 1e-2).astype(np.float32)
            y = random_floor1_y(rng, 1, len(floors[0][1]))
            pk = L.DecodedPacket(int(bf[i]), res, [y, y], prev[i], nxt[i])
            pk.pack()
            t0 = time.perf_counter()
            L.read_audio_packet_generic(su, pk, pwr, **kw)
            ts[int(bf[i])].append((time.perf_counter() - t0) * 1e6)
        out[name] = {("long" if k else "short"): {"median_us": float(np.median(v[5:])), "p10_us": float(np.percentile(v[5:], 10)),
                                                   "p90_us": float(np.percentile(v[5:], 90)), "packets": len(v)} for k, v in ts.items() if len(v) > 10}
        pwr.close()
    pwr = L.PreviousWindowRight(su)
    ts = []
    for i in range(200):
        spec = (rng.standard_normal((2, 1024)) * 1e-2).astype(np.float32)
        t0 = time.perf_counter()
        L.decode_spectrum(su, 1, spec, pwr)
        ts.append((time.perf_counter() - t0) * 1e6)
    out["spectrum_f32_planar"] = {"long": {"median_us": float(np.median(ts[5:])), "p10_us": float(np.percentile(ts[5:], 10)),
                                           "p90_us": float(np.percentile(ts[5:], 90)), "packets": len(ts)}}
    out["note"] = ("wall clock around the Python call (ctypes marshalling of one packet included, a few us); one synchronous "
                   "lwb_decode_packet = H2D of the residue + floor rows, front stages + synthesis kernels, D2H, stream sync")
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
