#!/usr/bin/env python3
"""Whole-bitstream decode at scale (BASELINE configs 1 / 3 shape): S logical Vorbis streams, P audio
packets each, real entropy-coded packets (tests/vorbis_packer.py, D distinct bitstreams replicated),
through lwf_batcher: host entropy decode on T threads -> pinned arenas -> ONE batched CUDA synthesis
call (residue entry, host memory, H2D/D2H inside).  One JSON line per thread count: where the time
goes (entropy decode vs synthesis) and the end-to-end rate."""
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
    import vorbis_packer as vp
    from bench import host_threads
    from lewton_b200 import frontend as fe

    rng = np.random.default_rng(77)
    C, P, D = 2, int(os.environ.get("SB_PACKETS", 24)), 6
    S = int(os.environ.get("SB_STREAMS", 1024))
    spec = vp.StreamSpec(rng, channels=C, residue_types=[1, 2], cascade_p=float(os.environ.get("SB_CASCADE_P", 0.12)))
    hdr = fe.Headers(spec.ident_packet(), spec.comment_packet(), spec.setup_packet())
    distinct = []
    t0 = time.perf_counter()
    for d in range(D):
        bf = rng.random(P) >= 0.1
        pk = []
        for i in range(P):
            long_modes = [m for m, (b, _) in enumerate(spec.modes) if b]
            short_modes = [m for m, (b, _) in enumerate(spec.modes) if not b]
            mode = int(rng.choice(long_modes if bf[i] else short_modes))
            prev = int(bf[i - 1]) if i else 1
            nxt = int(bf[i + 1]) if i + 1 < P else 1
            pk.append(spec.audio_packet(mode, prev, nxt, p_unused=0.02)[0])
        distinct.append(pk)
    pack_s = time.perf_counter() - t0
    ctx = L.Context(0)
    su = hdr.make_setup(ctx)
    stride = P * 1024
    pcm_host = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_float * (S * C * stride)).from_address(
        L._cabi.lib().lwb_host_alloc(S * C * stride * 4)))
    packet_bytes = sum(len(p) for d in range(D) for p in distinct[d]) / D
    # SB_ENTRY=vq: the residue crosses the boundary as VQ records (LWB_ENTRY_VQ) instead of dense vectors
    entry_name = os.environ.get("SB_ENTRY", "residue")
    entry = L._cabi.ENTRY_VQ if entry_name == "vq" else L._cabi.ENTRY_RESIDUE
    for threads in sorted({1, 4, host_threads()}):
        pwrs = [L.PreviousWindowRight(su) for _ in range(S)]
        jobs = [(pwrs[s], distinct[s % D]) for s in range(S)]
        bt = fe.StreamBatcher(ctx, hdr, threads=threads, entry=entry)
        res = bt.decode(jobs, pcm_host, stride)          # first call: streams start empty
        samples = sum(r[0] for r in res) * C
        assert all(r[2] == 0 for r in res)
        ent, syn, wall = [], [], []
        for _ in range(3):
            for p in pwrs:
                p.reset()
            t0 = time.perf_counter()
            bt.run(pcm_host)
            wall.append(time.perf_counter() - t0)
            ent.append(bt.entropy_seconds)
            syn.append(bt.synthesis_seconds)
        w, e, s_ = min(wall), min(ent), min(syn)
        print(json.dumps({"entry": entry_name, "streams": S, "packets_per_stream": P, "channels": C, "host_threads": threads,
                          "avg_packet_bytes": packet_bytes, "channel_samples": samples,
                          "wall_ms": w * 1e3, "entropy_decode_ms": e * 1e3, "synthesis_call_ms": s_ * 1e3,
                          "e2e_msamples_per_s": samples / w / 1e6,
                          "entropy_decode_msamples_per_s": samples / e / 1e6,
                          "entropy_decode_msamples_per_s_per_thread": samples / e / 1e6 / threads,
                          "synthesis_msamples_per_s": samples / s_ / 1e6,
                          "compressed_mbit_per_s_in": S * P * packet_bytes * 8 / w / 1e6,
                          "packer_seconds": pack_s}), flush=True)
        bt.close()
        for p in pwrs:
            p.close()
    ctx.close()


if __name__ == "__main__":
    main()
