#!/usr/bin/env python3
"""Host entropy decode alone (lwf_packet_decode, one thread, no GPU): microseconds per stereo long packet
of the synthetic ~300-byte stream shape used by stream_bench.py.  One JSON line."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import vorbis_packer as vp
    from lewton_b200 import _cabi as cabi
    from lewton_b200 import frontend as fe

    rng = np.random.default_rng(77)
    spec = vp.StreamSpec(rng, channels=2, residue_types=[1, 2], cascade_p=0.12)
    hdr = fe.Headers(spec.ident_packet(), spec.comment_packet(), spec.setup_packet())
    long_modes = [m for m, (b, _) in enumerate(spec.modes) if b]
    pk = [spec.audio_packet(int(rng.choice(long_modes)), 1, 1, p_unused=0.02)[0] for _ in range(40)]
    kinds, ys = np.zeros(2, np.uint8), np.zeros((2, 65), np.uint32)
    dense, res = np.zeros((2, 1024), np.float32), np.zeros((2, 1024), np.float32)
    dp = fe._DecodedPacket()
    dp.floor_kind, dp.floor1_y = kinds.ctypes.data_as(cabi.u8p), ys.ctypes.data_as(cabi.u32p)
    dp.dense_floor, dp.residue = dense.ctypes.data_as(cabi.fp), res.ctypes.data_as(cabi.fp)
    L = fe.lib()
    for p in pk:
        assert L.lwf_packet_decode(hdr._h, p, len(p), C.byref(dp)) == 0
    t0 = time.perf_counter()
    n = 0
    for _ in range(300):
        for p in pk:
            L.lwf_packet_decode(hdr._h, p, len(p), C.byref(dp))
            n += 1
    dt = time.perf_counter() - t0
    # the same inside one C call (no ctypes overhead), dense and VQ-record output
    arr = (C.c_char_p * len(pk))(*pk)
    lens = (C.c_size_t * len(pk))(*[len(p) for p in pk])
    L.lwf_debug_decode_loop.restype = C.c_double
    L.lwf_debug_decode_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    inner = {}
    for name, vq in (("dense", 0), ("vq", 1)):
        L.lwf_debug_decode_loop(hdr._h, arr, lens, len(pk), 50, vq)
        sec = min(L.lwf_debug_decode_loop(hdr._h, arr, lens, len(pk), 500, vq) for _ in range(5))
        inner[name + "_us_per_packet"] = sec / (500 * len(pk)) * 1e6
        inner[name + "_msamples_per_s_per_thread"] = 500 * len(pk) * 2048 / sec / 1e6
    print(json.dumps(inner))
    print(json.dumps({"us_per_packet": dt / n * 1e6, "packets_per_s": n / dt, "msamples_per_s_per_thread": n * 2048 / dt / 1e6,
                      "avg_packet_bytes": sum(map(len, pk)) / len(pk), "note": "includes ~1 us of ctypes call overhead per packet"}))


if __name__ == "__main__":
    main()
