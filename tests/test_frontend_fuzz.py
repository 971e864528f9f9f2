"""Mutation fuzzing of the host front half (tests/fuzz/fe_fuzz.cpp) under ASan + UBSan: hostile
headers, audio packets and Ogg bytes must produce status codes, never memory errors."""
import os
import struct
import subprocess

import numpy as np

import vorbis_packer as vp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frontend_survives_mutated_streams(tmp_path):
    exe = tmp_path / "fe_fuzz"
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                    "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "fuzz", "fe_fuzz.cpp"),
                    "-o", str(exe), "-lpthread"], check=True)
    for seed, channels, floor0 in ((1, 2, False), (2, 1, True), (3, 3, False)):
        rng = np.random.default_rng(seed)
        spec = vp.StreamSpec(rng, channels=channels, floor0=floor0, cascade_p=0.3)
        packets = []
        for k in range(6):
            mode = int(rng.integers(0, len(spec.modes)))
            packets.append(spec.audio_packet(mode, 1, 1)[0])
        ogg = vp.ogg_stream(9, [spec.ident_packet(), spec.comment_packet(), spec.setup_packet()], packets, [1000, 2000], 3)
        corpus = tmp_path / f"corpus{seed}.bin"
        with open(corpus, "wb") as f:
            for item in [spec.ident_packet(), spec.comment_packet(), spec.setup_packet(), ogg] + packets:
                f.write(struct.pack("<I", len(item)) + item)
        env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
        r = subprocess.run([str(exe), str(corpus), "1500"], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, (seed, r.stdout[-2000:], r.stderr[-4000:])
        fields = dict(zip(r.stdout.split()[::2], r.stdout.split()[1::2]))
        assert int(fields["parsed_ok"]) > 50 and int(fields["decoded_ok"]) > 200 and int(fields["ogg_packets"]) > 200, r.stdout
        # the batcher's thread pool over mutated packets (the stubbed synthesis call then reports "no device")
        r = subprocess.run([str(exe), str(corpus), "6", "batch"], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0 and "ok" in r.stdout, (seed, r.stdout[-500:], r.stderr[-3000:])
