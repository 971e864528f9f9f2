"""Host front half (include/lewton_frontend.h) on the CPU: known-answer tests taken from the
reference's own unit tests (bitpacking.rs, huffman_tree.rs, header.rs) and decode-what-was-packed
tests against the synthetic packer (tests/vorbis_packer.py), including truncated packets and Ogg
framing.  None of this needs a GPU: the library loads and the lwf_* entry points are pure host code."""
import ctypes as C

import numpy as np
import pytest

import vorbis_packer as vp
from lewton_b200 import _cabi as cabi
from lewton_b200 import frontend as fe


def read_bits(data, widths):
    out = (C.c_uint64 * len(widths))()
    n = fe.lib().lwf_debug_read_bits(bytes(data), len(data), bytes(widths), len(widths), out)
    return [int(out[i]) for i in range(n)]


def test_symbols_exported_and_struct_layouts(tmp_path):
    """Every lwf_* function include/lewton_frontend.h declares is exported, and the ctypes mirrors of
    its structs have the C compiler's sizes."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "lewton_frontend.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lwf_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(fe.SYMBOLS), declared ^ set(fe.SYMBOLS)
    nm = subprocess.run(["nm", "-D", "--defined-only", cabi.SO_PATH], capture_output=True, text=True, check=True).stdout
    assert declared <= set(re.findall(r" T (lwf_[a-z0-9_]+)", nm))
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "lewton_frontend.h"\nint main(void){printf("%zu %zu %zu\\n",'
                   'sizeof(lwf_info), sizeof(lwf_decoded_packet), sizeof(lwf_ogg_packet));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(fe.Info), C.sizeof(fe._DecodedPacket), C.sizeof(fe._OggPacket)]


def test_bitreader_spec_vectors():
    """bitpacking.rs:488-600 (vectors of Vorbis I spec 2.1.6 and the regression cases)."""
    arr = [0b11111100, 0b01001000, 0b11001110, 0b00000110]
    assert read_bits(arr, [4, 3, 7, 13]) == [12, 7, 17, 6969]
    assert read_bits([93, 92], [10]) == [93]
    assert read_bits(arr, [4, 0, 0, 3, 0, 7, 0, 0, 13, 0]) == [12, 0, 0, 7, 0, 17, 0, 0, 6969, 0]
    assert read_bits([0, 0, 0, 0, 1], [32, 8]) == [0, 1]
    assert read_bits([0x09, 0x02, 0, 0, 0, 0, 1], [1, 7, 8, 32, 8]) == [1, 4, 2, 0, 1]
    assert read_bits([0x42, 0x43, 0x56], [24]) == [0x564342]
    assert read_bits([0x28, 0x81, 0xd0, 0x90, 0x55, 0x00, 0x00], [5, 5, 4, 24, 16])[2:] == [0, 0x564342, 1]
    # a read past the end fails and leaves the cursor where it was: a shorter read still succeeds
    assert read_bits([0xff], [5, 5]) == [31]
    assert read_bits([0xff], [5, 5, 3]) == [31]          # stops at the first failure
    assert read_bits([0xff], [5, 3]) == [31, 7]
    assert read_bits([], [0, 0]) == [0, 0]
    assert read_bits([0xaa, 0x55, 0xaa, 0x55, 0xaa, 0x55, 0xaa, 0x55], [64]) == [0x55aa55aa55aa55aa]


def test_float32_unpack_vectors():
    """bitpacking.rs:316-358"""
    f = fe.lib().lwf_debug_float32_unpack
    table = [(1611661312, 1.0), (1616117760, 5.0), (1618345984, 11.0), (1620115456, 17.0), (1627381760, 255.0),
             (3759144960, -1.0), (3761242112, -2.0), (3763339264, -4.0), (3763601408, -5.0), (3765436416, -8.0),
             (3765829632, -11.0), (3768451072, -30.0), (3772628992, -119.0), (3780634624, -1530.0),
             (1628434432, 255.0), (1621655552, 17.0), (1619722240, 11.0), (1613234176, 1.0), (3760717824, -1.0),
             (3762814976, -2.0), (3764912128, -4.0), (3765043200, -5.0), (3767009280, -8.0), (3767205888, -11.0),
             (3769565184, -30.0), (3773751296, -119.0), (3781948416, -1530.0)]
    for v, want in table:
        assert f(v) == want, v
    for x in (0.0, 0.125, -3.5, 1024.0, 7.0 / 8):
        assert f(vp.float32_pack(x)) == x


def test_lookup1_values_and_ilog_vectors():
    """header.rs:650-671, lib.rs ilog"""
    f = fe.lib().lwf_debug_lookup1_values
    assert [f(1025, 10), f(1024, 10), f(1023, 10)] == [2, 2, 1]
    assert [f(3126, 5), f(3125, 5), f(3124, 5)] == [5, 5, 4]
    assert [f(1, 1), f(0, 15), f(0, 0), f(1, 0), f(400, 0)] == [1, 0, 0, 0xffffffff, 0xffffffff]
    for entries, dims in ((81, 4), (100, 2), (8, 1), (16, 8), (625, 4), (6561, 8)):
        lv = f(entries, dims)
        assert lv ** dims <= entries < (lv + 1) ** dims
    il = fe.lib().lwf_debug_ilog
    assert [il(0), il(1), il(2), il(3), il(4), il(7), il(255), il(256)] == [0, 1, 2, 2, 3, 3, 8, 9]


def huff(lengths, bits=None, max_out=64):
    L = fe.lib()
    data = b""
    if bits is not None:
        w = vp.BitWriter()
        w.write_bits(bits)
        data = w.bytes()
    out = (C.c_uint32 * max_out)()
    n = C.c_size_t()
    rc = L.lwf_debug_huffman(bytes(lengths), len(lengths), data, len(data), out, max_out, C.byref(n))
    return rc, [int(out[i]) for i in range(n.value)]


def test_huffman_reference_vectors():
    """huffman_tree.rs:262-330"""
    lengths = [2, 4, 4, 4, 4, 2, 3, 3]
    paths = [([0, 0], 0), ([0, 1, 0, 0], 1), ([0, 1, 0, 1], 2), ([0, 1, 1, 0], 3), ([0, 1, 1, 1], 4), ([1, 0], 5),
             ([1, 1, 0], 6), ([1, 1, 1], 7)]
    for bits, want in paths:
        rc, syms = huff(lengths, bits + [1] * 0)
        assert rc == 0 and syms[0] == want, (bits, syms)
    assert huff(list(range(1, 33)) + [32])[0] == 0
    assert huff([0] * 625)[0] == 0                                  # test_issue_8: loads (nothing to decode)
    assert huff([2, 4, 4, 4, 4, 2, 3])[0] != 0                      # underpopulated
    assert huff([2, 4, 4, 4, 2, 3, 3])[0] != 0
    assert huff([2, 4, 4, 4, 4, 2, 3, 3, 3])[0] != 0                # overspecified
    rc, syms = huff([1], [0, 1, 1, 0])
    assert rc == 0 and syms[:4] == [0, 0, 0, 0]                       # single entry: both bit values decode to it
    rc, syms = huff([0, 0, 1, 0], [0, 1])
    assert rc == 0 and syms[:2] == [2, 2]
    assert huff([2])[0] != 0 and huff([0, 3, 0])[0] != 0            # single entry of another length is invalid


@pytest.mark.parametrize("seed", range(6))
def test_huffman_random_codes_round_trip(seed):
    rng = np.random.default_rng(seed)
    cb = vp.Codebook(rng, int(rng.choice([5, 16, 100, 256])), 1, 0, sparse_unused=int(rng.integers(0, 3)), max_len=int(rng.integers(8, 25)))
    entries = [cb.random_entry(rng) for _ in range(200)]
    w = vp.BitWriter()
    for e in entries:
        cb.emit(w, e)
    out = (C.c_uint32 * 400)()
    n = C.c_size_t()
    data = w.bytes()
    rc = fe.lib().lwf_debug_huffman(bytes(cb.lengths), len(cb.lengths), data, len(data), out, 400, C.byref(n))
    assert rc == 0
    assert [int(out[i]) for i in range(200)] == entries


def make_stream(seed, **kw):
    rng = np.random.default_rng(seed)
    spec = vp.StreamSpec(rng, **kw)
    hdr = fe.Headers(spec.ident_packet(), spec.comment_packet(), spec.setup_packet())
    return spec, hdr


@pytest.mark.parametrize("seed,channels", [(1, 1), (2, 2), (3, 2), (4, 6), (5, 3)])
def test_headers_round_trip(seed, channels):
    spec, hdr = make_stream(seed, channels=channels, floor0=(seed % 2 == 1))
    assert (hdr.audio_channels, hdr.blocksize_0, hdr.blocksize_1, hdr.audio_sample_rate) == (channels, 8, 11, 44100)
    assert hdr.bitrate_nominal == 128000
    assert (hdr.n_codebooks, hdr.n_floors, hdr.n_residues, hdr.n_mappings, hdr.n_modes) == (
        len(spec.books), len(spec.floors), len(spec.residues), len(spec.mappings), len(spec.modes))
    assert hdr.vendor == spec.vendor and hdr.comment_list == spec.comments


def test_comment_header_tolerates_bad_entries():
    """header.rs:329-345: comments that are not UTF-8 or lack '=' are skipped, not errors."""
    rng = np.random.default_rng(9)
    spec = vp.StreamSpec(rng)
    hdr = fe.Headers(spec.ident_packet(), spec.comment_packet(extra_raw=[b"\xff\xfe=bad utf8", b"no equals sign", b"K=v=w"]),
                     spec.setup_packet())
    assert hdr.comment_list == spec.comments + [("K", "v=w")]


def test_header_errors():
    rng = np.random.default_rng(10)
    spec = vp.StreamSpec(rng)
    ident, comment, setup = spec.ident_packet(), spec.comment_packet(), spec.setup_packet()

    def code(i, c, s):
        with pytest.raises(fe.HeaderReadError) as e:
            fe.Headers(i, c, s)
        return e.value.code

    assert code(comment, comment, setup) == fe.ERR_HEADER_BAD_TYPE               # wrong packet type
    assert code(b"\x00" + ident[1:], comment, setup) == fe.ERR_HEADER_IS_AUDIO   # first bit clear
    assert code(b"\x01vorbiz" + ident[7:], comment, setup) == fe.ERR_NOT_VORBIS_HEADER
    assert code(ident[:12], comment, setup) == fe.ERR_END_OF_PACKET
    bad_version = bytearray(ident)
    bad_version[7] = 1
    assert code(bytes(bad_version), comment, setup) == fe.ERR_UNSUPPORTED_VERSION
    bad_bs = bytearray(ident)
    bad_bs[28] = 0x8b                                                            # blocksize_0 = 11 > blocksize_1 = 8
    assert code(bytes(bad_bs), comment, setup) == fe.ERR_HEADER_BAD_FORMAT
    no_framing = bytearray(ident)
    no_framing[29] = 0
    assert code(bytes(no_framing), comment, setup) == fe.ERR_HEADER_BAD_FORMAT
    assert code(ident, comment[:-1] + b"\x00", setup) == fe.ERR_HEADER_BAD_FORMAT
    assert code(ident, comment[:20], setup) == fe.ERR_END_OF_PACKET
    assert code(ident, comment, setup[: len(setup) // 2]) == fe.ERR_END_OF_PACKET
    broken = bytearray(setup)
    broken[8] ^= 0xff                                                            # first codebook's sync pattern
    assert code(ident, comment, bytes(broken)) == fe.ERR_HEADER_BAD_FORMAT


def libm():
    m = C.CDLL("libm.so.6")
    for f in ("cosf", "expf", "sqrtf"):
        getattr(m, f).argtypes = [C.c_float]
        getattr(m, f).restype = C.c_float
    return m


def floor0_expected(fl, amp, rows, blockflag, n2, bs0, bs1):
    """Independent f32 restatement of Vorbis I 6.2.2-6.2.3 in the reference's evaluation order
    (audio.rs:109-212, header_cached.rs:129-158), libm called through ctypes."""
    m = libm()
    atanf = C.CDLL("libm.so.6").atanf
    atanf.argtypes, atanf.restype = [C.c_float], C.c_float
    f32 = np.float32

    def bark(x):
        x = f32(x)
        return f32(f32(f32(13.1) * f32(atanf(f32(f32(0.00074) * x)))) + f32(f32(2.24) * f32(atanf(f32(f32(f32(0.0000000185) * x) * x))))) + f32(f32(0.0001) * x)

    n = 1 << ((bs1 if blockflag else bs0) - 1)
    hfl = f32(f32(fl.rate) / f32(2.0))
    hfl_dn = f32(hfl / f32(n))
    const = f32(f32(fl.bark_map_size) / f32(bark(hfl)))
    bms_m1 = f32(f32(fl.bark_map_size) - f32(1.0))
    omega_factor = f32(f32(np.pi) / f32(fl.bark_map_size))
    cos_omega = []
    for i in range(n):
        fb = f32(np.floor(f32(f32(bark(f32(f32(i) * hfl_dn))) * const)))
        cos_omega.append(f32(m.cosf(f32(min(fb, bms_m1) * omega_factor))))
    # coefficient cosines
    coeffs = []
    last = f32(0)
    for row in rows:
        last_new = last
        for e in row:
            coeffs.append(f32(m.cosf(f32(last + f32(e)))))
            last_new = f32(e)
            if len(coeffs) == fl.order:
                break
        last = f32(last + last_new)
        if len(coeffs) >= fl.order:
            break
    common = f32(f32(f32(amp) * f32(fl.amplitude_offset)) / f32((1 << fl.amplitude_bits) - 1))
    out = np.zeros(n2, f32)
    i = 0
    while i < n2:
        co = cos_omega[i]
        if fl.order & 1:
            pu, qu = (fl.order - 3) // 2, (fl.order - 1) // 2
            p, q = f32(f32(1.0) - f32(co * co)), f32(0.25)
        else:
            pu = qu = (fl.order - 2) // 2
            p, q = f32(f32(f32(1.0) - co) / f32(2.0)), f32(f32(f32(1.0) + co) / f32(2.0))
        for j in range(pu + 1):
            pm = f32(coeffs[2 * j + 1] - co)
            p = f32(p * f32(f32(f32(4.0) * pm) * pm))
        for j in range(qu + 1):
            qm = f32(coeffs[2 * j] - co)
            q = f32(q * f32(f32(f32(4.0) * qm) * qm))
        lfv = f32(m.expf(f32(f32(0.11512925) * f32(f32(common / f32(m.sqrtf(f32(p + q)))) - f32(fl.amplitude_offset)))))
        while i < n2 and cos_omega[i] == co:
            out[i] = lfv
            i += 1
    return out


def check_packet(spec, hdr, pkt, info, nbytes=None):
    want_floors, want_res = spec.expected(info, nbytes)
    data = pkt if nbytes is None else pkt[:nbytes]
    got = hdr.decode_packet(data)
    assert got.mode_number == info["mode"] and got.n == info["n"]
    assert (got.prev_window_flag, got.next_window_flag) == (bool(info["prev"]), bool(info["next"]))
    n2 = info["n"] // 2
    for c in range(spec.channels):
        w = want_floors[c]
        g = got.floors[c]
        if w is None:
            assert g is None, (c, "floor should be unused")
        elif w[0] == "one":
            assert g is not None and g.dtype == np.uint32, c
            assert list(g[: len(w[1])]) == w[1], c
        else:
            curve = floor0_expected(w[3], w[1], w[2], info["blockflag"], n2, spec.bs0, spec.bs1)
            assert g is not None and g.dtype == np.float32
            assert np.array_equal(g.view(np.uint32), curve.view(np.uint32)), (c, np.abs(g - curve).max())
    assert np.array_equal(got.residue.view(np.uint32), want_res.view(np.uint32)), np.abs(got.residue - want_res).max()


@pytest.mark.parametrize("seed,channels,floor0", [(20, 1, False), (21, 2, False), (22, 2, True), (23, 6, False), (24, 3, True),
                                                  (25, 2, False), (26, 4, False), (27, 2, True)])
def test_packet_decode_matches_what_was_packed(seed, channels, floor0):
    """Random setup, random packets of every mode: mode / flags / floor posts (or floor-0 curves) /
    residue vectors bit-identical to what the packer encoded (residue types 0, 1, 2; VQ lookup 1, 2;
    sparse / ordered books; submaps; coupling-driven do-not-decode propagation)."""
    spec, hdr = make_stream(seed, channels=channels, floor0=floor0)
    rng = spec.rng
    for k in range(12):
        mode = int(rng.integers(0, len(spec.modes)))
        pkt, info = spec.audio_packet(mode, int(rng.integers(0, 2)), int(rng.integers(0, 2)), p_unused=0.25)
        check_packet(spec, hdr, pkt, info)
        assert hdr.decoded_sample_count(pkt) == _sample_count(spec, info)


def _sample_count(spec, info):
    n, n0 = info["n"], 1 << spec.bs0
    ls = 0 if info["prev"] else (n - n0) >> 2
    rs = n >> 1 if info["next"] else (n * 3 - n0) >> 2
    return rs - ls


@pytest.mark.parametrize("seed,rtype", [(30, 0), (31, 1), (32, 2), (33, None)])
def test_truncated_packets_end_of_packet_rules(seed, rtype):
    """audio.rs:82-104, :640-716: a packet cut anywhere is not an error once its header bits are there --
    the floor being read (and every later one) becomes unused, residue decode stops where the data ends
    and keeps what was accumulated."""
    spec, hdr = make_stream(seed, channels=2, residue_types=[rtype] if rtype is not None else None)
    rng = spec.rng
    for k in range(4):
        mode = int(rng.integers(0, len(spec.modes)))
        pkt, info = spec.audio_packet(mode, 1, 1, p_unused=0.1)
        cuts = sorted(c for c in set([1, 2, 3, len(pkt) - 1] + rng.integers(1, max(2, len(pkt)), 12).tolist()) if 0 < c < len(pkt))
        for nb in cuts:
            if nb * 8 < info["header_bits"]:
                with pytest.raises(fe.AudioReadError) as e:
                    hdr.decode_packet(pkt[:nb])
                assert e.value.kind == "EndOfPacket"
                continue
            check_packet(spec, hdr, pkt, info, nb)
    with pytest.raises(fe.AudioReadError) as e:
        hdr.decode_packet(b"")
    assert e.value.kind == "EndOfPacket"
    with pytest.raises(fe.AudioReadError) as e:
        hdr.decode_packet(spec.ident_packet())
    assert e.value.kind == "AudioIsHeader"


def test_mode_number_out_of_range_is_bad_format():
    """audio.rs:926-930"""
    rng = np.random.default_rng(40)
    spec = vp.StreamSpec(rng, n_modes=3)           # 2 mode bits, mode 3 does not exist
    hdr = fe.Headers(spec.ident_packet(), spec.comment_packet(), spec.setup_packet())
    w = vp.BitWriter()
    w.write(0, 1)
    w.write(3, 2)
    w.write(0, 16)
    with pytest.raises(fe.AudioReadError) as e:
        hdr.decode_packet(w.bytes())
    assert e.value.kind == "AudioBadFormat"


def test_ogg_paging_round_trip_and_errors():
    rng = np.random.default_rng(50)
    spec = vp.StreamSpec(rng)
    hdrs = [spec.ident_packet(), spec.comment_packet(), spec.setup_packet()]
    pkts = [bytes(rng.integers(0, 256, int(n), dtype=np.uint8)) for n in (10, 255, 510, 0, 1, 300, 70000, 5)]
    # hand-made pages: packet 6 (70000 bytes) spans three pages
    pages = [vp.ogg_page(7, 0, 0, [(hdrs[0], True)], bos=True), vp.ogg_page(7, 1, 0, [(hdrs[1], True), (hdrs[2], True)]),
             vp.ogg_page(7, 2, 1000, [(pkts[0], True), (pkts[1], True), (pkts[2], True)]),
             vp.ogg_page(7, 3, 2000, [(pkts[3], True), (pkts[4], True), (pkts[5], True), (pkts[6][:255 * 200], False)]),
             vp.ogg_page(7, 4, 2000, [(pkts[6][255 * 200: 255 * 255], False)], continued=True),
             vp.ogg_page(7, 5, 3000, [(pkts[6][255 * 255:], True), (pkts[7], True)], continued=True, eos=True)]
    data = b"".join(pages)
    rd = fe.OggPacketReader(data)
    got = []
    while True:
        p = rd.read_packet()
        if p is None:
            break
        got.append(p)
    assert [g.data for g in got] == hdrs + pkts
    assert all(g.stream_serial == 7 for g in got)
    assert got[0].first_in_stream and not any(g.first_in_stream for g in got[1:])
    assert got[-1].last_in_stream and not any(g.last_in_stream for g in got[:-1])
    assert [g.absgp_page for g in got] == [0, 0, 0, 1000, 1000, 1000, 2000, 2000, 2000, 3000, 3000]
    assert [g.last_in_page for g in got] == [True, False, True, False, False, True, False, False, True, False, True]
    # a flipped payload bit fails the page CRC
    bad = bytearray(data)
    bad[len(pages[0]) + len(pages[1]) + 40] ^= 1
    rd = fe.OggPacketReader(bytes(bad))
    for _ in range(3):
        rd.read_packet()
    with pytest.raises(fe.OggReadError):
        rd.read_packet()
    with pytest.raises(fe.OggReadError):
        fe.OggPacketReader(b"NotAnOggFileAtAllButLongEnoughToHoldAHeader").read_packet()
    hs, serial = fe.read_headers(fe.OggPacketReader(data))
    assert serial == 7 and hs.audio_channels == spec.channels


@pytest.mark.parametrize("seed,bs0,bs1,channels,floor0", [(60, 6, 8, 2, True), (61, 9, 13, 1, False), (62, 7, 7, 3, True),
                                                           (63, 6, 13, 2, False)])
def test_packet_decode_other_blocksizes(seed, bs0, bs1, channels, floor0):
    """Blocksizes other than 256/2048 (bark maps, residue limits and floor ranges all scale with them)."""
    rng = np.random.default_rng(seed)
    spec = vp.StreamSpec(rng, channels=channels, bs0=bs0, bs1=bs1, floor0=floor0)
    hdr = fe.Headers(spec.ident_packet(), spec.comment_packet(), spec.setup_packet())
    assert (hdr.blocksize_0, hdr.blocksize_1) == (bs0, bs1)
    for k in range(8):
        mode = int(rng.integers(0, len(spec.modes)))
        pkt, info = spec.audio_packet(mode, int(rng.integers(0, 2)), int(rng.integers(0, 2)), p_unused=0.2)
        check_packet(spec, hdr, pkt, info)
        assert hdr.decoded_sample_count(pkt) == _sample_count(spec, info)
        if len(pkt) > 4:
            check_packet(spec, hdr, pkt, info, int(rng.integers(2, len(pkt))))


def vq_accumulate(spec, hdr, dp, runs, ents):
    """What the device does with a packet's VQ runs (kernel_prologue.cuh d_vq_accumulate): per coefficient the f32
    additions happen pass by pass; kinds 0 / 1 / 2 = residue types 1 / 0 / 2 (audio.rs:587-618, :744-756)."""
    C, n2 = spec.channels, dp.n // 2
    acc = np.zeros(C * n2, np.float32)
    mp = spec.mappings[spec.modes[dp.mode_number][1]]
    for p in range(8):
        touched = np.zeros(C * n2, bool)
        for r in runs:
            pk = int(r["pass_kind"])
            if pk & 7 != p:
                continue
            book = spec.books[int(r["book"])]
            kind, pos, first = (pk >> 3) & 3, int(r["pos"]), int(r["first"])
            for q in range(int(r["count"])):
                v = np.asarray(book.vq[int(ents[first + q])], np.float32)
                if kind == 0:
                    idx = pos + q * book.dims + np.arange(book.dims)
                elif kind == 1:
                    step = spec.residues[int(r["aux"])].partition_size // book.dims
                    idx = pos + q + np.arange(book.dims) * step
                else:
                    chs = [c for c in range(C) if mp["mux"][c] == int(r["aux"])]
                    t = pos + q * book.dims + np.arange(book.dims)
                    keep = (t // len(chs)) < n2
                    t, v = t[keep], v[keep]
                    idx = np.array([chs[i % len(chs)] for i in t], np.int64) * n2 + t // len(chs)
                assert not touched[idx].any(), "two vectors of one pass overlap"
                touched[idx] = True
                acc[idx] = (acc[idx] + v).astype(np.float32)
    return acc.reshape(C, n2)


@pytest.mark.parametrize("seed,channels,rtype", [(60, 2, 0), (61, 2, 1), (62, 2, 2), (63, 6, None), (64, 1, None), (65, 3, 2)])
def test_vq_records_reproduce_the_dense_residue(seed, channels, rtype):
    """LWB_ENTRY_VQ's host side: lwf_packet_decode_vq emits one run per partition read and one 16-bit entry per VQ
    vector; adding them up the way the
    device does (pass by pass) gives, bit for bit, the residue vectors lwf_packet_decode accumulates on the host --
    for whole packets and for packets cut at arbitrary bytes (the reference keeps what was decoded, audio.rs:640-716).
    Floors and mode bits are the same as the dense decode's."""
    spec, hdr = make_stream(seed, channels=channels, residue_types=[rtype] if rtype is not None else None)
    assert hdr.vq_capable()
    rng = spec.rng
    for k in range(6):
        mode = int(rng.integers(0, len(spec.modes)))
        pkt, info = spec.audio_packet(mode, 1, 1, p_unused=0.15)
        cuts = [len(pkt)] + [c for c in rng.integers(1, max(2, len(pkt)), 5).tolist() if c * 8 >= info["header_bits"]]
        for nb in cuts:
            dense = hdr.decode_packet(pkt[:nb])
            dp, runs, ents = hdr.decode_packet_vq(pkt[:nb])
            assert int(runs["count"].sum()) == len(ents)
            assert (dp.mode_number, dp.prev_window_flag, dp.next_window_flag, dp.n) == (dense.mode_number, dense.prev_window_flag,
                                                                                      dense.next_window_flag, dense.n)
            for a, b in zip(dp.floors, dense.floors):
                assert (a is None) == (b is None) and (a is None or np.array_equal(np.asarray(a), np.asarray(b)))
            got = vq_accumulate(spec, hdr, dp, runs, ents)
            assert np.array_equal(got.view(np.uint32), dense.residue.view(np.uint32)), (k, nb, np.abs(got - dense.residue).max())
