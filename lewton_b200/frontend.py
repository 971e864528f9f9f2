"""Host front half (include/lewton_frontend.h) -- ctypes mirror of the reference's public reading API:

  read_headers / Headers          header.rs:221, :309, :1082  (IdentHeader, CommentHeader, SetupHeader)
  Headers.decode_packet           audio.rs:919-986            (front half of read_audio_packet_generic)
  Headers.decoded_sample_count    audio.rs:874-909            (get_decoded_sample_count)
  OggPacketReader                 ogg::PacketReader as inside_ogg.rs uses it
  OggStreamReader                 inside_ogg.rs:60-227        (read_dec_packet, read_dec_packet_itl, get_last_absgp)

The entropy decode is CPU work by nature and runs on the host; synthesis goes through the CUDA back
half (lwb_decode_packet / lwb_decode_chains)."""
import ctypes as C

import numpy as np

from . import _cabi as cabi
from .api import AudioReadError, DecodedPacket, Setup

(ERR_END_OF_PACKET, ERR_NOT_VORBIS_HEADER, ERR_UNSUPPORTED_VERSION, ERR_HEADER_BAD_FORMAT, ERR_HEADER_BAD_TYPE,
 ERR_HEADER_IS_AUDIO, ERR_UTF8, ERR_AUDIO_IS_HEADER, ERR_OGG, ERR_NO_MORE_PACKETS) = range(16, 26)

SYMBOLS = ["lwf_headers_parse", "lwf_headers_destroy", "lwf_headers_info", "lwf_headers_comment", "lwf_headers_make_setup",
           "lwf_packet_decode", "lwf_packet_decode_vq", "lwf_headers_vq_capable", "lwf_decoded_sample_count", "lwf_ogg_open", "lwf_ogg_close", "lwf_ogg_next_packet",
           "lwf_reader_open", "lwf_reader_close", "lwf_reader_headers", "lwf_reader_read_dec_packet", "lwf_reader_last_absgp",
           "lwf_reader_skip_samples_linear", "lwf_reader_seek_absgp_pg",
           "lwf_batcher_create", "lwf_batcher_destroy", "lwf_batcher_set_entry", "lwf_batcher_decode", "lwf_batcher_last_timing",
           "lwf_debug_float32_unpack", "lwf_debug_lookup1_values", "lwf_debug_ilog", "lwf_debug_read_bits", "lwf_debug_huffman",
           "lwf_debug_decode_loop"]


VQ_RUN_DTYPE = np.dtype([("pos", np.uint16), ("first", np.uint16), ("book", np.uint8), ("pass_kind", np.uint8), ("aux", np.uint8),
                         ("count", np.uint8)])                                              # lwb_vq_run


class HeaderReadError(Exception):
    """header.rs:35-44"""

    def __init__(self, code):
        names = {ERR_END_OF_PACKET: "EndOfPacket", ERR_NOT_VORBIS_HEADER: "NotVorbisHeader",
                 ERR_UNSUPPORTED_VERSION: "UnsupportedVorbisVersion", ERR_HEADER_BAD_FORMAT: "HeaderBadFormat",
                 ERR_HEADER_BAD_TYPE: "HeaderBadType", ERR_HEADER_IS_AUDIO: "HeaderIsAudio", ERR_UTF8: "Utf8DecodeError",
                 cabi.ERR_BUFFER: "BufferNotAddressable"}
        super().__init__(names.get(code, "code %d" % code))
        self.code = code


class OggReadError(Exception):
    pass


class Info(C.Structure):
    _fields_ = [("audio_channels", C.c_uint8), ("blocksize_0", C.c_uint8), ("blocksize_1", C.c_uint8),
                ("audio_sample_rate", C.c_uint32), ("bitrate_maximum", C.c_int32), ("bitrate_nominal", C.c_int32),
                ("bitrate_minimum", C.c_int32), ("n_codebooks", C.c_uint32), ("n_floors", C.c_uint32),
                ("n_residues", C.c_uint32), ("n_mappings", C.c_uint32), ("n_modes", C.c_uint32), ("n_comments", C.c_uint32)]


class _DecodedPacket(C.Structure):
    _fields_ = [("mode_number", C.c_uint8), ("blockflag", C.c_uint8), ("prev_window_flag", C.c_uint8),
                ("next_window_flag", C.c_uint8), ("n", C.c_uint32), ("floor_kind", cabi.u8p), ("floor1_y", cabi.u32p),
                ("dense_floor", cabi.fp), ("residue", cabi.fp)]


class _OggPacket(C.Structure):
    _fields_ = [("data", cabi.u8p), ("len", C.c_size_t), ("stream_serial", C.c_uint32), ("absgp_page", C.c_uint64),
                ("first_in_stream", C.c_uint8), ("last_in_stream", C.c_uint8), ("first_in_page", C.c_uint8),
                ("last_in_page", C.c_uint8)]


class _StreamJob(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("n_packets", C.c_uint32), ("packets", C.POINTER(C.c_char_p)),
                ("lengths", C.POINTER(C.c_size_t)), ("out_offset", C.c_uint64), ("out_stride", C.c_uint64),
                ("n_samples", C.c_uint32), ("packets_done", C.c_uint32), ("status", C.c_int32)]


_declared = False


def lib():
    global _declared
    L = cabi.lib()
    if not _declared:
        vp, sz = C.c_void_p, C.c_size_t
        L.lwf_headers_parse.argtypes = [C.c_char_p, sz, C.c_char_p, sz, C.c_char_p, sz, C.POINTER(vp)]
        L.lwf_headers_destroy.argtypes = [vp]
        L.lwf_headers_destroy.restype = None
        L.lwf_headers_info.argtypes = [vp, C.POINTER(Info)]
        L.lwf_headers_comment.argtypes = [vp, C.c_int, C.c_char_p, sz]
        L.lwf_headers_comment.restype = sz
        L.lwf_headers_make_setup.argtypes = [vp, vp, C.POINTER(vp)]
        L.lwf_packet_decode.argtypes = [vp, C.c_char_p, sz, C.POINTER(_DecodedPacket)]
        L.lwf_decoded_sample_count.argtypes = [vp, C.c_char_p, sz, C.POINTER(sz)]
        L.lwf_ogg_open.argtypes = [C.c_char_p, sz, C.POINTER(vp)]
        L.lwf_ogg_close.argtypes = [vp]
        L.lwf_ogg_close.restype = None
        L.lwf_ogg_next_packet.argtypes = [vp, C.POINTER(_OggPacket)]
        L.lwf_reader_open.argtypes = [vp, C.c_char_p, sz, C.POINTER(vp)]
        L.lwf_reader_close.argtypes = [vp]
        L.lwf_reader_close.restype = None
        L.lwf_reader_headers.argtypes = [vp]
        L.lwf_reader_headers.restype = vp
        L.lwf_reader_read_dec_packet.argtypes = [vp, C.c_int, vp, sz, C.POINTER(sz)]
        L.lwf_reader_last_absgp.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.lwf_reader_skip_samples_linear.argtypes = [vp, sz, C.c_int, vp, sz, C.POINTER(sz), C.POINTER(sz), C.POINTER(C.c_int)]
        L.lwf_reader_seek_absgp_pg.argtypes = [vp, C.c_uint64]
        L.lwf_batcher_create.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
        L.lwf_batcher_destroy.argtypes = [vp]
        L.lwf_batcher_destroy.restype = None
        L.lwf_batcher_set_entry.argtypes = [vp, C.c_int]
        L.lwf_headers_vq_capable.argtypes = [vp]
        L.lwf_packet_decode_vq.argtypes = [vp, C.c_char_p, sz, vp, vp, sz, C.POINTER(sz), vp, sz, C.POINTER(sz)]
        L.lwf_batcher_decode.argtypes = [vp, C.POINTER(_StreamJob), sz, C.c_int, vp]
        L.lwf_batcher_last_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.lwf_batcher_last_timing.restype = None
        L.lwf_debug_float32_unpack.argtypes = [C.c_uint32]
        L.lwf_debug_float32_unpack.restype = C.c_float
        L.lwf_debug_lookup1_values.argtypes = [C.c_uint32, C.c_uint16]
        L.lwf_debug_lookup1_values.restype = C.c_uint32
        L.lwf_debug_ilog.argtypes = [C.c_uint64]
        L.lwf_debug_ilog.restype = C.c_uint8
        L.lwf_debug_read_bits.argtypes = [C.c_char_p, sz, C.c_char_p, sz, C.POINTER(C.c_uint64)]
        L.lwf_debug_read_bits.restype = sz
        L.lwf_debug_huffman.argtypes = [C.c_char_p, sz, C.c_char_p, sz, cabi.u32p, sz, C.POINTER(sz)]
        _declared = True
    return L


class Headers:
    """The three parsed Vorbis headers."""

    def __init__(self, ident, comment, setup, _handle=None):
        self._own = _handle is None
        if _handle is None:
            h = C.c_void_p()
            rc = lib().lwf_headers_parse(ident, len(ident), comment, len(comment), setup, len(setup), C.byref(h))
            if rc:
                raise HeaderReadError(rc)
            _handle = h.value
        self._h = _handle
        info = Info()
        lib().lwf_headers_info(self._h, C.byref(info))
        self.info = info
        for f, _ in Info._fields_:
            setattr(self, f, getattr(info, f))

    def _str(self, index):
        n = lib().lwf_headers_comment(self._h, index, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().lwf_headers_comment(self._h, index, buf, n + 1)
        return buf.raw[:n].decode("utf-8")

    @property
    def vendor(self):
        return self._str(-1)

    @property
    def comment_list(self):
        out = []
        for i in range(self.n_comments):
            k, _, v = self._str(i).partition("=")
            out.append((k, v))
        return out

    def make_setup(self, ctx):
        """The device-side header constants (lwb_setup) for these headers."""
        h = C.c_void_p()
        ctx.check(lib().lwf_headers_make_setup(self._h, ctx._h, C.byref(h)))
        return Setup._adopt(ctx, h.value, self.audio_channels, self.blocksize_0, self.blocksize_1)

    def decode_packet(self, packet):
        """audio.rs:919-986: returns an api.DecodedPacket (mode, window flags, per-channel floors, residue)."""
        Cn, n2max = self.audio_channels, (1 << self.blocksize_1) // 2
        kinds = np.zeros(Cn, np.uint8)
        ys = np.zeros((Cn, cabi.MAX_POSTS), np.uint32)
        dense = np.zeros((Cn, n2max), np.float32)
        res = np.zeros((Cn, n2max), np.float32)
        dp = _DecodedPacket()
        dp.floor_kind = kinds.ctypes.data_as(cabi.u8p)
        dp.floor1_y = ys.ctypes.data_as(cabi.u32p)
        dp.dense_floor = dense.ctypes.data_as(cabi.fp)
        dp.residue = res.ctypes.data_as(cabi.fp)
        rc = lib().lwf_packet_decode(self._h, bytes(packet), len(packet), C.byref(dp))
        if rc == cabi.ERR_BAD_FORMAT:
            raise AudioReadError(rc)
        if rc:
            e = AudioReadError(rc)
            e.kind = {ERR_END_OF_PACKET: "EndOfPacket", ERR_AUDIO_IS_HEADER: "AudioIsHeader"}.get(rc, e.kind)
            raise e
        n2 = dp.n // 2
        floors = []
        flat_dense = dense.ravel()
        flat_res = res.ravel()
        for c in range(Cn):
            if kinds[c] == cabi.FLOOR_UNUSED:
                floors.append(None)
            elif kinds[c] == cabi.FLOOR_ONE:
                floors.append(ys[c].copy())
            else:
                floors.append(flat_dense[c * n2:(c + 1) * n2].copy())
        residue = flat_res[: Cn * n2].reshape(Cn, n2).copy()
        out = DecodedPacket(dp.mode_number, residue, floors, dp.prev_window_flag, dp.next_window_flag)
        out.blockflag, out.n = bool(dp.blockflag), dp.n
        return out

    def vq_capable(self):
        """True if LWB_ENTRY_VQ applies to this stream (lwf_headers_vq_capable)."""
        return bool(lib().lwf_headers_vq_capable(self._h))

    def decode_packet_vq(self, packet):
        """The same front half with the residue left as VQ runs: (DecodedPacket with a zero residue, runs, entries):
        runs a structured array (lwb_vq_run), entries the uint16 codebook entries they index."""
        Cn, n2max = self.audio_channels, (1 << self.blocksize_1) // 2
        kinds = np.zeros(Cn, np.uint8)
        ys = np.zeros((Cn, cabi.MAX_POSTS), np.uint32)
        dense = np.zeros((Cn, n2max), np.float32)
        dp = _DecodedPacket()
        dp.floor_kind = kinds.ctypes.data_as(cabi.u8p)
        dp.floor1_y = ys.ctypes.data_as(cabi.u32p)
        dp.dense_floor = dense.ctypes.data_as(cabi.fp)
        cap = len(packet) * 8 + 16
        runs, ents = np.zeros(cap, VQ_RUN_DTYPE), np.zeros(cap, np.uint16)
        n, ne = C.c_size_t(), C.c_size_t()
        rc = lib().lwf_packet_decode_vq(self._h, bytes(packet), len(packet), C.byref(dp), runs.ctypes.data, cap, C.byref(n),
                                        ents.ctypes.data, cap, C.byref(ne))
        if rc == cabi.ERR_BAD_FORMAT:
            raise AudioReadError(rc)
        if rc:
            e = AudioReadError(rc)
            e.kind = {ERR_END_OF_PACKET: "EndOfPacket", ERR_AUDIO_IS_HEADER: "AudioIsHeader"}.get(rc, e.kind)
            raise e
        n2 = dp.n // 2
        floors = []
        flat_dense = dense.ravel()
        for c in range(Cn):
            if kinds[c] == cabi.FLOOR_UNUSED:
                floors.append(None)
            elif kinds[c] == cabi.FLOOR_ONE:
                floors.append(ys[c].copy())
            else:
                floors.append(flat_dense[c * n2:(c + 1) * n2].copy())
        out = DecodedPacket(dp.mode_number, np.zeros((Cn, n2), np.float32), floors, dp.prev_window_flag, dp.next_window_flag)
        out.blockflag, out.n = bool(dp.blockflag), dp.n
        return out, runs[: n.value].copy(), ents[: ne.value].copy()

    def decoded_sample_count(self, packet):
        n = C.c_size_t()
        rc = lib().lwf_decoded_sample_count(self._h, bytes(packet), len(packet), C.byref(n))
        if rc:
            raise AudioReadError(rc, "get_decoded_sample_count")
        return n.value

    def close(self):
        if self._h and self._own:
            lib().lwf_headers_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OggPacket:
    def __init__(self, p):
        self.data = C.string_at(p.data, p.len) if p.len else b""
        self.stream_serial, self.absgp_page = p.stream_serial, p.absgp_page
        self.first_in_stream, self.last_in_stream = bool(p.first_in_stream), bool(p.last_in_stream)
        self.first_in_page, self.last_in_page = bool(p.first_in_page), bool(p.last_in_page)


class OggPacketReader:
    def __init__(self, data):
        self._data = bytes(data)
        h = C.c_void_p()
        rc = lib().lwf_ogg_open(self._data, len(self._data), C.byref(h))
        if rc:
            raise OggReadError("open: %d" % rc)
        self._h = h.value

    def read_packet(self):
        """Next packet or None at the end of the data."""
        p = _OggPacket()
        rc = lib().lwf_ogg_next_packet(self._h, C.byref(p))
        if rc == ERR_NO_MORE_PACKETS:
            return None
        if rc:
            raise OggReadError("framing error (%d)" % rc)
        return OggPacket(p)

    def close(self):
        if self._h:
            lib().lwf_ogg_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_headers(reader):
    """inside_ogg.rs:19-39 on an OggPacketReader: (Headers, stream_serial)."""
    pk = reader.read_packet()
    ident, serial = pk.data, pk.stream_serial
    pk = reader.read_packet()
    while pk.stream_serial != serial:
        pk = reader.read_packet()
    comment = pk.data
    pk = reader.read_packet()
    while pk.stream_serial != serial:
        pk = reader.read_packet()
    return Headers(ident, comment, pk.data), serial


class OggStreamReader:
    """inside_ogg.rs:60-227: ogg/vorbis bytes in, PCM packets out (synthesis on the GPU)."""

    def __init__(self, ctx, data):
        self.ctx = ctx
        self._data = bytes(data)
        h = C.c_void_p()
        rc = lib().lwf_reader_open(ctx._h, self._data, len(self._data), C.byref(h))
        if rc:
            if 16 <= rc <= 22:
                raise HeaderReadError(rc)
            if rc >= ERR_OGG:
                raise OggReadError("code %d" % rc)
            ctx.check(rc)
        self._h = h.value
        self._buf = None
        # more than one logical stream in the data: the next one may have any channel count / blocksize
        self._may_chain = self._data.count(b"OggS\x00\x02") > 1
        ctx._children.add(self)
        self._refresh()

    def _refresh(self):
        self.headers = Headers(None, None, None, _handle=lib().lwf_reader_headers(self._h))
        self.ident_hdr = self.headers

    def _read(self, fmt, dtype, interleaved, skip=None):
        total = 255 * 8192 if self._may_chain else self.headers.audio_channels << self.headers.blocksize_1
        if self._buf is None or self._buf.size < total or self._buf.dtype != dtype:
            self._buf = np.zeros(total, dtype)
        buf = self._buf
        n = C.c_size_t()
        if skip is None:
            rc = lib().lwf_reader_read_dec_packet(self._h, fmt, buf.ctypes.data, buf.size, C.byref(n))
        else:
            left, got = C.c_size_t(), C.c_int()
            rc = lib().lwf_reader_skip_samples_linear(self._h, skip, fmt, buf.ctypes.data, buf.size, C.byref(n), C.byref(left),
                                                      C.byref(got))
            self._skip_left = left.value
            if rc == 0 and not got.value:
                return None
        if rc == ERR_NO_MORE_PACKETS:
            return None
        if rc == cabi.ERR_BAD_FORMAT:
            raise AudioReadError(rc)
        if rc in (ERR_END_OF_PACKET, ERR_AUDIO_IS_HEADER):
            e = AudioReadError(rc)
            e.kind = {ERR_END_OF_PACKET: "EndOfPacket", ERR_AUDIO_IS_HEADER: "AudioIsHeader"}[rc]
            raise e
        if 16 <= rc <= 23:               # the headers of a chained stream (inside_ogg.rs:118-137): VorbisError::BadHeader
            raise HeaderReadError(rc)
        if rc >= ERR_OGG:
            raise OggReadError("code %d" % rc)
        self.ctx.check(rc)
        if lib().lwf_reader_headers(self._h) != self.headers._h:
            self._refresh()                      # a chained stream started
        Cn = self.headers.audio_channels
        if interleaved:
            return buf[: n.value * Cn].copy()
        cap = buf.size // Cn
        return [buf[c * cap: c * cap + n.value].copy() for c in range(Cn)]

    def read_dec_packet(self):
        """Vec<Vec<i16>> or None"""
        return self._read(cabi.OUT_I16_PLANAR, np.int16, False)

    def read_dec_packet_itl(self):
        """interleaved Vec<i16> or None"""
        return self._read(cabi.OUT_I16_INTERLEAVED, np.int16, True)

    def read_dec_packet_f32(self):
        """read_dec_packet_generic::<Vec<Vec<f32>>>"""
        return self._read(cabi.OUT_F32_PLANAR, np.float32, False)

    def skip_samples_linear(self, to_skip, sample="f32"):
        """inside_ogg.rs:244-283 -> (Option<S>, usize): (planar packet or None, samples left to skip inside it)."""
        fmt, dt = (cabi.OUT_F32_PLANAR, np.float32) if sample == "f32" else (cabi.OUT_I16_PLANAR, np.int16)
        pck = self._read(fmt, dt, False, skip=int(to_skip))
        return pck, self._skip_left

    def seek_absgp_pg(self, absgp):
        """inside_ogg.rs:307-313: page-granular seek to a position <= absgp; resets cur_absgp and PreviousWindowRight."""
        rc = lib().lwf_reader_seek_absgp_pg(self._h, int(absgp))
        if rc >= ERR_OGG:
            raise OggReadError("code %d" % rc)
        self.ctx.check(rc)

    def get_last_absgp(self):
        v = C.c_uint64()
        return v.value if lib().lwf_reader_last_absgp(self._h, C.byref(v)) == 0 else None

    def close(self):
        if self._h:
            lib().lwf_reader_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StreamBatcher:
    """lwf_batcher: entropy-decode the packets of many streams (one shared set of headers) on a host
    thread pool and synthesise them with one batched call.  jobs: list of (PreviousWindowRight,
    [packet bytes, ...]); PCM lands planar in `pcm` at out_offset = job index * channels * stride."""

    def __init__(self, ctx, headers, threads=0, entry=cabi.ENTRY_RESIDUE):
        self.ctx, self.headers = ctx, headers
        h = C.c_void_p()
        ctx.check(lib().lwf_batcher_create(ctx._h, headers._h, threads, C.byref(h)))
        self._h = h.value
        ctx._children.add(self)
        if entry != cabi.ENTRY_RESIDUE:        # LWB_ENTRY_VQ: VQ records instead of dense residue vectors cross the boundary
            rc = lib().lwf_batcher_set_entry(self._h, entry)
            if rc:
                raise AudioReadError(rc, "this stream does not qualify for LWB_ENTRY_VQ (lwf_headers_vq_capable)")

    def decode(self, jobs, pcm, stride, out_format=cabi.OUT_F32_PLANAR):
        n = len(jobs)
        arr = (_StreamJob * n)()
        keep = []
        Cn = self.headers.audio_channels
        for j, (pwr, packets) in enumerate(jobs):
            pk = (C.c_char_p * len(packets))(*packets)
            ln = (C.c_size_t * len(packets))(*[len(p) for p in packets])
            keep.append((pk, ln))
            arr[j].stream = pwr._h
            arr[j].n_packets = len(packets)
            arr[j].packets = pk
            arr[j].lengths = ln
            arr[j].out_offset = j * Cn * stride
            arr[j].out_stride = stride
        self.prepared = (arr, keep, n)
        return self.run(pcm, out_format)

    def run(self, pcm, out_format=cabi.OUT_F32_PLANAR):
        """Decode the jobs of the last decode() again (same packets; benchmarking)."""
        arr, _, n = self.prepared
        self.ctx.check(lib().lwf_batcher_decode(self._h, arr, n, out_format, pcm.ctypes.data))
        e, s = C.c_double(), C.c_double()
        lib().lwf_batcher_last_timing(self._h, C.byref(e), C.byref(s))
        self.entropy_seconds, self.synthesis_seconds = e.value, s.value
        return [(arr[j].n_samples, arr[j].packets_done, arr[j].status) for j in range(n)]

    def close(self):
        if self._h:
            lib().lwf_batcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
