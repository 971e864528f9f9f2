/*
 * lewton_frontend.h -- C ABI of the HOST front half that feeds lewton_b200.h: Vorbis header
 * parsing, audio-packet entropy decode up to the cut at audio.rs:986, Ogg paging, and the
 * OggStreamReader-style loop around them.
 *
 * This is CPU code by nature (bit-serial Huffman / VQ decode); it is the part of lewton that stays
 * on the host in the drop-in design.  In a Rust build the crate's own front half
 * plays this role; this C++ restatement exists because Rust is not available in the build image,
 * so that whole streams can be decoded end to end and the batch / residue entry points of the
 * CUDA back end can be driven by real bitstreams.
 *
 * Reference interfaces mirrored (file:line in /root/reference/src):
 *   lwf_headers_parse            header.rs:221 read_header_ident, :309 read_header_comment,
 *                                :1082 read_header_setup
 *   lwf_packet_decode            audio.rs:919-986 (front half of read_audio_packet_generic),
 *                                :109-158 floor_zero_decode, :160-212 floor_zero_compute_curve,
 *                                :215-251 floor_one_decode, :557-760 floor/residue decode
 *   lwf_decoded_sample_count     audio.rs:874-909 get_decoded_sample_count
 *   lwf_ogg_*                    ogg 0.8.0 PacketReader as used by inside_ogg.rs:16-143
 *   lwf_reader_*                 inside_ogg.rs:60-313 OggStreamReader (read_dec_packet[_itl], skip_samples_linear, seek_absgp_pg,
 *                                end-of-stream truncation :219-222, absgp accounting :223-227,
 *                                chained streams :118-141)
 * Status codes are lewton_b200.h's LWB_* plus the LWF_* header errors below.
 */
#ifndef LEWTON_FRONTEND_H
#define LEWTON_FRONTEND_H

#include <stddef.h>
#include <stdint.h>

#include "lewton_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* header.rs:35-44 HeaderReadError, audio.rs:26-41 AudioReadError (values continue LWB_*) */
enum {
    LWF_ERR_END_OF_PACKET = 16,     /* HeaderReadError::EndOfPacket / AudioReadError::EndOfPacket   */
    LWF_ERR_NOT_VORBIS_HEADER = 17,
    LWF_ERR_UNSUPPORTED_VERSION = 18,
    LWF_ERR_HEADER_BAD_FORMAT = 19,
    LWF_ERR_HEADER_BAD_TYPE = 20,
    LWF_ERR_HEADER_IS_AUDIO = 21,
    LWF_ERR_UTF8 = 22,
    LWF_ERR_AUDIO_IS_HEADER = 23,   /* AudioReadError::AudioIsHeader                               */
    LWF_ERR_OGG = 24,               /* framing: capture pattern, CRC, lacing, truncated page       */
    LWF_ERR_NO_MORE_PACKETS = 25    /* read_packet() == None                                       */
};

/* ---- headers --------------------------------------------------------------------------------- */
typedef struct lwf_headers lwf_headers;     /* IdentHeader + CommentHeader + SetupHeader           */

typedef struct lwf_info {                   /* header.rs:188-211 IdentHeader + setup counts        */
    uint8_t audio_channels, blocksize_0, blocksize_1;
    uint32_t audio_sample_rate;
    int32_t bitrate_maximum, bitrate_nominal, bitrate_minimum;
    uint32_t n_codebooks, n_floors, n_residues, n_mappings, n_modes, n_comments;
} lwf_info;

int lwf_headers_parse(const uint8_t *ident, size_t ident_len, const uint8_t *comment, size_t comment_len,
                      const uint8_t *setup, size_t setup_len, lwf_headers **out);
void lwf_headers_destroy(lwf_headers *h);
int lwf_headers_info(const lwf_headers *h, lwf_info *out);
/* vendor string (index < 0) or "key=value" of comment `index`; returns the length, copies <= cap */
size_t lwf_headers_comment(const lwf_headers *h, int index, char *buf, size_t cap);
/* what lwb_setup_create needs, filled from the parsed headers (tables via lwb_tables_generate) */
int lwf_headers_make_setup(const lwf_headers *h, lwb_ctx *ctx, lwb_setup **out);

/* ---- one audio packet: front half of read_audio_packet_generic ------------------------------- */
typedef struct lwf_decoded_packet {
    uint8_t mode_number;
    uint8_t blockflag;
    uint8_t prev_window_flag, next_window_flag;    /* 1 for short blocks (map_or(true, ..))        */
    uint32_t n;                                     /* blocksize of this packet                     */
    /* per channel, caller-provided storage: */
    uint8_t *floor_kind;        /* [channels]                LWB_FLOOR_*                            */
    uint32_t *floor1_y;         /* [channels][LWB_MAX_POSTS]                                        */
    float *dense_floor;         /* [channels][n/2]: floor-0 curves (only rows with LWB_FLOOR_DENSE) */
    float *residue;             /* [channels][n/2]: residue vectors before inverse coupling         */
} lwf_decoded_packet;

/* Buffers in `out` must hold channels x blocksize_1/2 floats.  Returns LWB_OK, LWF_ERR_AUDIO_IS_HEADER,
 * LWF_ERR_END_OF_PACKET (header bits missing) or LWB_ERR_BAD_FORMAT (audio.rs:926-930, :975). */
int lwf_packet_decode(const lwf_headers *h, const uint8_t *packet, size_t len, lwf_decoded_packet *out);
int lwf_decoded_sample_count(const lwf_headers *h, const uint8_t *packet, size_t len, size_t *n_samples);
/* The same front half with the residue left as VQ runs + entries (audio.rs:587-717): what
 * residue_packet_decode would have ADDED, partition by partition in decode order, for LWB_ENTRY_VQ batches (out->residue
 * is not touched and may be NULL).  LWB_ERR_BUFFER if the capacities do not suffice (a packet of L bytes never needs
 * more than 8 L of either).  lwf_headers_vq_capable: 1 if the stream qualifies (<= 8 channels, channels * n/2 <= 12288,
 * VQ books of <= 65536 entries whose dimension divides their residue's partition size), else the dense path is used. */
int lwf_headers_vq_capable(const lwf_headers *h);
int lwf_packet_decode_vq(const lwf_headers *h, const uint8_t *packet, size_t len, lwf_decoded_packet *out,
                         lwb_vq_run *runs, size_t run_capacity, size_t *n_runs, uint16_t *entries, size_t entry_capacity,
                         size_t *n_entries);

/* ---- Ogg paging -------------------------------------------------------------------------------- */
typedef struct lwf_ogg lwf_ogg;             /* PacketReader over a memory buffer (not copied)      */
typedef struct lwf_ogg_packet {
    const uint8_t *data;                    /* valid until the next lwf_ogg_next_packet             */
    size_t len;
    uint32_t stream_serial;
    uint64_t absgp_page;
    uint8_t first_in_stream, last_in_stream, first_in_page, last_in_page;
} lwf_ogg_packet;
int lwf_ogg_open(const uint8_t *data, size_t len, lwf_ogg **out);
void lwf_ogg_close(lwf_ogg *o);
int lwf_ogg_next_packet(lwf_ogg *o, lwf_ogg_packet *pkt);    /* LWF_ERR_NO_MORE_PACKETS at the end  */

/* ---- OggStreamReader --------------------------------------------------------------------------- */
typedef struct lwf_reader lwf_reader;
/* Reads the three headers, builds the device-side setup on `ctx`, opens a PreviousWindowRight. */
int lwf_reader_open(lwb_ctx *ctx, const uint8_t *data, size_t len, lwf_reader **out);
void lwf_reader_close(lwf_reader *r);
const lwf_headers *lwf_reader_headers(const lwf_reader *r);
/* read_dec_packet_generic: decodes the next audio packet through lwb_decode_packet.  `out_format`
 * LWB_OUT_*; `out` holds capacity_total elements in all: planar channel c starts at
 * c * (capacity_total / channels), with the channel count of the stream the packet belongs to (a
 * chained stream may change it: query lwf_reader_headers afterwards).  *n_samples = samples per
 * channel after end-of-stream truncation.  LWF_ERR_NO_MORE_PACKETS = Ok(None). */
int lwf_reader_read_dec_packet(lwf_reader *r, int out_format, void *out, size_t capacity_total,
                               size_t *n_samples);
int lwf_reader_last_absgp(const lwf_reader *r, uint64_t *absgp);   /* returns 0 and sets *absgp if Some */
/* skip_samples_linear (inside_ogg.rs:244-283): walks packets by their sample counts only, decodes the packet before
 * the target on a fresh PreviousWindowRight (dropped) and returns the target packet.  *got_packet = 0 <=> Ok((None, _))
 * (the stream ended first); *left_to_skip = the second element of the reference's tuple. */
int lwf_reader_skip_samples_linear(lwf_reader *r, size_t to_skip, int out_format, void *out, size_t capacity_total,
                                   size_t *n_samples, size_t *left_to_skip, int *got_packet);
/* seek_absgp_pg (inside_ogg.rs:307-313): page-granular seek inside the current logical stream to a position <= absgp;
 * afterwards get_last_absgp() is None and the next packet returns 0 samples (fresh PreviousWindowRight). */
int lwf_reader_seek_absgp_pg(lwf_reader *r, uint64_t absgp);

/* ---- many streams at once: host entropy decode on a thread pool, one batched synthesis call ---- */
/* The shape of a decode server (BASELINE configs 1/3 at scale): packets of many logical streams that
 * share one set of headers are entropy-decoded in parallel on the host straight into pinned arenas
 * (residue vectors, floor posts), then synthesised by ONE lwb_decode_chains call (residue entry,
 * host memory).  Streams are independent; packets within a stream keep their order. */
typedef struct lwf_stream_job {
    lwb_stream *stream;               /* PreviousWindowRight of this logical stream                  */
    uint32_t n_packets;
    const uint8_t *const *packets;    /* [n_packets] audio packets in stream order                   */
    const size_t *lengths;            /* [n_packets]                                                 */
    uint64_t out_offset;              /* element offset of this stream's PCM in `pcm`                */
    uint64_t out_stride;              /* planar formats: elements between channel planes             */
    /* results */
    uint32_t n_samples;               /* samples per channel produced                                */
    uint32_t packets_done;            /* packets synthesised (== n_packets unless status != 0)       */
    int32_t status;                   /* LWB_OK, or the error of packet `packets_done`               */
} lwf_stream_job;

typedef struct lwf_batcher lwf_batcher;
/* `setup` must come from lwf_headers_make_setup(h, ctx); threads <= 0: one per host CPU */
int lwf_batcher_create(lwb_ctx *ctx, const lwf_headers *h, int threads, lwf_batcher **out);
void lwf_batcher_destroy(lwf_batcher *b);
/* LWB_ENTRY_RESIDUE (default: dense residue vectors cross the boundary) or LWB_ENTRY_VQ (VQ records do; needs
 * lwf_headers_vq_capable) */
int lwf_batcher_set_entry(lwf_batcher *b, int entry);
int lwf_batcher_decode(lwf_batcher *b, lwf_stream_job *jobs, size_t n_jobs, int out_format, void *pcm);
/* wall-clock seconds of the last lwf_batcher_decode: host entropy decode, synthesis call */
void lwf_batcher_last_timing(const lwf_batcher *b, double *entropy_seconds, double *synthesis_seconds);

/* ---- debug taps (known-answer tests of the reference's unit-test vectors) ---------------------- */
float lwf_debug_float32_unpack(uint32_t v);                          /* bitpacking.rs:304-314        */
uint32_t lwf_debug_lookup1_values(uint32_t entries, uint16_t dims);  /* header.rs:616-649            */
uint8_t lwf_debug_ilog(uint64_t v);                                  /* lib.rs:166-172               */
size_t lwf_debug_read_bits(const uint8_t *data, size_t len, const uint8_t *widths, size_t n, uint64_t *out);
int lwf_debug_huffman(const uint8_t *lengths, size_t n, const uint8_t *data, size_t len, uint32_t *out,
                      size_t max_out, size_t *n_out);                /* huffman_tree.rs:113-214      */
/* timing aid: the n packets decoded `reps` times inside one call (vq != 0: records instead of dense residues);
 * seconds, < 0 on error (profiles/frontend_bench.py) */
double lwf_debug_decode_loop(const lwf_headers *h, const uint8_t *const *packets, const size_t *lens, size_t n, int reps, int vq);

#ifdef __cplusplus
}
#endif
#endif
