// lwb_common.h -- structures shared by the host side and the kernels.
//
// HBM layout (all device-resident, owned by the ctx / setup / stream objects):
//   * per setup:   DevSetup (tables for both blocksizes, floor-1 constants, mappings, modes)
//   * per stream:  state[channels][blocksize_1/2] f32 = PreviousWindowRight (audio.rs:847-861),
//                  the un-windowed right part of the previous block; (has, len) tracked on host
//   * per batch:   coeff arena  [packet][channel][n/2] f32   (spectrum or residue)
//                  pcm arena    planar [chain][channel][stride] or interleaved [chain][t][channel]
//                  DevPacket[]  one descriptor per packet (geometry resolved on the host)
#pragma once
#include <stdint.h>

#include "../../include/lewton_b200.h"

namespace lwb {

struct DevTables {            // CachedBlocksizeDerived, header_cached.rs:27-31 (device pointers)
    const float *a, *b, *c, *window;
    const uint32_t *bitrev;
    // fast-path twiddle pack for this blocksize (see kernel_long.cuh), or nullptr
    const float *pack;
    int bs;
    int pad;
};

struct DevFloor1 {            // FloorTypeOne (header.rs:415-424), synthesis-relevant fields only
    uint8_t type;             // LWB_FLOOR_TYPE_*
    uint8_t mult;             // floor1_multiplier
    uint8_t nposts;           // floor1_x_list.len()
    uint8_t pad;
    uint16_t x[LWB_MAX_POSTS];       // floor1_x_list (values <= 1<<15)
    uint8_t sorted[LWB_MAX_POSTS];   // floor1_x_list_sorted[i].0
    uint8_t lo[LWB_MAX_POSTS];       // low_neighbor(x_list, i).0   (audio.rs:285-287), i >= 2
    uint8_t hi[LWB_MAX_POSTS];       // high_neighbor(x_list, i).0  (audio.rs:290-292), i >= 2
};

struct DevMapping {           // Mapping (header.rs:384-390) with mux/submap_floors folded
    uint16_t n_coupling;
    uint16_t pad;
    uint8_t mag[LWB_MAX_COUPLING];
    uint8_t ang[LWB_MAX_COUPLING];
    uint8_t floor_of_channel[LWB_MAX_CHANNELS + 1];   // submap_floors[mux[ch]]
    // channels of every submap in ascending order (the order residue type 2 interleaves them in, audio.rs:957-986);
    // filled for setups with <= 8 channels (LWB_ENTRY_VQ)
    uint8_t sub_nch[LWB_MAX_SUBMAPS];
    uint8_t sub_ch[LWB_MAX_SUBMAPS][8];
};

struct DevBook {              // Codebook (header.rs:360-368): the value table of LWB_ENTRY_VQ
    const float *vq;          // [entries][dims], or nullptr
    uint32_t entries;
    uint16_t dims;
    uint16_t pad;
};
constexpr int kMaxResidues = 64;      // header.rs:973 residue count = read_u6 + 1

struct DevSetup {
    DevTables tab[2];
    const DevFloor1 *floors;
    const DevMapping *mappings;
    uint8_t channels, bs0, bs1, n_floors;
    uint8_t mode_blockflag[LWB_MAX_MODES];
    uint8_t mode_mapping[LWB_MAX_MODES];
    // LWB_ENTRY_VQ
    const DevBook *books;
    uint32_t n_books, n_residues;
    uint32_t res_psize[kMaxResidues];  // residue_partition_size
};

// One packet of a batch; every index/geometry decision is made on the host
// (audio.rs:1056-1073 window geometry, :1083-1154 which branch of the OLA block runs).
struct DevPacket {
    const DevSetup *setup;
    float *state;             // stream state [channels][state_stride]
    uint64_t coeff_off;       // element offset of [channels][n/2] in the coeff arena
    uint64_t x_off;           // element offset of [channels][n] in the IMDCT scratch (generic path)
    uint64_t out_off;         // element offset of this packet's first sample (channel 0) in the pcm arena
    uint64_t out_stride;      // planar: elements between channel planes
    uint64_t pkt_index;       // row in the per-packet floor arenas
    int32_t prev_packet;      // previous packet of the same chain in this launch, or -1: use `state`
    uint32_t state_stride;    // blocksize_1 / 2
    uint16_t n;               // blocksize of this packet
    uint16_t ls, rs, re;      // left_win_start, right_win_start, right_win_end
    uint16_t plen;            // length of the previous right half; 0 = no previous (no output)
    uint16_t prev_rs;         // right_win_start of prev_packet (where its saved half begins)
    uint8_t blockflag;
    uint8_t mapping;
    uint8_t slope_sel;        // which blocksize's window_slope the left window uses
    uint8_t channels;
    uint8_t save_state;       // last packet of its chain in this launch: write x[rs..re) to state
    uint8_t pad[3];
};

}  // namespace lwb
