// kernel_mid.cuh -- fused IMDCT + window + overlap-add for runs of blocks of n = 1024 (blocksize 10: the long block of
// 128/1024 and 256/1024 streams, and uniform 1024-point streams): the register-resident structure of kernel_long.cuh for
// the 8-bit index space.
//
// A 1024-point block has 256 complex values; a warp holds 512, so it transforms TWO blocks in lockstep -- two different
// runs of equal length, like k_short_g's positions.  The block index is made the TOP bit of kernel_long's 9-bit element
// index, c = 256 blk + c', which leaves every index map, both transposes and their swizzle exactly as they are:
//     phase A: slot = (blk, c'7, c'6)   step 0, step 2 (bit 7: slots j, j + 2), stage 0 (bit 6: slots j, j + 1)
//     phase B: slot = c'5 c'4 c'3       stages 1, 2, 3 -- kernel_long's phase_b, other twiddles
//     phase C: slot = c'2 c'1 c'0       ld654, renaming, step 7, step 8, window / OLA -- kernel_long's phase_c_fft
// with one change to the phase-C map: kernel_long's second half of a lane holds the element whose upper six bits are
// complemented (the step-7 partner, imdct.rs:533-580); complementing the block bit as well would pair the two blocks, so
// here only the five bits inside the block are complemented (elemC_m).  A lane then carries elements of block
// (lane & 1) only in phase C, its output samples are m = 32 rev3(slot) + (lane >> 1) and the mirror images, and the
// overlap state of "its" run stays in its registers from packet to packet.
//
// The lane functions compile for the host too: tests/emu/mid_emu.cpp runs the 32 lanes sequentially against the oracle.
#pragma once
#include "kernel_long.cuh"

namespace lwb {

constexpr int kMidBs = 10;
constexpr int kMidN = 1024;
constexpr int kMidN2 = 512;          // spectrum floats / PCM samples per block
constexpr int kMidC = 256;           // complex elements per block

// phase C: the element of (lane, slot, half): upper bits T = (blk, T'), half 1 complements T' only
LWB_HD int elemC_m(int lane, int slot, int half)
{
    const int T0 = rev6(lane);
    return 8 * (half ? (T0 ^ 31) : T0) + slot;
}
// block of a lane in phase C, and the output index m (0..255) of (lane, slot, half) AFTER the step-7 half swap
LWB_HD int blockC_m(int lane) { return lane & 1; }
LWB_HD int outIndex_m(int lane, int slot, int half)
{
    const int flip = (slot & 1) ? half : !half;
    return 32 * rev3(slot) + (flip ? 31 - (lane >> 1) : (lane >> 1));
}
LWB_HD int rev8(int c) { return (rev9(c & 255) >> 1); }          // bit reversal of 8 bits

// Host: per-lane pack from the blocksize-10 tables (a, b: 512; c: 256; w: 512).  Same slot layout as long_build_pack;
// P_S2W* uses entries 0, 1 (index = slot & 1), P_L0W* entry 0, P_L1W* is unused; P_L2 / P_L3 / P_L4 carry stages 1 / 2 / 3.
inline void mid_build_pack(const float *a, const float *b, const float *c, const float *w, float *pack)
{
    V *P = reinterpret_cast<V *>(pack);
    for (int i = 0; i < P_END * 32; i++) P[i] = V{0.f, 0.f};
    for (int lane = 0; lane < 32; lane++) {
        auto put = [&](int slot, float x, float y) { P[slot * 32 + lane] = V{x, y}; };
        float tx[2], ty[2];
        // phase A: c' = elemA & 255
        for (int j = 0; j < 8; j++) {                          // step 0 (imdct.rs:337-371)
            for (int h = 0; h < 2; h++) {
                const int cc = elemA(lane, j, h) & 255;
                const float s = cc < 128 ? -1.0f : 1.0f;       // (-x)*A == x*(-A)
                tx[h] = s * a[510 - 2 * cc];
                ty[h] = s * a[511 - 2 * cc];
            }
            put(P_S0W0 + j, tx[0], tx[1]);
            put(P_S0W1 + j, ty[0], ty[1]);
        }
        for (int u = 0; u < 2; u++) {                          // step 2 (imdct.rs:385-430): lower element c'7 = 0, slot u
            for (int h = 0; h < 2; h++) {
                const int cc = elemA(lane, u, h) & 255;
                tx[h] = a[508 - 4 * cc];
                ty[h] = a[509 - 4 * cc];
            }
            put(P_S2W0 + u, tx[0], tx[1]);
            put(P_S2W1 + u, ty[0], ty[1]);
        }
        for (int h = 0; h < 2; h++) {                          // stage 0: a = r * 8, r < n >> 4
            const int r = (~elemA(lane, 1, h)) & 63;
            tx[h] = a[8 * r];
            ty[h] = a[8 * r + 1];
        }
        put(P_L0W0, tx[0], tx[1]);
        put(P_L0W1, ty[0], ty[1]);
        // phase B (both halves share the twiddle: same low bits)
        for (int u = 0; u < 4; u++) {
            const int r = (~elemB(lane, 4 + u, 0)) & 31;       // stage 1: a = r * 16
            put(P_L2W0 + u, a[16 * r], a[16 * r]);
            put(P_L2W1 + u, a[16 * r + 1], a[16 * r + 1]);
        }
        for (int u = 0; u < 2; u++) {
            const int r = (~elemB(lane, 2 + u, 0)) & 15;       // stage 2: a = r * 32
            put(P_L3W0 + u, a[32 * r], a[32 * r]);
            put(P_L3W1 + u, a[32 * r + 1], a[32 * r + 1]);
        }
        {
            const int r = (~elemB(lane, 1, 0)) & 7;            // stage 3: a = r * 64
            put(P_L4W0, a[64 * r], a[64 * r]);
            put(P_L4W1, a[64 * r + 1], a[64 * r + 1]);
        }
        // phase C
        put(P_A2, a[kMidN >> 3], a[kMidN >> 3]);
        for (int jj = 0; jj < 4; jj++) {
            for (int h = 0; h < 2; h++) {
                const int p = 255 - rev8(elemC_m(lane, 2 * jj + 1, h) & 255);     // step-7 index of the D side
                tx[h] = c[2 * p];
                ty[h] = c[2 * p + 1];
            }
            put(P_S7C0 + jj, tx[0], tx[1]);
            put(P_S7C1 + jj, ty[0], ty[1]);
        }
        for (int j = 0; j < 8; j++) {
            float b0[2], b1[2], wl[2], wh[2];
            for (int h = 0; h < 2; h++) {
                const int m = outIndex_m(lane, j, h);
                const int cp = 255 - m;                        // V element feeding output m
                b0[h] = b[2 * cp];
                b1[h] = b[2 * cp + 1];
                wl[h] = w[m];
                wh[h] = w[511 - m];
            }
            put(P_B0 + j, b0[0], b0[1]);
            put(P_B1 + j, b1[0], b1[1]);
            put(P_WLO + j, wl[0], wl[1]);
            put(P_WHI + j, wh[0], wh[1]);
        }
    }
}

// Phase A.  tile[blk] = the block's 512 spectrum floats.  Quad #f (4 floats at 4f) yields element c' = f from (q1, q3) and
// c' = 255 - f from (q0, q2) (step 0, imdct.rs:337-371).  The lane reads quads #(lane + 64 m) and #(63 - lane + 64 m),
// m < 2, of either block: they feed slots m and 3 - m of the block's four.
template <class TW>
LWB_HD void phase_a_m(const float *const tile[2], int lane, TW tw, V O[8], V E[8])
{
#pragma unroll
    for (int blk = 0; blk < 2; blk++)
#pragma unroll
        for (int m = 0; m < 2; m++) {
            const Q4 f1 = ld_q4(tile[blk] + 4 * (lane + 64 * m));
            const Q4 f2 = ld_q4(tile[blk] + 4 * (63 - lane + 64 * m));
            {
                const int j = 4 * blk + m;
                const V w0 = tw(P_S0W0 + j), w1 = tw(P_S0W1 + j);
                const V qa = V{f1.w, f2.w}, qb = V{f1.y, f2.y};
                O[j] = vsub_p(vmul(qa, w0), vmul(qb, w1));
                E[j] = vadd_p(vmul(qa, w1), vmul(qb, w0));
            }
            {
                const int j = 4 * blk + 3 - m;
                const V w0 = tw(P_S0W0 + j), w1 = tw(P_S0W1 + j);
                const V qa = V{f2.x, f1.x}, qb = V{f2.z, f1.z};
                O[j] = vsub_p(vmul(qa, w0), vmul(qb, w1));
                E[j] = vadd_p(vmul(qa, w1), vmul(qb, w0));
            }
        }
    // step 2 (imdct.rs:385-430): bit 7 of c' = slot bit 1
#pragma unroll
    for (int blk = 0; blk < 2; blk++)
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int j = 4 * blk + u;
            bfly(O[j + 2], E[j + 2], O[j], E[j], tw(P_S2W0 + u), tw(P_S2W1 + u));
        }
    // stage 0 (imdct.rs:445-446): bit 6 = slot bit 0
    {
        const V w0 = tw(P_L0W0), w1 = tw(P_L0W1);
#pragma unroll
        for (int j = 1; j < 8; j += 2) bfly(O[j], E[j], O[j - 1], E[j - 1], w0, w1);
    }
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------
// device side: k_mid.  Descriptors: groups of two LongRun (48 B each; in_stride / out in 512-sample units of this
// blocksize, first_short / last_short unused) of equal n_packets -- the host pads an odd run with a dummy --, dealt to the
// warps statically (group g -> warp g mod W) like k_long_s: descriptors by cp.async kMidFetch groups ahead, the producer
// cursor kLongRing stages ahead of the consumer across group boundaries (a stage = the two runs' 2 KB tiles, which then
// serve as the E | O planes of the transposes), the two 2 KB state rows of a group with history requested as soon as the
// state tile is free.
// ---------------------------------------------------------------------------------------------
constexpr int kMidFetch = 3;
constexpr int kMidDescSlots = kMidFetch + kLongRing + 3;
constexpr int kMidTileBytes = kMidN2 * 4;                    // 2048
constexpr size_t kMidGroupBytes = 2 * sizeof(LongRun);       // 96
constexpr size_t kMidSmemBytes = 2048 + (size_t)kLongWarps * (kLongRing + 1) * kLongTileBytes + (size_t)kLongPackFloats * 4 +
                                 kLongWarps * (kLongRing + 2) * 8 + (size_t)kLongWarps * kMidDescSlots * kMidGroupBytes + 64;

// step 8 + window + overlap-add + stores of the lane's block, all 8 slots.  FIRST: packet 0 of the run (its previous right
// half comes from the state tile if has_prev, else nothing is emitted).  flags: bit0 has_prev, bit2 dummy.
template <bool FIRST, typename OutT>
__device__ __forceinline__ void out_stage_m(const TwMix &tw, int lane, const V O[8], const V E[8], V pe[8], uint32_t flags,
                                            OutT *out, const float *s_state)
{
    const int hl = lane >> 1;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int r32 = 32 * rev3(j);
        const bool nat = (j & 1);             // odd slots: half x -> hl, half y -> 31 - hl
        V plo = pe[j], phi = pe[j];
        bool emit = !(flags & 4u);
        if (FIRST) {
            emit = emit && (flags & 1u);
            if (flags & 1u) {                 // prev[m] and prev[511 - m] read separately: an imported state need not be symmetric
                const float *s_lo = s_state + hl, *s_hi = s_state + 31 - hl;
                const float ax = nat ? s_lo[r32] : s_hi[r32], ay = nat ? s_hi[r32] : s_lo[r32];
                const float bx = nat ? s_hi[480 - r32] : s_lo[480 - r32];
                const float by = nat ? s_lo[480 - r32] : s_hi[480 - r32];
                plo = V{ax, ay};
                phi = V{bx, by};
            }
        }
        V lo, hi, pev;
        step8_ola(tw(P_B0 + j), tw(P_B1 + j), tw(P_WLO + j), tw(P_WHI + j), O[j], E[j], plo, phi, lo, hi, pev);
        pe[j] = pev;
        if (emit) {
            OutT *o_lo = out + hl, *o_hi = out + 31 - hl;
            if (nat) {
                st_pcm(o_lo + r32, lo.x); st_pcm(o_hi + r32, lo.y);
                st_pcm(o_hi + 480 - r32, hi.x); st_pcm(o_lo + 480 - r32, hi.y);
            } else {
                st_pcm(o_hi + r32, lo.x); st_pcm(o_lo + r32, lo.y);
                st_pcm(o_lo + 480 - r32, hi.x); st_pcm(o_hi + 480 - r32, hi.y);
            }
        }
    }
}

template <typename OutT>
__global__ void __launch_bounds__(kLongWarps * 32, 1)
k_mid(const LongRun *__restrict__ runs, uint32_t n_groups, const float *__restrict__ pack)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int blk = lane & 1;                                  // the run of the group whose samples this lane ends up with
    const uint32_t raw_s = smem_u32(smem_raw);
    const uint32_t align_pad = (2048u - (raw_s & 2047u)) & 2047u;
    unsigned char *base = smem_raw + align_pad;
    constexpr size_t kTilesBytes = (size_t)kLongWarps * kLongRing * kLongTileBytes;
    constexpr size_t kStateBytes = (size_t)kLongWarps * kLongTileBytes;
    float *tiles = reinterpret_cast<float *>(base) + (size_t)warp * kLongRing * kLongN2;
    float *s_state = reinterpret_cast<float *>(base + kTilesBytes) + (size_t)warp * kLongN2;      // [2][512]
    V *s_pack = reinterpret_cast<V *>(base + kTilesBytes + kStateBytes);
    unsigned char *tail = base + kTilesBytes + kStateBytes + (size_t)kLongPackFloats * 4;
    LongRun *s_desc = reinterpret_cast<LongRun *>(tail) + warp * kMidDescSlots * 2;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail + (size_t)kLongWarps * kMidDescSlots * kMidGroupBytes) + warp * (kLongRing + 2);
    {
        const float4 *src = reinterpret_cast<const float4 *>(pack);
        float4 *dst = reinterpret_cast<float4 *>(s_pack);
        for (int i = threadIdx.x; i < kLongPackFloats / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (lane == 0) {
        for (int i = 0; i < kLongRing + 1; i++) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    V twR[kTwReg1 - kTwReg0 > 0 ? kTwReg1 - kTwReg0 : 1];
#pragma unroll
    for (int s = kTwReg0; s < kTwReg1; s++) twR[s - kTwReg0] = s_pack[s * 32 + lane];
    const TwMix tw{twR, s_pack + lane};

    const uint32_t tiles_s = smem_u32(tiles), bars_s = smem_u32(bars), desc_s = smem_u32(s_desc);
    const uint32_t bar_state = bars_s + 8 * kLongRing, state_s = smem_u32(s_state);
    const uint32_t lA0 = laneA(lane, 0), lA1 = laneA(lane, 1);
    const uint32_t lB = laneB(lane);
    const uint32_t lC0 = 4u * (uint32_t)swz(elemC_m(lane, 0, 0)), lC1 = 4u * (uint32_t)swz(elemC_m(lane, 0, 1));

    const uint32_t W = gridDim.x * kLongWarps, gw = blockIdx.x * kLongWarps + warp;
    if (gw >= n_groups) return;
    const uint4 *rq = reinterpret_cast<const uint4 *>(runs);
    constexpr uint32_t kQuads = (uint32_t)(kMidGroupBytes / 16);           // 6 quads per group: lanes 0..5 copy one each
    uint32_t f_grp = gw, f_slot = 0;
    auto fetch = [&]() {
        if ((uint32_t)lane < kQuads && f_grp < n_groups)
            cp_async16(desc_s + f_slot * (uint32_t)kMidGroupBytes + lane * 16, rq + (size_t)kQuads * f_grp + lane);
        cp_async_commit();
        f_grp += W;
        f_slot = (f_slot + 1 == (uint32_t)kMidDescSlots) ? 0 : f_slot + 1;
    };
#pragma unroll
    for (int i = 0; i <= kMidFetch; i++) fetch();
    cp_async_wait<kMidFetch>();
    __syncwarp();
    // ---- producer (warp-uniform cursor; lane 0 issues) ----
    uint32_t p_grp = gw, p_pkt = 0, p_slot = 0, p_stage = 0;
    const float *p_in0 = s_desc[0].in, *p_in1 = s_desc[1].in;
    uint32_t p_st0 = s_desc[0].in_stride, p_st1 = s_desc[1].in_stride, p_npk = s_desc[0].n_packets;
    auto produce = [&]() {
        if (lane == 0) {
            fence_proxy_async();          // the stage was written through the generic proxy (transposes) before
            const uint32_t bar = bars_s + 8 * p_stage, dst = tiles_s + p_stage * kLongTileBytes;
            mbar_expect_tx(bar, 2 * kMidTileBytes);
            tma_load_1d(dst, p_in0 + (size_t)p_pkt * p_st0, kMidTileBytes, bar);
            tma_load_1d(dst + kMidTileBytes, p_in1 + (size_t)p_pkt * p_st1, kMidTileBytes, bar);
        }
        p_stage = (p_stage + 1 == (uint32_t)kLongRing) ? 0 : p_stage + 1;
        if (++p_pkt >= p_npk) {
            p_grp += W;
            p_pkt = 0;
            p_slot = (p_slot + 1 == (uint32_t)kMidDescSlots) ? 0 : p_slot + 1;
            fetch();
            cp_async_wait<kMidFetch>();
            __syncwarp();
            if (p_grp < n_groups) {
                const LongRun *g = s_desc + 2 * p_slot;
                p_in0 = g[0].in; p_in1 = g[1].in; p_st0 = g[0].in_stride; p_st1 = g[1].in_stride; p_npk = g[0].n_packets;
            }
        }
    };
    for (int i = 0; i < kLongRing; i++)
        if (p_grp < n_groups) produce();

    // ---- state rows: st_grp = the group whose rows are in the tile or on their way (~0: the tile is free) ----
    uint32_t st_grp = ~0u;
    auto issue_state = [&](const float *row0, bool has0, const float *row1, bool has1, uint32_t grp) {
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar_state, ((has0 ? 1u : 0u) + (has1 ? 1u : 0u)) * (uint32_t)kMidTileBytes);
            if (has0) tma_load_1d(state_s, row0, kMidTileBytes, bar_state);
            if (has1) tma_load_1d(state_s + kMidTileBytes, row1, kMidTileBytes, bar_state);
        }
        st_grp = grp;
    };
    auto request_state = [&](uint32_t from_grp, uint32_t from_slot) {       // first group in [from_grp, p_grp] with history
        uint32_t g = from_grp, sl = from_slot;
        while (g < n_groups && g <= p_grp) {
            const LongRun *d = s_desc + 2 * sl;
            if (d[0].has_prev || d[1].has_prev) {
                issue_state(d[0].state, d[0].has_prev != 0, d[1].state, d[1].has_prev != 0, g);
                return;
            }
            g += W;
            sl = (sl + 1 == (uint32_t)kMidDescSlots) ? 0 : sl + 1;
        }
    };

    uint32_t phase_bits = 0, slot_i = 0, c_slot = 0;
    for (uint32_t c_grp = gw; c_grp < n_groups; c_grp += W) {
        const LongRun *g = s_desc + 2 * c_slot;
        const uint32_t npk = g[0].n_packets;
        const bool grp_state = g[0].has_prev || g[1].has_prev;
        const float *row0 = g[0].state, *row1 = g[1].state;
        const bool has0 = g[0].has_prev != 0, has1 = g[1].has_prev != 0;
        // this lane's run
        const LongRun &mr = g[blk];
        const uint32_t flags = (mr.has_prev ? 1u : 0u) | (mr.write_state ? 2u : 0u) | (mr.dummy ? 4u : 0u);
        OutT *out = static_cast<OutT *>(mr.out);
        float *state_g = mr.state;
        const uint32_t my_slot = c_slot;
        c_slot = (c_slot + 1 == (uint32_t)kMidDescSlots) ? 0 : c_slot + 1;
        if (st_grp == ~0u) request_state(c_grp, my_slot);
        V pe[8];
#pragma unroll
        for (int j = 0; j < 8; j++) pe[j] = V{0.f, 0.f};

        for (uint32_t p = 0; p < npk; p++) {
            const uint32_t stage_s = tiles_s + slot_i * kLongTileBytes;
            mbar_wait(bars_s + 8 * slot_i, (phase_bits >> slot_i) & 1u);
            phase_bits ^= 1u << slot_i;
            V O[1][8], E[1][8];
            {
                const float *tp[2];
                tp[0] = tiles + slot_i * kLongN2;
                tp[1] = tp[0] + kMidN2;
                phase_a_m(tp, lane, tw, O[0], E[0]);
            }
            __syncwarp();           // every lane has consumed its quads: the tiles become the scratch
            {
                const uint32_t a0 = stage_s + lA0, a1 = stage_s + lA1;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    sts_eo(a0 ^ LWB_KA(j), E[0][j].x, O[0][j].x);
                    sts_eo(a1 ^ LWB_KA(j), E[0][j].y, O[0][j].y);
                }
            }
            __syncwarp();
            {
                const uint32_t b0 = stage_s + lB;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    lds_eo(b0 ^ LWB_KB(j, 0), E[0][j].x, O[0][j].x);
                    lds_eo(b0 ^ LWB_KB(j, 1), E[0][j].y, O[0][j].y);
                }
            }
            __syncwarp();
            phase_b<1>(tw, O, E);
            {
                const uint32_t b0 = stage_s + lB;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    sts_eo(b0 ^ LWB_KB(j, 0), E[0][j].x, O[0][j].x);
                    sts_eo(b0 ^ LWB_KB(j, 1), E[0][j].y, O[0][j].y);
                }
            }
            __syncwarp();
            {
                const uint32_t c0 = stage_s + lC0, c1 = stage_s + lC1;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    lds_eo(c0 ^ LWB_KC(j), E[0][j].x, O[0][j].x);
                    lds_eo(c1 ^ LWB_KC(j), E[0][j].y, O[0][j].y);
                }
            }
            __syncwarp();
            if (p_grp < n_groups) produce();          // the stage is free again
            phase_c_fft<1>(tw, O, E);
            if (p > 0) {
                out_stage_m<false, OutT>(tw, lane, O[0], E[0], pe, flags, out, s_state + blk * kMidN2);
            } else {
                if (grp_state) {
                    if (st_grp != c_grp) issue_state(row0, has0, row1, has1, c_grp);
                    mbar_wait(bar_state, (phase_bits >> 30) & 1u);
                    phase_bits ^= 1u << 30;
                }
                out_stage_m<true, OutT>(tw, lane, O[0], E[0], pe, flags, out, s_state + blk * kMidN2);
                __syncwarp();
                if (grp_state) {                                            // state tile consumed: on to the next group that needs it
                    st_grp = ~0u;
                    request_state(c_grp + W, c_slot);
                }
            }
            if (p > 0 || (flags & 1u)) out += kMidN2;
            slot_i = (slot_i + 1 == (uint32_t)kLongRing) ? 0 : slot_i + 1;
        }
        if ((flags & 6u) == 2u) {             // write_state and not dummy: the lane's 16 values of its run's right half, twice
            const int hl = lane >> 1;
            float *s_lo = state_g + hl, *s_hi = state_g + 31 - hl;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int r32 = 32 * rev3(j);
                const float vx = (j & 1) ? pe[j].x : pe[j].y, vy = (j & 1) ? pe[j].y : pe[j].x;
                s_lo[r32] = vx; s_hi[r32] = vy;                 // state[m]
                s_hi[480 - r32] = vx; s_lo[480 - r32] = vy;     // state[511 - m]: same value (imdct.rs:622-649)
            }
        }
    }
}

inline void mid_kernel_configure()
{
    cudaFuncSetAttribute(k_mid<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMidSmemBytes);
    cudaFuncSetAttribute(k_mid<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMidSmemBytes);
}

inline int mid_launch(cudaStream_t stream, const LongRun *d_runs, uint32_t n_groups, const float *d_pack, int sm_count, bool i16_out)
{
    if (!n_groups) return 0;
    const uint32_t want = (n_groups + kLongWarps - 1) / kLongWarps;
    const uint32_t grid = want < (uint32_t)sm_count ? want : (uint32_t)sm_count;
    if (i16_out) k_mid<int16_t><<<grid, kLongWarps * 32, kMidSmemBytes, stream>>>(d_runs, n_groups, d_pack);
    else k_mid<float><<<grid, kLongWarps * 32, kMidSmemBytes, stream>>>(d_runs, n_groups, d_pack);
    return cudaGetLastError() != cudaSuccess;
}
#endif  // __CUDACC__

}  // namespace lwb
