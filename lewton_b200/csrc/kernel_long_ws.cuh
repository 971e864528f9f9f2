// kernel_long_ws.cuh -- warp-specialised version of the fused long-block kernel.
//
// Same arithmetic, same element maps, same twiddle pack as k_long (kernel_long.cuh); what changes
// is who does what.  k_long is bound by per-warp latency: a warp needs ~240 registers (data +
// resident twiddles), so only 8 fit on an SM, 2 per scheduler, and their dependent phases
// (wait tile -> phase A -> transpose -> phase B -> transpose -> phase C -> overlap-add) leave the
// schedulers idle 45 % of the time although neither HBM, nor the FP pipe, nor shared memory is
// saturated (profiles/r1h_*).  Here every block flows through a PAIR of warps:
//
//   front warp F : wait TMA tile -> phase A -> transpose 1 -> phase B -> transpose-2 stores
//   back  warp K : transpose-2 loads -> phase C (ld654, step 7) -> step 8 -> window/OLA -> stores
//
// The hand-off is the second transpose itself (F writes the scratch, K reads it), so no extra data
// moves.  Each role needs only its own twiddles (F: phases A/B, K: phase C) and K alone carries the
// overlap state, which halves the registers per warp: 16 warps per SM (4 per scheduler), twice the
// independent instruction streams, and a two-stage pipeline per pair.
//
// Synchronisation: per ring slot three mbarriers (count 1): full_in (TMA landed, F waits),
// full_t2 (F's transpose-2 stores done, K waits), empty (K's loads done, F's lane 0 waits before it
// refills the slot by TMA).  F's lane 0 is the only thread that knows about runs: it draws tickets,
// fetches descriptors (TMA), issues tiles in processing order across run boundaries and writes a
// small per-tile record (`WsMeta`: kind, flags, where the PCM goes) that both warps read.  The
// stream state a run overlaps with travels through the ring as the run's first tile (kind STATE);
// the end of work as a tile of kind END.
#pragma once
#include "kernel_long.cuh"

#if defined(__CUDACC__)
namespace lwb {

#ifndef LWB_WS_PAIRS
#define LWB_WS_PAIRS 8
#endif
#ifndef LWB_WS_RING
#define LWB_WS_RING 6
#endif
#ifndef LWB_WS_AHEAD
#define LWB_WS_AHEAD 3          // tiles issued ahead of the one F is processing (<= ring - 2)
#endif
constexpr int kWsPairs = LWB_WS_PAIRS;
constexpr int kWsRing = LWB_WS_RING;
constexpr int kWsAhead = LWB_WS_AHEAD;
static_assert(kWsAhead <= kWsRing - 2, "leave at least one slot of slack between the two warps");

struct alignas(16) WsMeta {
    void *out;                  // this packet's PCM (when EMIT is set)
    float *state;               // the run's state row (for LAST_WSTATE)
    uint32_t kind;              // 0 packet, 1 state row, 2 end of work
    uint32_t flags;
    uint32_t pad[2];
};
enum { WS_PACKET = 0, WS_STATE = 1, WS_END = 2 };
enum { WS_FIRST = 1, WS_HAS_PREV = 2, WS_EMIT = 4, WS_LAST_WSTATE = 8 };

constexpr size_t kWsPairBytes = (size_t)kWsRing * kLongTileBytes;
// [tiles: pairs x ring x 4 KB, 2 KB aligned][pack][per pair: meta[ring], descriptor, barriers[3*ring+1]]
constexpr size_t kWsTailPerPair =
    ((kWsRing * sizeof(WsMeta) + sizeof(LongRun) + (3 * kWsRing + 1) * 8 + 15) / 16) * 16;    // keeps every pair's block 16-aligned
constexpr size_t kWsSmemBytes = 2048 + kWsPairs * kWsPairBytes + (size_t)kLongPackFloats * 4 + kWsPairs * kWsTailPerPair + 64;

// registers: F keeps the phase A/B twiddles [kWsF0, kWsF1), K the phase C ones [kWsK0, kWsK1)
#ifndef LWB_WS_F0
#define LWB_WS_F0 16            // step-0 pairs (slots 0..15) are fetched per block
#endif
#ifndef LWB_WS_F1
#define LWB_WS_F1 44
#endif
#ifndef LWB_WS_K0
#define LWB_WS_K0 44
#endif
#ifndef LWB_WS_K1
#define LWB_WS_K1 69            // window pairs (slots 69..84) are fetched per block
#endif
template <int R0, int R1>
struct TwRange {
    const V *r;
    const V *lane_base;
    __device__ __forceinline__ V operator()(int slot) const
    {
        return (slot >= R0 && slot < R1) ? r[slot - R0] : lane_base[slot * 32];
    }
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <typename OutT>
__global__ void __launch_bounds__(kWsPairs * 64, 1)
k_long_ws(const LongRun *__restrict__ runs, uint32_t n_runs, const float *__restrict__ pack,
          unsigned int *__restrict__ ticket)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = warp % kWsPairs;
    const bool is_front = warp < kWsPairs;          // warps 0..P-1 front, P..2P-1 back: both roles on every scheduler
    const uint32_t raw_s = smem_u32(smem_raw);
    const uint32_t align_pad = (2048u - (raw_s & 2047u)) & 2047u;
    unsigned char *base = smem_raw + align_pad;
    float *tiles = reinterpret_cast<float *>(base + (size_t)pair * kWsPairBytes);
    V *s_pack = reinterpret_cast<V *>(base + kWsPairs * kWsPairBytes);
    unsigned char *tail = base + kWsPairs * kWsPairBytes + (size_t)kLongPackFloats * 4 + (size_t)pair * kWsTailPerPair;
    WsMeta *meta = reinterpret_cast<WsMeta *>(tail);
    LongRun *s_desc = reinterpret_cast<LongRun *>(tail + kWsRing * sizeof(WsMeta));
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail + kWsRing * sizeof(WsMeta) + sizeof(LongRun));
    if (n_runs == 0) return;

    {
        const float4 *src = reinterpret_cast<const float4 *>(pack);
        float4 *dst = reinterpret_cast<float4 *>(s_pack);
        for (int i = threadIdx.x; i < kLongPackFloats / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (is_front && lane == 0) {
        for (int i = 0; i < 3 * kWsRing + 1; i++) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t tiles_s = smem_u32(tiles);
    const uint32_t bars_s = smem_u32(bars);
    auto bar_in = [&](uint32_t s) { return bars_s + 8u * s; };
    auto bar_t2 = [&](uint32_t s) { return bars_s + 8u * (kWsRing + s); };
    auto bar_empty = [&](uint32_t s) { return bars_s + 8u * (2 * kWsRing + s); };
    const uint32_t bar_desc = bars_s + 8u * (3 * kWsRing);
    const V *pk = s_pack + lane;

    if (is_front) {
        // =========================================== front ===========================================
        V twR[LWB_WS_F1 - LWB_WS_F0];
#pragma unroll
        for (int s = LWB_WS_F0; s < LWB_WS_F1; s++) twR[s - LWB_WS_F0] = pk[s * 32];
        const TwRange<LWB_WS_F0, LWB_WS_F1> tw{twR, pk};
        const uint32_t lA0 = laneA(lane, 0), lA1 = laneA(lane, 1), lB = laneB(lane);

        // ---- lane 0: the issue cursor ----------------------------------------------------------------
        const float *ir_in = nullptr;
        char *ir_out = nullptr;
        float *ir_state = nullptr;
        uint32_t ir_stride = 0, ir_npk = 0, ir_hp = 0, ir_ws = 0, ir_dummy = 0, ir_ti = 0, ir_tiles = 0;
        uint32_t nx_idx = 0, nx_stage = 3, desc_parity = 0;      // 0 ticket drawn, 1 descriptor in flight, 3 none
        uint32_t t_issue = 0;
        bool ended = false;
        auto load_run = [&](const LongRun &r) {
            ir_in = r.in; ir_out = static_cast<char *>(r.out); ir_state = r.state;
            ir_stride = r.in_stride; ir_npk = r.n_packets; ir_hp = r.has_prev ? 1u : 0u;
            ir_ws = r.write_state ? 1u : 0u; ir_dummy = r.dummy ? 1u : 0u;
            ir_ti = 0; ir_tiles = ir_hp + ir_npk;
        };
        auto issue_next = [&]() {                   // lane 0 only: put the next tile of the sequence into its slot
            const uint32_t s = t_issue % kWsRing;
            if (t_issue >= (uint32_t)kWsRing) mbar_wait(bar_empty(s), ((t_issue / kWsRing) - 1u) & 1u);
            if (ir_ti >= ir_tiles) {                // current run fully issued: move to the next one
                if (nx_stage == 0) {
                    if (nx_idx < n_runs) {
                        fence_proxy_async();
                        mbar_expect_tx(bar_desc, (uint32_t)sizeof(LongRun));
                        tma_load_1d(smem_u32(s_desc), runs + nx_idx, (uint32_t)sizeof(LongRun), bar_desc);
                        nx_stage = 1;
                    } else {
                        nx_stage = 3;
                    }
                }
                if (nx_stage == 1) {
                    mbar_wait(bar_desc, desc_parity);
                    desc_parity ^= 1u;
                    load_run(*s_desc);
                    nx_idx = atomicAdd(ticket, 1u);         // ticket for the run after this one
                    nx_stage = 0;
                } else {                                     // no more runs: END tile
                    meta[s].kind = WS_END;
                    meta[s].flags = 0;
                    mbar_arrive(bar_in(s));
                    ended = true;
                    t_issue++;
                    return;
                }
            } else if (nx_stage == 0 && ir_ti >= 2 && nx_idx < n_runs) {
                // the ticket drawn two tiles ago has arrived: fetch that run's descriptor in the background
                fence_proxy_async();
                mbar_expect_tx(bar_desc, (uint32_t)sizeof(LongRun));
                tma_load_1d(smem_u32(s_desc), runs + nx_idx, (uint32_t)sizeof(LongRun), bar_desc);
                nx_stage = 1;
            }
            WsMeta m;
            const float *src;
            if (ir_hp && ir_ti == 0) {
                m.kind = WS_STATE; m.flags = 0; m.out = nullptr; m.state = ir_state;
                src = ir_state;
            } else {
                const uint32_t p = ir_ti - ir_hp;
                const bool first = p == 0, emit = !ir_dummy && (p > 0 || ir_hp);
                m.kind = WS_PACKET;
                m.flags = (first ? WS_FIRST : 0u) | ((first && ir_hp) ? WS_HAS_PREV : 0u) | (emit ? WS_EMIT : 0u) |
                          ((p + 1 == ir_npk && ir_ws && !ir_dummy) ? WS_LAST_WSTATE : 0u);
                m.out = ir_out;
                m.state = ir_state;
                if (p > 0 || ir_hp) ir_out += kLongN2 * sizeof(OutT);
                src = ir_in + (size_t)p * ir_stride;
            }
            m.pad[0] = m.pad[1] = 0;
            meta[s] = m;
            fence_proxy_async();
            mbar_expect_tx(bar_in(s), kLongTileBytes);      // (release: publishes meta[s] with the barrier)
            tma_load_1d(tiles_s + s * kLongTileBytes, src, kLongTileBytes, bar_in(s));
            ir_ti++;
            t_issue++;
        };

        if (lane == 0) {
            const uint32_t idx = atomicAdd(ticket, 1u);
            if (idx < n_runs) {
                load_run(runs[idx]);
                nx_idx = atomicAdd(ticket, 1u);
                nx_stage = 0;
            } else {
                ir_ti = ir_tiles = 0;
                nx_stage = 3;
            }
        }

        for (uint32_t t = 0;; t++) {
            if (lane == 0)
                while (!ended && t_issue <= t + kWsAhead) issue_next();
            __syncwarp();
            const uint32_t s = t % kWsRing;
            mbar_wait(bar_in(s), (t / kWsRing) & 1u);
            const uint32_t kind = meta[s].kind;
            if (kind != WS_PACKET) {                 // state rows and the end marker pass straight through
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_t2(s));
                if (kind == WS_END) break;
                continue;
            }
            V O[1][8], E[1][8];
            {
                const float *tp[1] = {tiles + s * kLongN2};
                phase_a<1>(tp, lane, tw, O, E);
            }
            __syncwarp();
            const uint32_t tile_s = tiles_s + s * kLongTileBytes;
            const uint32_t a0 = tile_s + lA0, a1 = tile_s + lA1, b0 = tile_s + lB;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                sts_eo(a0 ^ LWB_KA(j), E[0][j].x, O[0][j].x);
                sts_eo(a1 ^ LWB_KA(j), E[0][j].y, O[0][j].y);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j++) {
                lds_eo(b0 ^ LWB_KB(j, 0), E[0][j].x, O[0][j].x);
                lds_eo(b0 ^ LWB_KB(j, 1), E[0][j].y, O[0][j].y);
            }
            __syncwarp();
            phase_b<1>(tw, O, E);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                sts_eo(b0 ^ LWB_KB(j, 0), E[0][j].x, O[0][j].x);
                sts_eo(b0 ^ LWB_KB(j, 1), E[0][j].y, O[0][j].y);
            }
            __syncwarp();                            // every lane's stores precede lane 0's release
            if (lane == 0) mbar_arrive(bar_t2(s));
        }
    } else {
        // =========================================== back ============================================
        V twR[LWB_WS_K1 - LWB_WS_K0];
#pragma unroll
        for (int s = LWB_WS_K0; s < LWB_WS_K1; s++) twR[s - LWB_WS_K0] = pk[s * 32];
        const TwRange<LWB_WS_K0, LWB_WS_K1> tw{twR, pk};
        const uint32_t lC0 = laneC(lane, 0), lC1 = laneC(lane, 1);
        V pe[8];
#pragma unroll
        for (int j = 0; j < 8; j++) pe[j] = V{0.f, 0.f};
        uint32_t st_slot = 0;

        for (uint32_t t = 0;; t++) {
            const uint32_t s = t % kWsRing;
            mbar_wait(bar_t2(s), (t / kWsRing) & 1u);
            const WsMeta m = meta[s];
            if (m.kind == WS_END) break;
            if (m.kind == WS_STATE) {                // keep the slot until the run's first overlap-add has read it
                st_slot = s;
                continue;
            }
            V O[1][8], E[1][8];
            {
                const uint32_t tile_s = tiles_s + s * kLongTileBytes;
                const uint32_t c0 = tile_s + lC0, c1 = tile_s + lC1;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    lds_eo(c0 ^ LWB_KC(j), E[0][j].x, O[0][j].x);
                    lds_eo(c1 ^ LWB_KC(j), E[0][j].y, O[0][j].y);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty(s));           // the slot may be refilled
            phase_c_fft<1>(tw, O, E);
            const bool first = m.flags & WS_FIRST, has_prev = m.flags & WS_HAS_PREV, emit = m.flags & WS_EMIT;
            OutT *out = static_cast<OutT *>(m.out);
            const float *st_tile = tiles + st_slot * kLongN2;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int r64 = 64 * rev3(j);
                const bool nat = (j & 1);
                V plo = pe[j], phi = pe[j];
                if (has_prev) {
                    const float *s_lo = st_tile + lane, *s_hi = st_tile + 63 - lane;
                    const float ax = nat ? s_lo[r64] : s_hi[r64], ay = nat ? s_hi[r64] : s_lo[r64];
                    const float bx = nat ? s_hi[960 - r64] : s_lo[960 - r64];
                    const float by = nat ? s_lo[960 - r64] : s_hi[960 - r64];
                    plo = V{ax, ay};
                    phi = V{bx, by};
                }
                V lo, hi, pev;
                step8_ola(tw(P_B0 + j), tw(P_B1 + j), tw(P_WLO + j), tw(P_WHI + j), O[0][j], E[0][j], plo, phi, lo, hi, pev);
                pe[j] = pev;
                if (emit) {
                    OutT *o_lo = out + lane, *o_hi = out + 63 - lane;
                    if (nat) {
                        st_pcm(o_lo + r64, lo.x); st_pcm(o_hi + r64, lo.y);
                        st_pcm(o_hi + 960 - r64, hi.x); st_pcm(o_lo + 960 - r64, hi.y);
                    } else {
                        st_pcm(o_hi + r64, lo.x); st_pcm(o_lo + r64, lo.y);
                        st_pcm(o_lo + 960 - r64, hi.x); st_pcm(o_hi + 960 - r64, hi.y);
                    }
                }
            }
            (void)first;
            if (has_prev) {
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty(st_slot));  // state row consumed
            }
            if (m.flags & WS_LAST_WSTATE) {
                float *s_lo = m.state + lane, *s_hi = m.state + 63 - lane;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int r64 = 64 * rev3(j);
                    const float vx = (j & 1) ? pe[j].x : pe[j].y, vy = (j & 1) ? pe[j].y : pe[j].x;
                    s_lo[r64] = vx; s_hi[r64] = vy;
                    s_hi[960 - r64] = vx; s_lo[960 - r64] = vy;
                }
            }
        }
    }
}

inline void long_ws_configure()
{
    cudaFuncSetAttribute(k_long_ws<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsSmemBytes);
    cudaFuncSetAttribute(k_long_ws<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsSmemBytes);
}

inline int long_ws_launch(cudaStream_t stream, const LongRun *d_runs, uint32_t n_runs, const float *d_pack,
                          unsigned int *ticket, int sm_count, bool i16_out)
{
    const uint32_t want = (n_runs + kWsPairs - 1) / kWsPairs;
    const uint32_t grid = want < (uint32_t)sm_count ? want : (uint32_t)sm_count;
    if (i16_out) k_long_ws<int16_t><<<grid, kWsPairs * 64, kWsSmemBytes, stream>>>(d_runs, n_runs, d_pack, ticket);
    else k_long_ws<float><<<grid, kWsPairs * 64, kWsSmemBytes, stream>>>(d_runs, n_runs, d_pack, ticket);
    return cudaGetLastError() != cudaSuccess;
}

}  // namespace lwb
#endif  // __CUDACC__
