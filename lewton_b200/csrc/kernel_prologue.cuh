// kernel_prologue.cuh -- the residue entry's front stages for batches (audio.rs:991-1039), as two kernels that
// replace the per-packet-CTA k_prologue of kernels_generic.cuh wherever a batch has <= 8 channels:
//
//   k_floor1_segments : floor-1 step 1 (post unwrap, audio.rs:391-435) with one THREAD per (packet, channel) row --
//                       the unwrap is serial over <= 65 posts, so 32 independent rows per warp is the only way to
//                       keep its lanes busy -- then every flagged segment of the row packed for per-bin evaluation
//                       (floor1_eval.cuh: Seg4, division-free closed form of render_line, audio.rs:503-524);
//   k_prologue_fused  : one CTA per packet, one thread per 4 bins, all channels: the rows' segment tables and a
//                       bin -> segment bitmap in shared memory, floor value per bin (closed form -> dB table,
//                       audio.rs:552-554; unused floor = zero curve, :1021-1024; dense = host-computed floor-0),
//                       inverse coupling in registers / shared memory (steps in reverse, audio.rs:991-1002),
//                       multiply (:1035-1037), float4 loads and stores.  The kernel is HBM-bound (4 B in + 4 B out per
//                       coefficient); the per-bin floor arithmetic rides in its idle issue slots.
//
// History (profiles/r2*): the per-packet-CTA kernel took 5x its roofline time (serial unwrap on one lane per warp with
// the CTA waiting, 8-way predicated register arrays).  A first split rendered the curve to a byte arena in a separate
// kernel (16 bins per work item): its per-bin segment-crossing branches diverged (20 of 32 lanes active, 173 M
// warp-instructions for 134 M bins, 0.27 ms) and the curve cost 2 B per coefficient of extra traffic.
#pragma once
#include "kernels_generic.cuh"

namespace lwb {

constexpr int kSegRows = 64;           // (packet, channel) rows per CTA of k_floor1_segments
constexpr int kSegThreads = 256;
constexpr int kSegStride = LWB_MAX_POSTS + 3;   // 68 Seg4 per row: <= 66 segments (65 posts + flat tail) + sentinel

// grid = ceil(n_pk * C / kSegRows).  Row r = (packet ordinal in pkts) * C + channel.  seg_cnt[r] = segments of the row
// (0 when its floor kind is not LWB_FLOOR_ONE); segtab[r * kSegStride + j], j <= count (the last one a sentinel).
__global__ void __launch_bounds__(kSegThreads)
k_floor1_segments(const DevPacket *__restrict__ pkts, uint32_t n_rows, int C, const uint8_t *__restrict__ floor_kind,
                  const uint32_t *__restrict__ floor1_y, uint4 *__restrict__ segtab, uint8_t *__restrict__ seg_cnt)
{
    __shared__ uint16_t s_x[kSegRows][kSegStride];
    __shared__ uint16_t s_y[kSegRows][kSegStride];
    __shared__ int s_m[kSegRows];
    const int tid = threadIdx.x;
    const uint32_t row0 = blockIdx.x * kSegRows;
    if (tid < kSegRows) {
        const uint32_t row = row0 + tid;
        int m = 0;
        if (row < n_rows) {
            const uint32_t pk = row / (uint32_t)C, c = row - pk * (uint32_t)C;
            const DevPacket &p = pkts[pk];
            const uint64_t frow = p.pkt_index * (uint64_t)C + c;
            if (floor_kind[frow] == LWB_FLOOR_ONE) {
                const DevSetup &su = *p.setup;
                const DevFloor1 &fl = su.floors[su.mappings[p.mapping].floor_of_channel[c]];
                m = d_floor1_posts(fl, floor1_y + frow * LWB_MAX_POSTS, p.n >> 1, s_x[tid], s_y[tid]);
            }
            seg_cnt[row] = (uint8_t)(m > 1 ? m - 1 : 0);
        }
        s_m[tid] = m;
    }
    __syncthreads();
    for (int i = tid; i < kSegRows * kSegStride; i += kSegThreads) {
        const int r = i / kSegStride, j = i - r * kSegStride;
        const int nseg = s_m[r] - 1;
        if (j <= nseg && nseg > 0) {                    // j == nseg: sentinel = a copy of the last segment
            const Seg4 sg = d_floor1_pack_segment(s_x[r], s_y[r], j < nseg ? j : nseg - 1);
            segtab[(size_t)(row0 + r) * kSegStride + j] = make_uint4(sg.x, sg.y, sg.z, sg.w);
        }
    }
}

constexpr int kPfThreads = 256;
constexpr int kPfMaxWords = 128;       // bitmap words per channel: n/2 <= 4096 bins
// dynamic shared memory: segment tables, bitmaps + prefix counts, and (more than one channel) the coupling staging
inline size_t prologue_fused_smem(int channels)
{
    return (size_t)channels * (kSegStride * sizeof(uint4) + 2 * kPfMaxWords * sizeof(uint32_t)) +
           (channels > 1 ? (size_t)channels * kPfThreads * sizeof(float4) : 0);
}

// floor values of the 4 bins [k0, k0 + 4) of one channel
__device__ __forceinline__ float4 d_floor_quad(int kind, const float *__restrict__ s_db, const uint4 *__restrict__ tab,
                                               const uint32_t *__restrict__ bm, const uint32_t *__restrict__ pre, int k0,
                                               const float *__restrict__ dense, uint64_t e)
{
    if (kind == LWB_FLOOR_ONE) {
        // segment of bin k0 = (number of flagged posts with x <= k0) - 1; the post at x = 0 is always flagged
        int seg = (int)pre[k0 >> 5] + __popc(bm[k0 >> 5] & (0xffffffffu >> (31 - (k0 & 31)))) - 1;
        float f[4];
        uint4 P = tab[seg];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            if (b) {
                seg += (k0 + b >= (int)(P.y >> 16));            // branch-free: reload (the same entry, mostly)
                P = tab[seg];
            }
            const Seg4 sg{P.x, P.y, P.z, P.w};
            f[b] = s_db[d_floor1_seg_y(sg, k0 + b) & 255u];
        }
        return make_float4(f[0], f[1], f[2], f[3]);
    }
    if (kind == LWB_FLOOR_DENSE) return *reinterpret_cast<const float4 *>(dense + e);
    return make_float4(0.f, 0.f, 0.f, 0.f);                       // audio.rs:1021-1024
}

// Persistent CTAs striding over the packets (grid = min(packets, a few CTAs per SM)).  Requires every coeff_off
// (and the arena bases) to be multiples of 4 elements, <= 8 channels, a uniform channel count C.
__global__ void __launch_bounds__(kPfThreads)
k_prologue_fused(const DevPacket *__restrict__ pkts, uint32_t n_pk, const float *__restrict__ residue, const float *__restrict__ dense_floor,
                 const uint8_t *__restrict__ floor_kind, const uint4 *__restrict__ segtab, const uint8_t *__restrict__ seg_cnt,
                 float *__restrict__ spec)
{
    extern __shared__ __align__(16) unsigned char pf_smem[];
    __shared__ float s_db[256];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    s_db[tid] = c_inverse_db[tid];
    if (blockIdx.x >= n_pk) return;
    const int C = pkts[blockIdx.x].channels;            // uniform over the batch
    uint4 *s_tab = reinterpret_cast<uint4 *>(pf_smem);                                    // [C][kSegStride]
    uint32_t *s_bm = reinterpret_cast<uint32_t *>(s_tab + (size_t)C * kSegStride);        // [C][kPfMaxWords]
    uint32_t *s_pre = s_bm + (size_t)C * kPfMaxWords;                                     // [C][kPfMaxWords]
    float4 *s_r = reinterpret_cast<float4 *>(s_pre + (size_t)C * kPfMaxWords);            // [C][kPfThreads] when C > 1
    for (uint32_t pk = blockIdx.x; pk < n_pk; pk += gridDim.x) {
        const DevPacket &p = pkts[pk];
        const DevSetup &su = *p.setup;
        const DevMapping &mp = su.mappings[p.mapping];
        const int n2 = p.n >> 1, nsteps = mp.n_coupling, nwords = (n2 + 31) >> 5;
        const uint8_t *kinds = floor_kind + p.pkt_index * C;
        const uint64_t base = p.coeff_off;
        __syncthreads();                                 // the previous packet is done with the shared tables
        for (int i = tid; i < C * kPfMaxWords; i += kPfThreads) s_bm[i] = 0u;
        for (int i = tid; i < C * kSegStride; i += kPfThreads) {
            const int c = i / kSegStride, j = i - c * kSegStride;
            const size_t row = (size_t)pk * C + c;
            if (j <= (int)seg_cnt[row] && seg_cnt[row]) s_tab[i] = segtab[row * kSegStride + j];
        }
        __syncthreads();
        for (int i = tid; i < C * kSegStride; i += kPfThreads) {
            const int c = i / kSegStride, j = i - c * kSegStride;
            if (j < (int)seg_cnt[(size_t)pk * C + c]) {
                const uint32_t x0 = s_tab[i].y & 0xffffu;
                if ((int)x0 < n2) atomicOr(&s_bm[c * kPfMaxWords + (x0 >> 5)], 1u << (x0 & 31));
            }
        }
        __syncthreads();
        if (warp < C) {                                  // exclusive prefix popcount over the channel's bitmap words
            uint32_t run = 0;
            for (int w0 = 0; w0 < nwords; w0 += 32) {
                const uint32_t v = (w0 + lane < nwords) ? (uint32_t)__popc(s_bm[warp * kPfMaxWords + w0 + lane]) : 0u;
                uint32_t inc = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += t;
                }
                if (w0 + lane < nwords) s_pre[warp * kPfMaxWords + w0 + lane] = run + inc - v;
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        __syncthreads();
        if (C <= 2 && nsteps <= 1) {
            const int k0 = kinds[0], k1 = C == 2 ? kinds[1] : LWB_FLOOR_UNUSED;
            const bool swapped = nsteps == 1 && mp.mag[0] == 1;          // (magnitude, angle) = (1, 0)
            for (int q = tid; q < (n2 >> 2); q += kPfThreads) {
                const uint64_t e0 = base + 4 * (uint64_t)q, e1 = e0 + n2;
                float4 r0 = *reinterpret_cast<const float4 *>(residue + e0);
                float4 r1 = C == 2 ? *reinterpret_cast<const float4 *>(residue + e1) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (nsteps == 1) {
                    if (swapped) {
                        d_inverse_couple(r1.x, r0.x); d_inverse_couple(r1.y, r0.y);
                        d_inverse_couple(r1.z, r0.z); d_inverse_couple(r1.w, r0.w);
                    } else {
                        d_inverse_couple(r0.x, r1.x); d_inverse_couple(r0.y, r1.y);
                        d_inverse_couple(r0.z, r1.z); d_inverse_couple(r0.w, r1.w);
                    }
                }
                const float4 f0 = d_floor_quad(k0, s_db, s_tab, s_bm, s_pre, 4 * q, dense_floor, e0);
                *reinterpret_cast<float4 *>(spec + e0) =
                    make_float4(__fmul_rn(f0.x, r0.x), __fmul_rn(f0.y, r0.y), __fmul_rn(f0.z, r0.z), __fmul_rn(f0.w, r0.w));
                if (C == 2) {
                    const float4 f1 = d_floor_quad(k1, s_db, s_tab + kSegStride, s_bm + kPfMaxWords, s_pre + kPfMaxWords, 4 * q,
                                                   dense_floor, e1);
                    *reinterpret_cast<float4 *>(spec + e1) =
                        make_float4(__fmul_rn(f1.x, r1.x), __fmul_rn(f1.y, r1.y), __fmul_rn(f1.z, r1.z), __fmul_rn(f1.w, r1.w));
                }
            }
            continue;
        }
        // general case: the thread's quads of all channels sit in shared memory (dynamic channel indices of the
        // coupling steps without predicated register arrays); every thread touches only its own column
        for (int q = tid; q < (n2 >> 2); q += kPfThreads) {
            const uint64_t e = base + 4 * (uint64_t)q;
            for (int c = 0; c < C; c++) s_r[c * kPfThreads + tid] = *reinterpret_cast<const float4 *>(residue + e + (uint64_t)c * n2);
            for (int s = nsteps - 1; s >= 0; s--) {                      // audio.rs:991-1002
                float4 m4 = s_r[mp.mag[s] * kPfThreads + tid], a4 = s_r[mp.ang[s] * kPfThreads + tid];
                d_inverse_couple(m4.x, a4.x); d_inverse_couple(m4.y, a4.y);
                d_inverse_couple(m4.z, a4.z); d_inverse_couple(m4.w, a4.w);
                s_r[mp.mag[s] * kPfThreads + tid] = m4;
                s_r[mp.ang[s] * kPfThreads + tid] = a4;
            }
            for (int c = 0; c < C; c++) {
                const uint64_t ec = e + (uint64_t)c * n2;
                const float4 r = s_r[c * kPfThreads + tid];
                const float4 f = d_floor_quad(kinds[c], s_db, s_tab + c * kSegStride, s_bm + c * kPfMaxWords, s_pre + c * kPfMaxWords,
                                              4 * q, dense_floor, ec);
                *reinterpret_cast<float4 *>(spec + ec) =
                    make_float4(__fmul_rn(f.x, r.x), __fmul_rn(f.y, r.y), __fmul_rn(f.z, r.z), __fmul_rn(f.w, r.w));
            }
        }
    }
}

inline void prologue_kernel_configure()
{
    cudaFuncSetAttribute(k_prologue_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prologue_fused_smem(8));
}

}  // namespace lwb
