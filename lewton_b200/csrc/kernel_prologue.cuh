// kernel_prologue.cuh -- the residue entry's front stages for batches (audio.rs:991-1039), as two
// bandwidth-shaped kernels that replace the per-packet-CTA k_prologue of kernels_generic.cuh wherever a
// batch has <= 8 channels:
//
//   k_floor1_curves : floor-1 step 1 (post unwrap, audio.rs:391-435) with one THREAD per (packet, channel) --
//                     the unwrap is serial over <= 65 posts, so 32 independent rows per warp is the only way
//                     to keep its lanes busy -- then step 2 (render_line, audio.rs:503-555) as a closed form,
//                     16 bins per work item, one byte per bin into a curve arena laid out like the
//                     coefficient arena (curve[e] belongs to coefficient element e);
//   k_prologue3     : one thread per 4 bins of a packet, all channels: inverse coupling in registers / shared
//                     memory (steps in reverse, audio.rs:991-1002), floor lookup (curve byte -> dB table,
//                     audio.rs:552-554; unused floor = zero curve, :1021-1024; dense = host-computed floor-0),
//                     multiply (:1035-1037), float4 loads and stores.
//
// HBM bytes per coefficient: residue 4 in, curve 1 out + 1 in, spectrum 4 out (+ ~0.26 for the posts); the
// old kernel moved the same 8 but took 5x its roofline time (serial unwrap on one lane per warp with the CTA
// waiting, 8-way predicated register arrays).
#pragma once
#include "kernels_generic.cuh"

namespace lwb {

constexpr int kCurveRows = 64;         // (packet, channel) rows per CTA of k_floor1_curves
constexpr int kCurveThreads = 256;
constexpr int kCurveSeg = LWB_MAX_POSTS + 3;   // 68: flagged posts + flat tail, padded

// grid = ceil(n_pk * C / kCurveRows).  Rows whose floor kind is not LWB_FLOOR_ONE are skipped (their curve
// bytes are never read).  curve: byte arena indexed by coefficient element offset (DevPacket::coeff_off).
__global__ void __launch_bounds__(kCurveThreads)
k_floor1_curves(const DevPacket *__restrict__ pkts, uint32_t n_rows, int C, const uint8_t *__restrict__ floor_kind,
                const uint32_t *__restrict__ floor1_y, uint8_t *__restrict__ curve)
{
    __shared__ uint16_t s_x[kCurveRows][kCurveSeg];
    __shared__ uint16_t s_y[kCurveRows][kCurveSeg];
    __shared__ uint32_t s_mg[kCurveRows][kCurveSeg];
    __shared__ int s_m[kCurveRows];
    __shared__ int s_n2[kCurveRows];
    __shared__ unsigned long long s_off[kCurveRows];
    const int tid = threadIdx.x;
    const uint32_t row0 = blockIdx.x * kCurveRows;
    if (tid < kCurveRows) {
        const uint32_t row = row0 + tid;
        int m = 0;
        if (row < n_rows) {
            const uint32_t pk = row / (uint32_t)C, c = row - pk * (uint32_t)C;
            const DevPacket &p = pkts[pk];
            const int n2 = p.n >> 1;
            const uint64_t frow = p.pkt_index * (uint64_t)C + c;
            if (floor_kind[frow] == LWB_FLOOR_ONE) {
                const DevSetup &su = *p.setup;
                const DevFloor1 &fl = su.floors[su.mappings[p.mapping].floor_of_channel[c]];
                m = d_floor1_posts(fl, floor1_y + frow * LWB_MAX_POSTS, n2, s_x[tid], s_y[tid]);
            }
            s_n2[tid] = n2;
            s_off[tid] = p.coeff_off + (uint64_t)c * n2;
        }
        s_m[tid] = m;
    }
    __syncthreads();
    for (int i = tid; i < kCurveRows * kCurveSeg; i += kCurveThreads) {
        const int r = i / kCurveSeg, j = i - r * kCurveSeg;
        if (j + 1 < s_m[r]) d_floor1_prepare_segment(s_x[r], s_y[r], s_mg[r], j);
    }
    __syncthreads();
    const int w = tid >> 5, lane = tid & 31;
    for (int r = w; r < kCurveRows; r += kCurveThreads / 32) {
        const int m = s_m[r];
        if (m < 2) continue;
        const int chunks = s_n2[r] >> 4;
        uint8_t *dst = curve + s_off[r];
        const bool a16 = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
        for (int ch = lane; ch < chunks; ch += 32) {
            uint32_t o[4];
            d_floor1_render16(s_x[r], s_y[r], s_mg[r], m, ch * 16, o);
            if (a16) {
                *reinterpret_cast<uint4 *>(dst + ch * 16) = make_uint4(o[0], o[1], o[2], o[3]);
            } else {
                uint32_t *d32 = reinterpret_cast<uint32_t *>(dst + ch * 16);     // offsets are multiples of 4 elements
                d32[0] = o[0]; d32[1] = o[1]; d32[2] = o[2]; d32[3] = o[3];
            }
        }
    }
}

constexpr int kPro3Threads = 256;
// (two channels with more than one coupling step take the general path too, so they get the staging as well)
inline size_t prologue3_smem(int channels) { return channels > 1 ? (size_t)channels * kPro3Threads * sizeof(float4) : 0; }

__device__ __forceinline__ float4 d_floor_quad(int kind, const float *__restrict__ s_db, const uint8_t *__restrict__ curve,
                                               const float *__restrict__ dense, uint64_t e)
{
    if (kind == LWB_FLOOR_ONE) {
        const uint32_t wv = *reinterpret_cast<const uint32_t *>(curve + e);
        return make_float4(s_db[wv & 255u], s_db[(wv >> 8) & 255u], s_db[(wv >> 16) & 255u], s_db[wv >> 24]);
    }
    if (kind == LWB_FLOOR_DENSE) return *reinterpret_cast<const float4 *>(dense + e);
    return make_float4(0.f, 0.f, 0.f, 0.f);                       // audio.rs:1021-1024
}

// Persistent CTAs striding over the packets (grid = min(packets, a few CTAs per SM)).  Requires every coeff_off
// (and the arena bases) to be multiples of 4 elements.
__global__ void __launch_bounds__(kPro3Threads)
k_prologue3(const DevPacket *__restrict__ pkts, uint32_t n_pk, const float *__restrict__ residue, const float *__restrict__ dense_floor,
            const uint8_t *__restrict__ floor_kind, const uint8_t *__restrict__ curve, float *__restrict__ spec)
{
    extern __shared__ float4 s_r[];                      // [C][kPro3Threads] when C > 1
    __shared__ float s_db[256];
    const int tid = threadIdx.x;
    s_db[tid] = c_inverse_db[tid];
    __syncthreads();
    for (uint32_t pk = blockIdx.x; pk < n_pk; pk += gridDim.x) {
        const DevPacket &p = pkts[pk];
        const DevSetup &su = *p.setup;
        const DevMapping &mp = su.mappings[p.mapping];
        const int C = p.channels, n2 = p.n >> 1, nsteps = mp.n_coupling;
        const uint8_t *kinds = floor_kind + p.pkt_index * C;
        const uint64_t base = p.coeff_off;
        if (C <= 2 && nsteps <= 1) {
            const int k0 = kinds[0], k1 = C == 2 ? kinds[1] : LWB_FLOOR_UNUSED;
            const bool swapped = nsteps == 1 && mp.mag[0] == 1;          // (magnitude, angle) = (1, 0)
            for (int q = tid; q < (n2 >> 2); q += kPro3Threads) {
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
                const float4 f0 = d_floor_quad(k0, s_db, curve, dense_floor, e0);
                *reinterpret_cast<float4 *>(spec + e0) =
                    make_float4(__fmul_rn(f0.x, r0.x), __fmul_rn(f0.y, r0.y), __fmul_rn(f0.z, r0.z), __fmul_rn(f0.w, r0.w));
                if (C == 2) {
                    const float4 f1 = d_floor_quad(k1, s_db, curve, dense_floor, e1);
                    *reinterpret_cast<float4 *>(spec + e1) =
                        make_float4(__fmul_rn(f1.x, r1.x), __fmul_rn(f1.y, r1.y), __fmul_rn(f1.z, r1.z), __fmul_rn(f1.w, r1.w));
                }
            }
            continue;
        }
        // general case: the thread's quads of all channels sit in shared memory (dynamic channel indices of the
        // coupling steps without predicated register arrays); every thread touches only its own column
        for (int q = tid; q < (n2 >> 2); q += kPro3Threads) {
            const uint64_t e = base + 4 * (uint64_t)q;
            for (int c = 0; c < C; c++) s_r[c * kPro3Threads + tid] = *reinterpret_cast<const float4 *>(residue + e + (uint64_t)c * n2);
            for (int s = nsteps - 1; s >= 0; s--) {                      // audio.rs:991-1002
                float4 m4 = s_r[mp.mag[s] * kPro3Threads + tid], a4 = s_r[mp.ang[s] * kPro3Threads + tid];
                d_inverse_couple(m4.x, a4.x); d_inverse_couple(m4.y, a4.y);
                d_inverse_couple(m4.z, a4.z); d_inverse_couple(m4.w, a4.w);
                s_r[mp.mag[s] * kPro3Threads + tid] = m4;
                s_r[mp.ang[s] * kPro3Threads + tid] = a4;
            }
            for (int c = 0; c < C; c++) {
                const uint64_t ec = e + (uint64_t)c * n2;
                const float4 r = s_r[c * kPro3Threads + tid];
                const float4 f = d_floor_quad(kinds[c], s_db, curve, dense_floor, ec);
                *reinterpret_cast<float4 *>(spec + ec) =
                    make_float4(__fmul_rn(f.x, r.x), __fmul_rn(f.y, r.y), __fmul_rn(f.z, r.z), __fmul_rn(f.w, r.w));
            }
        }
    }
}

}  // namespace lwb
