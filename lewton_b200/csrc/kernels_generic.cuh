// kernels_generic.cuh -- the general synthesis path: any blocksize 6..13, any channel count,
// mixed short/long sequences, all output formats.  Four kernels, fully parallel over packets:
//
//   k_prologue  : inverse coupling -> floor-1 render -> floor x residue   (audio.rs:991-1039)
//   k_imdct     : inverse MDCT of one (packet, channel) block in shared memory (imdct.rs:291-659)
//   k_overlap   : window / overlap-add / slice / sample conversion         (audio.rs:1079-1157)
//   k_save_state: PreviousWindowRight update                                (audio.rs:1121,1154)
//
// Bandwidth: this path round-trips the spectrum and the un-windowed IMDCT output through HBM
// (about 32 B per output sample instead of the algorithmic 8); the fused kernel in
// kernel_long.cuh is the hot path for the headline configuration, this one is the general
// fallback every configuration is correct on.
//
// Arithmetic rules (parity with the reference): binary32, round-to-nearest, no FMA contraction
// (-fmad=false), denormals kept (-ftz=false); every butterfly uses the reference's operand order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernel_long.cuh"
#include "lwb_common.h"
#include "floor1_eval.cuh"

namespace lwb {

__constant__ float c_inverse_db[256] = {
#include "floor1_inverse_db.inc"
};

// audio.rs:762-777, branch-free.  With s = (a > 0), t = (m > 0) (both false for NaN, like the reference's `> 0.`):
//   v = m + (t == s ? -a : a)   -- m - a and m + (-a) are the same IEEE operation --
//   s: (m, a) <- (m, v);  !s: (m, a) <- (v, m)
__device__ __forceinline__ void d_inverse_couple(float &m, float &a)
{
    const float m0 = m, a0 = a;
    const bool s = a0 > 0.f, t = m0 > 0.f;
    const float v = __fadd_rn(m0, (t == s) ? -a0 : a0);
    m = s ? m0 : v;
    a = s ? v : m0;
}

constexpr int kPrologueThreads = 256;
constexpr int kPrologueGroup = 8;            // channels whose floor posts sit in smem at once
// dynamic shared memory of k_prologue: one curve byte per bin for up to 8 channels of the largest block
inline size_t prologue_smem(int channels, int blocksize_1) { return (size_t)(channels < kPrologueGroup ? channels : kPrologueGroup) << (blocksize_1 - 1); }

// grid.x = packets.  spec[packet] = [channels][n/2] receives floor x decoupled residue.
__global__ void __launch_bounds__(kPrologueThreads)
k_prologue(const DevPacket *__restrict__ pkts, const float *__restrict__ residue,
           const float *__restrict__ dense_floor, const uint8_t *__restrict__ floor_kind,
           const uint32_t *__restrict__ floor1_y, float *__restrict__ spec)
{
    const DevPacket &p = pkts[blockIdx.x];
    const DevSetup &su = *p.setup;
    const DevMapping &mp = su.mappings[p.mapping];
    const int C = p.channels;
    const int n2 = p.n >> 1;
    const float *res = residue + p.coeff_off;
    float *out = spec + p.coeff_off;
    const int nsteps = mp.n_coupling;

    __shared__ uint16_t s_x[kPrologueGroup][LWB_MAX_POSTS + 1];
    __shared__ uint16_t s_y[kPrologueGroup][LWB_MAX_POSTS + 1];
    __shared__ int s_m[kPrologueGroup];
    const uint8_t *kinds = floor_kind + p.pkt_index * C;

    if (C <= kPrologueGroup) {
        // Up to 8 channels: the floor curves are rendered once into shared memory (one byte per bin) and a
        // single pass over the bins does the rest, so every value crosses the memory system once.
        //   a. posts: one lane per channel, serial over <= 65 posts (audio.rs:391-435);
        //   b. curve: warp w renders channel w, one flagged segment per lane, with the reference's own
        //      integer DDA (render_line, audio.rs:503-524) -- no search and no division per bin;
        //   c. bins: load the residues, inverse-couple them in registers (steps in reverse,
        //      audio.rs:991-1002), multiply by the floor (audio.rs:1006-1039), store.
        extern __shared__ uint8_t s_curve_raw[];            // [C][n2] bytes: the host passes min(C, 8) * blocksize_1 / 2
        uint8_t *s_curve = s_curve_raw;
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0 && w < C && kinds[w] == LWB_FLOOR_ONE) {
            const DevFloor1 &fl = su.floors[mp.floor_of_channel[w]];
            s_m[w] = d_floor1_posts(fl, floor1_y + (p.pkt_index * C + w) * LWB_MAX_POSTS, n2, s_x[w], s_y[w]);
        }
        __syncwarp();
        if (w < C && kinds[w] == LWB_FLOOR_ONE)
            for (int seg = lane; seg + 1 < s_m[w]; seg += 32) d_floor1_render_segment(s_x[w], s_y[w], seg, n2, s_curve + (size_t)w * n2);
        __syncthreads();
        if (C == 2 && nsteps <= 1) {
            // the common stereo shape: at most one coupling step, no register-array juggling
            const bool swapped = nsteps == 1 && mp.mag[0] == 1;       // (magnitude, angle) = (1, 0)
            const int k0 = kinds[0], k1 = kinds[1];
            for (int k = threadIdx.x; k < n2; k += kPrologueThreads) {
                float r0 = res[k], r1 = res[(size_t)n2 + k];
                if (nsteps == 1) {
                    if (swapped) d_inverse_couple(r1, r0);
                    else d_inverse_couple(r0, r1);
                }
                float f0 = 0.f, f1 = 0.f;                              // audio.rs:1021-1024
                if (k0 == LWB_FLOOR_ONE) f0 = c_inverse_db[s_curve[k]];
                else if (k0 == LWB_FLOOR_DENSE) f0 = dense_floor[p.coeff_off + k];
                if (k1 == LWB_FLOOR_ONE) f1 = c_inverse_db[s_curve[(size_t)n2 + k]];
                else if (k1 == LWB_FLOOR_DENSE) f1 = dense_floor[p.coeff_off + (size_t)n2 + k];
                out[k] = __fmul_rn(f0, r0);                            // audio.rs:1035-1037
                out[(size_t)n2 + k] = __fmul_rn(f1, r1);
            }
            return;
        }
        for (int k = threadIdx.x; k < n2; k += kPrologueThreads) {
            float r[8];
#pragma unroll
            for (int c = 0; c < 8; c++) r[c] = c < C ? res[(size_t)c * n2 + k] : 0.f;
            for (int s = nsteps - 1; s >= 0; s--) {
                const int mi = mp.mag[s], ai = mp.ang[s];
                float mv = 0.f, av = 0.f;
#pragma unroll
                for (int c = 0; c < 8; c++) { if (c == mi) mv = r[c]; if (c == ai) av = r[c]; }
                d_inverse_couple(mv, av);
#pragma unroll
                for (int c = 0; c < 8; c++) { if (c == mi) r[c] = mv; if (c == ai) r[c] = av; }
            }
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (c < C) {
                    const int kind = kinds[c];
                    float f = 0.f;
                    if (kind == LWB_FLOOR_ONE) f = c_inverse_db[s_curve[(size_t)c * n2 + k]];
                    else if (kind == LWB_FLOOR_DENSE) f = dense_floor[p.coeff_off + (size_t)c * n2 + k];
                    out[(size_t)c * n2 + k] = __fmul_rn(f, r[c]);
                }
            }
        }
        return;
    }

    // more than 8 channels: two passes through `out`
    // 1. inverse coupling, steps in reverse (audio.rs:991-1002).  Every thread owns its bins
    //    through all steps, so the working copy in `out` needs no synchronisation.
    for (int k = threadIdx.x; k < n2; k += kPrologueThreads) {
        for (int c = 0; c < C; c++) out[(size_t)c * n2 + k] = res[(size_t)c * n2 + k];
        for (int s = nsteps - 1; s >= 0; s--) {
            float mv = out[(size_t)mp.mag[s] * n2 + k], av = out[(size_t)mp.ang[s] * n2 + k];
            d_inverse_couple(mv, av);
            out[(size_t)mp.mag[s] * n2 + k] = mv;
            out[(size_t)mp.ang[s] * n2 + k] = av;
        }
    }

    // 2. floor curve x residue (audio.rs:1006-1039), kPrologueGroup channels at a time
    for (int c0 = 0; c0 < C; c0 += kPrologueGroup) {
        __syncthreads();
        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0 && c0 + w < C && kinds[c0 + w] == LWB_FLOOR_ONE) {
            const int c = c0 + w;
            const DevFloor1 &fl = su.floors[mp.floor_of_channel[c]];
            s_m[w] = d_floor1_posts(fl, floor1_y + (p.pkt_index * C + c) * LWB_MAX_POSTS, n2,
                                    s_x[w], s_y[w]);
        }
        __syncthreads();
        for (int g = 0; g < kPrologueGroup && c0 + g < C; g++) {
            const int c = c0 + g;
            const int kind = kinds[c];
            float *oc = out + (size_t)c * n2;
            for (int k = threadIdx.x; k < n2; k += kPrologueThreads) {
                float f;
                if (kind == LWB_FLOOR_ONE)
                    f = c_inverse_db[d_floor1_y_at(s_x[g], s_y[g], s_m[g], k) & 255u];
                else if (kind == LWB_FLOOR_DENSE)
                    f = dense_floor[p.coeff_off + (size_t)c * n2 + k];
                else
                    f = 0.f;                           // audio.rs:1021-1024
                oc[k] = __fmul_rn(f, oc[k]);           // audio.rs:1035-1037
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// inverse MDCT, imdct.rs:291-659, stage by stage in shared memory
// ---------------------------------------------------------------------------------------------
// One rotate-and-sum butterfly of step 3 (imdct.rs:36-41 / 94-99 / 161-166): `hi` is the odd
// index of the upper pair, `lo` of the lower pair.
__device__ __forceinline__ void d_bfly(float *e, int hi, int lo, float w0, float w1)
{
    const float k00 = __fsub_rn(e[hi], e[lo]);
    const float k01 = __fsub_rn(e[hi - 1], e[lo - 1]);
    e[hi] = __fadd_rn(e[hi], e[lo]);
    e[hi - 1] = __fadd_rn(e[hi - 1], e[lo - 1]);
    e[lo] = __fsub_rn(__fmul_rn(k00, w0), __fmul_rn(k01, w1));
    e[lo - 1] = __fadd_rn(__fmul_rn(k01, w0), __fmul_rn(k00, w1));
}

// imdct.rs:201-232, z7 = &zm7[7]
__device__ __forceinline__ void d_iter_54(float *z7)
{
    const float k00 = __fsub_rn(z7[0], z7[-4]);
    const float y0 = __fadd_rn(z7[0], z7[-4]);
    const float y2 = __fadd_rn(z7[-2], z7[-6]);
    const float k22 = __fsub_rn(z7[-2], z7[-6]);
    z7[0] = __fadd_rn(y0, y2);
    z7[-2] = __fsub_rn(y0, y2);
    const float k33 = __fsub_rn(z7[-3], z7[-7]);
    z7[-4] = __fadd_rn(k00, k33);
    z7[-6] = __fsub_rn(k00, k33);
    const float k11 = __fsub_rn(z7[-1], z7[-5]);
    const float y1 = __fadd_rn(z7[-1], z7[-5]);
    const float y3 = __fadd_rn(z7[-3], z7[-7]);
    z7[-1] = __fadd_rn(y1, y3);
    z7[-3] = __fsub_rn(y1, y3);
    z7[-5] = __fsub_rn(k11, k22);
    z7[-7] = __fadd_rn(k11, k22);
}

constexpr int kImdctThreads = 128;

// Steps 0-7 of inverse_mdct (imdct.rs:337-580) for one block, cooperatively by NT threads (`tid` in
// [0, NT)) that `sync()` synchronises (a warp, a named-barrier group, or the CTA).
// X: n/2 spectrum coefficients (global or shared; may alias U); on return V holds the post-step-7
// buffer that step 8 reads.  Literal schedule, including the reference's behaviour for n = 64/128.
template <class Sync>
__device__ __forceinline__ void d_imdct_to_v(const DevTables &tb, int n, const float *X, float *U, float *V, int tid,
                                             int NT, Sync sync)
{
    const int n2 = n >> 1, n4 = n >> 2, n8 = n >> 3;
    const int ld = tb.bs;
    const float *__restrict__ A = tb.a;
    const float *__restrict__ Cc = tb.c;

    // step 0 (imdct.rs:337-371): V <- rotated, reflected spectrum
    for (int t = tid; t < n8; t += NT) {
        const float x0 = X[4 * t], x2 = X[4 * t + 2];
        const float a0 = A[2 * t], a1 = A[2 * t + 1];
        const int d = n4 - 2 - 2 * t, ao = n4 + 2 * t, e = n2 - 3 - 4 * t;
        const float ne2 = -X[e + 2], ne0 = -X[e];
        const float b0 = A[ao], b1 = A[ao + 1];
        V[n2 - 1 - 2 * t] = __fsub_rn(__fmul_rn(x0, a0), __fmul_rn(x2, a1));
        V[n2 - 2 - 2 * t] = __fadd_rn(__fmul_rn(x0, a1), __fmul_rn(x2, a0));
        V[d + 1] = __fsub_rn(__fmul_rn(ne2, b0), __fmul_rn(ne0, b1));
        V[d] = __fadd_rn(__fmul_rn(ne2, b1), __fmul_rn(ne0, b0));
    }
    sync();
    // step 2 (imdct.rs:385-430): U <- V
    for (int t = tid; t < (n >> 4); t += NT) {
        const int ao = n2 - 8 - 8 * t, hi = n4 + 4 * t, lo = 4 * t;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int o = 2 * h;                  // pair (0,1) uses A[ao+4..5], pair (2,3) A[ao..ao+1]
            const float w0 = A[ao + 4 - 4 * h], w1 = A[ao + 5 - 4 * h];
            const float v1 = __fsub_rn(V[hi + o + 1], V[lo + o + 1]);
            const float v0 = __fsub_rn(V[hi + o], V[lo + o]);
            U[hi + o + 1] = __fadd_rn(V[hi + o + 1], V[lo + o + 1]);
            U[hi + o] = __fadd_rn(V[hi + o], V[lo + o]);
            U[lo + o + 1] = __fsub_rn(__fmul_rn(v1, w0), __fmul_rn(v0, w1));
            U[lo + o] = __fadd_rn(__fmul_rn(v0, w0), __fmul_rn(v1, w1));
        }
    }
    sync();
    // step 3 (imdct.rs:445-477), literal schedule: stage 0, stage 1 (a no-op for n = 64),
    // stages 2..ld-7.  For n = 64/128 this overlaps what ld654 does again below -- the
    // reference's behaviour, kept bit for bit.
    for (int l = 0; l < 2 || l <= ld - 7; l++) {
        if (l == 1 && n < 128) continue;          // r_loop(lim = n >> 5 = 2): lim >> 2 == 0 iterations
        const int k0 = n >> (l + 2), k1 = 1 << (l + 3);
        const int rbits = ld - l - 4;             // r < n >> (l+4)
        for (int q = tid; q < n8; q += NT) {
            const int r = q & ((1 << rbits) - 1), s = q >> rbits;
            const int i = n2 - 1 - k0 * s - 2 * r;
            d_bfly(U, i, i - (k0 >> 1), A[r * k1], A[r * k1 + 1]);
        }
        sync();
    }
    // imdct.rs:234-288 (ld654): last three stages per 16-float group
    {
        const float a2 = A[n >> 3];
        for (int g = tid; g < (n >> 5); g += NT) {
            float *z = U + (n2 - 1 - 16 * g);
            float k00, k11;
            k00 = __fsub_rn(z[0], z[-8]);   k11 = __fsub_rn(z[-1], z[-9]);
            z[0] = __fadd_rn(z[0], z[-8]);  z[-1] = __fadd_rn(z[-1], z[-9]);
            z[-8] = k00;                    z[-9] = k11;
            k00 = __fsub_rn(z[-2], z[-10]); k11 = __fsub_rn(z[-3], z[-11]);
            z[-2] = __fadd_rn(z[-2], z[-10]); z[-3] = __fadd_rn(z[-3], z[-11]);
            z[-10] = __fmul_rn(__fadd_rn(k00, k11), a2);
            z[-11] = __fmul_rn(__fsub_rn(k11, k00), a2);
            k00 = __fsub_rn(z[-12], z[-4]); k11 = __fsub_rn(z[-5], z[-13]);
            z[-4] = __fadd_rn(z[-4], z[-12]); z[-5] = __fadd_rn(z[-5], z[-13]);
            z[-12] = k11;                   z[-13] = k00;
            k00 = __fsub_rn(z[-14], z[-6]); k11 = __fsub_rn(z[-7], z[-15]);
            z[-6] = __fadd_rn(z[-6], z[-14]); z[-7] = __fadd_rn(z[-7], z[-15]);
            z[-14] = __fmul_rn(__fadd_rn(k00, k11), a2);
            z[-15] = __fmul_rn(__fsub_rn(k00, k11), a2);
            d_iter_54(z);
            d_iter_54(z - 8);
        }
    }
    sync();
    // steps 4-6 (imdct.rs:490-528): bit-reverse shuffle U -> V
    for (int q = tid; q < (n >> 4); q += NT) {
        const int d0 = n4 - 4 - 4 * q, d1 = n2 - 4 - 4 * q;
        int k4 = tb.bitrev[2 * q];
        V[d1 + 3] = U[k4 + 0]; V[d1 + 2] = U[k4 + 1]; V[d0 + 3] = U[k4 + 2]; V[d0 + 2] = U[k4 + 3];
        k4 = tb.bitrev[2 * q + 1];
        V[d1 + 1] = U[k4 + 0]; V[d1 + 0] = U[k4 + 1]; V[d0 + 1] = U[k4 + 2]; V[d0 + 0] = U[k4 + 3];
    }
    sync();
    // step 7 (imdct.rs:533-580), in place on V
    for (int t = tid; t < (n >> 4); t += NT) {
        const int d = 4 * t, e = n2 - 4 - 4 * t, co = 4 * t;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int dd = d + 2 * h, ee = e + 2 - 2 * h;
            const float c0 = Cc[co + 2 * h], c1 = Cc[co + 2 * h + 1];
            const float a02 = __fsub_rn(V[dd], V[ee]);
            const float a11 = __fadd_rn(V[dd + 1], V[ee + 1]);
            const float b0 = __fadd_rn(__fmul_rn(c1, a02), __fmul_rn(c0, a11));
            const float b1 = __fsub_rn(__fmul_rn(c1, a11), __fmul_rn(c0, a02));
            const float b2 = __fadd_rn(V[dd], V[ee]);
            const float b3 = __fsub_rn(V[dd + 1], V[ee + 1]);
            V[dd] = __fadd_rn(b2, b0);
            V[dd + 1] = __fadd_rn(b3, b1);
            V[ee] = __fsub_rn(b2, b0);
            V[ee + 1] = __fsub_rn(b1, b3);
        }
    }
    sync();
}

// The same transform for NP independent blocks of one size at once (block q: spectrum X + q * xs,
// buffers U + q * bs and V + q * bs).  Arithmetic per block is identical to d_imdct_to_v; every stage
// first loads the operands of all NP blocks, then computes, then stores, so that a thread has NP
// independent dependency chains in flight instead of one -- the transform is latency-bound for small
// n, where a stage is a single butterfly per lane (ncu: one instruction per ~54 cycles per warp).
template <int NP, class Sync>
__device__ __forceinline__ void d_imdct_to_v_np(const DevTables &tb, int n, const float *X, size_t xs, float *U, float *V, int bs,
                                                int tid, int NT, Sync sync)
{
    const int n2 = n >> 1, n4 = n >> 2, n8 = n >> 3;
    const int ld = tb.bs;
    const float *__restrict__ A = tb.a;
    const float *__restrict__ Cc = tb.c;
    // step 0
    for (int t = tid; t < n8; t += NT) {
        const float a0 = A[2 * t], a1 = A[2 * t + 1];
        const int d = n4 - 2 - 2 * t, ao = n4 + 2 * t, e = n2 - 3 - 4 * t;
        const float b0 = A[ao], b1 = A[ao + 1];
        float x0[NP], x2[NP], ne2[NP], ne0[NP];
#pragma unroll
        for (int q = 0; q < NP; q++) {
            const float *Xq = X + q * xs;
            x0[q] = Xq[4 * t]; x2[q] = Xq[4 * t + 2]; ne2[q] = -Xq[e + 2]; ne0[q] = -Xq[e];
        }
#pragma unroll
        for (int q = 0; q < NP; q++) {
            float *Vq = V + q * bs;
            Vq[n2 - 1 - 2 * t] = __fsub_rn(__fmul_rn(x0[q], a0), __fmul_rn(x2[q], a1));
            Vq[n2 - 2 - 2 * t] = __fadd_rn(__fmul_rn(x0[q], a1), __fmul_rn(x2[q], a0));
            Vq[d + 1] = __fsub_rn(__fmul_rn(ne2[q], b0), __fmul_rn(ne0[q], b1));
            Vq[d] = __fadd_rn(__fmul_rn(ne2[q], b1), __fmul_rn(ne0[q], b0));
        }
    }
    sync();
    // step 2
    for (int t = tid; t < (n >> 4); t += NT) {
        const int ao = n2 - 8 - 8 * t, hi = n4 + 4 * t, lo = 4 * t;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int o = 2 * h;
            const float w0 = A[ao + 4 - 4 * h], w1 = A[ao + 5 - 4 * h];
            float h1[NP], l1[NP], h0[NP], l0[NP];
#pragma unroll
            for (int q = 0; q < NP; q++) {
                const float *Vq = V + q * bs;
                h1[q] = Vq[hi + o + 1]; l1[q] = Vq[lo + o + 1]; h0[q] = Vq[hi + o]; l0[q] = Vq[lo + o];
            }
#pragma unroll
            for (int q = 0; q < NP; q++) {
                float *Uq = U + q * bs;
                const float v1 = __fsub_rn(h1[q], l1[q]), v0 = __fsub_rn(h0[q], l0[q]);
                Uq[hi + o + 1] = __fadd_rn(h1[q], l1[q]);
                Uq[hi + o] = __fadd_rn(h0[q], l0[q]);
                Uq[lo + o + 1] = __fsub_rn(__fmul_rn(v1, w0), __fmul_rn(v0, w1));
                Uq[lo + o] = __fadd_rn(__fmul_rn(v0, w0), __fmul_rn(v1, w1));
            }
        }
    }
    sync();
    // step 3, literal schedule
    for (int l = 0; l < 2 || l <= ld - 7; l++) {
        if (l == 1 && n < 128) continue;
        const int k0 = n >> (l + 2), k1 = 1 << (l + 3);
        const int rbits = ld - l - 4;
        for (int qq = tid; qq < n8; qq += NT) {
            const int r = qq & ((1 << rbits) - 1), s = qq >> rbits;
            const int i = n2 - 1 - k0 * s - 2 * r, lo = i - (k0 >> 1);
            const float w0 = A[r * k1], w1 = A[r * k1 + 1];
            float eh[NP], el[NP], eh1[NP], el1[NP];
#pragma unroll
            for (int q = 0; q < NP; q++) {
                const float *Uq = U + q * bs;
                eh[q] = Uq[i]; el[q] = Uq[lo]; eh1[q] = Uq[i - 1]; el1[q] = Uq[lo - 1];
            }
#pragma unroll
            for (int q = 0; q < NP; q++) {
                float *Uq = U + q * bs;
                const float k00 = __fsub_rn(eh[q], el[q]), k01 = __fsub_rn(eh1[q], el1[q]);
                Uq[i] = __fadd_rn(eh[q], el[q]);
                Uq[i - 1] = __fadd_rn(eh1[q], el1[q]);
                Uq[lo] = __fsub_rn(__fmul_rn(k00, w0), __fmul_rn(k01, w1));
                Uq[lo - 1] = __fadd_rn(__fmul_rn(k01, w0), __fmul_rn(k00, w1));
            }
        }
        sync();
    }
    // ld654: 16-float groups; NP blocks x (n >> 5) groups share the threads
    {
        const float a2 = A[n >> 3];
        const int groups = n >> 5;
        for (int gq = tid; gq < groups * NP; gq += NT) {
            const int q = gq / groups, g = gq - q * groups;
            float *z = U + q * bs + (n2 - 1 - 16 * g);
            float k00, k11;
            k00 = __fsub_rn(z[0], z[-8]);   k11 = __fsub_rn(z[-1], z[-9]);
            z[0] = __fadd_rn(z[0], z[-8]);  z[-1] = __fadd_rn(z[-1], z[-9]);
            z[-8] = k00;                    z[-9] = k11;
            k00 = __fsub_rn(z[-2], z[-10]); k11 = __fsub_rn(z[-3], z[-11]);
            z[-2] = __fadd_rn(z[-2], z[-10]); z[-3] = __fadd_rn(z[-3], z[-11]);
            z[-10] = __fmul_rn(__fadd_rn(k00, k11), a2);
            z[-11] = __fmul_rn(__fsub_rn(k11, k00), a2);
            k00 = __fsub_rn(z[-12], z[-4]); k11 = __fsub_rn(z[-5], z[-13]);
            z[-4] = __fadd_rn(z[-4], z[-12]); z[-5] = __fadd_rn(z[-5], z[-13]);
            z[-12] = k11;                   z[-13] = k00;
            k00 = __fsub_rn(z[-14], z[-6]); k11 = __fsub_rn(z[-7], z[-15]);
            z[-6] = __fadd_rn(z[-6], z[-14]); z[-7] = __fadd_rn(z[-7], z[-15]);
            z[-14] = __fmul_rn(__fadd_rn(k00, k11), a2);
            z[-15] = __fmul_rn(__fsub_rn(k00, k11), a2);
            d_iter_54(z);
            d_iter_54(z - 8);
        }
    }
    sync();
    // steps 4-6: bit-reverse shuffle U -> V
    for (int qq = tid; qq < (n >> 4); qq += NT) {
        const int d0 = n4 - 4 - 4 * qq, d1 = n2 - 4 - 4 * qq;
        const int ka = tb.bitrev[2 * qq], kb = tb.bitrev[2 * qq + 1];
        float a[NP][4], b[NP][4];
#pragma unroll
        for (int q = 0; q < NP; q++) {
            const float *Uq = U + q * bs;
#pragma unroll
            for (int k = 0; k < 4; k++) { a[q][k] = Uq[ka + k]; b[q][k] = Uq[kb + k]; }
        }
#pragma unroll
        for (int q = 0; q < NP; q++) {
            float *Vq = V + q * bs;
            Vq[d1 + 3] = a[q][0]; Vq[d1 + 2] = a[q][1]; Vq[d0 + 3] = a[q][2]; Vq[d0 + 2] = a[q][3];
            Vq[d1 + 1] = b[q][0]; Vq[d1 + 0] = b[q][1]; Vq[d0 + 1] = b[q][2]; Vq[d0 + 0] = b[q][3];
        }
    }
    sync();
    // step 7, in place on V
    for (int t = tid; t < (n >> 4); t += NT) {
        const int d = 4 * t, e = n2 - 4 - 4 * t, co = 4 * t;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int dd = d + 2 * h, ee = e + 2 - 2 * h;
            const float c0 = Cc[co + 2 * h], c1 = Cc[co + 2 * h + 1];
            float vd[NP], vd1[NP], ve[NP], ve1[NP];
#pragma unroll
            for (int q = 0; q < NP; q++) {
                const float *Vq = V + q * bs;
                vd[q] = Vq[dd]; vd1[q] = Vq[dd + 1]; ve[q] = Vq[ee]; ve1[q] = Vq[ee + 1];
            }
#pragma unroll
            for (int q = 0; q < NP; q++) {
                float *Vq = V + q * bs;
                const float a02 = __fsub_rn(vd[q], ve[q]);
                const float a11 = __fadd_rn(vd1[q], ve1[q]);
                const float b0 = __fadd_rn(__fmul_rn(c1, a02), __fmul_rn(c0, a11));
                const float b1 = __fsub_rn(__fmul_rn(c1, a11), __fmul_rn(c0, a02));
                const float b2 = __fadd_rn(vd[q], ve[q]);
                const float b3 = __fsub_rn(vd1[q], ve1[q]);
                Vq[dd] = __fadd_rn(b2, b0);
                Vq[dd + 1] = __fadd_rn(b3, b1);
                Vq[ee] = __fsub_rn(b2, b0);
                Vq[ee + 1] = __fsub_rn(b1, b3);
            }
        }
    }
    sync();
}

// grid = (packets, max channels); dynamic smem = n floats (U and V halves).
// in: spectrum [channels][n/2] at coeff_off; out: x [channels][n] at x_off.
__global__ void __launch_bounds__(kImdctThreads)
k_imdct(const DevPacket *__restrict__ pkts, const float *__restrict__ spec, float *__restrict__ xout)
{
    const DevPacket &p = pkts[blockIdx.x];
    const int ch = blockIdx.y;
    if (ch >= p.channels) return;
    const DevTables &tb = p.setup->tab[p.blockflag];
    const int n = p.n, n2 = n >> 1, n4 = n >> 2;
    const float *__restrict__ B = tb.b;
    const float *__restrict__ X = spec + p.coeff_off + (size_t)ch * n2;
    float *__restrict__ out = xout + p.x_off + (size_t)ch * n;
    extern __shared__ float smem[];
    float *U = smem, *V = smem + n2;
    const int tid = threadIdx.x;
    d_imdct_to_v(tb, n, X, U, V, tid, kImdctThreads, [] { __syncthreads(); });
    // step 8 + decode (imdct.rs:589-658)
    for (int m = tid; m < n4; m += kImdctThreads) {
        const int ee = n2 - 2 - 2 * m;            // V/B pair consumed for output index m
        const float v0 = V[ee], v1 = V[ee + 1];
        const float b0 = B[ee], b1 = B[ee + 1];
        const float p_odd = __fsub_rn(__fmul_rn(v0, b1), __fmul_rn(v1, b0));
        const float p_even = __fsub_rn(__fmul_rn(-v0, b0), __fmul_rn(v1, b1));
        out[m] = p_odd;
        out[n2 - 1 - m] = -p_odd;
        out[n2 + m] = p_even;
        out[n - 1 - m] = p_even;
    }
}

// ---------------------------------------------------------------------------------------------
// window / overlap-add / slice / sample conversion, audio.rs:1079-1157 + samples.rs
// ---------------------------------------------------------------------------------------------
// d_sample_i16 (samples.rs:92-103) lives in kernel_long.cuh, shared by both paths

constexpr int kOverlapThreads = 256;

template <int FORMAT>
__global__ void __launch_bounds__(kOverlapThreads)
k_overlap(const DevPacket *__restrict__ pkts, const float *__restrict__ x, void *__restrict__ pcm)
{
    const DevPacket &p = pkts[blockIdx.x];
    const int ch = blockIdx.y;
    if (ch >= p.channels || p.plen == 0) return;      // audio.rs:1140-1151: no previous -> no output
    const int n = p.n;
    const float *__restrict__ xc = x + p.x_off + (size_t)ch * n;
    const float *__restrict__ prev;
    if (p.prev_packet >= 0) {
        const DevPacket &q = pkts[p.prev_packet];
        prev = x + q.x_off + (size_t)ch * q.n + p.prev_rs;
    } else {
        prev = p.state + (size_t)ch * p.state_stride;
    }
    const float *__restrict__ w = p.setup->tab[p.slope_sel].window;
    const int plen = p.plen, ls = p.ls, olen = p.rs - p.ls;
    for (int i = threadIdx.x; i < olen; i += kOverlapThreads) {
        float v = xc[ls + i];
        if (i < plen)                                  // audio.rs:1116-1118
            v = __fadd_rn(__fmul_rn(v, w[i]), __fmul_rn(prev[i], w[plen - 1 - i]));
        if (FORMAT == LWB_OUT_F32_PLANAR)
            ((float *)pcm)[p.out_off + (size_t)ch * p.out_stride + i] = v;
        else if (FORMAT == LWB_OUT_I16_PLANAR)
            ((int16_t *)pcm)[p.out_off + (size_t)ch * p.out_stride + i] = d_sample_i16(v);
        else if (FORMAT == LWB_OUT_F32_INTERLEAVED)
            ((float *)pcm)[p.out_off + (size_t)i * p.channels + ch] = v;
        else
            ((int16_t *)pcm)[p.out_off + (size_t)i * p.channels + ch] = d_sample_i16(v);
    }
}

// grid = (packets, max channels): packets flagged save_state copy x[rs..re) into the stream state.
__global__ void __launch_bounds__(kOverlapThreads)
k_save_state(const DevPacket *__restrict__ pkts, const float *__restrict__ x)
{
    const DevPacket &p = pkts[blockIdx.x];
    const int ch = blockIdx.y;
    if (!p.save_state || ch >= p.channels) return;
    const float *__restrict__ xc = x + p.x_off + (size_t)ch * p.n + p.rs;
    float *__restrict__ st = p.state + (size_t)ch * p.state_stride;
    for (int i = threadIdx.x; i < p.re - p.rs; i += kOverlapThreads) st[i] = xc[i];
}

}  // namespace lwb
