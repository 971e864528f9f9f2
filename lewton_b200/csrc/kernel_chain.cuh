// kernel_chain.cuh -- the general synthesis kernel: any blocksize 6..13, mixed short/long
// sequences, up to 8 channels, spectrum or residue entry, all output formats, in ONE launch.
//
// One CTA owns one chain (consecutive packets of one stream), one group of `wpc` warps per channel
// (1 warp for small blocks, up to 8 for n = 8192; the group synchronises on its own named barrier).
// The group keeps its channel's working buffers (U, V: n/2 floats each) and the previous block's right half
// (PreviousWindowRight, audio.rs:847-861) in shared memory for the whole chain, so HBM sees only the
// algorithmic traffic: coefficients in, PCM out, the stream state once per chain.  Per packet:
//   residue entry: floor-1 posts per channel (one lane, serial, <= 65 posts) -> the whole CTA does
//                  inverse coupling + floor x residue bin-parallel straight into the channels'
//                  shared buffers (audio.rs:991-1039);
//   every warp:    inverse MDCT stage by stage in shared memory (imdct.rs:291-659, literal schedule),
//                  window / overlap-add / slice / sample conversion (audio.rs:1056-1157) with step 8
//                  evaluated on the fly per output sample.
// Window geometry (audio.rs:1056-1073) is recomputed on the device from the three mode bytes of a
// packet; the host has already walked the chain once to catch the OLA guard and size the output.
//
// This is the fallback for everything the fused long-block kernel (kernel_long.cuh) does not take;
// it replaces the four-kernel path (kernels_generic.cuh) wherever channels <= 8 and the buffers fit
// in shared memory, and is ~10-30x faster than it (profiles/r1_sweep_*).
#pragma once
#include "kernels_generic.cuh"

namespace lwb {

struct ChainDesc {
    const DevSetup *setup;
    float *state;              // [channels][state_stride]
    uint64_t coeff_off;        // first packet's [channels][n/2] in the coefficient arena
    uint64_t out_off;          // chain's first PCM element
    uint64_t out_stride;       // planar: elements between channel planes
    uint64_t pkt_index;        // first row in the per-packet floor arenas
    uint32_t n_packets;
    uint32_t byte_off;         // offset of this chain's (mode, prev, next) triples
    uint32_t state_stride;
    uint16_t plen0;            // length of the stream's saved right half when the chain starts
    uint8_t has0;              // 1 if the stream has history
    uint8_t channels;
};

// x[j] of the IMDCT output, from the post-step-7 buffer V (step 8, imdct.rs:589-658):
//   out[m] = p_odd, out[n2-1-m] = -p_odd, out[n2+m] = p_even, out[n-1-m] = p_even,  m < n/4
__device__ __forceinline__ float d_x_at(const float *V, const float *__restrict__ B, int n, int j)
{
    const int n2 = n >> 1, n4 = n >> 2;
    if (j < n2) {
        const bool mir = j >= n4;
        const int m = mir ? n2 - 1 - j : j;
        const int ee = n2 - 2 - 2 * m;
        const float p_odd = __fsub_rn(__fmul_rn(V[ee], __ldg(B + ee + 1)), __fmul_rn(V[ee + 1], __ldg(B + ee)));
        return mir ? -p_odd : p_odd;
    }
    const int jj = j - n2;
    const int m = jj >= n4 ? n2 - 1 - jj : jj;
    const int ee = n2 - 2 - 2 * m;
    return __fsub_rn(__fmul_rn(-V[ee], __ldg(B + ee)), __fmul_rn(V[ee + 1], __ldg(B + ee + 1)));
}

// MULTI = false: one warp per channel (<= 8 warps, compile-time group size, warp-level syncs);
// MULTI = true: `wpc` warps per channel, named barriers.
// np: blocks a channel group transforms together (spectrum entry, !MULTI; the host sizes shared memory for it).
template <int FORMAT, int ENTRY, bool MULTI>
__global__ void __launch_bounds__(MULTI ? 1024 : 256)
k_chain(const ChainDesc *__restrict__ chains, const uint8_t *__restrict__ pkt_bytes, const float *__restrict__ coeffs,
        const float *__restrict__ dense_floor, const uint8_t *__restrict__ floor_kind,
        const uint32_t *__restrict__ floor1_y, void *__restrict__ pcm, int n1max, int wpc, int np)
{
    extern __shared__ float ch_smem[];
    const ChainDesc cd = chains[blockIdx.x];
    const DevSetup &su = *cd.setup;
    const int C = cd.channels;
    const int gt = MULTI ? wpc * 32 : 32;                      // threads per channel group
    const int warp = threadIdx.x / gt, lane = threadIdx.x % gt, W = blockDim.x / gt;   // "warp" = channel group
    const bool active = warp < C;
    auto gsync = [&]() {
        if (!MULTI) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(warp + 1), "r"(gt) : "memory");
    };
    // per channel group: `np` blocks of U | V (n1max floats each), then the previous right half
    const int per_warp = np * n1max + (n1max >> 1);
    float *U = ch_smem + (size_t)warp * per_warp, *prev = U + np * n1max;
    // floor posts of up to 8 channels (residue entry)
    uint16_t *s_x = reinterpret_cast<uint16_t *>(ch_smem + (size_t)W * per_warp);
    uint16_t *s_y = s_x + 8 * (LWB_MAX_POSTS + 1);
    int *s_m = reinterpret_cast<int *>(s_y + 8 * (LWB_MAX_POSTS + 1));

    const int n0 = 1 << su.bs0;
    bool has = cd.has0;
    int plen = cd.plen0;
    if (active && has)
        for (int i = lane; i < plen; i += gt) prev[i] = cd.state[(size_t)warp * cd.state_stride + i];
    uint64_t coeff = cd.coeff_off;
    uint64_t pos = 0;                     // samples per channel emitted so far
    const uint8_t *bytes = pkt_bytes + cd.byte_off;

    for (uint32_t p = 0; p < cd.n_packets;) {
        const int blockflag = su.mode_blockflag[bytes[3 * p]];
        const DevTables &tb = su.tab[blockflag];
        const int n = 1 << tb.bs, n2 = n >> 1;
        // consecutive packets of one blocksize are transformed together (they are independent until the
        // overlap-add): up to `np` blocks in flight per channel group
        uint32_t g = 1;
        if (ENTRY == LWB_ENTRY_SPECTRUM && !MULTI)
            while (g < (uint32_t)np && p + g < cd.n_packets && su.mode_blockflag[bytes[3 * (p + g)]] == blockflag) g++;

        if (ENTRY == LWB_ENTRY_RESIDUE) {
            const int mode = bytes[3 * p];
            const DevMapping &mp = su.mappings[su.mode_mapping[mode]];
            const uint64_t row = (cd.pkt_index + p) * C;
            __syncthreads();              // previous packet finished with U/V and the post arrays
            if (active && lane == 0 && floor_kind[row + warp] == LWB_FLOOR_ONE) {
                const DevFloor1 &fl = su.floors[mp.floor_of_channel[warp]];
                s_m[warp] = d_floor1_posts(fl, floor1_y + (row + warp) * LWB_MAX_POSTS, n2,
                                           s_x + warp * (LWB_MAX_POSTS + 1), s_y + warp * (LWB_MAX_POSTS + 1));
            }
            __syncthreads();
            const int nsteps = mp.n_coupling;
            for (int k = threadIdx.x; k < n2; k += blockDim.x) {
                float r[8];
#pragma unroll
                for (int c = 0; c < 8; c++) r[c] = c < C ? coeffs[coeff + (size_t)c * n2 + k] : 0.f;
                for (int s = nsteps - 1; s >= 0; s--) {       // audio.rs:991-1002
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
                        const int kind = floor_kind[row + c];
                        float f = 0.f;                          // audio.rs:1021-1024
                        if (kind == LWB_FLOOR_ONE)
                            f = c_inverse_db[d_floor1_y_at(s_x + c * (LWB_MAX_POSTS + 1), s_y + c * (LWB_MAX_POSTS + 1),
                                                           s_m[c], k) & 255u];
                        else if (kind == LWB_FLOOR_DENSE)
                            f = dense_floor[coeff + (size_t)c * n2 + k];
                        ch_smem[(size_t)c * per_warp + k] = __fmul_rn(f, r[c]);      // channel c's U
                    }
                }
            }
            __syncthreads();
        }

        if (active) {
            float *V0 = U + (n1max >> 1);
            if (ENTRY == LWB_ENTRY_RESIDUE) {
                d_imdct_to_v(tb, n, U, U, V0, lane, gt, gsync);
            } else {
                const float *X = coeffs + coeff + (size_t)warp * n2;
                const size_t xs = (size_t)C * n2;
                uint32_t q = 0;
                while (q < g) {                                  // pieces of 4, 2, 1 blocks
                    const uint32_t rem = g - q;
                    if (rem >= 4) { d_imdct_to_v_np<4>(tb, n, X + q * xs, xs, U + q * n1max, V0 + q * n1max, n1max, lane, gt, gsync); q += 4; }
                    else if (rem >= 2) { d_imdct_to_v_np<2>(tb, n, X + q * xs, xs, U + q * n1max, V0 + q * n1max, n1max, lane, gt, gsync); q += 2; }
                    else { d_imdct_to_v(tb, n, X + q * xs, U + q * n1max, V0 + q * n1max, lane, gt, gsync); q += 1; }
                }
            }
        }
        for (uint32_t q = 0; q < g; q++) {
            const bool pf = blockflag ? bytes[3 * (p + q) + 1] != 0 : true;       // short blocks: map_or(true, ..)
            const bool nf = blockflag ? bytes[3 * (p + q) + 2] != 0 : true;
            // audio.rs:1056-1073
            const int ls = pf ? 0 : (n - n0) >> 2;
            const int slope_sel = pf ? blockflag : 0;
            const int rs = nf ? n2 : (n * 3 - n0) >> 2;
            const int re = nf ? n : (n * 3 + n0) >> 2;
            if (active) {
                const float *V = U + q * n1max + (n1max >> 1);
                const float *__restrict__ B = tb.b;
                const int olen = rs - ls;
                if (has) {
                    const float *__restrict__ w = su.tab[slope_sel].window;
                    for (int i = lane; i < olen; i += gt) {
                        float v = d_x_at(V, B, n, ls + i);
                        if (i < plen)                                  // audio.rs:1116-1118
                            v = __fadd_rn(__fmul_rn(v, __ldg(w + i)), __fmul_rn(prev[i], __ldg(w + plen - 1 - i)));
                        if (FORMAT == LWB_OUT_F32_PLANAR)
                            ((float *)pcm)[cd.out_off + (size_t)warp * cd.out_stride + pos + i] = v;
                        else if (FORMAT == LWB_OUT_I16_PLANAR)
                            ((int16_t *)pcm)[cd.out_off + (size_t)warp * cd.out_stride + pos + i] = d_sample_i16(v);
                        else if (FORMAT == LWB_OUT_F32_INTERLEAVED)
                            ((float *)pcm)[cd.out_off + (pos + i) * C + warp] = v;
                        else
                            ((int16_t *)pcm)[cd.out_off + (pos + i) * C + warp] = d_sample_i16(v);
                    }
                    gsync();
                }
                plen = re - rs;                                        // audio.rs:1121
                for (int i = lane; i < plen; i += gt) prev[i] = d_x_at(V, B, n, rs + i);
                gsync();
                if (has) pos += olen;
            } else {
                plen = re - rs;
                if (has) pos += rs - ls;
            }
            has = true;
            coeff += (uint64_t)C * n2;
        }
        p += g;
    }
    if (active)
        for (int i = lane; i < plen; i += gt) cd.state[(size_t)warp * cd.state_stride + i] = prev[i];
}

}  // namespace lwb
