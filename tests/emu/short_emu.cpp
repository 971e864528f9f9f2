// short_emu.cpp -- CPU emulation of lewton_b200/csrc/kernel_short.cuh (TEST INFRASTRUCTURE ONLY).
//
// Runs the warp's 32 lanes sequentially, phase by phase, through the same shared-memory transpose (same
// swizzle, same element maps, same twiddle pack), the same octet loop and the same neighbour-lane hand-over of
// the previous right half as the device code, so that `pytest -m "not gpu"` checks the short-block kernel's
// index mathematics and operation order bit-for-bit against the oracle without a GPU.  Reports the worst
// shared-memory bank-conflict degree of the transpose and of the tile reads.  The product never runs this.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../lewton_b200/csrc/kernel_short.cuh"

using namespace lwb;

namespace {
struct TwHostS {
    const V *pack;
    int lane;
    V operator()(int slot) const { return pack[sp_of(slot) * 32 + lane]; }
};
int g_conf = 0;
void note_banks(const int idx[32])
{
    int cnt[32] = {0};
    for (int i = 0; i < 32; i++) cnt[idx[i] & 31]++;
    for (int b = 0; b < 32; b++)
        if (cnt[b] > g_conf) g_conf = cnt[b];
}
// a 128-bit access is served one quarter warp at a time: 8 lanes x 16 bytes must hit 8 distinct 16-byte bank groups
void note_quads(const int byte_addr[32])
{
    for (int q = 0; q < 4; q++) {
        int cnt[8] = {0};
        for (int i = 0; i < 8; i++) cnt[(byte_addr[8 * q + i] >> 4) & 7]++;
        for (int g = 0; g < 8; g++)
            if (cnt[g] > g_conf) g_conf = cnt[g];
    }
}
}  // namespace

extern "C" int lwb_emu_short_pack_floats(void) { return kShortPackFloats; }

extern "C" void lwb_emu_short_build_pack(const float *a, const float *b, const float *c, const float *w, float *pack)
{
    short_build_pack(a, b, c, w, pack);
}

// One run of n_packets short packets of one channel.  spectrum: [n_packets][128]; state: [128] (read if has_prev,
// written at the end); out: emitted packets packed from the front.  Returns the worst bank-conflict degree.
extern "C" int lwb_emu_short_run(const float *pack_f, const float *spectrum, int n_packets, int has_prev, float *state, float *out)
{
    const V *pack = reinterpret_cast<const V *>(pack_f);
    g_conf = 0;
    static V O[32][8], E[32][8], pe[32][8], carry[32][8], podd[32][8];
    std::memset(carry, 0, sizeof(carry));
    const int tile_stride = 576 / 4;                           // floats (kShortTileStride)
    std::vector<float> stage(8 * tile_stride + 1024, 0.f);
    const int koff = has_prev ? 0 : 1;
    const int n_oct = (n_packets + 7) / 8;
    for (int o = 0; o < n_oct; o++) {
        const int nb = n_packets - 8 * o < 8 ? n_packets - 8 * o : 8;
        for (int b = 0; b < nb; b++) std::memcpy(&stage[(size_t)b * tile_stride], spectrum + (size_t)(8 * o + b) * 128, 512);
        int addr[32];
        for (int m = 0; m < 4; m++)
            for (int which = 0; which < 2; which++) {
                for (int lane = 0; lane < 32; lane++) {
                    const int l = lane & 3, b = lane >> 2;
                    addr[lane] = 4 * (b * tile_stride + 4 * (which ? 7 - l + 8 * m : l + 8 * m));
                }
                note_quads(addr);
            }
        for (int lane = 0; lane < 32; lane++)
            phase_a_s(&stage[(size_t)(lane >> 2) * tile_stride], lane & 3, TwHostS{pack, lane}, O[lane], E[lane]);
        float *pe_plane = stage.data(), *po_plane = stage.data() + 512;
        int idx[32];
        for (int j = 0; j < 8; j++)
            for (int h = 0; h < 2; h++) {
                for (int lane = 0; lane < 32; lane++) {
                    const int i = swzS(lane >> 2, elemA_s(lane & 3, j, h));
                    idx[lane] = i;
                    pe_plane[i] = h ? E[lane][j].y : E[lane][j].x;
                    po_plane[i] = h ? O[lane][j].y : O[lane][j].x;
                }
                note_banks(idx);
            }
        for (int j = 0; j < 8; j++)
            for (int h = 0; h < 2; h++) {
                for (int lane = 0; lane < 32; lane++) {
                    const int i = swzS(lane >> 2, elemC_s(lane & 3, j, h));
                    idx[lane] = i;
                    (h ? E[lane][j].y : E[lane][j].x) = pe_plane[i];
                    (h ? O[lane][j].y : O[lane][j].x) = po_plane[i];
                }
                note_banks(idx);
            }
        for (int lane = 0; lane < 32; lane++) {
            const TwHostS tw{pack, lane};
            phase_c_fft<1>(tw, &O[lane], &E[lane]);
            for (int j = 0; j < 8; j++) step8_s(tw(P_B0 + j), tw(P_B1 + j), O[lane][j], E[lane][j], podd[lane][j], pe[lane][j]);
        }
        for (int lane = 0; lane < 32; lane++) {
            const TwHostS tw{pack, lane};
            const int l = lane & 3, b = lane >> 2;
            const int k = 8 * o + b;
            const bool first0 = (o == 0 && b == 0);
            const bool emit = k < n_packets && !(first0 && !has_prev);
            float *ob = out + (ptrdiff_t)(k - koff) * 128;
            for (int j = 0; j < 8; j++) {
                const int mx = outIndex_s(l, j, 0), my = outIndex_s(l, j, 1);
                V plo = b ? pe[lane - 4][j] : carry[28 + l][j], phi = plo;
                if (first0 && has_prev) {
                    plo = V{state[mx], state[my]};
                    phi = V{state[127 - mx], state[127 - my]};
                }
                V lo, hi;
                ola_s(podd[lane][j], tw(P_WLO + j), tw(P_WHI + j), plo, phi, lo, hi);
                if (emit) {
                    ob[mx] = lo.x; ob[my] = lo.y;
                    ob[127 - mx] = hi.x; ob[127 - my] = hi.y;
                }
            }
        }
        std::memcpy(carry, pe, sizeof(pe));
    }
    const int bl = (n_packets - 1) & 7;
    for (int lane = 4 * bl; lane < 4 * bl + 4; lane++)
        for (int j = 0; j < 8; j++) {
            const int mx = outIndex_s(lane & 3, j, 0), my = outIndex_s(lane & 3, j, 1);
            state[mx] = pe[lane][j].x; state[my] = pe[lane][j].y;
            state[127 - mx] = pe[lane][j].x; state[127 - my] = pe[lane][j].y;
        }
    return g_conf;
}

// Worst bank-conflict degree of the PCM staging of kernel_short.cuh (esz = 4: f32, 2: i16): the 32 stores of a
// (slot, value) and the per-packet vector loads.  Mirrors the address arithmetic of k_short (wP0 / wP1 ^ chunk).
extern "C" int lwb_emu_short_staging_conflicts(int esz)
{
    int worst = 0;
    auto banks = [&](const int addr[32], int bytes) {
        // an access of `bytes` per lane is served in groups of 128 / bytes... conservatively: count distinct 32-bit words per bank
        int lanes_per_phase = bytes >= 16 ? 8 : bytes == 8 ? 16 : 32;
        for (int ph = 0; ph < 32 / lanes_per_phase; ph++) {
            int cnt[32] = {0}, word[32][8] = {{0}}, nw[32] = {0};
            for (int i = 0; i < lanes_per_phase; i++) {
                const int a = addr[ph * lanes_per_phase + i];
                for (int w = a >> 2; w < (a + (bytes > 4 ? bytes : 4) + 3) >> 2 && w <= (a + bytes - 1) >> 2; w++) {
                    const int bk = w & 31;
                    bool seen = false;
                    for (int q = 0; q < nw[bk]; q++) seen |= word[bk][q] == w;
                    if (!seen && nw[bk] < 8) { word[bk][nw[bk]++] = w; cnt[bk]++; }
                }
            }
            for (int b = 0; b < 32; b++) worst = cnt[b] > worst ? cnt[b] : worst;
        }
    };
    const int CH = 4 * esz;
    for (int j = 0; j < 8; j++) {
        const int r = rev3(j);
        const int cA = CH * (2 * r), cB = CH * (2 * r + 1), cC = CH * (31 - 2 * r), cD = CH * (30 - 2 * r);
        const int consts[4] = {cA, cB, cC, cD};
        const int which[4] = {0, 1, 1, 0};               // p0 or p1
        for (int v = 0; v < 4; v++) {
            int addr[32];
            for (int lane = 0; lane < 32; lane++) {
                const int l = lane & 3, b = lane >> 2;
                const int p = (128 * esz + 4 * esz) * b + esz * (which[v] ? 3 - l : l);
                addr[lane] = p ^ consts[v];
                // the staged sample must land where the copy-out expects it: row b, chunk (m >> 2) ^ b, position m & 3
                const int m_nat = v == 0 ? 8 * r + l : v == 1 ? 8 * r + 7 - l : v == 2 ? 127 - (8 * r + l) : 127 - (8 * r + 7 - l);
                const int want = 128 * esz * b + 4 * esz * (((m_nat >> 2) ^ b)) + esz * (m_nat & 3);
                if (addr[lane] != want) return -1;
            }
            banks(addr, esz);
        }
    }
    for (int i = 0; i < 8; i++) {
        int addr[32];
        for (int lane = 0; lane < 32; lane++) addr[lane] = 128 * esz * i + 4 * esz * (lane ^ i);
        banks(addr, 4 * esz);
    }
    return worst;
}
