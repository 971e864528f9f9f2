// mid_emu.cpp -- CPU emulation of lewton_b200/csrc/kernel_mid.cuh (TEST INFRASTRUCTURE ONLY): the warp's 32 lanes run
// sequentially through the kernel's phase functions, transposes, element maps and twiddle pack; see long_emu.cpp.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../lewton_b200/csrc/kernel_mid.cuh"

using namespace lwb;

namespace {
struct TwHost {
    const V *pack;
    int lane;
    V operator()(int slot) const { return pack[slot * 32 + lane]; }
};
int g_max_conflict = 0;
void note_banks(const int idx[32])
{
    int cnt[32] = {0};
    for (int l = 0; l < 32; l++) cnt[idx[l] & 31]++;
    for (int b = 0; b < 32; b++)
        if (cnt[b] > g_max_conflict) g_max_conflict = cnt[b];
}
}  // namespace

extern "C" int lwb_emu_mid_pack_floats(void) { return kLongPackFloats; }
extern "C" void lwb_emu_mid_build_pack(int kb, const float *a, const float *b, const float *c, const float *w, float *pack)
{
    if (kb == 1) mid_build_pack<1>(a, b, c, w, pack);
    else mid_build_pack<2>(a, b, c, w, pack);
}

// NB = 2^KB runs of n_packets packets in lockstep.  spectrum: [NB][n_packets][N2]; state: [NB][N2] (read if has_prev[b],
// written at the end); out: [NB][n_packets][N2] (emitted packets packed from the front).  Returns the worst bank-conflict
// degree of the transposes.
template <int KB>
static int emu_run(const float *pack_f, const float *spectrum, int n_packets, const int *has_prev, float *state, float *out_all)
{
    using M = Mid<KB>;
    const V *pack = reinterpret_cast<const V *>(pack_f);
    g_max_conflict = 0;
    std::vector<float> tiles(1024);                   // NB tiles of N2 floats = the E | O planes of the transposes
    static V O[32][8], E[32][8], pe[32][8];
    std::memset(pe, 0, sizeof(pe));
    float *out[M::NB];
    for (int b = 0; b < M::NB; b++) out[b] = out_all + (size_t)b * n_packets * M::N2;
    for (int p = 0; p < n_packets; p++) {
        const float *tp[M::NB];
        for (int b = 0; b < M::NB; b++) {
            std::memcpy(&tiles[(size_t)b * M::N2], spectrum + ((size_t)b * n_packets + p) * M::N2, M::N2 * 4);
            tp[b] = &tiles[(size_t)b * M::N2];
        }
        for (int lane = 0; lane < 32; lane++) phase_a_m<KB>(tp, lane, TwHost{pack, lane}, O[lane], E[lane]);
        int idx[32];
        float *pe_plane = &tiles[0], *po_plane = pe_plane + 512;
        for (int j = 0; j < 8; j++)
            for (int h = 0; h < 2; h++) {
                for (int lane = 0; lane < 32; lane++) {
                    const int i = swz(elemA(lane, j, h));
                    idx[lane] = i;
                    pe_plane[i] = h ? E[lane][j].y : E[lane][j].x;
                    po_plane[i] = h ? O[lane][j].y : O[lane][j].x;
                }
                note_banks(idx);
            }
        for (int j = 0; j < 8; j++)
            for (int h = 0; h < 2; h++) {
                for (int lane = 0; lane < 32; lane++) {
                    const int i = swz(elemB(lane, j, h));
                    idx[lane] = i;
                    (h ? E[lane][j].y : E[lane][j].x) = pe_plane[i];
                    (h ? O[lane][j].y : O[lane][j].x) = po_plane[i];
                }
                note_banks(idx);
            }
        for (int lane = 0; lane < 32; lane++) {
            V (*Ob)[8] = reinterpret_cast<V (*)[8]>(O[lane]);
            V (*Eb)[8] = reinterpret_cast<V (*)[8]>(E[lane]);
            phase_b<1>(TwHost{pack, lane}, Ob, Eb);
        }
        for (int j = 0; j < 8; j++)
            for (int h = 0; h < 2; h++)
                for (int lane = 0; lane < 32; lane++) {
                    const int i = swz(elemB(lane, j, h));
                    pe_plane[i] = h ? E[lane][j].y : E[lane][j].x;
                    po_plane[i] = h ? O[lane][j].y : O[lane][j].x;
                }
        for (int j = 0; j < 8; j++)
            for (int h = 0; h < 2; h++) {
                for (int lane = 0; lane < 32; lane++) {
                    const int i = swz(elemC_m<KB>(lane, j, h));
                    idx[lane] = i;
                    (h ? E[lane][j].y : E[lane][j].x) = pe_plane[i];
                    (h ? O[lane][j].y : O[lane][j].x) = po_plane[i];
                }
                note_banks(idx);
            }
        for (int lane = 0; lane < 32; lane++) {
            const TwHost tw{pack, lane};
            V (*Ob)[8] = reinterpret_cast<V (*)[8]>(O[lane]);
            V (*Eb)[8] = reinterpret_cast<V (*)[8]>(E[lane]);
            phase_c_fft<1>(tw, Ob, Eb);
            const int b = blockC_m<KB>(lane);
            for (int j = 0; j < 8; j++) {
                const int mx = outIndex_m<KB>(lane, j, 0), my = outIndex_m<KB>(lane, j, 1);
                const V b0 = tw(P_B0 + j), b1 = tw(P_B1 + j), wlo = tw(P_WLO + j), whi = tw(P_WHI + j);
                const bool emit = p > 0 || has_prev[b];
                const bool from_state = p == 0 && has_prev[b];
                const float *st = state + (size_t)b * M::N2;
                V plo = pe[lane][j], phi = pe[lane][j];
                if (from_state) {
                    plo = V{st[mx], st[my]};
                    phi = V{st[M::N2 - 1 - mx], st[M::N2 - 1 - my]};
                }
                V lo, hi, pev;
                step8_ola(b0, b1, wlo, whi, O[lane][j], E[lane][j], plo, phi, lo, hi, pev);
                pe[lane][j] = pev;
                if (emit) {
                    out[b][mx] = lo.x; out[b][my] = lo.y;
                    out[b][M::N2 - 1 - mx] = hi.x; out[b][M::N2 - 1 - my] = hi.y;
                }
            }
        }
        for (int b = 0; b < M::NB; b++)
            if (p > 0 || has_prev[b]) out[b] += M::N2;
    }
    for (int lane = 0; lane < 32; lane++)
        for (int j = 0; j < 8; j++) {
            const int b = blockC_m<KB>(lane);
            const int mx = outIndex_m<KB>(lane, j, 0), my = outIndex_m<KB>(lane, j, 1);
            float *st = state + (size_t)b * M::N2;
            st[mx] = pe[lane][j].x; st[my] = pe[lane][j].y;
            st[M::N2 - 1 - mx] = pe[lane][j].x; st[M::N2 - 1 - my] = pe[lane][j].y;
        }
    return g_max_conflict;
}

extern "C" int lwb_emu_mid_run(int kb, const float *pack_f, const float *spectrum, int n_packets, const int *has_prev, float *state,
                               float *out_all)
{
    return kb == 1 ? emu_run<1>(pack_f, spectrum, n_packets, has_prev, state, out_all)
                   : emu_run<2>(pack_f, spectrum, n_packets, has_prev, state, out_all);
}
