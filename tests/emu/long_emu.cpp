// long_emu.cpp -- CPU emulation of lewton_b200/csrc/kernel_long.cuh (TEST INFRASTRUCTURE ONLY).
//
// The fused kernel's per-lane phase functions are plain inline functions that also compile for
// the host.  This harness runs the warp's 32 lanes sequentially, phase by phase, through the
// same shared-memory transposes (same swizzle, same element maps, same twiddle pack) and the
// same packet loop as the device code, so that `pytest -m "not gpu"` can check the kernel's
// index mathematics and operation order bit-for-bit against the oracle without a GPU.  It also
// reports the worst shared-memory bank-conflict degree of every transpose access.
// The product library never contains or calls this code.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../lewton_b200/csrc/kernel_long.cuh"

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

extern "C" int lwb_emu_pack_floats(void) { return kLongPackFloats; }

extern "C" void lwb_emu_build_pack(const float *a, const float *b, const float *c, const float *w, float *pack)
{
    long_build_pack(a, b, c, w, pack);
}

// One run (NB = 1) or two runs in lockstep (NB = 2, both of n_packets packets), exactly like a
// warp of k_long.  spectrum: [NB][n_packets][1024]; state: [NB][1024] (read if has_prev[b], written
// at the end); out: [NB][n_packets][1024] (emitted packets packed from the front).  Returns the
// worst bank-conflict degree seen (1 = conflict-free).
template <int NB>
static int emu_run(const float *pack_f, const float *spectrum, int n_packets, const int *has_prev, float *state,
                   float *out_all)
{
    const V *pack = reinterpret_cast<const V *>(pack_f);
    g_max_conflict = 0;
    std::vector<float> tiles((size_t)NB * 1024);
    static V O[32][NB][8], E[32][NB][8], pe[32][NB][8];
    std::memset(pe, 0, sizeof(pe));
    float *out[NB];
    for (int b = 0; b < NB; b++) out[b] = out_all + (size_t)b * n_packets * 1024;
    for (int p = 0; p < n_packets; p++) {
        const float *tp[NB];
        for (int b = 0; b < NB; b++) {
            std::memcpy(&tiles[(size_t)b * 1024], spectrum + ((size_t)b * n_packets + p) * 1024, 4096);
            tp[b] = &tiles[(size_t)b * 1024];
        }
        for (int lane = 0; lane < 32; lane++) phase_a<NB>(tp, lane, TwHost{pack, lane}, O[lane], E[lane]);
        int idx[32];
        for (int b = 0; b < NB; b++) {
            float *pe_plane = &tiles[(size_t)b * 1024], *po_plane = pe_plane + 512;
            for (int j = 0; j < 8; j++)
                for (int h = 0; h < 2; h++) {
                    for (int lane = 0; lane < 32; lane++) {
                        const int i = swz(elemA(lane, j, h));
                        idx[lane] = i;
                        pe_plane[i] = h ? E[lane][b][j].y : E[lane][b][j].x;
                        po_plane[i] = h ? O[lane][b][j].y : O[lane][b][j].x;
                    }
                    note_banks(idx);
                }
            for (int j = 0; j < 8; j++)
                for (int h = 0; h < 2; h++) {
                    for (int lane = 0; lane < 32; lane++) {
                        const int i = swz(elemB(lane, j, h));
                        idx[lane] = i;
                        (h ? E[lane][b][j].y : E[lane][b][j].x) = pe_plane[i];
                        (h ? O[lane][b][j].y : O[lane][b][j].x) = po_plane[i];
                    }
                    note_banks(idx);
                }
        }
        for (int lane = 0; lane < 32; lane++) phase_b<NB>(TwHost{pack, lane}, O[lane], E[lane]);
        for (int b = 0; b < NB; b++) {
            float *pe_plane = &tiles[(size_t)b * 1024], *po_plane = pe_plane + 512;
            for (int j = 0; j < 8; j++)
                for (int h = 0; h < 2; h++)
                    for (int lane = 0; lane < 32; lane++) {
                        const int i = swz(elemB(lane, j, h));
                        pe_plane[i] = h ? E[lane][b][j].y : E[lane][b][j].x;
                        po_plane[i] = h ? O[lane][b][j].y : O[lane][b][j].x;
                    }
            for (int j = 0; j < 8; j++)
                for (int h = 0; h < 2; h++) {
                    for (int lane = 0; lane < 32; lane++) {
                        const int i = swz(elemC(lane, j, h));
                        idx[lane] = i;
                        (h ? E[lane][b][j].y : E[lane][b][j].x) = pe_plane[i];
                        (h ? O[lane][b][j].y : O[lane][b][j].x) = po_plane[i];
                    }
                    note_banks(idx);
                }
        }
        for (int lane = 0; lane < 32; lane++) {
            const TwHost tw{pack, lane};
            phase_c_fft<NB>(tw, O[lane], E[lane]);
            for (int j = 0; j < 8; j++) {
                const int mx = outIndex(lane, j, 0), my = outIndex(lane, j, 1);
                const V b0 = tw(P_B0 + j), b1 = tw(P_B1 + j), wlo = tw(P_WLO + j), whi = tw(P_WHI + j);
                for (int b = 0; b < NB; b++) {
                    const bool emit = p > 0 || has_prev[b];
                    const bool from_state = p == 0 && has_prev[b];
                    const float *st = state + (size_t)b * 1024;
                    V plo = pe[lane][b][j], phi = pe[lane][b][j];
                    if (from_state) {
                        plo = V{st[mx], st[my]};
                        phi = V{st[1023 - mx], st[1023 - my]};
                    }
                    V lo, hi, pev;
                    step8_ola(b0, b1, wlo, whi, O[lane][b][j], E[lane][b][j], plo, phi, lo, hi, pev);
                    pe[lane][b][j] = pev;
                    if (emit) {
                        out[b][mx] = lo.x; out[b][my] = lo.y;
                        out[b][1023 - mx] = hi.x; out[b][1023 - my] = hi.y;
                    }
                }
            }
        }
        for (int b = 0; b < NB; b++)
            if (p > 0 || has_prev[b]) out[b] += 1024;
    }
    for (int b = 0; b < NB; b++)
        for (int lane = 0; lane < 32; lane++)
            for (int j = 0; j < 8; j++) {
                const int mx = outIndex(lane, j, 0), my = outIndex(lane, j, 1);
                float *st = state + (size_t)b * 1024;
                st[mx] = pe[lane][b][j].x; st[my] = pe[lane][b][j].y;
                st[1023 - mx] = pe[lane][b][j].x; st[1023 - my] = pe[lane][b][j].y;
            }
    return g_max_conflict;
}

extern "C" int lwb_emu_long_run(const float *pack_f, const float *spectrum, int n_packets, int has_prev,
                                float *state, float *out)
{
    return emu_run<1>(pack_f, spectrum, n_packets, &has_prev, state, out);
}

extern "C" int lwb_emu_long_run2(const float *pack_f, const float *spectrum, int n_packets, const int *has_prev,
                                 float *state, float *out)
{
    return emu_run<2>(pack_f, spectrum, n_packets, has_prev, state, out);
}
