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

// spectrum: [n_packets][1024]; state: 1024 floats (read if has_prev, written at the end);
// out: [emitted][1024].  Returns the worst bank-conflict degree seen (1 = conflict-free).
extern "C" int lwb_emu_long_run(const float *pack_f, const float *spectrum, int n_packets, int has_prev,
                                float *state, float *out)
{
    const V *pack = reinterpret_cast<const V *>(pack_f);
    g_max_conflict = 0;
    std::vector<float> tile(1024);
    V O[32][8], E[32][8], pe[32][8];
    std::memset(pe, 0, sizeof(pe));
    for (int p = 0; p < n_packets; p++) {
        std::memcpy(tile.data(), spectrum + (size_t)p * 1024, 4096);
        for (int lane = 0; lane < 32; lane++) {
            Q4 F1[4], F2[4];
            for (int m = 0; m < 4; m++) {
                const float *u = &tile[4 * (lane + 64 * m)], *v = &tile[4 * (63 - lane + 64 * m)];
                F1[m] = Q4{u[0], u[1], u[2], u[3]};
                F2[m] = Q4{v[0], v[1], v[2], v[3]};
            }
            phase_a(F1, F2, TwHost{pack, lane}, O[lane], E[lane]);
        }
        float *pe_plane = tile.data(), *po_plane = tile.data() + 512;
        int idx[32];
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
        for (int lane = 0; lane < 32; lane++) phase_b(TwHost{pack, lane}, O[lane], E[lane]);
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
                    const int i = swz(elemC(lane, j, h));
                    idx[lane] = i;
                    (h ? E[lane][j].y : E[lane][j].x) = pe_plane[i];
                    (h ? O[lane][j].y : O[lane][j].x) = po_plane[i];
                }
                note_banks(idx);
            }
        const bool emit = p > 0 || has_prev;
        const bool from_state = p == 0 && has_prev;
        for (int lane = 0; lane < 32; lane++) {
            const TwHost tw{pack, lane};
            phase_c_fft(tw, O[lane], E[lane]);
            for (int j = 0; j < 8; j++) {
                const int mx = outIndex(lane, j, 0), my = outIndex(lane, j, 1);
                V plo = pe[lane][j], phi = pe[lane][j];
                if (from_state) {
                    plo = V{state[mx], state[my]};
                    phi = V{state[1023 - mx], state[1023 - my]};
                }
                V lo, hi, pev;
                phase_c_out(tw, j, O[lane][j], E[lane][j], plo, phi, lo, hi, pev);
                pe[lane][j] = pev;
                if (emit) {
                    out[mx] = lo.x; out[my] = lo.y;
                    out[1023 - mx] = hi.x; out[1023 - my] = hi.y;
                }
            }
        }
        if (emit) out += 1024;
    }
    for (int lane = 0; lane < 32; lane++)
        for (int j = 0; j < 8; j++) {
            const int mx = outIndex(lane, j, 0), my = outIndex(lane, j, 1);
            state[mx] = pe[lane][j].x; state[my] = pe[lane][j].y;
            state[1023 - mx] = pe[lane][j].x; state[1023 - my] = pe[lane][j].y;
        }
    return g_max_conflict;
}
