// lwb_api.cu -- the C ABI (include/lewton_b200.h): context, setup, stream state and batch
// submission.  Host logic mirrors the control flow of lewton's read_audio_packet_generic back
// half (src/audio.rs:988-1157): which window shape a packet has, whether a previous right half
// exists, what the packet returns -- all of that is decided here on the host from the mode bits
// (it never depends on sample values), so the kernels receive fully resolved descriptors and the
// device never has to be synchronised to learn a length.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "kernel_long.cuh"
#include "kernels_generic.cuh"
#include "kernel_chain.cuh"
#include "lwb_common.h"

namespace lwb {
int generate_tables(int bs, float *a, float *b, float *c, float *window, uint32_t *bitrev);
int prepare_floor1(const lwb_floor_desc &d, DevFloor1 *out);
}  // namespace lwb

using namespace lwb;

// ---------------------------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct lwb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_in[64] = {}, ev_done[65] = {};      // per chunk of a host-memory batch; [64] orders the copy streams
    // fused path: descriptor arrays are double buffered and uploaded on the copy stream so that the
    // upload of step k+1 overlaps kernel k; tickets come from a pool zeroed once per wrap
    DevBuf runs_buf[2];
    cudaEvent_t ev_desc[2] = {}, ev_kdone[2] = {};
    int runs_par = 0;
    uint32_t ticket_next = 0;
    uint64_t state_gen = 1;        // bumped whenever any stream's (has, len) changes: plans key on it
    std::string err;
    uint64_t launches = 0;
    // grow-only device arenas
    DevBuf coeffs, dense, pcm, spec, x, desc, kinds, ys, chains, ticket, cdesc, cbytes;
    // pinned staging for descriptors
    void *h_desc = nullptr;
    size_t h_desc_cap = 0;
    size_t x_cap_elems = (size_t)64 << 20;     // IMDCT scratch per round of the generic path (256 MiB)
};

struct lwb_setup {
    lwb_ctx *ctx = nullptr;
    DevSetup host;                 // device pointers inside
    DevSetup *d_setup = nullptr;
    std::vector<void *> allocs;
    uint8_t channels = 0, bs0 = 0, bs1 = 0;
    uint32_t n_modes = 0;
    uint32_t n_mappings = 0;
    std::vector<DevMapping> mappings;   // host copy (validation)
};

struct MixRound { size_t r0, nr, c0, nc; };
struct MixLaunch {
    char *db; size_t off_cd, off_by;
    const float *pack, *w_short; int ls; bool i16, residue; int out_format; unsigned warps; size_t smem; int n1max, wpc, np;
    const float *coeffs, *dense; const uint8_t *kinds; const uint32_t *ys; void *pcm;
};

struct lwb_plan {
    lwb_ctx *ctx = nullptr;
    lwb_chain *chains = nullptr;
    size_t n_chains = 0;
    lwb_batch_io io;
    // captured fused-path launch (valid while ctx->state_gen == gen)
    bool captured = false;
    uint64_t gen = 0;
    DevBuf runs;
    uint32_t n_groups = 0;
    const float *pack = nullptr;
    bool i16 = false;
    // captured mixed-path launch sequence (valid while ctx->state_gen == gen)
    bool mixed_captured = false;
    DevBuf mix;
    MixLaunch mix_launch;
    std::vector<MixRound> mix_rounds;
};

struct lwb_stream {
    lwb_ctx *ctx = nullptr;
    const lwb_setup *setup = nullptr;
    float *d_state = nullptr;      // [channels][n1/2]
    bool has = false;              // PreviousWindowRight.data.is_some()
    uint32_t plen = 0;             // per-channel length of the saved right half
    uint64_t busy_epoch = 0;       // guards against one stream appearing twice in a batch
};

static inline void set_stream_state(lwb_stream *s, bool has, uint32_t plen)
{
    if (s->has != has || s->plen != plen) {
        s->has = has;
        s->plen = plen;
        s->ctx->state_gen++;
    }
}

static int fail(lwb_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (ctx) {
        ctx->err = what;
        if (e != cudaSuccess) {
            ctx->err += ": ";
            ctx->err += cudaGetErrorString(e);
        }
    }
    return code;
}

#define CU(ctx, call)                                                        \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return fail((ctx), LWB_ERR_CUDA, #call, e__); \
    } while (0)

static int ensure(lwb_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return LWB_OK;
    if (b.p) {
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        CU(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 4096;
    CU(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return LWB_OK;
}

static int ensure_pinned(lwb_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->h_desc_cap) return LWB_OK;
    if (ctx->h_desc) {
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFreeHost(ctx->h_desc);
        ctx->h_desc = nullptr;
        ctx->h_desc_cap = 0;
    }
    size_t want = bytes * 2 + 4096;
    CU(ctx, cudaHostAlloc(&ctx->h_desc, want, cudaHostAllocDefault));
    ctx->h_desc_cap = want;
    return LWB_OK;
}

// ---------------------------------------------------------------------------------------------
// library / context
// ---------------------------------------------------------------------------------------------
extern "C" int lwb_abi_version(void) { return LWB_ABI_VERSION; }

extern "C" int lwb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int lwb_ctx_create(int device, lwb_ctx **out)
{
    if (!out) return LWB_ERR_INVALID;
    *out = nullptr;
    int n = lwb_device_count();
    if (n <= 0 || device < 0 || device >= n) return LWB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LWB_ERR_NO_DEVICE;
    if (prop.major != 10) return LWB_ERR_NO_DEVICE;       // kernels are built for sm_100a only
    lwb_ctx *ctx = new (std::nothrow) lwb_ctx();
    if (!ctx) return LWB_ERR_BUFFER;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return LWB_ERR_CUDA;
    }
    if (const char *e = getenv("LWB_SCRATCH_MB")) {
        long mb = atol(e);
        if (mb >= 1) ctx->x_cap_elems = (size_t)mb << 18;
    }
    long_kernel_configure();
    *out = ctx;
    return LWB_OK;
}

extern "C" void lwb_ctx_destroy(lwb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (DevBuf *b : {&ctx->coeffs, &ctx->dense, &ctx->pcm, &ctx->spec, &ctx->x, &ctx->desc,
                      &ctx->kinds, &ctx->ys, &ctx->chains, &ctx->ticket, &ctx->runs_buf[0], &ctx->runs_buf[1],
                      &ctx->cdesc, &ctx->cbytes})
        if (b->p) cudaFree(b->p);
    if (ctx->h_desc) cudaFreeHost(ctx->h_desc);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_in);
    cudaStreamDestroy(ctx->copy_out);
    delete ctx;
}

extern "C" int lwb_ctx_synchronize(lwb_ctx *ctx)
{
    if (!ctx) return LWB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return LWB_OK;
}

extern "C" const char *lwb_last_error(const lwb_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }
extern "C" void *lwb_ctx_cuda_stream(lwb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" uint64_t lwb_ctx_launch_count(const lwb_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" void *lwb_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void lwb_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" int lwb_device_alloc(lwb_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return LWB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMalloc(out, bytes ? bytes : 1));
    return LWB_OK;
}
extern "C" void lwb_device_free(lwb_ctx *ctx, void *p)
{
    if (!ctx || !p) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(p);
}
extern "C" int lwb_memcpy_h2d(lwb_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return LWB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return LWB_OK;
}
extern "C" int lwb_memcpy_d2h(lwb_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return LWB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return LWB_OK;
}

extern "C" int lwb_tables_generate(int bs, float *a, float *b, float *c, float *window, uint32_t *bitrev)
{
    return generate_tables(bs, a, b, c, window, bitrev);
}

// ---------------------------------------------------------------------------------------------
// setup
// ---------------------------------------------------------------------------------------------
template <typename T>
static int upload(lwb_setup *su, const T *host, size_t count, const T **dev)
{
    void *p = nullptr;
    lwb_ctx *ctx = su->ctx;
    CU(ctx, cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16)));
    su->allocs.push_back(p);
    if (count) CU(ctx, cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *dev = (const T *)p;
    return LWB_OK;
}

extern "C" void lwb_setup_destroy(lwb_setup *su)
{
    if (!su) return;
    cudaSetDevice(su->ctx->device);
    cudaStreamSynchronize(su->ctx->stream);
    for (void *p : su->allocs) cudaFree(p);
    delete su;
}

extern "C" int lwb_setup_create(lwb_ctx *ctx, const lwb_setup_desc *d, lwb_setup **out)
{
    if (!ctx || !d || !out) return LWB_ERR_INVALID;
    *out = nullptr;
    // header.rs:239-243 (blocksizes, channels)
    if (d->blocksize_0 < 6 || d->blocksize_0 > 13 || d->blocksize_1 < 6 || d->blocksize_1 > 13 ||
        d->blocksize_0 > d->blocksize_1 || d->audio_channels == 0)
        return fail(ctx, LWB_ERR_BAD_FORMAT, "setup: blocksizes/channels out of range");
    if (d->n_modes == 0 || d->n_modes > LWB_MAX_MODES || d->n_mappings == 0 || d->n_mappings > 64 ||
        d->n_floors == 0 || d->n_floors > 64 || !d->modes || !d->mappings || !d->floors)
        return fail(ctx, LWB_ERR_INVALID, "setup: counts out of range");
    CU(ctx, cudaSetDevice(ctx->device));
    lwb_setup *su = new (std::nothrow) lwb_setup();
    if (!su) return LWB_ERR_BUFFER;
    su->ctx = ctx;
    su->channels = d->audio_channels;
    su->bs0 = d->blocksize_0;
    su->bs1 = d->blocksize_1;
    su->n_modes = d->n_modes;
    su->n_mappings = d->n_mappings;
    std::memset(&su->host, 0, sizeof(su->host));
    int rc = LWB_OK;
    // tables (header_cached.rs:33-41): the caller's own, or generated here
    for (int i = 0; i < 2 && rc == LWB_OK; i++) {
        const int bs = i ? d->blocksize_1 : d->blocksize_0;
        const size_t n = (size_t)1 << bs;
        std::vector<float> a(n / 2), b(n / 2), c(n / 4), w(n / 2);
        std::vector<uint32_t> br(n / 8);
        const lwb_tables_ref &t = d->tables[i];
        if (t.a) {
            if (!t.b || !t.c || !t.window || !t.bitrev) { rc = LWB_ERR_INVALID; break; }
            std::copy(t.a, t.a + n / 2, a.begin());
            std::copy(t.b, t.b + n / 2, b.begin());
            std::copy(t.c, t.c + n / 4, c.begin());
            std::copy(t.window, t.window + n / 2, w.begin());
            std::copy(t.bitrev, t.bitrev + n / 8, br.begin());
        } else {
            generate_tables(bs, a.data(), b.data(), c.data(), w.data(), br.data());
        }
        DevTables &dt = su->host.tab[i];
        dt.bs = bs;
        if ((rc = upload(su, a.data(), a.size(), &dt.a)) || (rc = upload(su, b.data(), b.size(), &dt.b)) ||
            (rc = upload(su, c.data(), c.size(), &dt.c)) || (rc = upload(su, w.data(), w.size(), &dt.window)) ||
            (rc = upload(su, br.data(), br.size(), &dt.bitrev)))
            break;
        dt.pack = nullptr;
        if (bs == kLongBs) {
            std::vector<float> pack(kLongPackFloats);
            long_build_pack(a.data(), b.data(), c.data(), w.data(), pack.data());
            if ((rc = upload(su, pack.data(), pack.size(), &dt.pack))) break;
        }
        // the synchronous copies above read from vectors that die here
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = LWB_ERR_CUDA;
    }
    std::vector<DevFloor1> floors(d->n_floors);
    for (uint32_t i = 0; i < d->n_floors && rc == LWB_OK; i++) rc = prepare_floor1(d->floors[i], &floors[i]);
    su->mappings.resize(d->n_mappings);
    for (uint32_t i = 0; i < d->n_mappings && rc == LWB_OK; i++) {
        const lwb_mapping_desc &m = d->mappings[i];
        DevMapping &dm = su->mappings[i];
        std::memset(&dm, 0, sizeof(dm));
        if (m.coupling_steps > LWB_MAX_COUPLING || m.submaps == 0 || m.submaps > LWB_MAX_SUBMAPS) {
            rc = LWB_ERR_BAD_FORMAT;
            break;
        }
        dm.n_coupling = m.coupling_steps;
        for (int s = 0; s < m.coupling_steps; s++) {
            // header.rs:1006-1011
            if (m.magnitudes[s] == m.angles[s] || m.magnitudes[s] >= d->audio_channels ||
                m.angles[s] >= d->audio_channels) {
                rc = LWB_ERR_BAD_FORMAT;
                break;
            }
            dm.mag[s] = m.magnitudes[s];
            dm.ang[s] = m.angles[s];
        }
        for (int c = 0; c < d->audio_channels && rc == LWB_OK; c++) {
            if (m.mux[c] >= m.submaps || m.submap_floors[m.mux[c]] >= d->n_floors) {
                rc = LWB_ERR_BAD_FORMAT;      // header.rs:1023-1026, 1043-1047
                break;
            }
            dm.floor_of_channel[c] = m.submap_floors[m.mux[c]];
        }
    }
    for (uint32_t i = 0; i < d->n_modes && rc == LWB_OK; i++) {
        if (d->modes[i].mapping >= d->n_mappings) { rc = LWB_ERR_BAD_FORMAT; break; }   // header.rs:1067-1072
        su->host.mode_blockflag[i] = d->modes[i].blockflag ? 1 : 0;
        su->host.mode_mapping[i] = d->modes[i].mapping;
    }
    if (rc == LWB_OK) rc = upload(su, floors.data(), floors.size(), &su->host.floors);
    if (rc == LWB_OK) rc = upload(su, su->mappings.data(), su->mappings.size(), &su->host.mappings);
    su->host.channels = d->audio_channels;
    su->host.bs0 = d->blocksize_0;
    su->host.bs1 = d->blocksize_1;
    su->host.n_floors = (uint8_t)d->n_floors;
    if (rc == LWB_OK) {
        const DevSetup *dp = nullptr;
        rc = upload(su, &su->host, 1, &dp);
        su->d_setup = const_cast<DevSetup *>(dp);
    }
    if (rc == LWB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = LWB_ERR_CUDA;
    if (rc != LWB_OK) {
        lwb_setup_destroy(su);
        if (ctx->err.empty() || rc != LWB_ERR_CUDA) ctx->err = "setup: rejected (see header.rs validation rules)";
        return rc;
    }
    *out = su;
    return LWB_OK;
}

// ---------------------------------------------------------------------------------------------
// stream state
// ---------------------------------------------------------------------------------------------
static size_t state_stride(const lwb_setup *su) { return (size_t)1 << (su->bs1 - 1); }

extern "C" int lwb_stream_open(lwb_ctx *ctx, const lwb_setup *su, lwb_stream **out)
{
    if (!ctx || !su || !out || su->ctx != ctx) return LWB_ERR_INVALID;
    CU(ctx, cudaSetDevice(ctx->device));
    lwb_stream *s = new (std::nothrow) lwb_stream();
    if (!s) return LWB_ERR_BUFFER;
    s->ctx = ctx;
    s->setup = su;
    cudaError_t e = cudaMalloc((void **)&s->d_state, su->channels * state_stride(su) * sizeof(float));
    if (e != cudaSuccess) {
        delete s;
        return fail(ctx, LWB_ERR_CUDA, "stream_open: cudaMalloc", e);
    }
    *out = s;
    return LWB_OK;
}

extern "C" void lwb_stream_destroy(lwb_stream *s)
{
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    cudaFree(s->d_state);
    delete s;
}

extern "C" int lwb_stream_reset(lwb_stream *s)
{
    if (!s) return LWB_ERR_INVALID;
    set_stream_state(s, false, 0);
    return LWB_OK;
}
extern "C" int lwb_stream_is_empty(const lwb_stream *s) { return (!s || !s->has) ? 1 : 0; }
extern "C" uint32_t lwb_stream_state_len(const lwb_stream *s) { return (s && s->has) ? s->plen : 0; }

extern "C" int lwb_stream_clone(const lwb_stream *s, lwb_stream **out)
{
    if (!s || !out) return LWB_ERR_INVALID;
    int rc = lwb_stream_open(s->ctx, s->setup, out);
    if (rc) return rc;
    (*out)->has = s->has;
    (*out)->plen = s->plen;
    lwb_ctx *ctx = s->ctx;
    CU(ctx, cudaMemcpyAsync((*out)->d_state, s->d_state,
                            s->setup->channels * state_stride(s->setup) * sizeof(float),
                            cudaMemcpyDeviceToDevice, ctx->stream));
    return LWB_OK;
}

extern "C" int lwb_stream_export_state(lwb_stream *s, float *out)
{
    if (!s || !out) return LWB_ERR_INVALID;
    if (!s->has) return LWB_OK;
    lwb_ctx *ctx = s->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpy2DAsync(out, s->plen * sizeof(float), s->d_state, state_stride(s->setup) * sizeof(float),
                              s->plen * sizeof(float), s->setup->channels, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return LWB_OK;
}

extern "C" int lwb_stream_import_state(lwb_stream *s, const float *data, uint32_t len)
{
    if (!s || (!data && len)) return LWB_ERR_INVALID;
    if (len > state_stride(s->setup)) return LWB_ERR_BUFFER;
    lwb_ctx *ctx = s->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (len) {
        CU(ctx, cudaMemcpy2DAsync(s->d_state, state_stride(s->setup) * sizeof(float), data, len * sizeof(float),
                                  len * sizeof(float), s->setup->channels, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    s->ctx->state_gen++;           // contents changed even if the shape did not
    s->has = true;
    s->plen = len;
    return LWB_OK;
}

// ---------------------------------------------------------------------------------------------
// window geometry, audio.rs:1056-1073 (and its twin :889-908)
// ---------------------------------------------------------------------------------------------
struct Geom {
    uint32_t n, ls, le, rs, re;
    uint8_t blockflag, slope_sel, mapping;
};

static int geometry(const lwb_setup *su, uint8_t mode, int prev_flag, int next_flag, Geom *g)
{
    if (mode >= su->n_modes) return LWB_ERR_BAD_FORMAT;          // audio.rs:926-930
    const bool lng = su->host.mode_blockflag[mode] != 0;
    const uint32_t n = 1u << (lng ? su->bs1 : su->bs0);
    const uint32_t n0 = 1u << su->bs0;
    const bool prev = lng ? (prev_flag != 0) : true;             // short blocks: map_or(true, ..)
    const bool next = lng ? (next_flag != 0) : true;
    g->n = n;
    g->blockflag = lng;
    g->mapping = su->host.mode_mapping[mode];
    if (prev) { g->ls = 0; g->le = n >> 1; g->slope_sel = lng; }
    else { g->ls = (n - n0) >> 2; g->le = (n + n0) >> 2; g->slope_sel = 0; }
    if (next) { g->rs = n >> 1; g->re = n; }
    else { g->rs = (n * 3 - n0) >> 2; g->re = (n * 3 + n0) >> 2; }
    return LWB_OK;
}

extern "C" int lwb_decoded_sample_count(const lwb_setup *su, uint8_t mode, int prev_flag, int next_flag,
                                        uint32_t *n_samples)
{
    if (!su || !n_samples) return LWB_ERR_INVALID;
    Geom g;
    int rc = geometry(su, mode, prev_flag, next_flag, &g);
    if (rc) return rc;
    *n_samples = g.rs - g.ls;
    return LWB_OK;
}

// ---------------------------------------------------------------------------------------------
// batch planning
// ---------------------------------------------------------------------------------------------
struct PlanPacket {
    Geom g;
    uint32_t plen;          // 0: no previous half -> 0 samples out
    uint64_t coeff_off;     // absolute element offset
    uint64_t sample_pos;    // samples (per channel) produced by the chain before this packet
};

struct PlanChain {
    lwb_chain *c;
    std::vector<PlanPacket> pk;
    bool end_has;           // stream state after the planned packets
    uint32_t end_plen;
    bool clear_after;       // OLA guard fired on packet pk.size(): state becomes empty
};

static size_t elem_size(int fmt) { return (fmt == LWB_OUT_F32_PLANAR || fmt == LWB_OUT_F32_INTERLEAVED) ? 4 : 2; }
static bool is_planar(int fmt) { return fmt == LWB_OUT_F32_PLANAR || fmt == LWB_OUT_I16_PLANAR; }

static int plan_chain(lwb_chain *c, PlanChain *pc)
{
    const lwb_stream *s = c->stream;
    const lwb_setup *su = s->setup;
    bool has = s->has;
    uint32_t plen = s->plen;
    uint64_t coeff = c->coeff_offset, pos = 0;
    pc->c = c;
    pc->clear_after = false;
    c->status = LWB_OK;
    pc->pk.reserve(c->n_packets);
    for (uint32_t i = 0; i < c->n_packets; i++) {
        PlanPacket pp;
        int rc = geometry(su, c->mode_numbers[i], c->prev_window_flags ? c->prev_window_flags[i] : 1,
                          c->next_window_flags ? c->next_window_flags[i] : 1, &pp.g);
        if (rc) { c->status = rc; break; }
        if (has) {
            const uint32_t slope_len = 1u << ((pp.g.slope_sel ? su->bs1 : su->bs0) - 1);
            if (slope_len < plen) {             // audio.rs:1107-1111; :1083 has already taken the state
                c->status = LWB_ERR_BAD_FORMAT;
                pc->clear_after = true;
                break;
            }
            if (pp.g.ls + plen > pp.g.n) {      // chan[range] would be out of bounds: a panic in the reference
                c->status = LWB_ERR_MISMATCH;
                break;
            }
        }
        pp.plen = has ? plen : 0;
        pp.coeff_off = coeff;
        pp.sample_pos = pos;
        coeff += (uint64_t)su->channels * (pp.g.n >> 1);
        if (has) pos += pp.g.rs - pp.g.ls;
        has = true;
        plen = pp.g.re - pp.g.rs;
        pc->pk.push_back(pp);
    }
    pc->end_has = pc->clear_after ? false : has;
    pc->end_plen = pc->clear_after ? 0 : plen;
    c->packets_done = (uint32_t)pc->pk.size();
    c->n_samples = (uint32_t)pos;
    return LWB_OK;
}

template <typename K, typename... Args>
static int launch(lwb_ctx *ctx, K kernel, dim3 grid, dim3 block, size_t smem, Args... args)
{
    kernel<<<grid, block, smem, ctx->stream>>>(args...);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, LWB_ERR_CUDA, "kernel launch", e);
    return LWB_OK;
}

struct DevArenas {
    const float *coeffs;      // device
    const float *dense;       // device or null
    const uint8_t *kinds;     // device or null
    const uint32_t *ys;       // device or null
    uint64_t kinds_row0;      // first packet row uploaded
    void *pcm;                // device
    uint64_t coeff_base;      // element offset that device coeffs[0] corresponds to
    uint64_t pcm_base;        // element offset that device pcm[0] corresponds to
};

// Generic path: rounds of packets bounded by the IMDCT scratch.
static int run_generic(lwb_ctx *ctx, std::vector<PlanChain> &plan, const lwb_batch_io *io, const DevArenas &ar)
{
    size_t maxp = 0;
    for (auto &pc : plan) maxp = std::max(maxp, pc.pk.size());
    if (maxp == 0) return LWB_OK;
    // x elements of one "packet column" (packet i of every chain), to size the rounds
    std::vector<uint32_t> start(plan.size(), 0);
    const bool planar = is_planar(io->out_format);
    while (true) {
        // pick how many packets per chain go into this round
        size_t x_elems = 0, n_desc = 0, spec_lo = ~(size_t)0, spec_hi = 0;
        std::vector<uint32_t> take(plan.size(), 0);
        bool any = false;
        for (uint32_t step = 0;; step++) {
            size_t add = 0;
            bool more = false;
            for (size_t ci = 0; ci < plan.size(); ci++) {
                const uint32_t i = start[ci] + step;
                if (i < plan[ci].pk.size() && take[ci] == step) {
                    add += (size_t)plan[ci].c->stream->setup->channels * plan[ci].pk[i].g.n;
                    more = true;
                }
            }
            if (!more) break;
            if (x_elems && x_elems + add > ctx->x_cap_elems) break;
            for (size_t ci = 0; ci < plan.size(); ci++) {
                const uint32_t i = start[ci] + step;
                if (i < plan[ci].pk.size() && take[ci] == step) { take[ci]++; n_desc++; }
            }
            x_elems += add;
            any = true;
        }
        if (!any) break;
        int rc;
        if ((rc = ensure_pinned(ctx, n_desc * sizeof(DevPacket)))) return rc;
        if ((rc = ensure(ctx, ctx->desc, n_desc * sizeof(DevPacket)))) return rc;
        if ((rc = ensure(ctx, ctx->x, x_elems * sizeof(float)))) return rc;
        // the pinned descriptor staging is reused every round: wait for the previous upload
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        DevPacket *hp = (DevPacket *)ctx->h_desc;
        size_t di = 0, xo = 0;
        unsigned maxc = 1, maxn = 64;
        for (size_t ci = 0; ci < plan.size(); ci++) {
            PlanChain &pc = plan[ci];
            const lwb_stream *s = pc.c->stream;
            const lwb_setup *su = s->setup;
            const unsigned C = su->channels;
            for (uint32_t k = 0; k < take[ci]; k++) {
                const PlanPacket &pp = pc.pk[start[ci] + k];
                DevPacket &d = hp[di];
                std::memset(&d, 0, sizeof(d));
                d.setup = su->d_setup;
                d.state = s->d_state;
                d.coeff_off = pp.coeff_off - ar.coeff_base;
                d.x_off = xo;
                d.out_stride = pc.c->out_stride;
                d.out_off = pc.c->out_offset - ar.pcm_base + (planar ? pp.sample_pos : pp.sample_pos * C);
                d.pkt_index = pc.c->packet_index + start[ci] + k - ar.kinds_row0;
                d.prev_packet = k ? (int32_t)(di - 1) : -1;
                d.prev_rs = k ? hp[di - 1].rs : 0;
                d.state_stride = (uint32_t)state_stride(su);
                d.n = (uint16_t)pp.g.n;
                d.ls = (uint16_t)pp.g.ls;
                d.rs = (uint16_t)pp.g.rs;
                d.re = (uint16_t)pp.g.re;
                d.plen = (uint16_t)pp.plen;
                d.blockflag = pp.g.blockflag;
                d.mapping = pp.g.mapping;
                d.slope_sel = pp.g.slope_sel;
                d.channels = (uint8_t)C;
                d.save_state = (k + 1 == take[ci]);
                xo += (size_t)C * pp.g.n;
                spec_lo = std::min<size_t>(spec_lo, d.coeff_off);
                spec_hi = std::max<size_t>(spec_hi, d.coeff_off + (size_t)C * (pp.g.n >> 1));
                maxc = std::max(maxc, C);
                maxn = std::max<unsigned>(maxn, pp.g.n);
                di++;
            }
            start[ci] += take[ci];
        }
        CU(ctx, cudaMemcpyAsync(ctx->desc.p, hp, n_desc * sizeof(DevPacket), cudaMemcpyHostToDevice, ctx->stream));
        const DevPacket *dp = (const DevPacket *)ctx->desc.p;
        const float *spec = ar.coeffs;
        if (io->entry == LWB_ENTRY_RESIDUE) {
            if ((rc = ensure(ctx, ctx->spec, spec_hi * sizeof(float)))) return rc;
            if ((rc = launch(ctx, k_prologue, dim3((unsigned)n_desc), dim3(kPrologueThreads), 0, dp, ar.coeffs,
                             ar.dense, ar.kinds, ar.ys, (float *)ctx->spec.p)))
                return rc;
            spec = (const float *)ctx->spec.p;
        }
        if ((rc = launch(ctx, k_imdct, dim3((unsigned)n_desc, maxc), dim3(kImdctThreads), maxn * sizeof(float), dp,
                         spec, (float *)ctx->x.p)))
            return rc;
        dim3 g2((unsigned)n_desc, maxc), b2(kOverlapThreads);
        switch (io->out_format) {
        case LWB_OUT_F32_PLANAR: rc = launch(ctx, k_overlap<LWB_OUT_F32_PLANAR>, g2, b2, 0, dp, (const float *)ctx->x.p, ar.pcm); break;
        case LWB_OUT_I16_PLANAR: rc = launch(ctx, k_overlap<LWB_OUT_I16_PLANAR>, g2, b2, 0, dp, (const float *)ctx->x.p, ar.pcm); break;
        case LWB_OUT_F32_INTERLEAVED: rc = launch(ctx, k_overlap<LWB_OUT_F32_INTERLEAVED>, g2, b2, 0, dp, (const float *)ctx->x.p, ar.pcm); break;
        default: rc = launch(ctx, k_overlap<LWB_OUT_I16_INTERLEAVED>, g2, b2, 0, dp, (const float *)ctx->x.p, ar.pcm); break;
        }
        if (rc) return rc;
        if ((rc = launch(ctx, k_save_state, g2, b2, 0, dp, (const float *)ctx->x.p))) return rc;
    }
    return LWB_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused path (kernel_long.cuh).  Eligible batches: spectrum entry, planar f32 out, every packet a
// long block of blocksize 2^11 with long neighbours, every stream either empty or holding a
// 1024-sample right half.  Planned directly from the chain list in O(chains + mode bytes) -- at
// 0.8 G blocks/s per GPU a per-packet host plan would be the bottleneck.
// ---------------------------------------------------------------------------------------------
struct LongItem {
    lwb_chain *c;
    uint32_t P;
    bool has_prev;
};

struct Staging {
    void *h = nullptr;
    size_t cap = 0;
    cudaEvent_t ev = nullptr;
    bool pending = false;
};
static Staging g_stage[4][3];          // per device ordinal (ctx is per device), ring of 3
static int g_stage_next[4];

static int acquire_staging(lwb_ctx *ctx, size_t bytes, Staging **out)
{
    const int d = ctx->device & 3;
    Staging &st = g_stage[d][g_stage_next[d]];
    g_stage_next[d] = (g_stage_next[d] + 1) % 3;
    if (!st.ev) CU(ctx, cudaEventCreateWithFlags(&st.ev, cudaEventDisableTiming));
    if (st.pending) {
        CU(ctx, cudaEventSynchronize(st.ev));      // waits for the descriptor copy only, not for kernels
        st.pending = false;
    }
    if (st.cap < bytes) {
        if (st.h) cudaFreeHost(st.h);
        st.h = nullptr;
        st.cap = 0;
        CU(ctx, cudaHostAlloc(&st.h, bytes * 2 + 4096, cudaHostAllocDefault));
        st.cap = bytes * 2 + 4096;
    }
    *out = &st;
    return LWB_OK;
}

// Appends the runs of one chain.  A chain (one channel of one stream) is cut into several runs
// when there are too few chains to fill the machine; every run after the first re-transforms the
// packet before its first one as a primer (its right half is all the run needs), which keeps
// runs independent at the cost of one extra IMDCT per cut.
static void long_runs_of(const LongItem &it, size_t cuts, const float *coeffs, uint64_t coeff_base, char *pcm,
                         uint64_t pcm_base, size_t esz, LongRun *&w)
{
    const lwb_stream *s = it.c->stream;
    const lwb_setup *su = s->setup;
    const unsigned C = su->channels;
    const size_t P = it.P;
    for (unsigned ch = 0; ch < C; ch++) {
        const float *in0 = coeffs + (it.c->coeff_offset - coeff_base) + (size_t)ch * kLongN2;
        char *out0 = pcm + ((it.c->out_offset - pcm_base) + (size_t)ch * it.c->out_stride) * esz;
        for (size_t k = 0; k < cuts; k++) {
            const size_t p0 = P * k / cuts, p1 = P * (k + 1) / cuts;   // this run emits packets [p0, p1)
            LongRun &r = *w++;
            std::memset(&r, 0, sizeof(r));
            r.in_stride = (uint32_t)(C * kLongN2);
            r.state = s->d_state + (size_t)ch * state_stride(su);
            r.write_state = (k + 1 == cuts);
            if (k == 0) {
                r.in = in0;
                r.n_packets = (uint32_t)(p1 - p0);
                r.has_prev = it.has_prev;
                r.out = out0;
            } else {
                r.in = in0 + (p0 - 1) * (size_t)r.in_stride;           // primer = packet p0 - 1
                r.n_packets = (uint32_t)(p1 - p0 + 1);
                r.has_prev = 0;
                // samples emitted before packet p0: packets 0..p0-1, minus the first if no state
                r.out = out0 + (size_t)(p0 - (it.has_prev ? 0 : 1)) * kLongN2 * esz;
            }
        }
    }
}

// Every packet a long block of the fast blocksize with long neighbours, every stream empty or
// holding a 1024-sample right half, arenas aligned: what the fused kernel takes.
static bool batch_is_uniform_long(lwb_ctx *ctx, const lwb_chain *chains, size_t n_chains, const lwb_batch_io *io)
{
    if (io->out_format != LWB_OUT_F32_PLANAR && io->out_format != LWB_OUT_I16_PLANAR) return false;
    const float *pack = nullptr;
    for (size_t i = 0; i < n_chains; i++) {
        const lwb_chain *c = &chains[i];
        if (!c->stream || c->stream->ctx != ctx || (c->n_packets && !c->mode_numbers)) return false;
        const lwb_stream *s = c->stream;
        const lwb_setup *su = s->setup;
        if (su->bs1 != kLongBs || !su->host.tab[1].pack) return false;
        if (pack && pack != su->host.tab[1].pack) return false;
        pack = su->host.tab[1].pack;
        if ((c->out_offset & 3) || (c->out_stride & 3) || (c->coeff_offset & 3)) return false;
        if (s->has && s->plen != (uint32_t)kLongN2) return false;
        for (uint32_t k = 0; k < c->n_packets; k++) {
            const uint8_t m = c->mode_numbers[k];
            if (m >= su->n_modes || !su->host.mode_blockflag[m]) return false;
            if (c->prev_window_flags && !c->prev_window_flags[k]) return false;
            if (c->next_window_flags && !c->next_window_flags[k]) return false;
        }
    }
    return true;
}

// `spectrum_dev`: when non-null the spectrum has already been formed on the device (residue entry:
// k_prologue wrote it to ctx->spec, element offset `spectrum_base` = its [0]); the input side of the
// batch is then neither validated as a spectrum entry nor copied.
static int try_long(lwb_ctx *ctx, lwb_chain *chains, size_t n_chains, const lwb_batch_io *io, uint64_t epoch,
                    bool *handled, const float *spectrum_dev = nullptr, uint64_t spectrum_base = 0,
                    lwb_plan *plan = nullptr)
{
    *handled = false;
    const uint64_t gen_at_entry = ctx->state_gen;
    if (plan) plan->captured = false;
    if (!spectrum_dev && io->entry != LWB_ENTRY_SPECTRUM) return LWB_OK;
    if (io->out_format != LWB_OUT_F32_PLANAR && io->out_format != LWB_OUT_I16_PLANAR) return LWB_OK;
    if (getenv("LWB_FORCE_GENERIC")) return LWB_OK;
    const bool i16 = io->out_format == LWB_OUT_I16_PLANAR;
    const size_t esz = i16 ? 2 : 4;
    std::vector<LongItem> items;
    items.reserve(n_chains);
    const float *pack = nullptr;
    size_t chan_chains = 0;
    for (size_t i = 0; i < n_chains; i++) {
        lwb_chain *c = &chains[i];
        if (!c->stream || c->stream->ctx != ctx || (c->n_packets && !c->mode_numbers)) return LWB_OK;   // generic path reports it
        const lwb_stream *s = c->stream;
        const lwb_setup *su = s->setup;
        if (su->bs1 != kLongBs || !su->host.tab[1].pack) return LWB_OK;
        if (pack && pack != su->host.tab[1].pack) return LWB_OK;          // one twiddle pack per launch
        pack = su->host.tab[1].pack;
        if ((c->out_offset & 3) || (c->out_stride & 3) || (c->coeff_offset & 3)) return LWB_OK;
        if (s->has && s->plen != (uint32_t)kLongN2) return LWB_OK;
        const uint32_t P = c->n_packets;
        for (uint32_t k = 0; k < P; k++) {
            const uint8_t m = c->mode_numbers[k];
            if (m >= su->n_modes || !su->host.mode_blockflag[m]) return LWB_OK;
            if (c->prev_window_flags && !c->prev_window_flags[k]) return LWB_OK;
            if (c->next_window_flags && !c->next_window_flags[k]) return LWB_OK;
        }
        items.push_back(LongItem{c, P, s->has});
        if (P) chan_chains += su->channels;
    }
    *handled = true;
    // from here on this path owns the batch
    uint64_t c_lo = ~0ull, c_hi = 0, o_lo = ~0ull, o_hi = 0;
    for (auto &it : items) {
        lwb_chain *c = it.c;
        if (!spectrum_dev) {       // (the residue path has already run this check while planning)
            if (c->stream->busy_epoch == epoch) return fail(ctx, LWB_ERR_INVALID, "a stream appears in two chains of one batch");
            c->stream->busy_epoch = epoch;
        }
        const unsigned C = c->stream->setup->channels;
        c->status = LWB_OK;
        c->packets_done = it.P;
        c->n_samples = it.P ? (uint32_t)((it.P - (it.has_prev ? 0 : 1)) * kLongN2) : 0;
        if (!it.P) continue;
        if (c->out_stride < c->n_samples) return fail(ctx, LWB_ERR_BUFFER, "chain: out_stride smaller than the samples produced");
        c_lo = std::min(c_lo, c->coeff_offset);
        c_hi = std::max(c_hi, c->coeff_offset + (uint64_t)it.P * C * kLongN2);
        o_lo = std::min(o_lo, c->out_offset);
        o_hi = std::max(o_hi, c->out_offset + (uint64_t)(C - 1) * c->out_stride + c->n_samples);
    }
    if (!chan_chains) return LWB_OK;
    const size_t warp_slots = (size_t)ctx->sm_count * kLongWarps * kLongNB;
    size_t target_runs = warp_slots * 4;                   // ~4 groups per warp evens out the tail
    if (const char *e = getenv("LWB_LONG_TARGET_RUNS")) target_runs = (size_t)atol(e);
    const size_t min_run = 8;                              // packets per run below which a cut costs > 12%
    int rc;
    constexpr uint32_t kTicketPool = 1024;
    if (!ctx->ticket.p) {
        if ((rc = ensure(ctx, ctx->ticket, kTicketPool * sizeof(unsigned int)))) return rc;
        for (int k = 0; k < 2; k++) {
            CU(ctx, cudaEventCreateWithFlags(&ctx->ev_desc[k], cudaEventDisableTiming));
            CU(ctx, cudaEventCreateWithFlags(&ctx->ev_kdone[k], cudaEventDisableTiming));
        }
    }

    const bool host = io->memory == LWB_MEM_HOST;          // the pcm arena is in host memory
    const bool in_host = host && !spectrum_dev;            // ... and so is the coefficient arena
    // host memory: chunks of chains, H2D / kernel / D2H of consecutive chunks overlap on three streams
    size_t n_chunks = 1;
    if (host) {
        const size_t bytes = (size_t)(c_hi - c_lo) * 4;
        n_chunks = std::min<size_t>(std::max<size_t>(1, bytes >> 25), std::min<size_t>(8, items.size()));   // profiles/e2e_chunks_r1.log
        if (const char *e = getenv("LWB_E2E_CHUNKS")) n_chunks = std::max<size_t>(1, std::min<size_t>((size_t)atol(e), std::min<size_t>(64, items.size())));
    }
    const float *d_coeffs = spectrum_dev ? spectrum_dev : io->coeffs;
    char *d_pcm = (char *)io->pcm;
    uint64_t cbase = spectrum_dev ? spectrum_base : 0, obase = 0;
    if (host) {
        if (in_host) {
            if ((rc = ensure(ctx, ctx->coeffs, (size_t)(c_hi - c_lo) * 4))) return rc;
            d_coeffs = (const float *)ctx->coeffs.p;
            cbase = c_lo;
        }
        if (o_hi > o_lo && (rc = ensure(ctx, ctx->pcm, (size_t)(o_hi - o_lo) * esz))) return rc;
        d_pcm = (char *)ctx->pcm.p;
        obase = o_lo;
        if (!ctx->ev_in[0])
            for (int k = 0; k < 65; k++) {
                if (k < 64) CU(ctx, cudaEventCreateWithFlags(&ctx->ev_in[k], cudaEventDisableTiming));
                CU(ctx, cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming));
            }
        // the copy streams must not run ahead of work already queued on the compute stream that
        // still reads/writes the arenas (previous call): order them behind it
        CU(ctx, cudaEventRecord(ctx->ev_done[64], ctx->stream));
        CU(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->ev_done[64], 0));
    }
    // count runs
    std::vector<size_t> cuts(items.size(), 1);
    size_t total_runs = 0;
    for (size_t i = 0; i < items.size(); i++) {
        if (!items[i].P) { cuts[i] = 0; continue; }
        // per launch (chunk) the machine should see >= target_runs runs
        const size_t per_launch = std::max<size_t>(1, chan_chains / n_chunks);
        size_t k = 1;
        if (per_launch < target_runs) k = (target_runs + per_launch - 1) / per_launch;
        cuts[i] = std::max<size_t>(1, std::min(k, items[i].P / min_run));
        total_runs += cuts[i] * items[i].c->stream->setup->channels;
    }
    // the kernel takes groups of kLongNB runs of equal length; unpaired runs get a dummy partner
    const size_t cap_runs = total_runs * (kLongNB > 1 ? 2 : 1) + kLongNB;
    Staging *st;
    if ((rc = acquire_staging(ctx, cap_runs * sizeof(LongRun), &st))) return rc;
    const int par = ctx->runs_par;
    ctx->runs_par ^= 1;
    // a plan (device-memory batches) owns its descriptor buffer so that later executions can reuse it
    const bool capture = plan && !host && !spectrum_dev && n_chunks == 1;
    DevBuf &rb = capture ? plan->runs : ctx->runs_buf[par];
    if ((rc = ensure(ctx, rb, cap_runs * sizeof(LongRun)))) return rc;
    LongRun *const d_runs_base = (LongRun *)rb.p;
    LongRun *h_runs = (LongRun *)st->h, *w = h_runs;
    std::vector<LongRun> tmp;
    struct ChunkPlan { size_t r0, nr; uint64_t kc_lo, kc_hi, ko_lo, ko_hi; };
    std::vector<ChunkPlan> cplan;
    std::vector<uint32_t> order;
    for (size_t k = 0; k < n_chunks; k++) {
        const size_t i0 = items.size() * k / n_chunks, i1 = items.size() * (k + 1) / n_chunks;
        LongRun *w0 = w;
        uint64_t kc_lo = ~0ull, kc_hi = 0, ko_lo = ~0ull, ko_hi = 0;
        // NB == 1: descriptors are written straight into the pinned staging; otherwise into a scratch
        // vector that is regrouped below
        size_t chunk_runs = 0;
        for (size_t i = i0; i < i1; i++)
            if (items[i].P) chunk_runs += cuts[i] * items[i].c->stream->setup->channels;
        LongRun *gen = w;
        if (kLongNB > 1) {
            tmp.resize(chunk_runs);
            gen = tmp.data();
        }
        for (size_t i = i0; i < i1; i++) {
            if (!items[i].P) continue;
            long_runs_of(items[i], cuts[i], d_coeffs, cbase, d_pcm, obase, esz, gen);
            const lwb_chain *c = items[i].c;
            const unsigned C = c->stream->setup->channels;
            kc_lo = std::min(kc_lo, c->coeff_offset);
            kc_hi = std::max(kc_hi, c->coeff_offset + (uint64_t)items[i].P * C * kLongN2);
            ko_lo = std::min(ko_lo, c->out_offset);
            ko_hi = std::max(ko_hi, c->out_offset + (uint64_t)(C - 1) * c->out_stride + c->n_samples);
        }
        if (!chunk_runs) continue;
        if (kLongNB == 1) {
            w = gen;
        } else {
            // group runs of equal packet count (consecutive channels of a stream already are)
            bool sorted = true;
            for (size_t i = 1; i < tmp.size() && sorted; i++) sorted = tmp[i].n_packets == tmp[0].n_packets;
            order.resize(tmp.size());
            for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
            if (!sorted)
                std::stable_sort(order.begin(), order.end(),
                                 [&](uint32_t a, uint32_t b) { return tmp[a].n_packets < tmp[b].n_packets; });
            size_t i = 0;
            while (i < order.size()) {
                size_t j = i;
                while (j < order.size() && tmp[order[j]].n_packets == tmp[order[i]].n_packets) j++;
                for (size_t q = i; q < j; q++) *w++ = tmp[order[q]];
                size_t fill = (kLongNB - (j - i) % kLongNB) % kLongNB;
                while (fill--) {
                    LongRun d = tmp[order[j - 1]];       // reads valid memory, stores nothing
                    d.dummy = 1;
                    d.write_state = 0;
                    d.has_prev = 0;
                    *w++ = d;
                }
                i = j;
            }
        }
        cplan.push_back(ChunkPlan{(size_t)(w0 - h_runs), (size_t)(w - w0), kc_lo, kc_hi, ko_lo, ko_hi});
    }
    // one descriptor upload for the whole call, on the copy stream, behind the kernel that last read
    // this half of the double buffer
    const size_t all_runs = (size_t)(w - h_runs);
    if (!all_runs) return LWB_OK;
    CU(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->ev_kdone[par], 0));
    CU(ctx, cudaMemcpyAsync(d_runs_base, h_runs, all_runs * sizeof(LongRun), cudaMemcpyHostToDevice, ctx->copy_out));
    CU(ctx, cudaEventRecord(ctx->ev_desc[par], ctx->copy_out));
    CU(ctx, cudaEventRecord(st->ev, ctx->copy_out));
    st->pending = true;
    CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_desc[par], 0));
    for (size_t k = 0; k < cplan.size(); k++) {
        const ChunkPlan &cp = cplan[k];
        if (in_host) {
            CU(ctx, cudaMemcpyAsync((float *)ctx->coeffs.p + (cp.kc_lo - cbase), io->coeffs + cp.kc_lo,
                                    (size_t)(cp.kc_hi - cp.kc_lo) * 4, cudaMemcpyHostToDevice, ctx->copy_in));
            CU(ctx, cudaEventRecord(ctx->ev_in[k], ctx->copy_in));
            CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[k], 0));
        }
        if (ctx->ticket_next % kTicketPool == 0)
            CU(ctx, cudaMemsetAsync(ctx->ticket.p, 0, kTicketPool * sizeof(unsigned int), ctx->stream));
        unsigned int *ticket = (unsigned int *)ctx->ticket.p + (ctx->ticket_next++ % kTicketPool);
        if (long_launch(ctx->stream, d_runs_base + cp.r0, (uint32_t)(cp.nr / kLongNB), pack, ticket, ctx->sm_count, i16))
            return fail(ctx, LWB_ERR_CUDA, "long kernel launch", cudaGetLastError());
        ctx->launches++;
        if (host && cp.ko_hi > cp.ko_lo) {
            CU(ctx, cudaEventRecord(ctx->ev_done[k], ctx->stream));
            CU(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->ev_done[k], 0));
            CU(ctx, cudaMemcpyAsync((char *)io->pcm + cp.ko_lo * esz, (char *)ctx->pcm.p + (cp.ko_lo - obase) * esz,
                                    (size_t)(cp.ko_hi - cp.ko_lo) * esz, cudaMemcpyDeviceToHost, ctx->copy_out));
        }
    }
    CU(ctx, cudaEventRecord(ctx->ev_kdone[par], ctx->stream));
    if (capture && cplan.size() == 1) {
        plan->captured = true;
        plan->gen = gen_at_entry;          // valid while no stream changed shape since planning
        plan->n_groups = (uint32_t)(cplan[0].nr / kLongNB);
        plan->pack = pack;
        plan->i16 = i16;
    }
    if (host) {
        CU(ctx, cudaStreamSynchronize(ctx->copy_out));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    for (auto &it : items)
        if (it.P) set_stream_state(it.c->stream, true, kLongN2);
    return LWB_OK;
}

// Residue-entry batches whose every packet is a long block with long neighbours (what the fused
// kernel takes) -- decided from the generic plan.
static bool plan_is_long(const std::vector<PlanChain> &plan, const lwb_batch_io *io)
{
    if (io->out_format != LWB_OUT_F32_PLANAR && io->out_format != LWB_OUT_I16_PLANAR) return false;
    if (getenv("LWB_FORCE_GENERIC")) return false;
    for (auto &pc : plan) {
        const lwb_setup *su = pc.c->stream->setup;
        if (su->bs1 != kLongBs || !su->host.tab[1].pack) return false;
        if (pc.c->status != LWB_OK) return false;
        for (auto &pp : pc.pk) {
            if (!pp.g.blockflag || pp.g.ls != 0 || pp.g.rs != (pp.g.n >> 1) || pp.g.re != pp.g.n) return false;
            if (pp.plen != 0 && pp.plen != (pp.g.n >> 1)) return false;
        }
    }
    return true;
}

// k_prologue over every packet of the plan: ctx->spec[coeff_off] <- floor x decoupled residue.
static int run_prologue_all(lwb_ctx *ctx, std::vector<PlanChain> &plan, const DevArenas &ar, size_t spec_elems)
{
    size_t n_desc = 0;
    for (auto &pc : plan) n_desc += pc.pk.size();
    if (!n_desc) return LWB_OK;
    int rc;
    if ((rc = ensure_pinned(ctx, n_desc * sizeof(DevPacket)))) return rc;
    if ((rc = ensure(ctx, ctx->desc, n_desc * sizeof(DevPacket)))) return rc;
    if ((rc = ensure(ctx, ctx->spec, spec_elems * sizeof(float)))) return rc;
    CU(ctx, cudaStreamSynchronize(ctx->stream));          // pinned descriptor staging is reused
    DevPacket *hp = (DevPacket *)ctx->h_desc;
    size_t di = 0;
    for (auto &pc : plan) {
        const lwb_setup *su = pc.c->stream->setup;
        for (size_t k = 0; k < pc.pk.size(); k++) {
            const PlanPacket &pp = pc.pk[k];
            DevPacket &d = hp[di++];
            std::memset(&d, 0, sizeof(d));
            d.setup = su->d_setup;
            d.coeff_off = pp.coeff_off - ar.coeff_base;
            d.pkt_index = pc.c->packet_index + k - ar.kinds_row0;
            d.n = (uint16_t)pp.g.n;
            d.blockflag = pp.g.blockflag;
            d.mapping = pp.g.mapping;
            d.channels = su->channels;
        }
    }
    CU(ctx, cudaMemcpyAsync(ctx->desc.p, hp, n_desc * sizeof(DevPacket), cudaMemcpyHostToDevice, ctx->stream));
    return launch(ctx, k_prologue, dim3((unsigned)n_desc), dim3(kPrologueThreads), 0, (const DevPacket *)ctx->desc.p,
                  ar.coeffs, ar.dense, ar.kinds, ar.ys, (float *)ctx->spec.p);
}

// ---------------------------------------------------------------------------------------------
// Chain kernel path (kernel_chain.cuh): everything the fused long-block kernel does not take,
// as long as channels <= 8 and the per-channel buffers fit in shared memory.
// ---------------------------------------------------------------------------------------------
// Shared memory of the chain kernel: per channel `np` blocks of U | V plus the previous right half, and the
// floor posts of up to 8 channels.  np (blocks a channel group transforms together) is 4 where that fits.
static size_t chain_smem(unsigned maxc, int n1max, int np)
{
    return (size_t)maxc * ((size_t)np * n1max + n1max / 2) * 4 + 8 * (LWB_MAX_POSTS + 1) * 2 * 2 + 64;
}
static int chain_np(unsigned maxc, int n1max, int wpc, bool residue)
{
    if (residue || wpc != 1 || getenv("LWB_CHAIN_NP1")) return 1;
    int np = 4;
    while (np > 1 && chain_smem(maxc, n1max, np) > 64 * 1024) np >>= 1;
    return np;
}

template <int ENTRY>
static int launch_chain(lwb_ctx *ctx, int fmt, unsigned n_chains, unsigned warps, size_t smem, const ChainDesc *d,
                        const uint8_t *bytes, const float *coeffs, const float *dense, const uint8_t *kinds,
                        const uint32_t *ys, void *pcm, int n1max, int wpc, int np)
{
#define LWB_CHAIN_CASE(F)                                                                                    \
    case F:                                                                                                  \
        if (wpc == 1) {                                                                                      \
            cudaFuncSetAttribute(k_chain<F, ENTRY, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            return launch(ctx, k_chain<F, ENTRY, false>, dim3(n_chains), dim3(warps * 32), smem, d, bytes, coeffs, dense, \
                          kinds, ys, pcm, n1max, wpc, np);                                                    \
        }                                                                                                    \
        cudaFuncSetAttribute(k_chain<F, ENTRY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        return launch(ctx, k_chain<F, ENTRY, true>, dim3(n_chains), dim3(warps * 32), smem, d, bytes, coeffs, dense, kinds, \
                      ys, pcm, n1max, wpc, 1);
    switch (fmt) {
        LWB_CHAIN_CASE(LWB_OUT_F32_PLANAR)
        LWB_CHAIN_CASE(LWB_OUT_I16_PLANAR)
        LWB_CHAIN_CASE(LWB_OUT_F32_INTERLEAVED)
        LWB_CHAIN_CASE(LWB_OUT_I16_INTERLEAVED)
    }
#undef LWB_CHAIN_CASE
    return LWB_ERR_INVALID;
}

// one launch of the fused kernel and one of the chain kernel per round, in stream order
static int mixed_launch_rounds(lwb_ctx *ctx, const MixLaunch &ml, const std::vector<MixRound> &rounds)
{
    constexpr uint32_t kTicketPool = 1024;
    cudaStream_t sm = ctx->stream;
    int rc = LWB_OK;
    for (const MixRound &rd : rounds) {
        if (rd.nr) {
            if (ctx->ticket_next % kTicketPool == 0)
                CU(ctx, cudaMemsetAsync(ctx->ticket.p, 0, kTicketPool * sizeof(unsigned int), sm));
            unsigned int *ticket = (unsigned int *)ctx->ticket.p + (ctx->ticket_next++ % kTicketPool);
            if (kLongNB != 1) return fail(ctx, LWB_ERR_INVALID, "mixed path needs one run per warp");
            if (long_launch(sm, (const LongRun *)ml.db + rd.r0, (uint32_t)rd.nr, ml.pack, ticket, ctx->sm_count, ml.i16, ml.w_short, ml.ls))
                return fail(ctx, LWB_ERR_CUDA, "long kernel launch", cudaGetLastError());
            ctx->launches++;
        }
        if (rd.nc) {
            const ChainDesc *dcd = (const ChainDesc *)(ml.db + ml.off_cd) + rd.c0;
            const uint8_t *dby = (const uint8_t *)(ml.db + ml.off_by);
            if (ml.residue)
                rc = launch_chain<LWB_ENTRY_RESIDUE>(ctx, ml.out_format, (unsigned)rd.nc, ml.warps, ml.smem, dcd, dby, ml.coeffs, ml.dense,
                                                     ml.kinds, ml.ys, ml.pcm, ml.n1max, ml.wpc, ml.np);
            else
                rc = launch_chain<LWB_ENTRY_SPECTRUM>(ctx, ml.out_format, (unsigned)rd.nc, ml.warps, ml.smem, dcd, dby, ml.coeffs, ml.dense,
                                                      ml.kinds, ml.ys, ml.pcm, ml.n1max, ml.wpc, ml.np);
            if (rc) return rc;
        }
    }
    return LWB_OK;
}

static int try_chain(lwb_ctx *ctx, lwb_chain *chains, size_t n_chains, const lwb_batch_io *io, uint64_t epoch,
                     bool *handled, lwb_plan *plan = nullptr)
{
    *handled = false;
    const uint64_t gen_at_entry = ctx->state_gen;
    if (plan) plan->mixed_captured = false;
    if (const char *e = getenv("LWB_FORCE_GENERIC"))
        if (std::strcmp(e, "1") == 0) return LWB_OK;          // "1": the four-kernel path; "2": no fused kernel only
    const bool residue = io->entry == LWB_ENTRY_RESIDUE;
    const bool planar = is_planar(io->out_format);
    const size_t esz = elem_size(io->out_format);
    unsigned maxc = 1;
    int n1max = 64;
    size_t total_packets = 0;
    for (size_t i = 0; i < n_chains; i++) {
        const lwb_chain *c = &chains[i];
        if (!c->stream || c->stream->ctx != ctx || (c->n_packets && !c->mode_numbers)) return LWB_OK;   // generic path reports it
        const lwb_setup *su = c->stream->setup;
        if (su->channels > 8) return LWB_OK;
        maxc = std::max<unsigned>(maxc, su->channels);
        n1max = std::max(n1max, 1 << su->bs1);
        total_packets += c->n_packets;
    }
    if (chain_smem(maxc, n1max, 1) > 200 * 1024) return LWB_OK;
    // warps per channel: one per 1024 samples of the largest block, at most 32 warps per CTA
    int wpc = std::max(1, std::min(8, n1max / 1024));
    while (wpc > 1 && (unsigned)wpc * maxc > 32) wpc >>= 1;
    const int np = chain_np(maxc, n1max, wpc, residue);
    const size_t smem = chain_smem(maxc, n1max, np);
    if (residue && !io->floor_kind) return fail(ctx, LWB_ERR_INVALID, "residue entry needs floor_kind");
    *handled = true;

    // light walk of every chain: geometry, OLA guard, output size (audio.rs:1056-1073, 1083-1154)
    int rc;
    Staging *st;
    const size_t desc_bytes = n_chains * sizeof(ChainDesc), byte_bytes = total_packets * 3 + 16;
    if ((rc = acquire_staging(ctx, desc_bytes + byte_bytes, &st))) return rc;
    ChainDesc *hd = (ChainDesc *)st->h;
    uint8_t *hb = (uint8_t *)st->h + desc_bytes;
    uint64_t c_lo = ~0ull, c_hi = 0, o_lo = ~0ull, o_hi = 0, r_lo = ~0ull, r_hi = 0;
    int uniform_c = -1;
    bool need_dense = false;
    size_t boff = 0, n_launch = 0;
    struct End { lwb_stream *s; bool has; uint32_t plen; bool touched; };
    std::vector<End> ends(n_chains);
    for (size_t i = 0; i < n_chains; i++) {
        lwb_chain *c = &chains[i];
        lwb_stream *s = c->stream;
        const lwb_setup *su = s->setup;
        if (s->busy_epoch == epoch) return fail(ctx, LWB_ERR_INVALID, "a stream appears in two chains of one batch");
        s->busy_epoch = epoch;
        const unsigned C = su->channels;
        if (residue) {
            if (uniform_c < 0) uniform_c = (int)C;
            if (uniform_c != (int)C) return fail(ctx, LWB_ERR_INVALID, "residue batches need one channel count");
        }
        bool has = s->has, clear_after = false;
        uint32_t plen = s->plen;
        uint64_t coeff = c->coeff_offset, pos = 0;
        uint32_t done = 0;
        c->status = LWB_OK;
        for (uint32_t k = 0; k < c->n_packets; k++) {
            Geom g;
            int grc = geometry(su, c->mode_numbers[k], c->prev_window_flags ? c->prev_window_flags[k] : 1,
                               c->next_window_flags ? c->next_window_flags[k] : 1, &g);
            if (grc) { c->status = grc; break; }
            if (has) {
                const uint32_t slope_len = 1u << ((g.slope_sel ? su->bs1 : su->bs0) - 1);
                if (slope_len < plen) { c->status = LWB_ERR_BAD_FORMAT; clear_after = true; break; }   // audio.rs:1107-1111
                if (g.ls + plen > g.n) { c->status = LWB_ERR_MISMATCH; break; }
                pos += g.rs - g.ls;
            }
            hb[boff + 3 * k] = c->mode_numbers[k];
            hb[boff + 3 * k + 1] = c->prev_window_flags ? c->prev_window_flags[k] : 1;
            hb[boff + 3 * k + 2] = c->next_window_flags ? c->next_window_flags[k] : 1;
            coeff += (uint64_t)C * (g.n >> 1);
            has = true;
            plen = g.re - g.rs;
            done++;
        }
        c->packets_done = done;
        c->n_samples = (uint32_t)pos;
        ends[i] = End{s, clear_after ? false : has, clear_after ? 0u : plen, done > 0 || clear_after};
        if (!done) continue;
        if (planar && c->out_stride < pos) return fail(ctx, LWB_ERR_BUFFER, "chain: out_stride smaller than the samples produced");
        ChainDesc &d = hd[n_launch++];
        std::memset(&d, 0, sizeof(d));
        d.setup = su->d_setup;
        d.state = s->d_state;
        d.coeff_off = c->coeff_offset;
        d.out_off = c->out_offset;
        d.out_stride = c->out_stride;
        d.pkt_index = c->packet_index;
        d.n_packets = done;
        d.byte_off = (uint32_t)boff;
        d.state_stride = (uint32_t)state_stride(su);
        d.plen0 = (uint16_t)s->plen;
        d.has0 = s->has;
        d.channels = (uint8_t)C;
        boff += (size_t)done * 3;
        c_lo = std::min(c_lo, c->coeff_offset);
        c_hi = std::max(c_hi, coeff);
        const uint64_t ext = planar ? (uint64_t)(C - 1) * c->out_stride + pos : pos * C;
        o_lo = std::min(o_lo, c->out_offset);
        o_hi = std::max(o_hi, c->out_offset + ext);
        if (residue) {
            r_lo = std::min(r_lo, c->packet_index);
            r_hi = std::max<uint64_t>(r_hi, c->packet_index + done);
            for (uint64_t r = c->packet_index * C; r < (c->packet_index + done) * C; r++) {
                const uint8_t kd = io->floor_kind[r];
                if (kd > LWB_FLOOR_DENSE) return fail(ctx, LWB_ERR_INVALID, "floor_kind out of range");
                if (kd == LWB_FLOOR_ONE && !io->floor1_y) return fail(ctx, LWB_ERR_INVALID, "floor1_y missing");
                if (kd == LWB_FLOOR_DENSE) need_dense = true;
            }
        }
    }
    if (need_dense && !io->dense_floor) return fail(ctx, LWB_ERR_INVALID, "dense_floor missing");
    if (n_launch) {
        const bool host = io->memory == LWB_MEM_HOST;
        const float *d_coeffs = io->coeffs, *d_dense = io->dense_floor;
        char *d_pcm = (char *)io->pcm;
        cudaStream_t sm = ctx->stream;
        if (host) {
            // arenas are addressed with the caller's element offsets: bias the device pointers instead of the descriptors
            if ((rc = ensure(ctx, ctx->coeffs, (size_t)(c_hi - c_lo) * 4))) return rc;
            if (o_hi > o_lo && (rc = ensure(ctx, ctx->pcm, (size_t)(o_hi - o_lo) * esz))) return rc;
            CU(ctx, cudaMemcpyAsync(ctx->coeffs.p, io->coeffs + c_lo, (size_t)(c_hi - c_lo) * 4, cudaMemcpyHostToDevice, sm));
            d_coeffs = (const float *)ctx->coeffs.p - c_lo;
            if (need_dense) {
                if ((rc = ensure(ctx, ctx->dense, (size_t)(c_hi - c_lo) * 4))) return rc;
                CU(ctx, cudaMemcpyAsync(ctx->dense.p, io->dense_floor + c_lo, (size_t)(c_hi - c_lo) * 4, cudaMemcpyHostToDevice, sm));
                d_dense = (const float *)ctx->dense.p - c_lo;
            }
            d_pcm = (char *)ctx->pcm.p - o_lo * esz;
        }
        const uint8_t *d_kinds = nullptr;
        const uint32_t *d_ys = nullptr;
        if (residue) {
            const size_t rows = (size_t)(r_hi - r_lo) * uniform_c;
            if ((rc = ensure(ctx, ctx->kinds, rows))) return rc;
            CU(ctx, cudaMemcpyAsync(ctx->kinds.p, io->floor_kind + r_lo * uniform_c, rows, cudaMemcpyHostToDevice, sm));
            d_kinds = (const uint8_t *)ctx->kinds.p - r_lo * uniform_c;
            if (io->floor1_y) {
                if ((rc = ensure(ctx, ctx->ys, rows * LWB_MAX_POSTS * 4))) return rc;
                CU(ctx, cudaMemcpyAsync(ctx->ys.p, io->floor1_y + r_lo * uniform_c * LWB_MAX_POSTS, rows * LWB_MAX_POSTS * 4,
                                        cudaMemcpyHostToDevice, sm));
                d_ys = (const uint32_t *)ctx->ys.p - r_lo * uniform_c * LWB_MAX_POSTS;
            }
        }
        // descriptors and mode bytes share one device buffer; a prepared batch (device memory, spectrum
        // entry) owns it and replays the launch while no stream changes shape
        const bool capture = plan && !host && !residue;
        DevBuf &dbuf = capture ? plan->mix : ctx->cdesc;
        const size_t used_desc = n_launch * sizeof(ChainDesc);
        if ((rc = ensure(ctx, dbuf, used_desc + boff + 16))) return rc;
        CU(ctx, cudaMemcpyAsync(dbuf.p, hd, used_desc, cudaMemcpyHostToDevice, sm));
        CU(ctx, cudaMemcpyAsync((char *)dbuf.p + used_desc, hb, boff + 16, cudaMemcpyHostToDevice, sm));
        CU(ctx, cudaEventRecord(st->ev, sm));
        st->pending = true;
        MixLaunch ml;
        ml.db = (char *)dbuf.p; ml.off_cd = 0; ml.off_by = used_desc; ml.pack = nullptr; ml.w_short = nullptr; ml.ls = 0;
        ml.i16 = false; ml.residue = residue; ml.out_format = io->out_format; ml.warps = maxc * wpc; ml.smem = smem;
        ml.n1max = n1max; ml.wpc = wpc; ml.np = np; ml.coeffs = d_coeffs; ml.dense = d_dense; ml.kinds = d_kinds; ml.ys = d_ys; ml.pcm = d_pcm;
        std::vector<MixRound> rounds(1, MixRound{0, 0, 0, n_launch});
        if ((rc = mixed_launch_rounds(ctx, ml, rounds))) return rc;
        if (capture) {
            plan->mixed_captured = true;
            plan->gen = gen_at_entry;
            plan->mix_launch = ml;
            plan->mix_rounds = std::move(rounds);
        }
        if (host) {
            if (o_hi > o_lo)
                CU(ctx, cudaMemcpyAsync((char *)io->pcm + o_lo * esz, ctx->pcm.p, (size_t)(o_hi - o_lo) * esz, cudaMemcpyDeviceToHost, sm));
            CU(ctx, cudaStreamSynchronize(sm));
        }
    }
    for (auto &e : ends)
        if (e.touched) set_stream_state(e.s, e.has, e.plen);
    return LWB_OK;
}

// ---------------------------------------------------------------------------------------------
// Mixed short/long streams (the standard 256/2048 Vorbis shape): each chain is cut into segments
// -- maximal runs of long blocks with long neighbours go to the fused kernel, everything else to
// the chain kernel -- and the segments of all chains are executed round by round, handing the
// overlap state over through the stream's device state (PreviousWindowRight) between launches.
// ---------------------------------------------------------------------------------------------
static int try_mixed(lwb_ctx *ctx, lwb_chain *chains, size_t n_chains, const lwb_batch_io *io, uint64_t epoch, bool *handled,
                     lwb_plan *plan = nullptr)
{
    *handled = false;
    const uint64_t gen_at_entry = ctx->state_gen;
    if (plan) plan->mixed_captured = false;
    if (getenv("LWB_FORCE_GENERIC") || getenv("LWB_NO_MIXED")) return LWB_OK;
    if (io->out_format != LWB_OUT_F32_PLANAR && io->out_format != LWB_OUT_I16_PLANAR) return LWB_OK;
    const bool residue = io->entry == LWB_ENTRY_RESIDUE;
    const bool i16 = io->out_format == LWB_OUT_I16_PLANAR;
    const size_t esz = i16 ? 2 : 4;
    unsigned maxc = 1;
    int n1max = 64, n0max = 64, bs0 = -1;
    size_t total_packets = 0, long_like = 0;
    const float *pack = nullptr, *w_short = nullptr;
    for (size_t i = 0; i < n_chains; i++) {
        const lwb_chain *c = &chains[i];
        if (!c->stream || c->stream->ctx != ctx || (c->n_packets && !c->mode_numbers)) return LWB_OK;
        const lwb_setup *su = c->stream->setup;
        if (su->channels > 8 || su->bs1 != kLongBs || !su->host.tab[1].pack) return LWB_OK;
        if (pack && pack != su->host.tab[1].pack) return LWB_OK;
        pack = su->host.tab[1].pack;
        // one short window for the whole batch (the fused kernel takes it as a launch argument)
        if (bs0 >= 0 && (bs0 != su->bs0 || w_short != su->host.tab[0].window)) return LWB_OK;
        bs0 = su->bs0;
        w_short = su->host.tab[0].window;
        if ((c->out_offset & 3) || (c->out_stride & 3) || (c->coeff_offset & 3)) return LWB_OK;
        maxc = std::max<unsigned>(maxc, su->channels);
        n1max = std::max(n1max, 1 << su->bs1);
        n0max = std::max(n0max, 1 << su->bs0);
        total_packets += c->n_packets;
        for (uint32_t k = 0; k < c->n_packets; k++) {
            const uint8_t m = c->mode_numbers[k];
            if (m < su->n_modes && su->host.mode_blockflag[m]) long_like++;
        }
    }
    // worth it only if the fused kernel gets a good share of the packets (every hand-over between the
    // kernels costs a launch): at least half the packets long blocks
    if (long_like * 2 < total_packets) return LWB_OK;
    const int ls_long = (kLongN - (1 << bs0)) >> 2, pl_short = 1 << (bs0 - 1);
    if (residue && !io->floor_kind) return fail(ctx, LWB_ERR_INVALID, "residue entry needs floor_kind");
    *handled = true;

    struct Seg { bool is_long, first_short, last_short; uint32_t p0, n; bool has; uint32_t plen; uint64_t coeff, pos; };
    bool chain_sees_long = false;       // the chain kernel's shared memory is sized for what it actually gets
    struct Walk { uint32_t seg0, n_seg; bool end_has; uint32_t end_plen; bool touched; uint32_t boff; };
    std::vector<Walk> walks(n_chains);
    std::vector<Seg> segs;
    segs.reserve(n_chains * 2);
    struct Pk { bool has; uint32_t plen; uint64_t coeff, pos; };
    std::vector<Pk> pk;
    std::vector<uint8_t> bytes(total_packets * 3 + 16);
    size_t boff = 0, max_rounds = 0;
    uint64_t c_lo = ~0ull, c_hi = 0, o_lo = ~0ull, o_hi = 0, r_lo = ~0ull, r_hi = 0;
    int uniform_c = -1;
    bool need_dense = false;
    std::vector<uint8_t> is_l;           // bit0 fused-kernel packet, bit1 follows a short block, bit2 precedes one
    for (size_t i = 0; i < n_chains; i++) {
        lwb_chain *c = &chains[i];
        lwb_stream *s = c->stream;
        const lwb_setup *su = s->setup;
        if (s->busy_epoch == epoch) return fail(ctx, LWB_ERR_INVALID, "a stream appears in two chains of one batch");
        s->busy_epoch = epoch;
        const unsigned C = su->channels;
        if (residue) {
            if (uniform_c < 0) uniform_c = (int)C;
            if (uniform_c != (int)C) return fail(ctx, LWB_ERR_INVALID, "residue batches need one channel count");
        }
        Walk &w = walks[i];
        w.boff = (uint32_t)boff;
        bool has = s->has, clear_after = false;
        uint32_t plen = s->plen, done = 0;
        uint64_t coeff = c->coeff_offset, pos = 0;
        c->status = LWB_OK;
        // pass 1: geometry + which packets the fused kernel may take (state entering them is empty or 1024)
        if (pk.size() < c->n_packets) { pk.resize(c->n_packets); is_l.resize(c->n_packets); }
        w.seg0 = (uint32_t)segs.size();
        w.n_seg = 0;
        for (uint32_t k = 0; k < c->n_packets; k++) {
            Geom g;
            int grc = geometry(su, c->mode_numbers[k], c->prev_window_flags ? c->prev_window_flags[k] : 1,
                               c->next_window_flags ? c->next_window_flags[k] : 1, &g);
            if (grc) { c->status = grc; break; }
            if (has) {
                const uint32_t slope_len = 1u << ((g.slope_sel ? su->bs1 : su->bs0) - 1);
                if (slope_len < plen) { c->status = LWB_ERR_BAD_FORMAT; clear_after = true; break; }
                if (g.ls + plen > g.n) { c->status = LWB_ERR_MISMATCH; break; }
            }
            pk[k] = Pk{has, plen, coeff, pos};
            is_l[k] = 0;
            if (g.blockflag && g.n == (uint32_t)kLongN) {
                const bool fs = g.ls != 0, lsf = g.re != g.n;
                if (!has || plen == (fs ? (uint32_t)pl_short : (uint32_t)kLongN2)) is_l[k] = 1 | (fs ? 2 : 0) | (lsf ? 4 : 0);
            }
            bytes[boff + 3 * k] = c->mode_numbers[k];
            bytes[boff + 3 * k + 1] = c->prev_window_flags ? c->prev_window_flags[k] : 1;
            bytes[boff + 3 * k + 2] = c->next_window_flags ? c->next_window_flags[k] : 1;
            if (has) pos += g.rs - g.ls;
            coeff += (uint64_t)C * (g.n >> 1);
            has = true;
            plen = g.re - g.rs;
            done++;
        }
        c->packets_done = done;
        c->n_samples = (uint32_t)pos;
        w.end_has = clear_after ? false : has;
        w.end_plen = clear_after ? 0u : plen;
        w.touched = done > 0 || clear_after;
        boff += (size_t)done * 3;
        if (!done) continue;
        if (c->out_stride < pos) return fail(ctx, LWB_ERR_BUFFER, "chain: out_stride smaller than the samples produced");
        // pass 2: segments.  A fused-kernel run starts at a long block that follows a short one and ends at
        // one that precedes a short one; everything else is handed to the chain kernel.
        uint32_t k = 0;
        while (k < done) {
            uint32_t j = k + 1;
            if (is_l[k]) {
                while (j < done && is_l[j] && !(is_l[j - 1] & 4) && !(is_l[j] & 2)) j++;
                segs.push_back(Seg{true, (is_l[k] & 2) != 0, (is_l[j - 1] & 4) != 0, k, j - k, pk[k].has, pk[k].plen, pk[k].coeff,
                                     pk[k].pos});
            } else {
                while (j < done && !is_l[j]) j++;
                for (uint32_t q = k; q < j; q++)
                    if (su->host.mode_blockflag[c->mode_numbers[q]]) chain_sees_long = true;
                segs.push_back(Seg{false, false, false, k, j - k, pk[k].has, pk[k].plen, pk[k].coeff, pk[k].pos});
            }
            k = j;
        }
        w.n_seg = (uint32_t)segs.size() - w.seg0;
        max_rounds = std::max<size_t>(max_rounds, w.n_seg);
        c_lo = std::min(c_lo, c->coeff_offset);
        c_hi = std::max(c_hi, coeff);
        o_lo = std::min(o_lo, c->out_offset);
        o_hi = std::max(o_hi, c->out_offset + (uint64_t)(C - 1) * c->out_stride + pos);
        if (residue) {
            r_lo = std::min(r_lo, c->packet_index);
            r_hi = std::max<uint64_t>(r_hi, c->packet_index + done);
            for (uint64_t r = c->packet_index * C; r < (c->packet_index + done) * C; r++) {
                const uint8_t kd = io->floor_kind[r];
                if (kd > LWB_FLOOR_DENSE) return fail(ctx, LWB_ERR_INVALID, "floor_kind out of range");
                if (kd == LWB_FLOOR_ONE && !io->floor1_y) return fail(ctx, LWB_ERR_INVALID, "floor1_y missing");
                if (kd == LWB_FLOOR_DENSE) need_dense = true;
            }
        }
    }
    if (need_dense && !io->dense_floor) return fail(ctx, LWB_ERR_INVALID, "dense_floor missing");
    int rc = LWB_OK;
    if (!chain_sees_long) n1max = n0max;
    int wpc = std::max(1, std::min(8, n1max / 1024));
    while (wpc > 1 && (unsigned)wpc * maxc > 32) wpc >>= 1;
    const int np = chain_np(maxc, n1max, wpc, residue);
    const size_t smem = chain_smem(maxc, n1max, np);
    if (max_rounds) {
        const bool host = io->memory == LWB_MEM_HOST;
        cudaStream_t sm = ctx->stream;
        const float *d_coeffs = io->coeffs, *d_dense = io->dense_floor;
        char *d_pcm = (char *)io->pcm;
        if (host) {
            if ((rc = ensure(ctx, ctx->coeffs, (size_t)(c_hi - c_lo) * 4))) return rc;
            if (o_hi > o_lo && (rc = ensure(ctx, ctx->pcm, (size_t)(o_hi - o_lo) * esz))) return rc;
            CU(ctx, cudaMemcpyAsync(ctx->coeffs.p, io->coeffs + c_lo, (size_t)(c_hi - c_lo) * 4, cudaMemcpyHostToDevice, sm));
            d_coeffs = (const float *)ctx->coeffs.p - c_lo;
            if (need_dense) {
                if ((rc = ensure(ctx, ctx->dense, (size_t)(c_hi - c_lo) * 4))) return rc;
                CU(ctx, cudaMemcpyAsync(ctx->dense.p, io->dense_floor + c_lo, (size_t)(c_hi - c_lo) * 4, cudaMemcpyHostToDevice, sm));
                d_dense = (const float *)ctx->dense.p - c_lo;
            }
            d_pcm = (char *)ctx->pcm.p - o_lo * esz;
        }
        const uint8_t *d_kinds = nullptr;
        const uint32_t *d_ys = nullptr;
        if (residue) {
            const size_t rows = (size_t)(r_hi - r_lo) * uniform_c;
            if ((rc = ensure(ctx, ctx->kinds, rows))) return rc;
            CU(ctx, cudaMemcpyAsync(ctx->kinds.p, io->floor_kind + r_lo * uniform_c, rows, cudaMemcpyHostToDevice, sm));
            d_kinds = (const uint8_t *)ctx->kinds.p - r_lo * uniform_c;
            if (io->floor1_y) {
                if ((rc = ensure(ctx, ctx->ys, rows * LWB_MAX_POSTS * 4))) return rc;
                CU(ctx, cudaMemcpyAsync(ctx->ys.p, io->floor1_y + r_lo * uniform_c * LWB_MAX_POSTS, rows * LWB_MAX_POSTS * 4,
                                        cudaMemcpyHostToDevice, sm));
                d_ys = (const uint32_t *)ctx->ys.p - r_lo * uniform_c * LWB_MAX_POSTS;
            }
        }
        // descriptors of every round: [LongRun...][ChainDesc...][DevPacket (prologue of the long segments)...][mode bytes]
        // A round with few fused-kernel runs leaves most of the 148 x 8 warps idle and lasts as long as its
        // longest run: such rounds cut their runs (each cut costs one extra IMDCT, the primer packet whose
        // right half is all the next piece needs), as the all-long path does.
        const size_t target_runs = (size_t)ctx->sm_count * kLongWarps * 2;
        constexpr uint32_t kMinCutRun = 6;
        std::vector<size_t> round_long(max_rounds, 0);
        for (size_t i = 0; i < n_chains; i++)
            for (uint32_t q = 0; q < walks[i].n_seg; q++)
                if (segs[walks[i].seg0 + q].is_long) round_long[q] += chains[i].stream->setup->channels;
        std::vector<uint32_t> round_cut(max_rounds, 1);
        if (!getenv("LWB_MIXED_NO_CUTS"))
            for (size_t r = 0; r < max_rounds; r++)
                if (round_long[r] && round_long[r] < target_runs)
                    round_cut[r] = (uint32_t)std::min<size_t>(16, (target_runs + round_long[r] - 1) / round_long[r]);
        auto cuts_of = [&](const Seg &sg, size_t r) { return std::max<uint32_t>(1, std::min(round_cut[r], sg.n / kMinCutRun)); };
        size_t n_runs = 0, n_cd = 0, n_pro = 0;
        for (size_t i = 0; i < n_chains; i++)
            for (uint32_t q = 0; q < walks[i].n_seg; q++) {
                const Seg &sg = segs[walks[i].seg0 + q];
                if (sg.is_long) { n_runs += (size_t)chains[i].stream->setup->channels * cuts_of(sg, q); if (residue) n_pro += sg.n; }
                else n_cd++;
            }
        // a prepared batch (device memory, spectrum entry) owns its descriptors so that later executions replay them
        const bool capture = plan && !host && !residue;
        DevBuf &dbuf = capture ? plan->mix : ctx->cdesc;
        const size_t off_cd = n_runs * sizeof(LongRun), off_pro = off_cd + n_cd * sizeof(ChainDesc);
        const size_t off_by = off_pro + n_pro * sizeof(DevPacket), total = off_by + boff + 16;
        Staging *st;
        if ((rc = acquire_staging(ctx, total, &st))) return rc;
        if ((rc = ensure(ctx, dbuf, total))) return rc;
        char *hb = (char *)st->h, *db = (char *)dbuf.p;
        LongRun *h_runs = (LongRun *)hb;
        ChainDesc *h_cd = (ChainDesc *)(hb + off_cd);
        DevPacket *h_pro = (DevPacket *)(hb + off_pro);
        std::memcpy(hb + off_by, bytes.data(), boff);
        const float *d_spec = nullptr;
        if (residue && n_pro) {
            if ((rc = ensure(ctx, ctx->spec, (size_t)(c_hi - c_lo) * 4))) return rc;
            d_spec = (const float *)ctx->spec.p - c_lo;          // same element offsets as the coefficient arena
        }
        std::vector<MixRound> rounds(max_rounds);
        size_t wr = 0, wc = 0, wp = 0;
        for (size_t r = 0; r < max_rounds; r++) {
            rounds[r].r0 = wr;
            rounds[r].c0 = wc;
            // fused-kernel runs first, longest first (three buckets): the kernel hands runs out in
            // descriptor order, and a 64-packet run started last would be the whole round's tail
            for (int bucket = 0; bucket < 3; bucket++)
                for (size_t i = 0; i < n_chains; i++) {
                    if (r >= walks[i].n_seg) continue;
                    const Seg &sg = segs[walks[i].seg0 + r];
                    if (!sg.is_long) continue;
                    const uint32_t cuts = cuts_of(sg, r), piece = sg.n / cuts;
                    if ((piece >= 32 ? 0 : piece >= 8 ? 1 : 2) != bucket) continue;
                    const lwb_chain *c = &chains[i];
                    const lwb_stream *s = c->stream;
                    const lwb_setup *su = s->setup;
                    const unsigned C = su->channels;
                    // samples packet 0 emits (0 without history; a block after a short one emits 1024 - ls)
                    const size_t first_emit = sg.has ? (sg.first_short ? (size_t)kLongN2 - ls_long : (size_t)kLongN2) : 0;
                    for (unsigned ch = 0; ch < C; ch++) {
                        const float *in0 = (residue ? d_spec : d_coeffs) + sg.coeff + (size_t)ch * kLongN2;
                        char *out0 = d_pcm + (c->out_offset + (size_t)ch * c->out_stride + sg.pos) * esz;
                        for (uint32_t k = 0; k < cuts; k++) {
                            const size_t p0 = (size_t)sg.n * k / cuts, p1 = (size_t)sg.n * (k + 1) / cuts;
                            LongRun &lr = h_runs[wr++];
                            std::memset(&lr, 0, sizeof(lr));
                            lr.in_stride = (uint32_t)(C * kLongN2);
                            lr.state = s->d_state + (size_t)ch * state_stride(su);
                            lr.write_state = (k + 1 == cuts);
                            lr.last_short = (k + 1 == cuts) && sg.last_short;
                            if (k == 0) {
                                lr.in = in0;
                                lr.out = out0;
                                lr.n_packets = (uint32_t)(p1 - p0);
                                lr.has_prev = sg.has;
                                lr.first_short = sg.first_short;
                            } else {
                                lr.in = in0 + (p0 - 1) * (size_t)lr.in_stride;         // primer = packet p0 - 1
                                lr.out = out0 + (first_emit + (p0 - 1) * (size_t)kLongN2) * esz;
                                lr.n_packets = (uint32_t)(p1 - p0 + 1);
                                lr.has_prev = 0;
                            }
                        }
                    }
                    if (residue)
                        for (uint32_t q = 0; q < sg.n; q++) {
                            DevPacket &d = h_pro[wp++];
                            std::memset(&d, 0, sizeof(d));
                            d.setup = su->d_setup;
                            d.coeff_off = sg.coeff + (uint64_t)q * C * kLongN2;
                            d.pkt_index = c->packet_index + sg.p0 + q;
                            d.n = kLongN;
                            d.blockflag = 1;
                            d.mapping = su->host.mode_mapping[c->mode_numbers[sg.p0 + q]];
                            d.channels = (uint8_t)C;
                        }
                }
            for (size_t i = 0; i < n_chains; i++) {
                if (r >= walks[i].n_seg) continue;
                const Seg &sg = segs[walks[i].seg0 + r];
                if (sg.is_long) continue;
                const lwb_chain *c = &chains[i];
                const lwb_stream *s = c->stream;
                const lwb_setup *su = s->setup;
                ChainDesc &d = h_cd[wc++];
                std::memset(&d, 0, sizeof(d));
                d.setup = su->d_setup;
                d.state = s->d_state;
                d.coeff_off = sg.coeff;
                d.out_off = c->out_offset + sg.pos;
                d.out_stride = c->out_stride;
                d.pkt_index = c->packet_index + sg.p0;
                d.n_packets = sg.n;
                d.byte_off = walks[i].boff + 3 * sg.p0;
                d.state_stride = (uint32_t)state_stride(su);
                d.plen0 = (uint16_t)sg.plen;
                d.has0 = sg.has;
                d.channels = (uint8_t)su->channels;
            }
            rounds[r].nr = wr - rounds[r].r0;
            rounds[r].nc = wc - rounds[r].c0;
        }
        CU(ctx, cudaMemcpyAsync(db, hb, total, cudaMemcpyHostToDevice, sm));
        CU(ctx, cudaEventRecord(st->ev, sm));
        st->pending = true;
        if (residue && n_pro)
            if ((rc = launch(ctx, k_prologue, dim3((unsigned)n_pro), dim3(kPrologueThreads), 0, (const DevPacket *)(db + off_pro),
                             d_coeffs, d_dense, d_kinds, d_ys, const_cast<float *>(d_spec))))
                return rc;
        constexpr uint32_t kTicketPool = 1024;
        if (!ctx->ticket.p) {
            if ((rc = ensure(ctx, ctx->ticket, kTicketPool * sizeof(unsigned int)))) return rc;
            for (int k = 0; k < 2; k++) {
                CU(ctx, cudaEventCreateWithFlags(&ctx->ev_desc[k], cudaEventDisableTiming));
                CU(ctx, cudaEventCreateWithFlags(&ctx->ev_kdone[k], cudaEventDisableTiming));
            }
        }
        MixLaunch ml;
        ml.db = db; ml.off_cd = off_cd; ml.off_by = off_by; ml.pack = pack; ml.w_short = w_short; ml.ls = ls_long;
        ml.i16 = i16; ml.residue = residue; ml.out_format = io->out_format; ml.warps = maxc * wpc; ml.smem = smem;
        ml.n1max = n1max; ml.wpc = wpc; ml.np = np; ml.coeffs = d_coeffs; ml.dense = d_dense; ml.kinds = d_kinds; ml.ys = d_ys; ml.pcm = d_pcm;
        if ((rc = mixed_launch_rounds(ctx, ml, rounds))) return rc;
        if (capture) {
            plan->mixed_captured = true;
            plan->gen = gen_at_entry;
            plan->mix_launch = ml;
            plan->mix_rounds = std::move(rounds);
        }
        if (host) {
            if (o_hi > o_lo)
                CU(ctx, cudaMemcpyAsync((char *)io->pcm + o_lo * esz, ctx->pcm.p, (size_t)(o_hi - o_lo) * esz, cudaMemcpyDeviceToHost, sm));
            CU(ctx, cudaStreamSynchronize(sm));
        }
    }
    for (size_t i = 0; i < n_chains; i++)
        if (walks[i].touched) set_stream_state(chains[i].stream, walks[i].end_has, walks[i].end_plen);
    return LWB_OK;
}

static int decode_chains_impl(lwb_ctx *ctx, lwb_chain *chains, size_t n_chains, const lwb_batch_io *io, lwb_plan *prepared)
{
    if (!ctx || (!chains && n_chains) || !io) return LWB_ERR_INVALID;
    if (io->entry != LWB_ENTRY_SPECTRUM && io->entry != LWB_ENTRY_RESIDUE) return fail(ctx, LWB_ERR_INVALID, "bad entry");
    if (io->memory != LWB_MEM_HOST && io->memory != LWB_MEM_DEVICE) return fail(ctx, LWB_ERR_INVALID, "bad memory space");
    if (io->out_format < 0 || io->out_format > LWB_OUT_I16_INTERLEAVED) return fail(ctx, LWB_ERR_INVALID, "bad out_format");
    if (n_chains == 0) return LWB_OK;
    if (!io->coeffs || !io->pcm) return fail(ctx, LWB_ERR_INVALID, "null arena");
    CU(ctx, cudaSetDevice(ctx->device));
    static uint64_t epoch = 0;
    epoch++;
    {
        bool handled = false;
        const char *fg = getenv("LWB_FORCE_GENERIC");
        const bool no_fused = fg && std::strcmp(fg, "2") == 0;
        int rc0 = no_fused ? LWB_OK : try_long(ctx, chains, n_chains, io, epoch, &handled, nullptr, 0, prepared);
        if (rc0 || handled) return rc0;
        // residue-entry batches of uniform long blocks go prologue + fused kernel (below); everything
        // else that fits goes to the chain kernel
        const bool residue_long = !no_fused && !fg && io->entry == LWB_ENTRY_RESIDUE && batch_is_uniform_long(ctx, chains, n_chains, io);
        if (!residue_long) {
            if (!no_fused) {
                rc0 = try_mixed(ctx, chains, n_chains, io, epoch, &handled, prepared);
                if (rc0 || handled) return rc0;
            }
            rc0 = try_chain(ctx, chains, n_chains, io, epoch, &handled, prepared);
            if (rc0 || handled) return rc0;
        }
    }
    const bool residue = io->entry == LWB_ENTRY_RESIDUE;
    const bool planar = is_planar(io->out_format);
    std::vector<PlanChain> plan(n_chains);
    uint64_t c_lo = ~0ull, c_hi = 0, o_lo = ~0ull, o_hi = 0, r_lo = ~0ull, r_hi = 0;
    int uniform_c = -1;
    bool need_dense = false;
    for (size_t i = 0; i < n_chains; i++) {
        lwb_chain *c = &chains[i];
        if (!c->stream || c->stream->ctx != ctx || (c->n_packets && !c->mode_numbers))
            return fail(ctx, LWB_ERR_INVALID, "chain: bad stream or mode list");
        if (c->stream->busy_epoch == epoch) return fail(ctx, LWB_ERR_INVALID, "a stream appears in two chains of one batch");
        c->stream->busy_epoch = epoch;
        const int C = c->stream->setup->channels;
        if (residue) {
            if (uniform_c < 0) uniform_c = C;
            if (uniform_c != C) return fail(ctx, LWB_ERR_INVALID, "residue batches need one channel count");
            if (!io->floor_kind) return fail(ctx, LWB_ERR_INVALID, "residue entry needs floor_kind");
        }
        plan_chain(c, &plan[i]);
        PlanChain &pc = plan[i];
        if (pc.pk.empty()) continue;
        const PlanPacket &last = pc.pk.back();
        c_lo = std::min(c_lo, c->coeff_offset);
        c_hi = std::max(c_hi, last.coeff_off + (uint64_t)C * (last.g.n >> 1));
        const uint64_t ext = planar ? (uint64_t)(C - 1) * c->out_stride + c->n_samples : (uint64_t)c->n_samples * C;
        if (planar && c->out_stride < c->n_samples) return fail(ctx, LWB_ERR_BUFFER, "chain: out_stride smaller than the samples produced");
        o_lo = std::min(o_lo, c->out_offset);
        o_hi = std::max(o_hi, c->out_offset + ext);
        if (residue) {
            r_lo = std::min(r_lo, c->packet_index);
            r_hi = std::max<uint64_t>(r_hi, c->packet_index + pc.pk.size());
            for (uint64_t r = c->packet_index * C; r < (c->packet_index + pc.pk.size()) * C; r++) {
                const uint8_t kd = io->floor_kind[r];
                if (kd > LWB_FLOOR_DENSE) return fail(ctx, LWB_ERR_INVALID, "floor_kind out of range");
                if (kd == LWB_FLOOR_ONE && !io->floor1_y) return fail(ctx, LWB_ERR_INVALID, "floor1_y missing");
                if (kd == LWB_FLOOR_DENSE) need_dense = true;
            }
        }
    }
    if (need_dense && !io->dense_floor) return fail(ctx, LWB_ERR_INVALID, "dense_floor missing");
    int rc = LWB_OK;
    if (c_hi > c_lo) {
        DevArenas ar;
        std::memset(&ar, 0, sizeof(ar));
        const size_t esz = elem_size(io->out_format);
        if (io->memory == LWB_MEM_HOST) {
            // stage: H2D of the used coefficient range, D2H of the used pcm range
            if ((rc = ensure(ctx, ctx->coeffs, (c_hi - c_lo) * sizeof(float)))) return rc;
            if (o_hi > o_lo && (rc = ensure(ctx, ctx->pcm, (o_hi - o_lo) * esz))) return rc;
            CU(ctx, cudaMemcpyAsync(ctx->coeffs.p, io->coeffs + c_lo, (c_hi - c_lo) * sizeof(float),
                                    cudaMemcpyHostToDevice, ctx->stream));
            ar.coeffs = (const float *)ctx->coeffs.p;
            ar.coeff_base = c_lo;
            if (need_dense) {
                if ((rc = ensure(ctx, ctx->dense, (c_hi - c_lo) * sizeof(float)))) return rc;
                CU(ctx, cudaMemcpyAsync(ctx->dense.p, io->dense_floor + c_lo, (c_hi - c_lo) * sizeof(float),
                                        cudaMemcpyHostToDevice, ctx->stream));
                ar.dense = (const float *)ctx->dense.p;
            }
            ar.pcm = ctx->pcm.p;
            ar.pcm_base = o_lo;
        } else {
            ar.coeffs = io->coeffs;
            ar.dense = io->dense_floor;
            ar.pcm = io->pcm;
        }
        if (residue) {
            const size_t rows = (size_t)(r_hi - r_lo) * uniform_c;
            if ((rc = ensure(ctx, ctx->kinds, rows))) return rc;
            CU(ctx, cudaMemcpyAsync(ctx->kinds.p, io->floor_kind + r_lo * uniform_c, rows, cudaMemcpyHostToDevice, ctx->stream));
            ar.kinds = (const uint8_t *)ctx->kinds.p;
            if (io->floor1_y) {
                if ((rc = ensure(ctx, ctx->ys, rows * LWB_MAX_POSTS * sizeof(uint32_t)))) return rc;
                CU(ctx, cudaMemcpyAsync(ctx->ys.p, io->floor1_y + r_lo * uniform_c * LWB_MAX_POSTS,
                                        rows * LWB_MAX_POSTS * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
                ar.ys = (const uint32_t *)ctx->ys.p;
            }
            ar.kinds_row0 = r_lo;
        }
        if (residue && plan_is_long(plan, io)) {
            // residue entry, uniform long blocks: k_prologue forms the spectrum on the device, the fused
            // kernel does the rest (one extra spectrum round trip compared with the spectrum entry)
            if ((rc = run_prologue_all(ctx, plan, ar, (size_t)(c_hi - c_lo)))) return rc;
            bool handled = false;
            rc = try_long(ctx, chains, n_chains, io, epoch, &handled, (const float *)ctx->spec.p, ar.coeff_base);
            if (rc) return rc;
            if (handled) return LWB_OK;            // try_long has committed results and stream states
        }
        rc = run_generic(ctx, plan, io, ar);
        if (rc) return rc;
        if (io->memory == LWB_MEM_HOST) {
            if (o_hi > o_lo)
                CU(ctx, cudaMemcpyAsync((char *)io->pcm + o_lo * esz, ctx->pcm.p, (o_hi - o_lo) * esz,
                                        cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    // commit the host-side view of every stream's state
    for (auto &pc : plan) {
        lwb_stream *s = pc.c->stream;
        if (!pc.pk.empty() || pc.clear_after) set_stream_state(s, pc.end_has, pc.end_plen);
    }
    return LWB_OK;
}

extern "C" int lwb_decode_chains(lwb_ctx *ctx, lwb_chain *chains, size_t n_chains, const lwb_batch_io *io)
{
    return decode_chains_impl(ctx, chains, n_chains, io, nullptr);
}

// ---------------------------------------------------------------------------------------------
// prepared batches
// ---------------------------------------------------------------------------------------------
extern "C" int lwb_plan_create(lwb_ctx *ctx, lwb_chain *chains, size_t n_chains, const lwb_batch_io *io, lwb_plan **out)
{
    if (!ctx || !out || (!chains && n_chains) || !io) return LWB_ERR_INVALID;
    lwb_plan *p = new (std::nothrow) lwb_plan();
    if (!p) return LWB_ERR_BUFFER;
    p->ctx = ctx;
    p->chains = chains;
    p->n_chains = n_chains;
    p->io = *io;
    *out = p;
    return LWB_OK;
}

extern "C" void lwb_plan_destroy(lwb_plan *p)
{
    if (!p) return;
    if (p->runs.p || p->mix.p) {
        cudaSetDevice(p->ctx->device);
        cudaStreamSynchronize(p->ctx->stream);
        if (p->runs.p) cudaFree(p->runs.p);
        if (p->mix.p) cudaFree(p->mix.p);
    }
    delete p;
}

extern "C" int lwb_plan_execute(lwb_plan *p)
{
    if (!p) return LWB_ERR_INVALID;
    lwb_ctx *ctx = p->ctx;
    if (p->captured && p->gen == ctx->state_gen && !getenv("LWB_FORCE_GENERIC")) {
        // steady state: nothing about the batch or the stream states has changed shape since the
        // descriptors were built -- the per-chain results in the caller's array are still right,
        // the stream states stay (has, 1024): just launch.
        CU(ctx, cudaSetDevice(ctx->device));
        constexpr uint32_t kTicketPool = 1024;
        if (ctx->ticket_next % kTicketPool == 0)
            CU(ctx, cudaMemsetAsync(ctx->ticket.p, 0, kTicketPool * sizeof(unsigned int), ctx->stream));
        unsigned int *ticket = (unsigned int *)ctx->ticket.p + (ctx->ticket_next++ % kTicketPool);
        if (long_launch(ctx->stream, (const LongRun *)p->runs.p, p->n_groups, p->pack, ticket, ctx->sm_count, p->i16))
            return fail(ctx, LWB_ERR_CUDA, "long kernel launch", cudaGetLastError());
        ctx->launches++;
        return LWB_OK;
    }
    if (p->mixed_captured && p->gen == ctx->state_gen && !getenv("LWB_FORCE_GENERIC")) {
        CU(ctx, cudaSetDevice(ctx->device));
        return mixed_launch_rounds(ctx, p->mix_launch, p->mix_rounds);
    }
    return decode_chains_impl(ctx, p->chains, p->n_chains, &p->io, p);
}

// ---------------------------------------------------------------------------------------------
// one packet
// ---------------------------------------------------------------------------------------------
static int one_packet(lwb_stream *s, int entry, uint8_t mode, int prev_flag, int next_flag, const float *coeffs,
                      const lwb_packet *pkt, int out_format, void *out, size_t cap, size_t *n_samples)
{
    if (!s || !coeffs || !out || !n_samples) return LWB_ERR_INVALID;
    *n_samples = 0;
    Geom g;
    int rc = geometry(s->setup, mode, prev_flag, next_flag, &g);
    if (rc) return rc;
    const size_t produce = s->has ? g.rs - g.ls : 0;
    if (produce > cap) return LWB_ERR_BUFFER;          // checked before anything is consumed
    uint8_t m = mode, pf = (uint8_t)(prev_flag != 0), nf = (uint8_t)(next_flag != 0);
    lwb_chain c;
    std::memset(&c, 0, sizeof(c));
    c.stream = s;
    c.n_packets = 1;
    c.mode_numbers = &m;
    c.prev_window_flags = &pf;
    c.next_window_flags = &nf;
    c.out_stride = cap;
    lwb_batch_io io;
    std::memset(&io, 0, sizeof(io));
    io.entry = entry;
    io.memory = LWB_MEM_HOST;
    io.coeffs = coeffs;
    if (pkt) {
        io.dense_floor = pkt->dense_floor;
        io.floor_kind = pkt->floor_kind;
        io.floor1_y = pkt->floor1_y;
    }
    io.out_format = out_format;
    io.pcm = out;
    rc = lwb_decode_chains(s->ctx, &c, 1, &io);
    if (rc) return rc;
    if (c.status) return c.status;
    *n_samples = c.n_samples;
    return LWB_OK;
}

extern "C" int lwb_decode_packet(lwb_stream *s, const lwb_packet *pkt, int out_format, void *out, size_t cap,
                                 size_t *n_samples)
{
    if (!pkt || !pkt->floor_kind || !pkt->residue) return LWB_ERR_INVALID;
    return one_packet(s, LWB_ENTRY_RESIDUE, pkt->mode_number, pkt->prev_window_flag, pkt->next_window_flag,
                      pkt->residue, pkt, out_format, out, cap, n_samples);
}

extern "C" int lwb_decode_spectrum(lwb_stream *s, uint8_t mode, int prev_flag, int next_flag, const float *spectrum,
                                   int out_format, void *out, size_t cap, size_t *n_samples)
{
    return one_packet(s, LWB_ENTRY_SPECTRUM, mode, prev_flag, next_flag, spectrum, nullptr, out_format, out, cap,
                      n_samples);
}

// ---------------------------------------------------------------------------------------------
// debug taps (lib.rs:56-94): intermediates of one packet, state untouched
// ---------------------------------------------------------------------------------------------
extern "C" int lwb_debug_packet_taps(lwb_stream *s, const lwb_packet *pkt, float *post_inverse, float *pre_mdct,
                                     float *post_mdct)
{
    if (!s || !pkt || !pkt->floor_kind || !pkt->residue) return LWB_ERR_INVALID;
    lwb_ctx *ctx = s->ctx;
    const lwb_setup *su = s->setup;
    Geom g;
    int rc = geometry(su, pkt->mode_number, pkt->prev_window_flag, pkt->next_window_flag, &g);
    if (rc) return rc;
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t C = su->channels, n2 = g.n >> 1;
    bool need_dense = false, need_y = false;
    for (size_t c = 0; c < C; c++) {
        if (pkt->floor_kind[c] > LWB_FLOOR_DENSE) return LWB_ERR_INVALID;
        need_dense |= pkt->floor_kind[c] == LWB_FLOOR_DENSE;
        need_y |= pkt->floor_kind[c] == LWB_FLOOR_ONE;
    }
    if ((need_dense && !pkt->dense_floor) || (need_y && !pkt->floor1_y)) return LWB_ERR_INVALID;
    if ((rc = ensure(ctx, ctx->coeffs, C * n2 * 4)) || (rc = ensure(ctx, ctx->spec, C * n2 * 4)) ||
        (rc = ensure(ctx, ctx->x, C * g.n * 4)) || (rc = ensure(ctx, ctx->kinds, C)) ||
        (rc = ensure(ctx, ctx->ys, C * LWB_MAX_POSTS * 4)) || (rc = ensure(ctx, ctx->dense, C * n2 * 4)) ||
        (rc = ensure(ctx, ctx->desc, sizeof(DevPacket))) || (rc = ensure_pinned(ctx, sizeof(DevPacket))))
        return rc;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    DevPacket *d = (DevPacket *)ctx->h_desc;
    std::memset(d, 0, sizeof(*d));
    d->setup = su->d_setup;
    d->state = s->d_state;
    d->prev_packet = -1;
    d->state_stride = (uint32_t)state_stride(su);
    d->n = (uint16_t)g.n;
    d->ls = (uint16_t)g.ls; d->rs = (uint16_t)g.rs; d->re = (uint16_t)g.re;
    d->blockflag = g.blockflag; d->mapping = g.mapping; d->slope_sel = g.slope_sel;
    d->channels = (uint8_t)C;
    cudaStream_t st = ctx->stream;
    CU(ctx, cudaMemcpyAsync(ctx->desc.p, d, sizeof(*d), cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->coeffs.p, pkt->residue, C * n2 * 4, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->kinds.p, pkt->floor_kind, C, cudaMemcpyHostToDevice, st));
    if (need_y) CU(ctx, cudaMemcpyAsync(ctx->ys.p, pkt->floor1_y, C * LWB_MAX_POSTS * 4, cudaMemcpyHostToDevice, st));
    if (need_dense) CU(ctx, cudaMemcpyAsync(ctx->dense.p, pkt->dense_floor, C * n2 * 4, cudaMemcpyHostToDevice, st));
    const DevPacket *dp = (const DevPacket *)ctx->desc.p;
    if (post_inverse) {
        // audio.rs:1004 tap: coupling only -- run the prologue with every floor "dense = 1.0"?  No:
        // the tap is taken by running the prologue on a copy with all floors unused replaced by a
        // unit curve, so that floor x residue leaves the decoupled residue unchanged.
        std::vector<float> ones(C * n2, 1.0f);
        std::vector<uint8_t> kd(C, LWB_FLOOR_DENSE);
        void *tmp_dense = nullptr, *tmp_kinds = nullptr;
        CU(ctx, cudaMalloc(&tmp_dense, C * n2 * 4));
        CU(ctx, cudaMalloc(&tmp_kinds, C));
        CU(ctx, cudaMemcpyAsync(tmp_dense, ones.data(), C * n2 * 4, cudaMemcpyHostToDevice, st));
        CU(ctx, cudaMemcpyAsync(tmp_kinds, kd.data(), C, cudaMemcpyHostToDevice, st));
        rc = launch(ctx, k_prologue, dim3(1), dim3(kPrologueThreads), 0, dp, (const float *)ctx->coeffs.p,
                    (const float *)tmp_dense, (const uint8_t *)tmp_kinds, (const uint32_t *)ctx->ys.p,
                    (float *)ctx->spec.p);
        if (!rc) {
            cudaError_t e = cudaMemcpyAsync(post_inverse, ctx->spec.p, C * n2 * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) rc = fail(ctx, LWB_ERR_CUDA, "tap copy", e);
        }
        cudaFree(tmp_dense);
        cudaFree(tmp_kinds);
        if (rc) return rc;
    }
    if ((rc = launch(ctx, k_prologue, dim3(1), dim3(kPrologueThreads), 0, dp, (const float *)ctx->coeffs.p,
                     (const float *)ctx->dense.p, (const uint8_t *)ctx->kinds.p, (const uint32_t *)ctx->ys.p,
                     (float *)ctx->spec.p)))
        return rc;
    if (pre_mdct) CU(ctx, cudaMemcpyAsync(pre_mdct, ctx->spec.p, C * n2 * 4, cudaMemcpyDeviceToHost, st));
    if (post_mdct) {
        if ((rc = launch(ctx, k_imdct, dim3(1, (unsigned)C), dim3(kImdctThreads), g.n * sizeof(float), dp,
                         (const float *)ctx->spec.p, (float *)ctx->x.p)))
            return rc;
        CU(ctx, cudaMemcpyAsync(post_mdct, ctx->x.p, C * g.n * 4, cudaMemcpyDeviceToHost, st));
    }
    CU(ctx, cudaStreamSynchronize(st));
    return LWB_OK;
}
