// fe_fuzz.cpp -- TEST INFRASTRUCTURE: mutation fuzzing of the host front half under
// AddressSanitizer + UndefinedBehaviorSanitizer.  Headers, audio packets and Ogg pages are attacker-
// controlled input in a decoder; whatever the bytes are, the parser must return a status, never read or
// write out of bounds, never unwind across the C ABI.  The corpus (valid headers / packets / an Ogg
// stream from tests/vorbis_packer.py) comes in a file; mutations are deterministic (xorshift).
// Build + run: tests/test_frontend_fuzz.py.
#define LWF_MAX_ENTRIES (1u << 14)
#define LWF_MAX_VQ_ELEMS (1ull << 18)
#include "../../lewton_b200/csrc/frontend.cpp"

#include <cstdio>
#include <cstdlib>
#include <string>

// the CUDA back half is not linked: the fuzz target is the host parser
extern "C" {
int lwb_setup_create(lwb_ctx *, const lwb_setup_desc *, lwb_setup **) { return LWB_ERR_NO_DEVICE; }
void lwb_setup_destroy(lwb_setup *) {}
int lwb_stream_open(lwb_ctx *, const lwb_setup *, lwb_stream **) { return LWB_ERR_NO_DEVICE; }
void lwb_stream_destroy(lwb_stream *) {}
int lwb_stream_reset(lwb_stream *) { return LWB_OK; }
int lwb_decode_packet(lwb_stream *, const lwb_packet *, int, void *, size_t, size_t *) { return LWB_ERR_NO_DEVICE; }
int lwb_decode_chains(lwb_ctx *, lwb_chain *, size_t, const lwb_batch_io *) { return LWB_ERR_NO_DEVICE; }
void *lwb_host_alloc(size_t n) { return std::malloc(n); }
void lwb_host_free(void *p) { std::free(p); }
}

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint64_t rnd()
{
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}

using Bytes = std::vector<uint8_t>;

static Bytes mutate(const Bytes &in)
{
    Bytes b = in;
    const int kind = (int)(rnd() % 10);
    if (b.empty()) return b;
    if (kind < 6) {                               // flip 1..8 bits
        const int k = 1 + (int)(rnd() % 8);
        for (int i = 0; i < k; i++) b[rnd() % b.size()] ^= (uint8_t)(1u << (rnd() % 8));
    } else if (kind < 8) {                        // truncate
        b.resize(rnd() % b.size());
    } else if (kind < 9) {                        // overwrite a run with random bytes
        const size_t at = rnd() % b.size(), n = 1 + rnd() % 16;
        for (size_t i = at; i < b.size() && i < at + n; i++) b[i] = (uint8_t)rnd();
    } else {                                      // duplicate a tail
        const size_t at = rnd() % b.size();
        b.insert(b.end(), b.begin() + at, b.end());
    }
    return b;
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    const long iters = std::atol(argv[2]);
    std::vector<Bytes> items;
    for (;;) {
        uint32_t n;
        if (std::fread(&n, 4, 1, f) != 1) break;
        Bytes b(n);
        if (n && std::fread(b.data(), 1, n, f) != n) return 2;
        items.push_back(std::move(b));
    }
    std::fclose(f);
    if (items.size() < 5) return 2;
    const Bytes &ident = items[0], &comment = items[1], &setup = items[2], &ogg = items[3];
    std::vector<Bytes> packets(items.begin() + 4, items.end());

    lwf_headers *good = nullptr;
    if (lwf_headers_parse(ident.data(), ident.size(), comment.data(), comment.size(), setup.data(), setup.size(), &good)) return 3;
    lwf_info info;
    lwf_headers_info(good, &info);
    long parsed_ok = 0, decoded_ok = 0, ogg_packets = 0;

    auto decode_all = [&](const lwf_headers *h, bool mutate_packets) {
        lwf_info inf;
        lwf_headers_info(h, &inf);
        const size_t C = inf.audio_channels, n2 = (size_t)1 << (inf.blocksize_1 - 1);
        std::vector<uint8_t> kinds(C);
        std::vector<uint32_t> ys(C * LWB_MAX_POSTS);
        std::vector<float> dense(C * n2), res(C * n2);
        for (const Bytes &pk : packets) {
            const Bytes m = mutate_packets ? mutate(pk) : pk;
            lwf_decoded_packet dp;
            std::memset(&dp, 0, sizeof(dp));
            dp.floor_kind = kinds.data();
            dp.floor1_y = ys.data();
            dp.dense_floor = dense.data();
            dp.residue = res.data();
            if (lwf_packet_decode(h, m.data(), m.size(), &dp) == LWB_OK) decoded_ok++;
            size_t cnt;
            lwf_decoded_sample_count(h, m.data(), m.size(), &cnt);
        }
    };

    if (argc > 3 && std::string(argv[3]) == "batch") {
        // the thread pool of lwf_batcher_decode (for ThreadSanitizer): many jobs over mutated packets; the stub
        // synthesis call fails with LWB_ERR_NO_DEVICE after the parallel entropy decode has run
        lwf_batcher *bt = nullptr;
        if (lwf_batcher_create(reinterpret_cast<lwb_ctx *>(0x10), good, 8, &bt)) return 4;
        std::vector<float> pcm(1);
        for (long it = 0; it < iters; it++) {
            const size_t n_jobs = 64;
            std::vector<std::vector<Bytes>> store(n_jobs);
            std::vector<std::vector<const uint8_t *>> ptrs(n_jobs);
            std::vector<std::vector<size_t>> lens(n_jobs);
            std::vector<lwf_stream_job> jobs(n_jobs);
            for (size_t j = 0; j < n_jobs; j++) {
                for (const Bytes &pk : packets) store[j].push_back((rnd() & 3) ? pk : mutate(pk));
                for (const Bytes &b : store[j]) { ptrs[j].push_back(b.data()); lens[j].push_back(b.size()); }
                std::memset(&jobs[j], 0, sizeof(jobs[j]));
                jobs[j].stream = reinterpret_cast<lwb_stream *>(0x20);
                jobs[j].n_packets = (uint32_t)store[j].size();
                jobs[j].packets = ptrs[j].data();
                jobs[j].lengths = lens[j].data();
            }
            const int rc = lwf_batcher_decode(bt, jobs.data(), n_jobs, LWB_OUT_F32_PLANAR, pcm.data());
            if (rc != LWB_ERR_NO_DEVICE) { std::printf("unexpected rc %d\n", rc); return 5; }
        }
        lwf_batcher_destroy(bt);
        lwf_headers_destroy(good);
        std::printf("batch iterations %ld ok\n", iters);
        return 0;
    }
    for (long it = 0; it < iters; it++) {
        const int what = (int)(rnd() % 4);
        if (what == 0) {                          // hostile setup header, then packets through it
            const Bytes s = mutate(setup);
            lwf_headers *h = nullptr;
            if (lwf_headers_parse(ident.data(), ident.size(), comment.data(), comment.size(), s.data(), s.size(), &h) == LWB_OK) {
                parsed_ok++;
                decode_all(h, (rnd() & 1) != 0);
                char buf[64];
                lwf_headers_comment(h, 0, buf, sizeof(buf));
                lwf_headers_destroy(h);
            }
        } else if (what == 1) {                   // hostile ident / comment headers
            const Bytes i2 = mutate(ident), c2 = mutate(comment);
            lwf_headers *h = nullptr;
            if (lwf_headers_parse(i2.data(), i2.size(), c2.data(), c2.size(), setup.data(), setup.size(), &h) == LWB_OK) {
                parsed_ok++;
                decode_all(h, false);
                lwf_headers_destroy(h);
            }
        } else if (what == 2) {                   // hostile audio packets against good headers
            decode_all(good, true);
        } else {                                  // hostile Ogg bytes
            const Bytes o = mutate(ogg);
            lwf_ogg *rd = nullptr;
            if (lwf_ogg_open(o.data(), o.size(), &rd) == LWB_OK) {
                lwf_ogg_packet pk;
                while (lwf_ogg_next_packet(rd, &pk) == LWB_OK) {
                    ogg_packets++;
                    volatile uint8_t sink = 0;
                    for (size_t i = 0; i < pk.len; i++) sink ^= pk.data[i];     // touch every byte handed out
                    (void)sink;
                }
                lwf_ogg_close(rd);
            }
            lwf_reader *r = nullptr;                // the reader must fail cleanly without a device
            if (lwf_reader_open(nullptr, o.data(), o.size(), &r) == LWB_OK) lwf_reader_close(r);
        }
    }
    lwf_headers_destroy(good);
    std::printf("iterations %ld parsed_ok %ld decoded_ok %ld ogg_packets %ld\n", iters, parsed_ok, decoded_ok, ogg_packets);
    return 0;
}
