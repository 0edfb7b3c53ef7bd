// C ABI of the B200 lash hot paths (see include/lash_gpu.h for the contract and the reference
// seams each entry point replaces).  Host-side orchestration only: staging, tile planning,
// stream double-buffering, launches.  No CPU fallback anywhere: every entry point either runs the
// CUDA kernels or returns an error.
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/lash_gpu.h"
#include "kernels.h"

using namespace lash;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) {                                                                          \
            return fail(e__ == cudaErrorMemoryAllocation ? LASH_E_NOMEM : LASH_E_CUDA,                     \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                              \
        }                                                                                                  \
    } while (0)

extern "C" const char* lash_gpu_last_error(void) { return g_err.c_str(); }
extern "C" int lash_gpu_abi_version(void) { return LASH_GPU_ABI_VERSION; }
extern "C" int lash_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------
// growable device / pinned buffers
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max(n, (size_t)1 << 16);
        want += want / 4;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max(n, (size_t)1 << 16);
        want += want / 4;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// Staging buffers of one sketcher slot.  They are handed back to the context when a sketcher closes and reused by the next
// one: cudaFree / cudaFreeHost of tens of MiB per lash_sketch_close showed up as 10-500 ms "drain" outliers in the
// FASTA ingest runs (a host that sketches batch after batch opens and closes a sketcher per batch).
struct SlotBufs {
    DevBuf packed, meta, mask, text, aux;
    PinBuf meta_host;
    void release() {
        packed.release(); meta.release(); mask.release(); text.release(); aux.release(); meta_host.release();
    }
};

struct lash_ctx {
    int device = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr;
    double dist_ms = 0.0;
    uint64_t dist_launches = 0;
    // optional wrapping sum of the output bit patterns of the last lash_dist* call (lash_dist_set_checksum)
    bool want_checksum = false;
    uint64_t checksum = 0, checksum_cells = 0;
    // scratch of lash_dist / lash_dist_stream, kept across calls (cudaMalloc/cudaFree per call cost
    // milliseconds of jitter on a 2.5 ms operation)
    DevBuf d_ref, d_qry, d_card, d_out[2], d_flags, d_regmin, d_ml, d_sum;
    DevBuf d_hmh_terms_r, d_hmh_terms_q, d_hmh_ec, d_hmh_idx;   // HMH small-sketch path (setup_hmh_ec)
    DevBuf d_hll_mm;                                            // HLL per-sketch register min / max (prepare_regmin)
    PinBuf h_out[2];
    std::mutex slot_mu;
    std::vector<SlotBufs> slot_cache;   // at most kSlotCacheMax sets
};
static constexpr size_t kSlotCacheMax = 4;

extern "C" int lash_ctx_create(int device, lash_ctx** out) {
    if (!out) return fail(LASH_E_INVALID, "lash_ctx_create: out is NULL");
    int n = lash_gpu_device_count();
    if (n <= 0) return fail(LASH_E_CUDA, "lash_ctx_create: no CUDA device visible (this library has no CPU fallback)");
    if (device < 0 || device >= n) return fail(LASH_E_INVALID, "lash_ctx_create: device index out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(LASH_E_CUDA, std::string("lash_ctx_create: device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                     ", this build carries sm_100a code only");
    lash_ctx* c = new (std::nothrow) lash_ctx();
    if (!c) return fail(LASH_E_NOMEM, "lash_ctx_create: out of host memory");
    c->device = device;
    c->n_sm = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(ensure_tables());
    *out = c;
    return LASH_OK;
}
extern "C" int lash_ctx_destroy(lash_ctx* c) {
    if (!c) return LASH_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamDestroy(c->stream);
    c->d_ref.release(); c->d_qry.release(); c->d_card.release(); c->d_out[0].release(); c->d_out[1].release();
    c->d_flags.release(); c->d_regmin.release(); c->d_ml.release(); c->d_sum.release(); c->h_out[0].release(); c->h_out[1].release();
    for (auto& b : c->slot_cache) b.release();
    c->d_hll_mm.release();
    c->d_hmh_terms_r.release(); c->d_hmh_terms_q.release(); c->d_hmh_ec.release(); c->d_hmh_idx.release();
    delete c;
    return LASH_OK;
}
extern "C" int lash_ctx_device(const lash_ctx* c) { return c ? c->device : -1; }

// Pin the calling thread to the CPUs next to `device` (sysfs local_cpulist of its PCI function).  Pinned staging
// memory is placed on the NUMA node of the thread that allocates it; on a two-socket box with one process per GPU,
// ranks whose staging memory sits on the far socket share the inter-socket link for every H2D byte.
static bool parse_cpulist(const std::string& text, cpu_set_t* set) {
    CPU_ZERO(set);
    int n = 0;
    size_t i = 0;
    while (i < text.size()) {
        while (i < text.size() && !isdigit((unsigned char)text[i])) ++i;
        if (i >= text.size()) break;
        long a = strtol(text.c_str() + i, nullptr, 10);
        while (i < text.size() && isdigit((unsigned char)text[i])) ++i;
        long b = a;
        if (i < text.size() && text[i] == '-') {
            ++i;
            b = strtol(text.c_str() + i, nullptr, 10);
            while (i < text.size() && isdigit((unsigned char)text[i])) ++i;
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) {
            CPU_SET((int)c, set);
            ++n;
        }
    }
    return n > 0;
}
extern "C" int lash_bind_thread_to_device(int device) {
    int n = lash_gpu_device_count();
    if (device < 0 || device >= n) return fail(LASH_E_INVALID, "lash_bind_thread_to_device: device index out of range");
    char bus[32] = {0};
    CU(cudaDeviceGetPCIBusId(bus, sizeof(bus), device));
    std::string id(bus);
    for (auto& ch : id) ch = (char)tolower((unsigned char)ch);
    const std::string path = "/sys/bus/pci/devices/" + id + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return fail(LASH_E_STATE, "lash_bind_thread_to_device: cannot read " + path);
    char buf[4096];
    const size_t got = fread(buf, 1, sizeof(buf) - 1, f);
    fclose(f);
    buf[got] = 0;
    cpu_set_t want, have, both;
    if (!parse_cpulist(buf, &want)) return fail(LASH_E_STATE, "lash_bind_thread_to_device: empty local_cpulist (no NUMA information)");
    // stay inside the CPUs this process is allowed to use (containers, taskset)
    if (sched_getaffinity(0, sizeof(have), &have) != 0) return fail(LASH_E_STATE, "lash_bind_thread_to_device: sched_getaffinity failed");
    CPU_AND(&both, &want, &have);
    if (CPU_COUNT(&both) == 0) return fail(LASH_E_STATE, "lash_bind_thread_to_device: none of the device-local CPUs is available to this process");
    if (sched_setaffinity(0, sizeof(both), &both) != 0) return fail(LASH_E_STATE, "lash_bind_thread_to_device: sched_setaffinity failed");
    return CPU_COUNT(&both);
}

extern "C" int lash_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(LASH_E_INVALID, "lash_host_alloc: out is NULL");
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return LASH_OK;
}
extern "C" int lash_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return LASH_OK;
}

static bool valid_algo_p(int algo, int p) {
    if (algo == LASH_ALGO_HMH) return true;
    if (algo == LASH_ALGO_HLL) return p >= 4 && p <= 18;   // streaming_algorithms threshold table range
    if (algo == LASH_ALGO_ULL) return p >= 3 && p <= 26;   // ultraloglog::new range
    return false;
}
extern "C" size_t lash_sketch_reg_bytes(int algo, int p) {
    if (algo == LASH_ALGO_HMH) return 32768;
    if (!valid_algo_p(algo, p)) return 0;
    return (size_t)1 << p;
}
extern "C" uint64_t lash_sketch_padded_bytes(uint64_t n_bases) {
    uint64_t b = (n_bases + 3) / 4;
    return ((b + 15) / 16) * 16 + 16;
}

// ------------------------------------------------------------------------------------------------
// sketcher
// ------------------------------------------------------------------------------------------------
static constexpr int kSlots = 2;  // double-buffered staging: copy of push n+1 overlaps kernels of push n

struct Slot : SlotBufs {
    cudaStream_t stream = nullptr;
    cudaEvent_t copied = nullptr;   // H2D of packed + metadata done
    cudaEvent_t meta_ready = nullptr;  // metadata upload (on the sketcher's meta stream) done
    cudaEvent_t k_start = nullptr, k_stop = nullptr;
    bool timing_pending = false;
    // SlotBufs: packed / meta / mask, text + aux (lash_sketch_push_ascii: raw text as the H2D target, block counts /
    // prefixes / kept bases per span), meta_host (pinned)
    uint64_t ticket = 0;
    bool used = false;
};

struct lash_sketcher {
    lash_ctx* ctx = nullptr;
    SketchParams sp;
    uint64_t n_genomes = 0;
    size_t reg_bytes = 0;
    uint32_t* acc = nullptr;  // [n_genomes][cell_words]
    Slot slot[kSlots];
    // tile / record tables go up on their own stream, so the upload of push n+1 overlaps the kernels of push n instead
    // of sitting in front of its own kernel (1.5 MB of tiles = 60 us per push at config 2)
    cudaStream_t meta_stream = nullptr;
    uint64_t next_ticket = 1;
    double kernel_ms = 0.0;
    uint64_t launches = 0;
    uint64_t min_chunk = 0;
    uint64_t per_iter = 0;  // k-mer starts one CTA iteration covers
    std::vector<SketchTile> plan_tiles;  // host-side tile plan of the current push (capacity reused across pushes)
    std::vector<SpanRecs> plan_mspans;
    std::vector<TextBlock> plan_tblocks;
    std::vector<TextSpanDev> plan_tspans;
    cudaStream_t ext_stream = nullptr;  // caller-provided stream (lash_sketch_set_stream)
};

static int harvest_timing(lash_sketcher* s, Slot& sl) {
    if (sl.timing_pending) {
        CU(cudaEventSynchronize(sl.k_stop));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, sl.k_start, sl.k_stop));
        s->kernel_ms += ms;
        sl.timing_pending = false;
    }
    return LASH_OK;
}

extern "C" int lash_sketch_open(lash_ctx* ctx, int algo, int p, int k, uint64_t seed, uint64_t n_genomes,
                                lash_sketcher** out) {
    if (!ctx || !out) return fail(LASH_E_INVALID, "lash_sketch_open: NULL argument");
    if (k < 1 || k > 32) return fail(LASH_E_INVALID, "k-mer length must be 1-32");  // utils.rs:500-502
    if (!valid_algo_p(algo, p)) return fail(LASH_E_INVALID, "lash_sketch_open: bad algorithm / precision");
    if (n_genomes == 0 || n_genomes > 0xffffffffull) return fail(LASH_E_INVALID, "lash_sketch_open: n_genomes out of range");
    CU(cudaSetDevice(ctx->device));
    lash_sketcher* s = new (std::nothrow) lash_sketcher();
    if (!s) return fail(LASH_E_NOMEM, "lash_sketch_open: out of host memory");
    s->ctx = ctx;
    s->n_genomes = n_genomes;
    s->reg_bytes = lash_sketch_reg_bytes(algo, p);
    s->sp.algo = algo;
    s->sp.p = algo == LASH_ALGO_HMH ? 14 : p;
    s->sp.k = k;
    s->sp.hc = make_hash_consts(seed);
    plan_sketch(s->sp);
    s->sp.n_sm = ctx->n_sm;
    s->per_iter = (uint64_t)s->sp.threads * kStartsPerThread;
    // a CTA should see enough k-mers to warm its private accumulator (>= ~64 per cell)
    s->min_chunk = std::max<uint64_t>(s->per_iter, 64ull * s->sp.n_cells);
    size_t acc_bytes = (size_t)s->sp.cell_words * 4 * n_genomes;
    cudaError_t e = cudaMalloc((void**)&s->acc, acc_bytes);
    if (e != cudaSuccess) {
        delete s;
        return fail(LASH_E_NOMEM, std::string("lash_sketch_open: accumulator allocation failed: ") + cudaGetErrorString(e));
    }
    CU(cudaMemsetAsync(s->acc, 0, acc_bytes, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    {
        std::lock_guard<std::mutex> g(ctx->slot_mu);
        for (int i = 0; i < kSlots && !ctx->slot_cache.empty(); ++i) {
            static_cast<SlotBufs&>(s->slot[i]) = ctx->slot_cache.back();
            ctx->slot_cache.pop_back();
        }
    }
    for (int i = 0; i < kSlots; ++i) {
        CU(cudaStreamCreateWithFlags(&s->slot[i].stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s->slot[i].copied, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s->slot[i].meta_ready, cudaEventDisableTiming));
        CU(cudaEventCreate(&s->slot[i].k_start));
        CU(cudaEventCreate(&s->slot[i].k_stop));
    }
    CU(cudaStreamCreateWithFlags(&s->meta_stream, cudaStreamNonBlocking));
    *out = s;
    return LASH_OK;
}

static int push_impl(lash_sketcher* s, const void* packed, bool packed_on_device, uint64_t n_bytes, const lash_span* spans,
                     uint32_t n_spans, const uint64_t* rec_start, uint64_t n_rec_entries, uint64_t* ticket_out) {
    if (!s) return fail(LASH_E_INVALID, "lash_sketch_push: NULL sketcher");
    if (n_spans && (!spans || !packed)) return fail(LASH_E_INVALID, "lash_sketch_push: NULL buffer");
    CU(cudaSetDevice(s->ctx->device));
    const int k = s->sp.k;
    static const bool trace = getenv("LASH_TRACE_PUSH") != nullptr;
    const auto tp0 = std::chrono::steady_clock::now();

    // ---- validate + plan tiles (host) --------------------------------------------------------
    uint64_t total_starts = 0;
    uint64_t mask_words = 0;
    uint32_t n_multi = 0;
    uint64_t n_multi_recs = 0;
    bool need_rec_table = false;
    for (uint32_t i = 0; i < n_spans; ++i) {
        const lash_span& sp = spans[i];
        if (sp.genome >= s->n_genomes) return fail(LASH_E_INVALID, "lash_sketch_push: span.genome out of range");
        if (sp.byte_off % 16) return fail(LASH_E_INVALID, "lash_sketch_push: span.byte_off must be a multiple of 16");
        if (sp.byte_off + (sp.n_bases + 3) / 4 > n_bytes) return fail(LASH_E_INVALID, "lash_sketch_push: span exceeds buffer");
        if (sp.n_rec > 1 && sp.rec_len != 0) {
            // uniform record length (fixed-length reads): boundaries are arithmetic, no table
            if ((uint64_t)sp.n_rec != (sp.n_bases + sp.rec_len - 1) / sp.rec_len)
                return fail(LASH_E_INVALID, "lash_sketch_push: n_rec does not match n_bases / record length");
            ++n_multi;
            n_multi_recs += sp.n_rec;
            mask_words += ((sp.n_bases + 63) / 64) * 2 + 2;
        } else if (sp.n_rec > 1) {
            need_rec_table = true;
            if (!rec_start || sp.rec_first + sp.n_rec + 1 > n_rec_entries)
                return fail(LASH_E_INVALID, "lash_sketch_push: rec_start table too short");
            const uint64_t* rs = rec_start + sp.rec_first;
            if (rs[0] != 0 || rs[sp.n_rec] != sp.n_bases)
                return fail(LASH_E_INVALID, "lash_sketch_push: rec_start must run from 0 to n_bases");
            for (uint32_t r = 0; r < sp.n_rec; ++r)
                if (rs[r] > rs[r + 1]) return fail(LASH_E_INVALID, "lash_sketch_push: rec_start not ascending");
            ++n_multi;
            n_multi_recs += sp.n_rec;
            mask_words += ((sp.n_bases + 63) / 64) * 2 + 2;
        }
        if (sp.n_bases >= (uint64_t)k) total_starts += sp.n_bases - k + 1;
    }
    // Tiles are the balancing granule of the persistent CTAs (each CTA takes a contiguous run of
    // them): aim for ~32 tiles per resident CTA, never less than one CTA iteration.
    const uint64_t target_tiles = (uint64_t)s->ctx->n_sm * 8 * 32;
    uint64_t chunk = (total_starts + target_tiles - 1) / target_tiles;
    chunk = std::max(chunk, s->per_iter);
    const uint64_t kStartsPerIter = s->per_iter;
    chunk = ((chunk + kStartsPerIter - 1) / kStartsPerIter) * kStartsPerIter;

    // plan buffers live in the sketcher: a fresh 1.4 MB vector per push cost ~1 ms of page faults + regrowth (seen
    // with LASH_TRACE_PUSH on the GPU box), which is exposed whenever the GPU is not already busy with an older push
    std::vector<SketchTile>& tiles = s->plan_tiles;
    std::vector<SpanRecs>& mspans = s->plan_mspans;
    tiles.clear();
    mspans.clear();
    tiles.reserve(total_starts / chunk + 2 * (size_t)n_spans + 16);
    mspans.reserve(n_multi);
    uint64_t mask_off = 0;
    for (uint32_t i = 0; i < n_spans; ++i) {
        const lash_span& sp = spans[i];
        uint64_t this_mask = ~0ull;
        if (sp.n_rec > 1) {
            SpanRecs sr;
            sr.mask_word_off = mask_off;
            sr.rec_first = sp.rec_first;
            sr.n_bases = sp.n_bases;
            sr.n_rec = sp.n_rec;
            sr.uniform_len = sp.rec_len;
            mspans.push_back(sr);
            this_mask = mask_off;
            mask_off += ((sp.n_bases + 63) / 64) * 2 + 2;
        }
        if (sp.n_bases < (uint64_t)k) continue;  // utils.rs:460-462
        const uint64_t starts = sp.n_bases - k + 1;
        const uint64_t n_t = (starts + chunk - 1) / chunk;
        // equalise the chunks of one span (multiples of the CTA iteration)
        uint64_t per = (starts + n_t - 1) / n_t;
        per = ((per + kStartsPerIter - 1) / kStartsPerIter) * kStartsPerIter;
        for (uint64_t b = 0; b < starts; b += per) {
            SketchTile t;
            t.word_off = sp.byte_off / 4;
            t.mask_word_off = this_mask;
            t.begin = b;
            t.end = std::min(starts, b + per);
            t.genome = (uint32_t)sp.genome;
            t.clip = 0xffffffffu;
            tiles.push_back(t);
        }
    }
    if (tiles.size() > 0x7fffffffull) return fail(LASH_E_INVALID, "lash_sketch_push: too many tiles in one push");

    const auto tp1 = std::chrono::steady_clock::now();
    // ---- stage ---------------------------------------------------------------------------------
    const uint64_t ticket = s->next_ticket++;
    Slot& sl = s->slot[ticket % kSlots];
    const cudaStream_t stream = s->ext_stream ? s->ext_stream : sl.stream;
    if (sl.used) {
        // host-side reuse of this slot's pinned metadata and timing events
        CU(cudaEventSynchronize(sl.copied));
        int rc = harvest_timing(s, sl);
        if (rc) return rc;
    }
    const size_t tiles_bytes = tiles.size() * sizeof(SketchTile);
    const size_t mspans_bytes = mspans.size() * sizeof(SpanRecs);
    const size_t recs_bytes = need_rec_table ? n_rec_entries * sizeof(uint64_t) : 0;
    const size_t off_mspans = (tiles_bytes + 15) / 16 * 16;
    const size_t off_recs = off_mspans + (mspans_bytes + 15) / 16 * 16;
    const size_t meta_bytes = off_recs + recs_bytes;
    if (meta_bytes) {
        CU(sl.meta_host.reserve(meta_bytes));
        CU(sl.meta.reserve(meta_bytes));
        char* mh = (char*)sl.meta_host.p;
        if (tiles_bytes) memcpy(mh, tiles.data(), tiles_bytes);
        if (mspans_bytes) memcpy(mh + off_mspans, mspans.data(), mspans_bytes);
        if (recs_bytes) memcpy(mh + off_recs, rec_start, recs_bytes);
        static const bool meta_inline = getenv("LASH_META_SAME_STREAM") != nullptr;  // A/B switch
        if (meta_inline) {
            CU(cudaMemcpyAsync(sl.meta.p, mh, meta_bytes, cudaMemcpyHostToDevice, stream));
        } else {
            // the slot's previous kernels (two pushes ago) may still be reading sl.meta
            if (sl.used) CU(cudaStreamWaitEvent(s->meta_stream, sl.k_stop, 0));
            CU(cudaMemcpyAsync(sl.meta.p, mh, meta_bytes, cudaMemcpyHostToDevice, s->meta_stream));
            CU(cudaEventRecord(sl.meta_ready, s->meta_stream));
            CU(cudaStreamWaitEvent(stream, sl.meta_ready, 0));
        }
    }
    const uint32_t* packed_dev = nullptr;
    if (packed_on_device) {
        packed_dev = (const uint32_t*)packed;
    } else if (n_bytes) {
        CU(sl.packed.reserve(n_bytes + 64));
        CU(cudaMemcpyAsync(sl.packed.p, packed, n_bytes, cudaMemcpyHostToDevice, stream));
        packed_dev = (const uint32_t*)sl.packed.p;
    }
    CU(cudaEventRecord(sl.copied, stream));
    const auto tp2 = std::chrono::steady_clock::now();
    uint32_t* mask_dev = nullptr;
    if (n_multi) {
        CU(sl.mask.reserve(mask_off * 4));
        mask_dev = (uint32_t*)sl.mask.p;
        CU(cudaMemsetAsync(mask_dev, 0, mask_off * 4, stream));
    }
    CU(cudaEventRecord(sl.k_start, stream));
    if (n_multi) {
        CU(launch_build_invalid_mask((const SpanRecs*)((char*)sl.meta.p + off_mspans), n_multi, n_multi_recs,
                                     (const uint64_t*)((char*)sl.meta.p + off_recs), mask_dev, k, s->ctx->n_sm, stream));
        s->launches += 1;
    }
    if (!tiles.empty()) {
        CU(launch_sketch(s->sp, packed_dev, mask_dev, (const SketchTile*)sl.meta.p, (uint32_t)tiles.size(), s->acc, stream));
        s->launches += 1;
    }
    CU(cudaEventRecord(sl.k_stop, stream));
    if (trace) {
        const auto tp3 = std::chrono::steady_clock::now();
        auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        fprintf(stderr, "[lash push] %zu tiles: plan %.0f us, slot wait + uploads %.0f us, launches %.0f us\n", tiles.size(),
                us(tp0, tp1), us(tp1, tp2), us(tp2, tp3));
    }
    sl.timing_pending = true;
    sl.used = true;
    sl.ticket = ticket;
    if (ticket_out) *ticket_out = ticket;
    return LASH_OK;
}

extern "C" int lash_sketch_push(lash_sketcher* s, const uint8_t* packed, uint64_t n_bytes, const lash_span* spans,
                                uint32_t n_spans, const uint64_t* rec_start, uint64_t n_rec_entries, uint64_t* ticket) {
    return push_impl(s, packed, false, n_bytes, spans, n_spans, rec_start, n_rec_entries, ticket);
}
extern "C" int lash_sketch_push_dev(lash_sketcher* s, const void* packed_dev, uint64_t n_bytes, const lash_span* spans,
                                    uint32_t n_spans, const uint64_t* rec_start, uint64_t n_rec_entries, uint64_t* ticket) {
    if (((uintptr_t)packed_dev) % 16) return fail(LASH_E_INVALID, "lash_sketch_push_dev: device buffer must be 16-byte aligned");
    return push_impl(s, packed_dev, true, n_bytes, spans, n_spans, rec_start, n_rec_entries, ticket);
}
// ------------------------------------------------------------------------------------------------
// lash_sketch_push_ascii: raw sequence text in, filter_out_n + 2-bit pack on the device (text_kernels.cu), then the same
// sketch kernel.  The host plans the tiles on the raw byte count (an upper bound of the bases a span keeps); the kernel
// clips every tile to the kept count the pack kernels leave in span_kept[].
// ------------------------------------------------------------------------------------------------
static int push_text_impl(lash_sketcher* s, const void* text, bool text_on_device, uint64_t n_bytes, const lash_text_span* spans,
                          uint32_t n_spans, uint64_t* ticket_out) {
    if (!s) return fail(LASH_E_INVALID, "lash_sketch_push_ascii: NULL sketcher");
    if (n_spans && (!spans || !text)) return fail(LASH_E_INVALID, "lash_sketch_push_ascii: NULL buffer");
    CU(cudaSetDevice(s->ctx->device));
    const int k = s->sp.k;
    uint64_t total_starts = 0, out_bytes = 0, mask_words = 0, n_blocks64 = 0;
    for (uint32_t i = 0; i < n_spans; ++i) {
        const lash_text_span& sp = spans[i];
        if (sp.genome >= s->n_genomes) return fail(LASH_E_INVALID, "lash_sketch_push_ascii: span.genome out of range");
        if (sp.byte_off % 16) return fail(LASH_E_INVALID, "lash_sketch_push_ascii: span.byte_off must be a multiple of 16");
        if (sp.byte_off + sp.n_bytes > n_bytes) return fail(LASH_E_INVALID, "lash_sketch_push_ascii: span exceeds buffer");
        if (sp.n_bytes >= (uint64_t)k) total_starts += sp.n_bytes - k + 1;
        out_bytes += lash_sketch_padded_bytes(sp.n_bytes);
        if (sp.n_rec > 1) mask_words += ((sp.n_bytes + 63) / 64) * 2 + 2;
        n_blocks64 += (sp.n_bytes + kTextBlockBytes - 1) / kTextBlockBytes;
    }
    if (n_blocks64 > 0x7fffffffull) return fail(LASH_E_INVALID, "lash_sketch_push_ascii: push too large");
    static const bool trace = getenv("LASH_TRACE_PUSH") != nullptr;
    const auto tp0 = std::chrono::steady_clock::now();
    const uint64_t target_tiles = (uint64_t)s->ctx->n_sm * 8 * 32;
    uint64_t chunk = (total_starts + target_tiles - 1) / target_tiles;
    chunk = std::max(chunk, s->per_iter);
    chunk = ((chunk + s->per_iter - 1) / s->per_iter) * s->per_iter;

    std::vector<SketchTile>& tiles = s->plan_tiles;
    std::vector<TextBlock>& tblocks = s->plan_tblocks;
    std::vector<TextSpanDev>& tspans = s->plan_tspans;
    tiles.clear();
    tblocks.clear();
    tspans.clear();
    tiles.reserve(total_starts / chunk + 2 * (size_t)n_spans + 16);
    tblocks.reserve(n_blocks64);
    tspans.reserve(n_spans);
    uint64_t out_off = 0, mask_off = 0;
    for (uint32_t i = 0; i < n_spans; ++i) {
        const lash_text_span& sp = spans[i];
        TextSpanDev td;
        td.out_word_off = out_off / 4;
        td.mask_word_off = ~0ull;
        if (sp.n_rec > 1) {
            td.mask_word_off = mask_off;
            mask_off += ((sp.n_bytes + 63) / 64) * 2 + 2;
        }
        td.first_block = (uint32_t)tblocks.size();
        for (uint64_t b = 0; b < sp.n_bytes; b += kTextBlockBytes) {
            TextBlock tb;
            tb.byte_begin = sp.byte_off + b;
            tb.n_bytes = (uint32_t)std::min<uint64_t>(kTextBlockBytes, sp.n_bytes - b);
            tb.span = i;
            tblocks.push_back(tb);
        }
        td.n_blocks = (uint32_t)tblocks.size() - td.first_block;
        tspans.push_back(td);
        if (sp.n_bytes >= (uint64_t)k) {
            const uint64_t starts = sp.n_bytes - k + 1;
            const uint64_t n_t = (starts + chunk - 1) / chunk;
            uint64_t per = (starts + n_t - 1) / n_t;
            per = ((per + s->per_iter - 1) / s->per_iter) * s->per_iter;
            for (uint64_t b = 0; b < starts; b += per) {
                SketchTile t;
                t.word_off = out_off / 4;
                t.mask_word_off = td.mask_word_off;
                t.begin = b;
                t.end = std::min(starts, b + per);
                t.genome = (uint32_t)sp.genome;
                t.clip = i;
                tiles.push_back(t);
            }
        }
        out_off += lash_sketch_padded_bytes(sp.n_bytes);
    }
    if (tiles.size() > 0x7fffffffull) return fail(LASH_E_INVALID, "lash_sketch_push_ascii: too many tiles in one push");

    const auto tp1 = std::chrono::steady_clock::now();
    const uint64_t ticket = s->next_ticket++;
    Slot& sl = s->slot[ticket % kSlots];
    const cudaStream_t stream = s->ext_stream ? s->ext_stream : sl.stream;
    if (sl.used) {
        CU(cudaEventSynchronize(sl.copied));
        int rc = harvest_timing(s, sl);
        if (rc) return rc;
    }
    const auto tp2 = std::chrono::steady_clock::now();
    const size_t tiles_bytes = tiles.size() * sizeof(SketchTile);
    const size_t off_blocks = (tiles_bytes + 15) / 16 * 16;
    const size_t off_spans = off_blocks + (tblocks.size() * sizeof(TextBlock) + 15) / 16 * 16;
    const size_t meta_bytes = off_spans + tspans.size() * sizeof(TextSpanDev);
    if (meta_bytes) {
        CU(sl.meta_host.reserve(meta_bytes));
        CU(sl.meta.reserve(meta_bytes));
        char* mh = (char*)sl.meta_host.p;
        if (tiles_bytes) memcpy(mh, tiles.data(), tiles_bytes);
        if (!tblocks.empty()) memcpy(mh + off_blocks, tblocks.data(), tblocks.size() * sizeof(TextBlock));
        if (!tspans.empty()) memcpy(mh + off_spans, tspans.data(), tspans.size() * sizeof(TextSpanDev));
        if (sl.used) CU(cudaStreamWaitEvent(s->meta_stream, sl.k_stop, 0));
        CU(cudaMemcpyAsync(sl.meta.p, mh, meta_bytes, cudaMemcpyHostToDevice, s->meta_stream));
        CU(cudaEventRecord(sl.meta_ready, s->meta_stream));
        CU(cudaStreamWaitEvent(stream, sl.meta_ready, 0));
    }
    const uint8_t* text_dev = nullptr;
    if (text_on_device) {
        text_dev = (const uint8_t*)text;
    } else if (n_bytes) {
        CU(sl.text.reserve(n_bytes + 64));
        CU(cudaMemcpyAsync(sl.text.p, text, n_bytes, cudaMemcpyHostToDevice, stream));
        text_dev = (const uint8_t*)sl.text.p;
    }
    CU(cudaEventRecord(sl.copied, stream));
    const auto tp3 = std::chrono::steady_clock::now();
    const uint32_t n_blocks = (uint32_t)tblocks.size();
    uint32_t* mask_dev = nullptr;
    uint64_t* aux = nullptr;
    if (n_blocks) {
        CU(sl.packed.reserve(out_off + 64));
        CU(cudaMemsetAsync(sl.packed.p, 0, out_off + 64, stream));
        if (mask_off) {
            CU(sl.mask.reserve(mask_off * 4));
            mask_dev = (uint32_t*)sl.mask.p;
            CU(cudaMemsetAsync(mask_dev, 0, mask_off * 4, stream));
        }
        CU(sl.aux.reserve(((size_t)2 * n_blocks + 1 + n_spans) * 8));
        aux = (uint64_t*)sl.aux.p;
    }
    CU(cudaEventRecord(sl.k_start, stream));
    if (n_blocks) {
        uint64_t* block_cnt = aux;
        uint64_t* block_prefix = aux + n_blocks + 1;
        uint64_t* span_kept = block_prefix + n_blocks;
        CU(launch_text_pack(text_dev, (const TextBlock*)((char*)sl.meta.p + off_blocks), n_blocks,
                            (const TextSpanDev*)((char*)sl.meta.p + off_spans), n_spans, block_cnt, block_prefix, span_kept,
                            (uint32_t*)sl.packed.p, mask_dev, k, s->ctx->n_sm, stream));
        s->launches += 3;
        if (!tiles.empty()) {
            CU(launch_sketch(s->sp, (const uint32_t*)sl.packed.p, mask_dev, (const SketchTile*)sl.meta.p, (uint32_t)tiles.size(), s->acc,
                             stream, span_kept));
            s->launches += 1;
        }
    }
    CU(cudaEventRecord(sl.k_stop, stream));
    if (trace) {
        const auto tp4 = std::chrono::steady_clock::now();
        auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        fprintf(stderr, "[lash push_ascii] %llu bytes, %zu tiles, %u blocks: plan %.0f us, slot wait %.0f us, uploads %.0f us, launches %.0f us\n",
                (unsigned long long)n_bytes, tiles.size(), n_blocks, us(tp0, tp1), us(tp1, tp2), us(tp2, tp3), us(tp3, tp4));
    }
    sl.timing_pending = true;
    sl.used = true;
    sl.ticket = ticket;
    if (ticket_out) *ticket_out = ticket;
    return LASH_OK;
}
extern "C" int lash_sketch_push_ascii(lash_sketcher* s, const uint8_t* text, uint64_t n_bytes, const lash_text_span* spans,
                                      uint32_t n_spans, uint64_t* ticket) {
    return push_text_impl(s, text, false, n_bytes, spans, n_spans, ticket);
}
extern "C" int lash_sketch_push_ascii_dev(lash_sketcher* s, const void* text_dev, uint64_t n_bytes, const lash_text_span* spans,
                                          uint32_t n_spans, uint64_t* ticket) {
    if (((uintptr_t)text_dev) % 16) return fail(LASH_E_INVALID, "lash_sketch_push_ascii_dev: device buffer must be 16-byte aligned");
    return push_text_impl(s, text_dev, true, n_bytes, spans, n_spans, ticket);
}
extern "C" int lash_sketch_wait_copied(lash_sketcher* s, uint64_t ticket) {
    if (!s) return fail(LASH_E_INVALID, "lash_sketch_wait_copied: NULL sketcher");
    if (ticket == 0 || ticket >= s->next_ticket) return fail(LASH_E_INVALID, "lash_sketch_wait_copied: unknown ticket");
    Slot& sl = s->slot[ticket % kSlots];
    // a newer push on the same slot implies the older copy completed (stream order)
    if (sl.used && sl.ticket >= ticket) CU(cudaEventSynchronize(sl.copied));
    return LASH_OK;
}
extern "C" int lash_sketch_sync(lash_sketcher* s) {
    if (!s) return fail(LASH_E_INVALID, "lash_sketch_sync: NULL sketcher");
    CU(cudaSetDevice(s->ctx->device));
    if (s->ext_stream) CU(cudaStreamSynchronize(s->ext_stream));
    for (int i = 0; i < kSlots; ++i) {
        CU(cudaStreamSynchronize(s->slot[i].stream));
        int rc = harvest_timing(s, s->slot[i]);
        if (rc) return rc;
    }
    return LASH_OK;
}
extern "C" int lash_sketch_fetch(lash_sketcher* s, uint64_t first, uint64_t n, void* regs_out) {
    if (!s || (!regs_out && n)) return fail(LASH_E_INVALID, "lash_sketch_fetch: NULL argument");
    if (first + n > s->n_genomes) return fail(LASH_E_INVALID, "lash_sketch_fetch: genome range out of bounds");
    int rc = lash_sketch_sync(s);
    if (rc) return rc;
    if (n) CU(cudaMemcpy(regs_out, (const char*)s->acc + first * s->reg_bytes, n * s->reg_bytes, cudaMemcpyDeviceToHost));
    return LASH_OK;
}
extern "C" int lash_sketch_regs_dev(lash_sketcher* s, void** regs_dev) {
    if (!s || !regs_dev) return fail(LASH_E_INVALID, "lash_sketch_regs_dev: NULL argument");
    *regs_dev = s->acc;
    return LASH_OK;
}
extern "C" int lash_sketch_set_stream(lash_sketcher* s, void* stream) {
    if (!s) return fail(LASH_E_INVALID, "lash_sketch_set_stream: NULL sketcher");
    int rc = lash_sketch_sync(s);
    if (rc) return rc;
    s->ext_stream = (cudaStream_t)stream;
    return LASH_OK;
}
extern "C" int lash_sketch_reset(lash_sketcher* s) {
    if (!s) return fail(LASH_E_INVALID, "lash_sketch_reset: NULL sketcher");
    if (s->ext_stream) {
        CU(cudaSetDevice(s->ctx->device));
        CU(cudaMemsetAsync(s->acc, 0, (size_t)s->sp.cell_words * 4 * s->n_genomes, s->ext_stream));
        return LASH_OK;
    }
    int rc = lash_sketch_sync(s);
    if (rc) return rc;
    CU(cudaMemset(s->acc, 0, (size_t)s->sp.cell_words * 4 * s->n_genomes));
    return LASH_OK;
}
extern "C" int lash_sketch_stats(lash_sketcher* s, double* kernel_ms, uint64_t* launches) {
    if (!s) return fail(LASH_E_INVALID, "lash_sketch_stats: NULL sketcher");
    int rc = lash_sketch_sync(s);
    if (rc) return rc;
    if (kernel_ms) *kernel_ms = s->kernel_ms;
    if (launches) *launches = s->launches;
    return LASH_OK;
}
extern "C" int lash_sketch_close(lash_sketcher* s) {
    if (!s) return LASH_OK;
    cudaSetDevice(s->ctx->device);
    for (int i = 0; i < kSlots; ++i) {
        Slot& sl = s->slot[i];
        if (sl.stream) cudaStreamSynchronize(sl.stream);
        if (s->ext_stream) cudaStreamSynchronize(s->ext_stream);
        {
            std::lock_guard<std::mutex> g(s->ctx->slot_mu);
            if (s->ctx->slot_cache.size() < kSlotCacheMax) {
                s->ctx->slot_cache.push_back(static_cast<SlotBufs&>(sl));
                static_cast<SlotBufs&>(sl) = SlotBufs();
            }
        }
        sl.release();
        if (sl.copied) cudaEventDestroy(sl.copied);
        if (sl.meta_ready) cudaEventDestroy(sl.meta_ready);
        if (sl.k_start) cudaEventDestroy(sl.k_start);
        if (sl.k_stop) cudaEventDestroy(sl.k_stop);
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    if (s->meta_stream) {
        cudaStreamSynchronize(s->meta_stream);
        cudaStreamDestroy(s->meta_stream);
    }
    if (s->acc) cudaFree(s->acc);
    delete s;
    return LASH_OK;
}

extern "C" int lash_sketch_merge_dev(lash_ctx* ctx, int algo, int p, void* dst_dev, const void* src_dev, uint64_t n_sketches,
                                     void* stream) {
    if (!ctx) return fail(LASH_E_INVALID, "lash_sketch_merge_dev: NULL ctx");
    if (!valid_algo_p(algo, p)) return fail(LASH_E_INVALID, "lash_sketch_merge_dev: bad algorithm / precision");
    if (n_sketches == 0) return LASH_OK;
    if (!dst_dev || !src_dev) return fail(LASH_E_INVALID, "lash_sketch_merge_dev: NULL device pointer");
    const size_t rb = lash_sketch_reg_bytes(algo, p);
    if ((rb * n_sketches) % 4 || ((uintptr_t)dst_dev | (uintptr_t)src_dev) % 4)
        return fail(LASH_E_INVALID, "lash_sketch_merge_dev: register arrays must be 4-byte aligned and a multiple of 4 bytes");
    CU(cudaSetDevice(ctx->device));
    CU(launch_merge(algo, (uint32_t*)dst_dev, (const uint32_t*)src_dev, rb * n_sketches / 4, ctx->n_sm,
                    stream ? (cudaStream_t)stream : ctx->stream));
    return LASH_OK;
}
extern "C" int lash_sketch_merge(lash_ctx* ctx, int algo, int p, void* dst_regs, const void* src_regs, uint64_t n_sketches) {
    if (!ctx) return fail(LASH_E_INVALID, "lash_sketch_merge: NULL ctx");
    if (!valid_algo_p(algo, p)) return fail(LASH_E_INVALID, "lash_sketch_merge: bad algorithm / precision");
    if (n_sketches == 0) return LASH_OK;
    if (!dst_regs || !src_regs) return fail(LASH_E_INVALID, "lash_sketch_merge: NULL argument");
    const size_t bytes = lash_sketch_reg_bytes(algo, p) * n_sketches;
    CU(cudaSetDevice(ctx->device));
    CU(ctx->d_ref.reserve((bytes + 3) / 4 * 4));
    CU(ctx->d_qry.reserve((bytes + 3) / 4 * 4));
    CU(cudaMemsetAsync(ctx->d_ref.p, 0, (bytes + 3) / 4 * 4, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_qry.p, 0, (bytes + 3) / 4 * 4, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_ref.p, dst_regs, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_qry.p, src_regs, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_merge(algo, (uint32_t*)ctx->d_ref.p, (const uint32_t*)ctx->d_qry.p, (bytes + 3) / 4, ctx->n_sm, ctx->stream));
    CU(cudaMemcpyAsync(dst_regs, ctx->d_ref.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return LASH_OK;
}

// ------------------------------------------------------------------------------------------------
// distance
// ------------------------------------------------------------------------------------------------
static int check_dist_args(int algo, int p, int k, int estimator, int model, uint64_t n_ref, uint64_t n_qry, int triangular) {
    if (!valid_algo_p(algo, p)) return fail(LASH_E_INVALID, "lash_dist: bad algorithm / precision");
    if (k < 1 || k > 32) return fail(LASH_E_INVALID, "k-mer length must be 1-32");
    if (model != 0 && model != 1 && model != LASH_MODEL_FRAC) return fail(LASH_E_INVALID, "model needs to be 0 or 1");  // main.rs:421
    if (algo == LASH_ALGO_ULL && estimator != 0 && estimator != 1)
        return fail(LASH_E_INVALID, "estimator needs to be either fgra or ml");  // utils.rs:217
    if (triangular && n_ref != n_qry) return fail(LASH_E_INVALID, "lash_dist: triangular needs the same set on both sides");
    return LASH_OK;
}

// smallest non-empty register of both sets, for the FGRA pair-table kernel (dist_kernels.cu)
static int prepare_regmin(lash_ctx* ctx, DistParams& dp, size_t rb, cudaStream_t st) {
    dp.n_sm = ctx->n_sm;
    dp.reg_mm_ref = dp.reg_mm_qry = nullptr;
    static const bool ml_table = [] { const char* v = getenv("LASH_ML_S"); return v && std::string(v) == "table"; }();   // A/B
    if (dp.algo == LASH_ALGO_HLL || (dp.algo == LASH_ALGO_ULL && dp.estimator == LASH_EST_ML && !ml_table)) {
        // K4i's windows / K4c's G-sum tiles: smallest and largest register of every sketch (16-byte loads: aligned arrays of
        // sketches of >= 16 registers)
        if ((((uintptr_t)dp.ref | (uintptr_t)dp.qry) & 15u) == 0 && rb % 16 == 0) {
            const bool same = dp.qry == dp.ref && dp.n_qry == dp.n_ref;
            if (ctx->d_hll_mm.reserve(4 * (dp.n_ref + (same ? 0 : dp.n_qry))) == cudaSuccess) {
                uint32_t* mm = (uint32_t*)ctx->d_hll_mm.p;
                CU(launch_hll_minmax(dp.ref, dp.n_ref, (uint32_t)rb, mm, st));
                if (!same) CU(launch_hll_minmax(dp.qry, dp.n_qry, (uint32_t)rb, mm + dp.n_ref, st));
                ctx->dist_launches += same ? 1 : 2;
                dp.reg_mm_ref = mm;
                dp.reg_mm_qry = same ? mm : mm + dp.n_ref;
            } else {
                cudaGetLastError();
            }
        }
        if (dp.algo == LASH_ALGO_HLL) return LASH_OK;
    }
    if (dp.algo != LASH_ALGO_ULL) return LASH_OK;
    CU(ctx->d_regmin.reserve(8));
    uint32_t* w = (uint32_t*)ctx->d_regmin.p;
    dp.tile_counter = w + 1;  // word 1: tile counter of the persistent pair-table kernels (zeroed per launch)
    if (dp.estimator != LASH_EST_FGRA) return LASH_OK;
    CU(cudaMemsetAsync(w, 0xff, 4, st));
    CU(launch_regmin(dp.ref, rb * dp.n_ref, w, ctx->n_sm, st));
    if (dp.qry != dp.ref) CU(launch_regmin(dp.qry, rb * dp.n_qry, w, ctx->n_sm, st));
    ctx->dist_launches += dp.qry != dp.ref ? 2 : 1;
    dp.regmin = w;
    return LASH_OK;
}

// ULL ML: scratch for the two-kernel form (tile kernel stores the pair statistics, ml_finish_kernel solves).  Called
// right before launch_dist, once dp's row range and output addressing are final.  Falls back to the fused kernel
// (ml_scratch = nullptr) when the scratch would be larger than LASH_ML_SCRATCH_MAX_MB (default 16384) or cannot be
// allocated, or when LASH_ML_KERNEL=fused asks for it (A/B measurements).
static void setup_ml_scratch(lash_ctx* ctx, DistParams& dp) {
    dp.ml_scratch = nullptr;
    if (dp.algo != LASH_ALGO_ULL || dp.estimator != LASH_EST_ML || dp.row_end <= dp.row_begin) return;
    static const bool fused = [] { const char* v = getenv("LASH_ML_KERNEL"); return v && std::string(v) == "fused"; }();
    static const uint64_t max_bytes = [] {
        const char* v = getenv("LASH_ML_SCRATCH_MAX_MB");
        return (uint64_t)(v ? strtoull(v, nullptr, 10) : 16384ull) << 20;
    }();
    if (fused) return;
    uint64_t o_base, o_end;
    if (dp.packed_tri) {
        o_base = dp.row_begin * (dp.row_begin + 1) / 2;
        o_end = dp.row_end * (dp.row_end + 1) / 2;
    } else {
        o_base = (dp.row_begin - dp.out_row0) * dp.n_qry;
        o_end = (dp.row_end - dp.out_row0) * dp.n_qry;
    }
    const uint64_t cells = o_end - o_base;
    const uint64_t bytes = cells * 4ull * ml_scratch_words(dp.p);
    if (cells == 0 || bytes > max_bytes) return;
    if (ctx->d_ml.reserve(bytes) != cudaSuccess) {
        cudaGetLastError();  // clear the sticky allocation error: the fused kernel needs no scratch
        return;
    }
    dp.ml_scratch = (uint32_t*)ctx->d_ml.p;
    dp.ml_cells = cells;
    dp.ml_o_base = o_base;
}

// HMH: expected-collision sums of all pairs of SMALL sketches (cardinality <= 2^19), before the tile kernel (dist_kernels.cu,
// "K4m, small sketches").  Counts the small sketches of the row range and of the query set (one 8-byte D2H + stream sync: the
// only place where the device-resident API synchronises, and only for HMH), sizes the scratch (41 x 1024 doubles per small
// sketch, one double per small pair; LASH_HMH_EC_MAX_MB caps it, default 24 GiB -- sketches beyond the cap keep the per-pair
// loop), and enqueues slots -> term vectors -> tile product on `st`.  LASH_HMH_EC=loop disables the path (A/B measurements).
static int setup_hmh_ec(lash_ctx* ctx, DistParams& dp, cudaStream_t st) {
    dp.hmh_ec = nullptr;
    if (dp.algo != LASH_ALGO_HMH || dp.row_end <= dp.row_begin || dp.n_qry == 0) return LASH_OK;
    static const bool off = [] { const char* v = getenv("LASH_HMH_EC"); return v && std::string(v) == "loop"; }();
    if (off) return LASH_OK;
    static const uint64_t max_bytes = [] {
        const char* v = getenv("LASH_HMH_EC_MAX_MB");
        return (uint64_t)(v ? strtoull(v, nullptr, 10) : 24576ull) << 20;
    }();
    const uint64_t n_rows = dp.row_end - dp.row_begin;
    // index scratch: [0] small rows, [1] small queries (counts), then slot_r[n_rows], slot_q[n_qry], src_r[n_rows], src_q[n_qry]
    const size_t idx_bytes = 16 + 4 * (2 * n_rows + 2 * dp.n_qry);
    if (ctx->d_hmh_idx.reserve(idx_bytes) != cudaSuccess) { cudaGetLastError(); return LASH_OK; }
    uint32_t* counts = (uint32_t*)ctx->d_hmh_idx.p;
    int32_t* slot_r = (int32_t*)(counts + 4);
    int32_t* slot_q = slot_r + n_rows;
    uint32_t* src_r = (uint32_t*)(slot_q + dp.n_qry);
    uint32_t* src_q = src_r + n_rows;
    CU(cudaMemsetAsync(counts, 0, 16, st));
    CU(launch_hmh_count_small(dp.card_ref, dp.row_begin, dp.row_end, counts, st));
    CU(launch_hmh_count_small(dp.card_qry, 0, dp.n_qry, counts + 1, st));
    uint32_t h[2] = {0, 0};
    CU(cudaMemcpyAsync(h, counts, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ctx->dist_launches += 2;
    if (h[0] == 0 || h[1] == 0) return LASH_OK;   // no pair of two small sketches: nothing runs the loop
    // caps: both below the grid limit, scratch within the budget (term vectors dominate; shrink the larger side first)
    const uint64_t per = (uint64_t)kHmhEcTermsPerSketch * 8;
    uint64_t cap_r = std::min<uint64_t>(h[0], 65535), cap_q = std::min<uint64_t>(h[1], 65535);
    while ((cap_r + cap_q) * per + cap_r * cap_q * 8 > max_bytes && (cap_r > 64 || cap_q > 64)) {
        if (cap_r >= cap_q) cap_r = std::max<uint64_t>(64, cap_r * 3 / 4);
        else cap_q = std::max<uint64_t>(64, cap_q * 3 / 4);
    }
    if (ctx->d_hmh_terms_r.reserve(cap_r * per) != cudaSuccess || ctx->d_hmh_terms_q.reserve(cap_q * per) != cudaSuccess ||
        ctx->d_hmh_ec.reserve(cap_r * cap_q * 8) != cudaSuccess) {
        cudaGetLastError();   // no scratch: the per-pair loop computes the same numbers
        return LASH_OK;
    }
    CU(launch_hmh_slots(dp.card_ref, dp.row_begin, dp.row_end, (uint32_t)cap_r, slot_r, src_r, counts + 2, st));
    CU(launch_hmh_slots(dp.card_qry, 0, dp.n_qry, (uint32_t)cap_q, slot_q, src_q, counts + 3, st));
    CU(launch_hmh_ec_fill(dp.card_ref, src_r, counts + 2, (uint32_t)cap_r, (double*)ctx->d_hmh_terms_r.p, st));
    CU(launch_hmh_ec_fill(dp.card_qry, src_q, counts + 3, (uint32_t)cap_q, (double*)ctx->d_hmh_terms_q.p, st));
    CU(launch_hmh_ec_gemm((const double*)ctx->d_hmh_terms_r.p, src_r, counts + 2, (uint32_t)cap_r, (const double*)ctx->d_hmh_terms_q.p, src_q,
                          counts + 3, (uint32_t)cap_q, dp.triangular, (double*)ctx->d_hmh_ec.p, (uint32_t)cap_q, st));
    ctx->dist_launches += 5;
    dp.hmh_slot_ref = slot_r;
    dp.hmh_slot_qry = slot_q;
    dp.hmh_ec = (const double*)ctx->d_hmh_ec.p;
    dp.hmh_row0 = dp.row_begin;
    dp.hmh_ec_ld = (uint32_t)cap_q;
    return LASH_OK;
}

extern "C" int lash_cardinality_dev(lash_ctx* ctx, int algo, int p, int estimator, const void* regs_dev, uint64_t n,
                                    double* card_dev, void* stream) {
    if (!ctx) return fail(LASH_E_INVALID, "lash_cardinality_dev: NULL ctx");
    if (!valid_algo_p(algo, p)) return fail(LASH_E_INVALID, "lash_cardinality_dev: bad algorithm / precision");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    CU(launch_cardinality(algo, algo == LASH_ALGO_HMH ? 14 : p, estimator, regs_dev, n, card_dev, nullptr, st));
    ctx->dist_launches += 1;
    return LASH_OK;
}

extern "C" int lash_cardinality(lash_ctx* ctx, int algo, int p, int estimator, const void* regs, uint64_t n, double* card_out) {
    if (!ctx || (n && (!regs || !card_out))) return fail(LASH_E_INVALID, "lash_cardinality: NULL argument");
    if (!valid_algo_p(algo, p)) return fail(LASH_E_INVALID, "lash_cardinality: bad algorithm / precision");
    if (n == 0) return LASH_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t rb = lash_sketch_reg_bytes(algo, p);
    DevBuf d, c;
    CU(d.reserve(rb * n));
    CU(c.reserve(8 * n));
    CU(cudaMemcpyAsync(d.p, regs, rb * n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = lash_cardinality_dev(ctx, algo, p, estimator, d.p, n, (double*)c.p, ctx->stream);
    if (rc == LASH_OK) {
        cudaError_t e = cudaMemcpyAsync(card_out, c.p, 8 * n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(LASH_E_CUDA, cudaGetErrorString(e));
    }
    d.release();
    c.release();
    return rc;
}

extern "C" int lash_dist_dev(lash_ctx* ctx, int algo, int p, int k, int estimator, int model, int fp32, const void* ref_dev,
                             uint64_t n_ref, const void* qry_dev, uint64_t n_qry, const double* card_ref_dev,
                             const double* card_qry_dev, int triangular, uint64_t row_begin, uint64_t row_end, void* out_dev,
                             uint32_t* flags_dev, void* stream) {
    if (!ctx) return fail(LASH_E_INVALID, "lash_dist_dev: NULL ctx");
    int rc = check_dist_args(algo, p, k, estimator, model, n_ref, n_qry, triangular);
    if (rc) return rc;
    if (row_end > n_ref || row_begin > row_end) return fail(LASH_E_INVALID, "lash_dist_dev: bad row range");
    if (row_begin == row_end || n_qry == 0) return LASH_OK;
    if (!ref_dev || !qry_dev || !out_dev || !card_ref_dev || !card_qry_dev)
        return fail(LASH_E_INVALID, "lash_dist_dev: NULL device pointer");
    CU(cudaSetDevice(ctx->device));
    DistParams dp;
    dp.algo = algo;
    dp.p = algo == LASH_ALGO_HMH ? 14 : p;
    dp.k = k;
    dp.estimator = estimator;
    dp.model = model;
    dp.fp32 = fp32 ? 1 : 0;
    dp.triangular = triangular ? 1 : 0;
    dp.ref = ref_dev;
    dp.qry = qry_dev;
    dp.n_ref = n_ref;
    dp.n_qry = n_qry;
    dp.card_ref = card_ref_dev;
    dp.card_qry = card_qry_dev;
    dp.row_begin = row_begin;
    dp.row_end = row_end;
    dp.out = out_dev;
    dp.packed_tri = triangular ? 1 : 0;
    dp.out_row0 = 0;
    dp.flags = flags_dev;
    rc = prepare_regmin(ctx, dp, lash_sketch_reg_bytes(algo, p), stream ? (cudaStream_t)stream : ctx->stream);
    if (rc) return rc;
    setup_ml_scratch(ctx, dp);
    rc = setup_hmh_ec(ctx, dp, stream ? (cudaStream_t)stream : ctx->stream);
    if (rc) return rc;
    uint32_t nl = 0;
    CU(launch_dist(dp, stream ? (cudaStream_t)stream : ctx->stream, &nl));
    ctx->dist_launches += nl;
    return LASH_OK;
}

// shared body of lash_dist / lash_dist_stream: upload, cardinalities, row blocks
static int dist_host(lash_ctx* ctx, int algo, int p, int k, int estimator, int model, int fp32, const void* ref_regs,
                     uint64_t n_ref, const void* qry_regs, uint64_t n_qry, int triangular, void* out, uint64_t rows_per_block,
                     lash_dist_block_cb cb, void* user, uint64_t row_begin = 0, uint64_t row_end = ~0ull) {
    if (!ctx) return fail(LASH_E_INVALID, "lash_dist: NULL ctx");
    int rc = check_dist_args(algo, p, k, estimator, model, n_ref, n_qry, triangular);
    if (rc) return rc;
    ctx->dist_ms = 0.0;
    ctx->dist_launches = 0;
    if (n_ref == 0 || n_qry == 0) return LASH_OK;
    if (!ref_regs || !qry_regs) return fail(LASH_E_INVALID, "lash_dist: NULL register array");
    if (!out && !cb) return fail(LASH_E_INVALID, "lash_dist: no output");
    CU(cudaSetDevice(ctx->device));
    const int pp = algo == LASH_ALGO_HMH ? 14 : p;
    const size_t rb = lash_sketch_reg_bytes(algo, p);
    const size_t esz = fp32 ? 4 : 8;
    const bool same = (ref_regs == qry_regs && n_ref == n_qry);
    cudaStream_t st = ctx->stream;
    DevBuf &d_ref = ctx->d_ref, &d_qry = ctx->d_qry, &d_card = ctx->d_card, &d_flags = ctx->d_flags;
    DevBuf* d_out = ctx->d_out;
    PinBuf* h_out = ctx->h_out;
    // Every per-call event and stream lives in this guard: whichever way the function leaves, both streams are drained
    // first (kernels and D2H copies may still target ctx->d_out / h_out, which the next call may reserve() = free) and
    // then everything is destroyed.
    struct Guard {
        cudaStream_t st, st2 = nullptr;
        std::vector<cudaEvent_t> evs;
        explicit Guard(cudaStream_t s) : st(s) {}
        cudaError_t event(cudaEvent_t* e, unsigned flags = cudaEventDefault) {
            cudaError_t rc = cudaEventCreateWithFlags(e, flags);
            if (rc == cudaSuccess) evs.push_back(*e);
            return rc;
        }
        ~Guard() {
            if (st2) cudaStreamSynchronize(st2);
            cudaStreamSynchronize(st);
            for (cudaEvent_t e : evs) cudaEventDestroy(e);
            if (st2) cudaStreamDestroy(st2);
        }
    } guard(st);
#define CUC(call)                                                                     \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            return fail(e__ == cudaErrorMemoryAllocation ? LASH_E_NOMEM : LASH_E_CUDA, \
                        std::string(#call) + ": " + cudaGetErrorString(e__));         \
        }                                                                             \
    } while (0)
    cudaEvent_t ev0, ev1;
    CUC(d_ref.reserve(rb * n_ref));
    if (!same) CUC(d_qry.reserve(rb * n_qry));
    CUC(d_card.reserve(8 * (n_ref + n_qry)));
    CUC(d_flags.reserve(4));
    CUC(guard.event(&ev0));
    CUC(guard.event(&ev1));
    CUC(cudaMemcpyAsync(d_ref.p, ref_regs, rb * n_ref, cudaMemcpyHostToDevice, st));
    if (!same) CUC(cudaMemcpyAsync(d_qry.p, qry_regs, rb * n_qry, cudaMemcpyHostToDevice, st));
    CUC(cudaMemsetAsync(d_flags.p, 0, 4, st));
    unsigned long long* d_sum = nullptr;
    ctx->checksum = ctx->checksum_cells = 0;
    if (ctx->want_checksum) {
        CUC(ctx->d_sum.reserve(16));
        d_sum = (unsigned long long*)ctx->d_sum.p;
        CUC(cudaMemsetAsync(d_sum, 0, 16, st));
    }
    const void* qdev = same ? d_ref.p : d_qry.p;
    double* card_r = (double*)d_card.p;
    double* card_q = same ? card_r : card_r + n_ref;
    float ms_total = 0.f;
    CUC(cudaEventRecord(ev0, st));
    CUC(launch_cardinality(algo, pp, estimator, d_ref.p, n_ref, card_r, (uint32_t*)d_flags.p, st));
    ctx->dist_launches += 1;
    if (!same) {
        CUC(launch_cardinality(algo, pp, estimator, d_qry.p, n_qry, card_q, (uint32_t*)d_flags.p, st));
        ctx->dist_launches += 1;
    }
    CUC(cudaEventRecord(ev1, st));

    DistParams dp;
    dp.algo = algo; dp.p = pp; dp.k = k; dp.estimator = estimator; dp.model = model; dp.fp32 = fp32 ? 1 : 0;
    dp.triangular = triangular ? 1 : 0;
    dp.ref = d_ref.p; dp.qry = qdev; dp.n_ref = n_ref; dp.n_qry = n_qry;
    dp.card_ref = card_r; dp.card_qry = card_q; dp.flags = (uint32_t*)d_flags.p;
    rc = prepare_regmin(ctx, dp, rb, st);
    if (rc) return rc;

    if (!cb) {
        // whole result on device, one D2H
        const uint64_t cells = triangular ? n_ref * (n_ref + 1) / 2 : n_ref * n_qry;
        CUC(d_out[0].reserve(cells * esz));
        dp.row_begin = 0; dp.row_end = n_ref; dp.out = d_out[0].p; dp.packed_tri = triangular ? 1 : 0; dp.out_row0 = 0;
        cudaEvent_t k0, k1;
        CUC(guard.event(&k0));
        CUC(guard.event(&k1));
        CUC(cudaEventRecord(k0, st));
        setup_ml_scratch(ctx, dp);
        rc = setup_hmh_ec(ctx, dp, st);
        if (rc) return rc;
        uint32_t nl = 0;
        CUC(launch_dist(dp, st, &nl));
        ctx->dist_launches += nl;
        CUC(cudaEventRecord(k1, st));
        if (d_sum) {
            CUC(launch_out_checksum(dp, d_sum, st));
            ctx->dist_launches += 1;
        }
        CUC(cudaMemcpyAsync(out, d_out[0].p, cells * esz, cudaMemcpyDeviceToHost, st));
        CUC(cudaStreamSynchronize(st));
        float ms = 0.f;
        CUC(cudaEventElapsedTime(&ms, k0, k1));
        ms_total += ms;
    } else {
        if (rows_per_block == 0) rows_per_block = std::max<uint64_t>(1, ((uint64_t)256 << 20) / (n_qry * esz));
        rows_per_block = std::min(rows_per_block, n_ref);
        const size_t blk_bytes = rows_per_block * n_qry * esz;
        CUC(cudaStreamCreateWithFlags(&guard.st2, cudaStreamNonBlocking));
        const cudaStream_t st2 = guard.st2;
        cudaEvent_t kstart, kdone[2], csum[2], copied[2];
        for (int i = 0; i < 2; ++i) {
            CUC(d_out[i].reserve(blk_bytes));
            CUC(h_out[i].reserve(blk_bytes));
            CUC(guard.event(&kdone[i]));
            CUC(guard.event(&csum[i], cudaEventDisableTiming));
            CUC(guard.event(&copied[i], cudaEventDisableTiming));
        }
        CUC(guard.event(&kstart));
        // kernels on st, copies on st2; callback for block b-1 runs on the host while block b computes
        struct Pending { uint64_t row0, nrows; int buf; bool valid; } pend = {0, 0, 0, false};
        int buf = 0;
        int cb_rc = 0;
        uint64_t nblocks = 0;
        if (row_end > n_ref) row_end = n_ref;
        for (uint64_t r0 = row_begin; r0 < row_end && cb_rc == 0; r0 += rows_per_block, buf ^= 1, ++nblocks) {
            const uint64_t r1 = std::min(row_end, r0 + rows_per_block);
            if (nblocks >= 2) CUC(cudaEventSynchronize(copied[buf]));  // device buffer free again (its D2H finished)
            dp.row_begin = r0; dp.row_end = r1; dp.out = d_out[buf].p; dp.packed_tri = 0; dp.out_row0 = r0;
            CUC(cudaEventRecord(kstart, st));
            setup_ml_scratch(ctx, dp);
            rc = setup_hmh_ec(ctx, dp, st);
            if (rc) return rc;
            uint32_t nl = 0;
            CUC(launch_dist(dp, st, &nl));
            ctx->dist_launches += nl;
            CUC(cudaEventRecord(kdone[buf], st));
            if (d_sum) {
                CUC(launch_out_checksum(dp, d_sum, st));
                ctx->dist_launches += 1;
                CUC(cudaEventRecord(csum[buf], st));
            }
            // deliver the previous block while this one computes
            if (pend.valid) {
                CUC(cudaEventSynchronize(copied[pend.buf]));
                cb_rc = cb(user, pend.row0, pend.nrows, h_out[pend.buf].p);
                pend.valid = false;
            }
            CUC(cudaStreamWaitEvent(st2, kdone[buf], 0));
            // triangular: nothing right of column r1-1 is defined in this block -> copy only the columns below the diagonal
            // (same pitch on both sides, so the block keeps its dense [n_rows][n_qry] addressing)
            const uint64_t ncols = triangular ? std::min<uint64_t>(n_qry, r1) : n_qry;
            CUC(cudaMemcpy2DAsync(h_out[buf].p, n_qry * esz, d_out[buf].p, n_qry * esz, ncols * esz, r1 - r0, cudaMemcpyDeviceToHost, st2));
            // the next kernel into this buffer (two blocks on) is ordered after `copied` by the host wait above; the checksum
            // kernel only reads, so it may overlap the copy
            CUC(cudaEventRecord(copied[buf], st2));
            CUC(cudaEventSynchronize(kdone[buf]));
            float ms = 0.f;
            CUC(cudaEventElapsedTime(&ms, kstart, kdone[buf]));
            ms_total += ms;
            pend = {r0, r1 - r0, buf, true};
        }
        if (pend.valid && cb_rc == 0) {
            CUC(cudaEventSynchronize(copied[pend.buf]));
            cb_rc = cb(user, pend.row0, pend.nrows, h_out[pend.buf].p);
        }
        CUC(cudaStreamSynchronize(st2));
        CUC(cudaStreamSynchronize(st));
        if (cb_rc != 0) return fail(LASH_E_STATE, "lash_dist_stream: callback returned non-zero");
    }
    float ms_card = 0.f;
    CUC(cudaEventElapsedTime(&ms_card, ev0, ev1));
    ctx->dist_ms = ms_total + ms_card;
    uint32_t flags = 0;
    CUC(cudaMemcpy(&flags, d_flags.p, 4, cudaMemcpyDeviceToHost));
    if (d_sum) {
        unsigned long long hs[2] = {0, 0};
        CUC(cudaMemcpy(hs, d_sum, 16, cudaMemcpyDeviceToHost));
        ctx->checksum = hs[0];
        ctx->checksum_cells = hs[1];
    }
#undef CUC
    return flags ? LASH_W_HLL_BIAS_REGIME : LASH_OK;
}

extern "C" int lash_dist(lash_ctx* ctx, int algo, int p, int k, int estimator, int model, int fp32, const void* ref_regs,
                         uint64_t n_ref, const void* qry_regs, uint64_t n_qry, int triangular, void* out) {
    if (!out && n_ref && n_qry) return fail(LASH_E_INVALID, "lash_dist: out is NULL");
    return dist_host(ctx, algo, p, k, estimator, model, fp32, ref_regs, n_ref, qry_regs, n_qry, triangular, out, 0, nullptr, nullptr);
}
extern "C" int lash_dist_stream(lash_ctx* ctx, int algo, int p, int k, int estimator, int model, int fp32, const void* ref_regs,
                                uint64_t n_ref, const void* qry_regs, uint64_t n_qry, int triangular, uint64_t rows_per_block,
                                lash_dist_block_cb cb, void* user) {
    if (!cb) return fail(LASH_E_INVALID, "lash_dist_stream: callback is NULL");
    return dist_host(ctx, algo, p, k, estimator, model, fp32, ref_regs, n_ref, qry_regs, n_qry, triangular, nullptr, rows_per_block, cb, user);
}
extern "C" int lash_dist_stream_rows(lash_ctx* ctx, int algo, int p, int k, int estimator, int model, int fp32, const void* ref_regs,
                                     uint64_t n_ref, const void* qry_regs, uint64_t n_qry, int triangular, uint64_t row_begin,
                                     uint64_t row_end, uint64_t rows_per_block, lash_dist_block_cb cb, void* user) {
    if (!cb) return fail(LASH_E_INVALID, "lash_dist_stream_rows: callback is NULL");
    if (row_begin > row_end || row_end > n_ref) return fail(LASH_E_INVALID, "lash_dist_stream_rows: bad row range");
    if (row_begin == row_end) return LASH_OK;
    return dist_host(ctx, algo, p, k, estimator, model, fp32, ref_regs, n_ref, qry_regs, n_qry, triangular, nullptr, rows_per_block, cb,
                     user, row_begin, row_end);
}
extern "C" int lash_dist_set_checksum(lash_ctx* ctx, int enable) {
    if (!ctx) return fail(LASH_E_INVALID, "lash_dist_set_checksum: NULL ctx");
    ctx->want_checksum = enable != 0;
    return LASH_OK;
}
extern "C" int lash_dist_checksum(lash_ctx* ctx, uint64_t* sum_bits, uint64_t* n_cells) {
    if (!ctx) return fail(LASH_E_INVALID, "lash_dist_checksum: NULL ctx");
    if (sum_bits) *sum_bits = ctx->checksum;
    if (n_cells) *n_cells = ctx->checksum_cells;
    return LASH_OK;
}
extern "C" int lash_dist_checksum_dev(lash_ctx* ctx, int fp32, const void* out_dev, uint64_t n_qry, int triangular,
                                      uint64_t row_begin, uint64_t row_end, uint64_t* sums_dev, void* stream) {
    if (!ctx || !out_dev || !sums_dev) return fail(LASH_E_INVALID, "lash_dist_checksum_dev: NULL argument");
    if (row_begin > row_end) return fail(LASH_E_INVALID, "lash_dist_checksum_dev: bad row range");
    CU(cudaSetDevice(ctx->device));
    DistParams dp;
    dp.algo = 0; dp.p = 0; dp.k = 0; dp.estimator = 0; dp.model = 0;
    dp.fp32 = fp32 ? 1 : 0;
    dp.triangular = triangular ? 1 : 0;
    dp.ref = dp.qry = nullptr; dp.n_ref = row_end; dp.n_qry = n_qry; dp.card_ref = dp.card_qry = nullptr;
    dp.row_begin = row_begin; dp.row_end = row_end; dp.out = const_cast<void*>(out_dev);
    dp.packed_tri = triangular ? 1 : 0; dp.out_row0 = 0; dp.flags = nullptr; dp.n_sm = ctx->n_sm;
    CU(launch_out_checksum(dp, (unsigned long long*)sums_dev, stream ? (cudaStream_t)stream : ctx->stream));
    ctx->dist_launches += 1;
    return LASH_OK;
}
extern "C" int lash_dist_stats(lash_ctx* ctx, double* kernel_ms, uint64_t* launches) {
    if (!ctx) return fail(LASH_E_INVALID, "lash_dist_stats: NULL ctx");
    if (kernel_ms) *kernel_ms = ctx->dist_ms;
    if (launches) *launches = ctx->dist_launches;
    return LASH_OK;
}
