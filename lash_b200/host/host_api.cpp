// extern "C" surface of the host layer (include/lash_host.h): thin wrappers over lash_host.hpp.
#include <cstring>
#include <map>
#include <new>

#include "fastx.hpp"
#include "lash_host.hpp"
#include "pack.hpp"
#include "sketch_io.hpp"

static thread_local std::string g_host_err;
static int fail(int code, const std::string& m) {
    g_host_err = m;
    return code;
}
static int from_status(const lash::Status& st) {
    if (st.code != 0) g_host_err = st.message;
    return st.code;
}

extern "C" const char* lash_host_last_error(void) { return g_host_err.c_str(); }

struct lash_fastx {
    lashhost::FastxReader rd;
    std::string id, seq;
};
extern "C" int lash_fastx_open(const char* path, lash_fastx** out) {
    if (!path || !out) return fail(LASH_E_INVALID, "lash_fastx_open: NULL argument");
    lash_fastx* r = new (std::nothrow) lash_fastx();
    if (!r) return fail(LASH_E_NOMEM, "out of host memory");
    if (!r->rd.open(path)) {
        const std::string e = r->rd.err();
        delete r;
        return fail(LASH_HOST_E_IO, "Invalid input file: " + e);
    }
    *out = r;
    return 0;
}
extern "C" int lash_fastx_next(lash_fastx* r, const char** id, size_t* id_len, const char** seq, size_t* seq_len) {
    if (!r) return fail(LASH_E_INVALID, "lash_fastx_next: NULL reader");
    const int rc = r->rd.next_record(r->id, r->seq);
    if (rc < 0) return fail(LASH_HOST_E_FORMAT, r->rd.err());
    if (rc == 0) return 0;
    if (id) *id = r->id.data();
    if (id_len) *id_len = r->id.size();
    if (seq) *seq = r->seq.data();
    if (seq_len) *seq_len = r->seq.size();
    return 1;
}
extern "C" int lash_fastx_close(lash_fastx* r) {
    delete r;
    return 0;
}

extern "C" int lash_host_pack_isa(void) { return lashhost::pack_has_avx512() ? 2 : lashhost::pack_has_simd() ? 1 : 0; }
extern "C" int lash_host_pack_has_simd(void) { return lashhost::pack_has_simd() ? 1 : 0; }

extern "C" int lash_host_filter_pack(const uint8_t* seq, size_t n, uint8_t* packed, uint64_t* n_bases, int use_simd) {
    if ((!seq && n) || !packed || !n_bases) return fail(LASH_E_INVALID, "lash_host_filter_pack: NULL argument");
    // Re-open the ABI-format stream as an LSB-first BaseStream: flip the bytes written so far, append, flip back.
    // (The streaming path never does this -- it keeps one BaseStream per span; this entry point exists for
    // bindings and tests.)
    const uint64_t have = *n_bases;
    std::vector<uint8_t> tmp(lashhost::BaseStream::padded_bytes(have + n) + 64);
    lashhost::BaseStream bs;
    bs.attach(tmp.data(), tmp.size());
    for (uint64_t i = 0; i < have; ++i) bs.push_base((unsigned)(packed[i >> 2] >> (6 - 2 * (i & 3))) & 3u);
    bs.append_filtered(seq, n, use_simd);
    const uint64_t total = bs.size();
    bs.finalize();
    memcpy(packed, tmp.data(), (total + 3) / 4);
    *n_bases = total;
    return 0;
}

static std::vector<std::string> to_vec(const char* const* files, uint64_t n) {
    std::vector<std::string> v;
    v.reserve(n);
    for (uint64_t i = 0; i < n; ++i) v.emplace_back(files[i] ? files[i] : "");
    return v;
}

extern "C" int lash_host_sketch_files_regs(lash_ctx* ctx, int algo, int p, int k, uint64_t seed, const char* const* files,
                                           uint64_t n_files, int threads, uint64_t chunk_bytes, void* regs_out,
                                           lash_sketch_files_stats* stats) {
    if ((!files && n_files) || (!regs_out && n_files)) return fail(LASH_E_INVALID, "lash_host_sketch_files_regs: NULL argument");
    if (k < 1) return fail(LASH_E_INVALID, "k-mer length must be 1-32");
    std::optional<uint32_t> prec;
    if (algo != LASH_ALGO_HMH) prec = (uint32_t)p;
    return from_status(lash::sketch_files_impl(ctx, algo, prec, to_vec(files, n_files), (size_t)k, nullptr,
                                               (uint32_t)std::max(threads, 0), seed, chunk_bytes, regs_out, stats));
}
extern "C" int lash_host_sketch_files(lash_ctx* ctx, int algo, int p, int k, uint64_t seed, const char* const* files,
                                      uint64_t n_files, const char* output_name, int threads, lash_sketch_files_stats* stats) {
    if ((!files && n_files) || !output_name) return fail(LASH_E_INVALID, "lash_host_sketch_files: NULL argument");
    if (k < 1) return fail(LASH_E_INVALID, "k-mer length must be 1-32");
    std::optional<uint32_t> prec;
    if (algo != LASH_ALGO_HMH) prec = (uint32_t)p;
    const std::string out = output_name;
    return from_status(lash::sketch_files_impl(ctx, algo, prec, to_vec(files, n_files), (size_t)k, &out, (uint32_t)std::max(threads, 0),
                                               seed, 0, nullptr, stats));
}

extern "C" int lash_host_pack_files_dry(const char* const* files, uint64_t n_files, int k, int threads, uint64_t chunk_bytes,
                                        lash_sketch_files_stats* stats) {
    if (!files && n_files) return fail(LASH_E_INVALID, "lash_host_pack_files_dry: NULL argument");
    if (k < 1) return fail(LASH_E_INVALID, "k-mer length must be 1-32");
    return from_status(lash::pack_files_dry(to_vec(files, n_files), (size_t)k, (uint32_t)std::max(threads, 0), chunk_bytes, stats));
}

extern "C" int lash_host_set_ingest_mode(int mode) {
    if (mode < 0 || mode > 2) return fail(LASH_E_INVALID, "lash_host_set_ingest_mode: 0 auto, 1 packed, 2 ascii");
    lash::set_ingest_mode(mode);
    return 0;
}
extern "C" int lash_host_release_pinned(void) {
    lash::release_pinned();
    return 0;
}

extern "C" int lash_host_write_parameters(const char* output_name, int algo, int p, int k, uint64_t seed) {
    if (!output_name) return fail(LASH_E_INVALID, "lash_host_write_parameters: NULL argument");
    // main.rs:249-276: every value is a string; serde_json's Map is a BTreeMap, i.e. keys come out sorted
    std::map<std::string, std::string> m;
    m["k"] = std::to_string(k);
    m["algorithm"] = algo == LASH_ALGO_HMH ? "hmh" : algo == LASH_ALGO_HLL ? "hll" : "ull";
    if (algo != LASH_ALGO_HMH) m["precision"] = std::to_string(p);
    m["seed"] = std::to_string(seed);
    m["molecule"] = "nucleotide";
    std::string err;
    if (!lashhost::write_file(std::string(output_name) + "_parameters.json", lashhost::json_pretty_string_map(m), err))
        return fail(LASH_HOST_E_IO, err);
    return 0;
}

extern "C" int lash_host_write_sketches(const char* path, int algo, int p, const void* regs, uint64_t n, int threads) {
    if (!path || (!regs && n)) return fail(LASH_E_INVALID, "lash_host_write_sketches: NULL argument");
    if (lash_sketch_reg_bytes(algo, p) == 0) return fail(LASH_E_INVALID, "lash_host_write_sketches: bad algorithm / precision");
    std::string err;
    if (!lashhost::write_sketches(path, algo, p, regs, n, threads, err)) return fail(LASH_HOST_E_IO, err);
    return 0;
}
extern "C" int lash_host_read_sketches(const char* path, int algo, int* p_inout, uint64_t n, void* regs_out) {
    if (!path || (!regs_out && n)) return fail(LASH_E_INVALID, "lash_host_read_sketches: NULL argument");
    std::vector<uint8_t> regs;
    std::string err;
    int p = p_inout ? *p_inout : 0;
    if (algo == LASH_ALGO_HMH) p = 14;
    if (!lashhost::read_sketches(path, algo, &p, n, regs, err)) return fail(LASH_HOST_E_FORMAT, err);
    if (!regs.empty()) memcpy(regs_out, regs.data(), regs.size());
    if (p_inout) *p_inout = p;
    return 0;
}

extern "C" int lash_host_dist(lash_ctx* ctx, const char* ref_prefix, const char* query_prefix, const char* output_file,
                              const char* estimator, int model, int dm, int fp32, int threads, int fused) {
    if (!ref_prefix || !query_prefix || !output_file) return fail(LASH_E_INVALID, "lash_host_dist: NULL argument");
    if (model < 0) return fail(LASH_E_INVALID, "model needs to be 0 or 1");
    return from_status(lash::dist_command(ctx, ref_prefix, query_prefix, output_file, estimator ? estimator : "fgra", (uint64_t)model,
                                          dm != 0, fp32 != 0, threads, fused != 0));
}

extern "C" int lash_host_dist_rows(lash_ctx* ctx, const char* ref_prefix, const char* query_prefix, const char* output_file,
                                   const char* estimator, int model, int dm, int fp32, int threads, int rank, int world) {
    if (!ref_prefix || !query_prefix || !output_file) return fail(LASH_E_INVALID, "lash_host_dist_rows: NULL argument");
    if (model < 0) return fail(LASH_E_INVALID, "model needs to be 0 or 1");
    return from_status(lash::dist_command(ctx, ref_prefix, query_prefix, output_file, estimator ? estimator : "fgra", (uint64_t)model,
                                          dm != 0, fp32 != 0, threads, true, rank, world));
}

extern "C" int lash_host_format_fixed6_f64(double v, char* out) {
    std::string s;
    lash::append_fixed6(s, v);
    memcpy(out, s.data(), s.size());
    return (int)s.size();
}
extern "C" size_t lash_host_format_fixed6_bulk(const double* v, size_t n, char* out) {
    return (v && out) ? lash::format_fixed6_bulk(v, n, out) : 0;
}
extern "C" int lash_host_format_fixed6_f32(float v, char* out) {
    std::string s;
    lash::append_fixed6(s, v);
    memcpy(out, s.data(), s.size());
    return (int)s.size();
}
