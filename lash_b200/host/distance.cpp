// Host side of the distance path: sketch loading, the reference's HashMap name semantics, the
// `{:.6}` writer and the `dist` sub-command body (reference src/utils.rs:84-373, src/main.rs:279-613).
// Every distance / frac value comes from lash_dist_stream (CUDA); nothing is estimated on the CPU.
#include <dirent.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <map>
#include <thread>
#include <unordered_map>

#include "io.hpp"
#include "lash_host.hpp"
#include "sketch_io.hpp"

namespace lash {

// ------------------------------------------------------------------------------------------------
// Rust `{:.6}`: the exact binary value, rounded half-to-even at the 6th decimal
// ------------------------------------------------------------------------------------------------
static void fixed6(std::string& out, double v) {
    if (std::isnan(v)) {
        out += "NaN";
        return;
    }
    if (std::signbit(v)) out.push_back('-');  // Rust prints the sign of -0.0 (e.g. -ln(1)/k, main.rs:419)
    v = std::fabs(v);
    if (std::isinf(v)) {
        out += "inf";
        return;
    }
    if (v >= 1048576.0) {  // far outside any distance; glibc's %f is exact too
        char tmp[400];
        const int n = snprintf(tmp, sizeof(tmp), "%.6f", v);
        out.append(tmp, (size_t)n);
        return;
    }
    uint64_t bits;
    memcpy(&bits, &v, 8);
    const int be = (int)(bits >> 52);
    uint64_t m = bits & ((1ull << 52) - 1);
    int e;  // v = m * 2^e
    if (be == 0) {
        e = -1074;
    } else {
        m |= 1ull << 52;
        e = be - 1075;
    }
    unsigned __int128 scaled = (unsigned __int128)m * 1000000u;  // < 2^73
    uint64_t q;
    if (e >= 0) {
        q = (uint64_t)(scaled << e);  // v < 2^20, so e <= -33 for non-integers; kept for completeness
    } else {
        const int s = -e;
        if (s >= 127) {
            q = 0;  // v * 1e6 < 2^-54: rounds to zero
        } else {
            const unsigned __int128 one = 1;
            const unsigned __int128 rem = scaled & ((one << s) - 1), half = one << (s - 1);
            q = (uint64_t)(scaled >> s);
            if (rem > half || (rem == half && (q & 1))) ++q;
        }
    }
    char tmp[32];
    int n = 32;
    uint64_t frac = q % 1000000u, ip = q / 1000000u;
    for (int i = 0; i < 6; ++i, frac /= 10) tmp[--n] = (char)('0' + frac % 10);
    tmp[--n] = '.';
    do {
        tmp[--n] = (char)('0' + ip % 10);
        ip /= 10;
    } while (ip);
    out.append(tmp + n, (size_t)(32 - n));
}
void append_fixed6(std::string& out, double v) { fixed6(out, v); }
void append_fixed6(std::string& out, float v) { fixed6(out, (double)v); }  // f32 -> f64 is exact

// Bulk-writer variant of the same formatting, into a raw buffer (>= 340 bytes of room).  For |v| < 1024 the product
// v * 1e6 is off by < 6e-8 from the exact value, so its ties-to-even rounding IS the exact answer unless it lies
// within 2e-7 of a half-integer; only those (and huge / non-finite values) take the exact 128-bit path above.
static const char kDigitPairs[] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263"
    "646566676869707172737475767778798081828384858687888990919293949596979899";
static inline char* fmt6(char* p, double v) {
    const double av = std::fabs(v);
    if (av < 1024.0) {
        const double x = av * 1e6;
        const double r = std::nearbyint(x);
        if (std::fabs(std::fabs(x - r) - 0.5) > 2e-7) {
            if (std::signbit(v)) *p++ = '-';
            const uint32_t q = (uint32_t)r;
            uint32_t ip = q / 1000000u;
            const uint32_t fr = q - ip * 1000000u;
            if (ip >= 10u) {
                char tmp[4];
                int n = 0;
                while (ip) {
                    tmp[n++] = (char)('0' + ip % 10u);
                    ip /= 10u;
                }
                while (n) *p++ = tmp[--n];
            } else {
                *p++ = (char)('0' + ip);
            }
            const uint32_t a = fr / 10000u, bc = fr - a * 10000u, b = bc / 100u, c = bc - b * 100u;
            p[0] = '.';
            memcpy(p + 1, kDigitPairs + 2 * a, 2);
            memcpy(p + 3, kDigitPairs + 2 * b, 2);
            memcpy(p + 5, kDigitPairs + 2 * c, 2);
            return p + 7;
        }
    }
    std::string slow;
    fixed6(slow, v);
    memcpy(p, slow.data(), slow.size());
    return p + slow.size();
}

// the bulk writer's formatter, '\n'-separated (test hook behind lash_host_format_fixed6_bulk)
size_t format_fixed6_bulk(const double* v, size_t n, char* out) {
    char* p = out;
    for (size_t i = 0; i < n; ++i) {
        p = fmt6(p, v[i]);
        *p++ = '\n';
    }
    return (size_t)(p - out);
}

// ------------------------------------------------------------------------------------------------
// inputs
// ------------------------------------------------------------------------------------------------
// HashMap::insert semantics (utils.rs:115,125,219,316): a repeated name keeps ONE entry holding the
// LAST sketch read for it.  Order: first occurrence (the reference's is hashbrown order).
static void dedup_last(const std::vector<std::string>& names, std::vector<std::string>& uniq, std::vector<uint64_t>& src) {
    std::unordered_map<std::string, size_t> pos;
    uniq.clear();
    src.clear();
    for (uint64_t i = 0; i < names.size(); ++i) {
        auto it = pos.find(names[i]);
        if (it == pos.end()) {
            pos.emplace(names[i], uniq.size());
            uniq.push_back(names[i]);
            src.push_back(i);
        } else {
            src[it->second] = i;
        }
    }
}

static Status load_side(int algo, int* p, const std::vector<std::string>& names, const std::string& file, std::vector<std::string>& uniq,
                        std::vector<uint8_t>& regs) {
    std::vector<uint8_t> all;
    std::string err;
    if (!lashhost::read_sketches(file, algo, p, names.size(), all, err)) return Status{LASH_HOST_E_FORMAT, "Error with reading from " + file + ": " + err};
    std::vector<uint64_t> src;
    dedup_last(names, uniq, src);
    const size_t rb = lashhost::reg_bytes(algo, *p ? *p : 14);
    if (uniq.size() == names.size()) {
        regs.swap(all);
    } else {
        regs.resize(uniq.size() * rb);
        for (size_t i = 0; i < uniq.size(); ++i) memcpy(regs.data() + i * rb, all.data() + src[i] * rb, rb);
    }
    return Status{};
}

Status load_dist_inputs(int algo, const std::vector<std::string>& reference_names, const std::string& ref_sketch_file,
                        const std::vector<std::string>& query_names, const std::string& query_sketch_file, DistInputs& in) {
    int p = algo == LASH_ALGO_HMH ? 14 : 0;
    // the reference reads the query side first (utils.rs:107), then the reference side
    Status st = load_side(algo, &p, query_names, query_sketch_file, in.qry_names, in.qry_regs);
    if (!st.ok()) return st;
    st = load_side(algo, &p, reference_names, ref_sketch_file, in.ref_names, in.ref_regs);
    if (!st.ok()) return st;
    in.p = p;
    return Status{};
}

namespace {
struct BlockCtx {
    const std::function<void(uint64_t, uint64_t, const void*)>* fn;
};
int block_trampoline(void* user, uint64_t row0, uint64_t n_rows, const void* block) {
    (*static_cast<BlockCtx*>(user)->fn)(row0, n_rows, block);
    return 0;
}
}  // namespace

std::pair<uint64_t, uint64_t> row_shard(uint64_t n_rows, int rank, int world, bool triangular) {
    auto cut = [&](int r) -> uint64_t {
        if (r <= 0) return 0;
        if (r >= world) return n_rows;
        if (!triangular) return n_rows * (uint64_t)r / (uint64_t)world;
        const double total = (double)n_rows * ((double)n_rows + 1.0) / 2.0;
        const double x = (std::sqrt(1.0 + 8.0 * total * (double)r / (double)world) - 1.0) / 2.0;
        const double rx = std::nearbyint(x);
        return rx < 0 ? 0 : (rx > (double)n_rows ? n_rows : (uint64_t)rx);
    };
    return {cut(rank), cut(rank + 1)};
}

Status distance_blocks(lash_ctx* ctx, int algo, int k, int estimator, int model, bool fp32, const DistInputs& in, bool same_files,
                       const std::function<void(uint64_t row0, uint64_t n_rows, const void* block)>& on_block, uint64_t row_begin,
                       uint64_t row_end) {
    const uint64_t n_ref = in.ref_names.size(), n_qry = in.qry_names.size();
    if (n_ref == 0 || n_qry == 0) return Status{};
    // same_files (main.rs:404) means both sides were loaded from the same files: the triangle rule of
    // utils.rs:158-160 then needs one index space, which only exists when the two lists coincide
    const bool tri = same_files && n_ref == n_qry;
    BlockCtx bc{&on_block};
    // ~64 MiB of output per block keeps the pinned staging modest and the callback granular
    const uint64_t esz = fp32 ? 4 : 8;
    uint64_t rows = std::max<uint64_t>(1, (64ull << 20) / (n_qry * esz));
    const void* q = (same_files && n_ref == n_qry) ? in.ref_regs.data() : in.qry_regs.data();
    if (row_end > n_ref) row_end = n_ref;
    if (row_begin >= row_end) return Status{};
    const int rc = lash_dist_stream_rows(ctx, algo, in.p, k, estimator, model, fp32 ? 1 : 0, in.ref_regs.data(), n_ref, q, n_qry,
                                         tri ? 1 : 0, row_begin, row_end, rows, block_trampoline, &bc);
    if (rc < 0) return Status{rc, lash_gpu_last_error()};
    return Status{rc, rc > 0 ? "some pairs fell in the HLL++ bias-table regime (see lash_gpu.h)" : ""};
}

// ------------------------------------------------------------------------------------------------
// `lash dist`
// ------------------------------------------------------------------------------------------------
namespace {

// main.rs:283-337: the three files are found by basename prefix + suffix
Status find_files(const std::string& prefix, std::map<std::string, std::string>& out) {
    std::string dir = ".", base = prefix;
    const size_t slash = prefix.find_last_of('/');
    if (slash != std::string::npos) {
        dir = slash == 0 ? "/" : prefix.substr(0, slash);
        base = prefix.substr(slash + 1);
    }
    DIR* d = opendir(dir.c_str());
    if (!d) return Status{LASH_HOST_E_IO, "cannot read directory " + dir + ": " + strerror(errno)};
    auto ends_with = [](const std::string& s, const char* suf) {
        const size_t n = strlen(suf);
        return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
    };
    while (dirent* e = readdir(d)) {
        const std::string name = e->d_name;
        if (name.compare(0, base.size(), base) != 0) continue;
        const std::string path = dir + "/" + name;
        struct stat sb;
        if (stat(path.c_str(), &sb) != 0 || !S_ISREG(sb.st_mode)) continue;
        if (ends_with(name, "parameters.json")) out["params"] = path;
        else if (ends_with(name, "files.json")) out["files"] = path;
        else if (ends_with(name, ".bin")) out["sketches"] = path;
    }
    closedir(d);
    if (out.size() != 3)
        return Status{LASH_HOST_E_IO, "There should be 3 files starting with " + base + " but " + std::to_string(out.size()) + " were found instead"};
    return Status{};
}

class OutFile {
  public:
    ~OutFile() {
        if (fd_ >= 0) ::close(fd_);
    }
    bool open(const std::string& path, std::string& err) {
        fd_ = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
        if (fd_ < 0) err = "cannot create " + path + ": " + strerror(errno);
        return fd_ >= 0;
    }
    bool write(const std::string& s) {
        const bool ok = write_at(pos_, s.data(), s.size());
        pos_ += s.size();
        return ok;
    }
    // positional write (thread-safe: several formatter threads store their parts concurrently)
    bool write_at(uint64_t at, const char* p, size_t n) const {
        size_t off = 0;
        while (off < n) {
            long w = ::pwrite(fd_, p + off, n - off, (off_t)(at + off));
            if (w < 0 && errno == EINTR) continue;
            if (w < 0) return false;
            off += (size_t)w;
        }
        return true;
    }
    uint64_t pos() const { return pos_; }
    void advance(uint64_t n) { pos_ += n; }

  private:
    int fd_ = -1;
    uint64_t pos_ = 0;
};

// growable raw text buffer of one formatter thread (capacity survives across blocks)
struct TextBuf {
    std::vector<char> mem;
    size_t n = 0;
    char* room(size_t need) {
        if (n + need > mem.size()) mem.resize(std::max(mem.size() * 2, n + need + (1u << 20)));
        return mem.data() + n;
    }
};

// format rows [r0, r1) of a block into `out` (fused path: values are final distances)
template <class T>
void format_rows(TextBuf& out, const T* block, uint64_t row0, uint64_t r0, uint64_t r1, uint64_t nq, bool tri, bool dm,
                 const std::vector<std::string>& ref_names, const std::vector<std::string>& qry_names, size_t max_qry_name,
                 const std::vector<std::vector<uint32_t>>& same_name_cols) {
    for (uint64_t r = r0; r < r1; ++r) {
        const uint64_t i = row0 + r;
        const uint64_t cols = tri ? i + 1 : nq;
        const T* v = block + r * nq;
        const std::vector<uint32_t>& zeros = same_name_cols[i];
        const std::string& rn = ref_names[i];
        size_t zi = 0;
        // worst case per cell: '\t' + 340 (a huge value through the exact path) is absurd for distances; reserve the
        // common bound and re-check inside for the exact-path cells
        const size_t per_cell = dm ? 16 : rn.size() + max_qry_name + 18;
        char* p = out.room(rn.size() + 2 + cols * per_cell + 400);
        if (dm && cols) {
            *p++ = '\n';
            memcpy(p, rn.data(), rn.size());
            p += rn.size();
        }
        for (uint64_t j = 0; j < cols; ++j) {
            T d = v[j];
            while (zi < zeros.size() && zeros[zi] < j) ++zi;
            if (zi < zeros.size() && zeros[zi] == j) d = (T)0;  // name equality => 0 (main.rs:452-453)
            if (!(std::fabs((double)d) < 1024.0)) {                  // inf / NaN / huge: may need up to 340 bytes
                out.n = (size_t)(p - out.mem.data());
                p = out.room((cols - j) * per_cell + 800);
            }
            if (!dm) {
                memcpy(p, rn.data(), rn.size());
                p += rn.size();
                *p++ = '\t';
                const std::string& qn = qry_names[j];
                memcpy(p, qn.data(), qn.size());
                p += qn.size();
                *p++ = '\t';
                p = fmt6(p, (double)d);
                *p++ = '\n';
            } else {
                *p++ = '\t';
                p = fmt6(p, (double)d);
            }
        }
        out.n = (size_t)(p - out.mem.data());
    }
}

template <class T>
Status run_dist(lash_ctx* ctx, int algo, int k, const std::string& estimator, uint64_t model, bool dm, bool same_files, int threads,
                bool fused, const std::vector<std::string>& reference_names, const std::string& ref_bin,
                const std::vector<std::string>& query_names, const std::string& query_bin, OutFile& file, int rank, int world) {
    Status wst;
    if (!fused && world == 1) {
        // the reference's own structure: *_distance(..., emit) with emit = print_dist (main.rs:474-606)
        std::string sink;
        PrintDist<T> print(sink, dm, (size_t)k, model);
        auto emit = [&](const DistRow<T>& rows) {
            print(rows);
            if (sink.size() >= (8u << 20)) {
                if (!file.write(sink)) wst = Status{LASH_HOST_E_IO, "Error writing to file"};
                sink.clear();
            }
        };
        Status st;
        if (algo == LASH_ALGO_HMH) st = hmh_distance<T>(ctx, reference_names, ref_bin, query_names, query_bin, dm, same_files, emit);
        else if (algo == LASH_ALGO_ULL) st = ull_distance<T>(ctx, reference_names, ref_bin, query_names, query_bin, estimator, dm, same_files, emit);
        else st = hll_distance<T>(ctx, reference_names, ref_bin, query_names, query_bin, dm, same_files, emit);
        if (!st.ok()) return st;
        if (!file.write(sink)) return Status{LASH_HOST_E_IO, "Error writing to file"};
        return wst.ok() ? st : wst;
    }
    // fused path: distances straight from the kernel, rows formatted by `threads` workers per block
    int est = LASH_EST_FGRA;
    if (algo == LASH_ALGO_ULL) {
        if (estimator == "ml") est = LASH_EST_ML;
        else if (estimator != "fgra") return Status{LASH_E_INVALID, "estimator needs to be either fgra or ml"};
    }
    DistInputs in;
    Status st = load_dist_inputs(algo, reference_names, ref_bin, query_names, query_bin, in);
    if (!st.ok()) return st;
    const uint64_t nq = in.qry_names.size();
    const bool tri = same_files && in.ref_names.size() == nq;
    std::unordered_map<std::string, std::vector<uint32_t>> by_name;
    for (uint32_t j = 0; j < nq; ++j) by_name[in.qry_names[j]].push_back(j);
    std::vector<std::vector<uint32_t>> same_name_cols(in.ref_names.size());
    for (size_t i = 0; i < in.ref_names.size(); ++i) {
        auto it = by_name.find(in.ref_names[i]);
        if (it != by_name.end()) same_name_cols[i] = it->second;
    }
    const std::pair<uint64_t, uint64_t> rows = row_shard(in.ref_names.size(), rank, world, tri);
    if (dm && rank == 0) {
        std::string hdr;
        for (const auto& q : in.qry_names) {
            hdr.push_back('\t');
            hdr += q;
        }
        if (!file.write(hdr)) return Status{LASH_HOST_E_IO, "Error writing columns for matrix output"};
    }
    const unsigned nt = (unsigned)std::max(1, threads);
    std::vector<TextBuf> parts(nt);
    size_t max_qry_name = 0;
    for (const auto& q : in.qry_names) max_qry_name = std::max(max_qry_name, q.size());
    std::atomic<bool> io_ok{true};
    st = distance_blocks(ctx, algo, k, est, (int)model, std::is_same<T, float>::value, in, same_files,
                         [&](uint64_t row0, uint64_t n_rows, const void* block) {
                             const T* b = static_cast<const T*>(block);
                             const unsigned use = (unsigned)std::min<uint64_t>(nt, n_rows);
                             // each worker formats its rows, learns its file offset from its predecessor (sizes are known
                             // only after formatting) and stores its part itself: formatting AND the copies into the page
                             // cache run in parallel, the file is still written in row order
                             std::vector<std::atomic<int64_t>> start(use + 1);
                             for (auto& a : start) a.store(-1, std::memory_order_relaxed);
                             start[0].store((int64_t)file.pos(), std::memory_order_release);
                             auto job = [&](unsigned t) {
                                 // triangular rows grow with i: cut the block so that workers get equal CELLS
                                 auto cut = [&](unsigned x) -> uint64_t {
                                     if (!tri) return n_rows * x / use;
                                     const double lo = (double)row0, hi = (double)(row0 + n_rows);
                                     const double area = (hi * (hi + 1) - lo * (lo + 1)) * (double)x / (double)use + lo * (lo + 1);
                                     const double rr = (std::sqrt(1.0 + 4.0 * area) - 1.0) / 2.0 - lo;
                                     const uint64_t c = rr <= 0 ? 0 : (uint64_t)std::llround(rr);
                                     return x == use ? n_rows : std::min<uint64_t>(c, n_rows);
                                 };
                                 const uint64_t r0 = cut(t), r1 = std::max(cut(t + 1), r0);
                                 parts[t].n = 0;
                                 format_rows<T>(parts[t], b, row0, r0, r1, nq, tri, dm, in.ref_names, in.qry_names, max_qry_name, same_name_cols);
                                 int64_t at;
                                 while ((at = start[t].load(std::memory_order_acquire)) < 0) std::this_thread::yield();
                                 start[t + 1].store(at + (int64_t)parts[t].n, std::memory_order_release);
                                 if (!file.write_at((uint64_t)at, parts[t].mem.data(), parts[t].n)) io_ok.store(false);
                             };
                             std::vector<std::thread> pool;
                             for (unsigned t = 1; t < use; ++t) pool.emplace_back(job, t);
                             job(0);
                             for (auto& th : pool) th.join();
                             file.advance((uint64_t)(start[use].load() - start[0].load()));
                         },
                         rows.first, rows.second);
    if (!io_ok.load()) wst = Status{LASH_HOST_E_IO, "Error writing to file"};
    if (!st.ok()) return st;
    return wst.ok() ? st : wst;
}

}  // namespace

Status dist_command(lash_ctx* ctx, const std::string& ref_prefix, const std::string& query_prefix, const std::string& output_file,
                    const std::string& estimator, uint64_t model, bool dm, bool fp32, int threads, bool fused, int rank, int world) {
    if (!ctx) return Status{LASH_E_INVALID, "dist: NULL ctx"};
    if (world < 1 || rank < 0 || rank >= world) return Status{LASH_E_INVALID, "dist: rank / world out of range"};
    if (world > 1 && !fused) return Status{LASH_E_INVALID, "dist: row sharding needs the fused writer"};
    if (model != 0 && model != 1) return Status{LASH_E_INVALID, "model needs to be 0 or 1"};  // main.rs:421
    std::map<std::string, std::string> ref_files, query_files;
    Status st = find_files(ref_prefix, ref_files);
    if (!st.ok()) return st;
    st = find_files(query_prefix, query_files);
    if (!st.ok()) return st;
    std::string err, text;
    std::map<std::string, std::string> ref_map, query_map;
    if (!lashhost::read_file(ref_files["params"], text, err) || !lashhost::json_parse_string_map(text, ref_map, err))
        return Status{LASH_HOST_E_FORMAT, ref_files["params"] + ": " + err};
    if (!lashhost::read_file(query_files["params"], text, err) || !lashhost::json_parse_string_map(text, query_map, err))
        return Status{LASH_HOST_E_FORMAT, query_files["params"] + ": " + err};
    // check that parameters match between ref and query genomes (main.rs:362-377)
    if (ref_map["k"] != query_map["k"]) return Status{LASH_HOST_E_PARAMS, "Genomes were not sketched with the same k"};
    if (ref_map["algorithm"] != query_map["algorithm"]) return Status{LASH_HOST_E_PARAMS, "Algorithms do not match in query and sketch genomes"};
    const std::string alg = ref_map["algorithm"];
    if ((alg == "ull" || alg == "hll") && ref_map["precision"] != query_map["precision"])
        return Status{LASH_HOST_E_PARAMS, alg + " was not sketched with same precision btwn genomes"};
    char* endp = nullptr;
    const unsigned long k = strtoul(ref_map["k"].c_str(), &endp, 10);
    if (ref_map["k"].empty() || *endp || k < 1 || k > 32) return Status{LASH_HOST_E_PARAMS, "invalid k in " + ref_files["params"]};
    int algo;
    if (alg == "hmh") algo = LASH_ALGO_HMH;
    else if (alg == "ull") algo = LASH_ALGO_ULL;
    else if (alg == "hll") algo = LASH_ALGO_HLL;
    else return Status{LASH_HOST_E_PARAMS, "Algorithm must be either hmh, ull, or hll"};

    std::vector<std::string> query_names, reference_names;
    if (!lashhost::read_file(query_files["files"], text, err) || !lashhost::json_parse_string_array(text, query_names, err))
        return Status{LASH_HOST_E_FORMAT, query_files["files"] + ": " + err};
    if (!lashhost::read_file(ref_files["files"], text, err) || !lashhost::json_parse_string_array(text, reference_names, err))
        return Status{LASH_HOST_E_FORMAT, ref_files["files"] + ": " + err};
    // main.rs:404 compares the two *_files.json names; the reference only ever looks in the CWD, so equal names there mean
    // the same file.  Here a prefix may carry a directory, so "sub/a" and "./sub/a" must compare equal too: same inode.
    bool same_files = query_files["files"] == ref_files["files"];
    if (!same_files) {
        struct stat sa, sb;
        same_files = stat(query_files["files"].c_str(), &sa) == 0 && stat(ref_files["files"].c_str(), &sb) == 0 &&
                     sa.st_dev == sb.st_dev && sa.st_ino == sb.st_ino;
    }

    OutFile file;
    std::string out_path = output_file;
    if (world > 1) {
        char suffix[16];
        snprintf(suffix, sizeof(suffix), ".part%04d", rank);
        out_path += suffix;
    }
    if (!file.open(out_path, err)) return Status{LASH_HOST_E_IO, err};
    if (!dm && rank == 0 && !file.write("Reference\tQuery\tDistance\n")) return Status{LASH_HOST_E_IO, "Error writing to file"};  // main.rs:409-412
    if (fp32)
        return run_dist<float>(ctx, algo, (int)k, estimator, model, dm, same_files, threads, fused, reference_names, ref_files["sketches"],
                               query_names, query_files["sketches"], file, rank, world);
    return run_dist<double>(ctx, algo, (int)k, estimator, model, dm, same_files, threads, fused, reference_names, ref_files["sketches"],
                            query_names, query_files["sketches"], file, rank, world);
}

}  // namespace lash
