// C++ host API above the C ABI (include/lash_gpu.h): the reference's own operator interface for
// the two hot paths, name for name, so that call sites and tests read like the reference's.
//
//   reference (Rust, src/utils.rs / src/main.rs)                   here
//   ---------------------------------------------------------------------------------------------
//   trait KmerSketch, impls for Sketch / HyperLogLog<i64> /        tag types lash::Hmh / Hll / Ull
//     UltraLogLog                          utils.rs:377-433          (the sketch state lives on the GPU)
//   sketch_files::<S>(precision, files, kmer_length, output_name,  lash::sketch_files<S>(ctx, ...)
//     threads, seed, aa)                   utils.rs:439-581
//   hmh_distance / ull_distance / hll_distance::<F, T>(...)        lash::hmh_distance<T> / ull_distance<T> /
//                                          utils.rs:84-373          hll_distance<T>(ctx, ..., emit)
//   compute_distance::<F>, print_dist::<T> main.rs:415-471         lash::compute_distance<T>, lash::PrintDist<T>
//   `dist` sub-command body                main.rs:279-613         lash::dist_command(...)
//
// Differences that are deliberate and documented in DESIGN.md: errors are returned (Status) instead
// of panicking; rows are emitted in list order (the reference's order is hashbrown iteration order);
// amino-acid sketching (dead code behind `let aa = false`, main.rs:198) is not provided.
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>
#include <optional>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/lash_host.h"

namespace lash {

struct Status {
    int code = 0;  // 0 ok, <0 error, >0 warning (LASH_W_HLL_BIAS_REGIME)
    std::string message;
    bool ok() const { return code >= 0; }
};

struct Hmh { static constexpr int algo = LASH_ALGO_HMH; static constexpr const char* name = "hmh"; };
struct Hll { static constexpr int algo = LASH_ALGO_HLL; static constexpr const char* name = "hll"; };
struct Ull { static constexpr int algo = LASH_ALGO_ULL; static constexpr const char* name = "ull"; };

using SketchFilesStats = lash_sketch_files_stats;

// ---- sketching ----------------------------------------------------------------------------------
// non-template core; regs_out (optional) receives the registers, output_name (optional) the files
Status sketch_files_impl(lash_ctx* ctx, int algo, std::optional<uint32_t> precision, const std::vector<std::string>& files,
                         size_t kmer_length, const std::string* output_name, uint32_t threads, uint64_t seed,
                         uint64_t chunk_bytes, void* regs_out, SketchFilesStats* stats);

// host ingest ceiling: the same parse + filter + pack, chunks dropped instead of pushed (no GPU work)
Status pack_files_dry(const std::vector<std::string>& files, size_t kmer_length, uint32_t threads, uint64_t chunk_bytes,
                      SketchFilesStats* stats);

// return the cached pinned staging blocks of sketch_files to the driver
void release_pinned();
// 0 auto, 1 host filter + pack (lash_sketch_push), 2 device filter + pack (lash_sketch_push_ascii)
void set_ingest_mode(int mode);

template <class S>
Status sketch_files(lash_ctx* ctx, std::optional<uint32_t> precision, const std::vector<std::string>& files, size_t kmer_length,
                    const std::string& output_name, uint32_t threads, uint64_t seed, SketchFilesStats* stats = nullptr) {
    return sketch_files_impl(ctx, S::algo, precision, files, kmer_length, &output_name, threads, seed, 0, nullptr, stats);
}

// ---- distance -----------------------------------------------------------------------------------
template <class T>
using DistRow = std::vector<std::tuple<const std::string*, const std::string*, T>>;

// core: emit_block(ref index of first row, rows, query count, values) with values = frac cast to T,
// dense [rows][n_qry]; names after the reference's HashMap de-duplication are returned in *_names_out.
struct DistInputs {
    std::vector<std::string> ref_names, qry_names;  // unique names, list order
    std::vector<uint8_t> ref_regs, qry_regs;
    int p = 0;
};
Status load_dist_inputs(int algo, const std::vector<std::string>& reference_names, const std::string& ref_sketch_file,
                        const std::vector<std::string>& query_names, const std::string& query_sketch_file, DistInputs& in);
// reference rows [row_begin, row_end) only (row_end = ~0: all) -- one process per GPU takes one row range
Status distance_blocks(lash_ctx* ctx, int algo, int k, int estimator, int model, bool fp32, const DistInputs& in, bool same_files,
                       const std::function<void(uint64_t row0, uint64_t n_rows, const void* block)>& on_block,
                       uint64_t row_begin = 0, uint64_t row_end = ~0ull);
// [begin, end) reference rows of process `rank` of `world`; triangular: cut so that every process gets the same
// number of pairs (rows [0, x) of a lower triangle hold x(x+1)/2 of them) -- same cuts as lash_b200/shard.py::row_shard
std::pair<uint64_t, uint64_t> row_shard(uint64_t n_rows, int rank, int world, bool triangular);

template <class T, class F>
Status distance_generic(lash_ctx* ctx, int algo, const std::string* estimator, const std::vector<std::string>& reference_names,
                        const std::string& ref_sketch_file, const std::vector<std::string>& query_names,
                        const std::string& query_sketch_file, bool create_matrix, bool same_files, F&& emit) {
    static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value, "T is f32 or f64");
    int est = LASH_EST_FGRA;
    if (estimator) {
        if (*estimator == "fgra") est = LASH_EST_FGRA;
        else if (*estimator == "ml") est = LASH_EST_ML;
        else return Status{LASH_E_INVALID, "estimator needs to be either fgra or ml"};  // utils.rs:217
    }
    DistInputs in;
    Status st = load_dist_inputs(algo, reference_names, ref_sketch_file, query_names, query_sketch_file, in);
    if (!st.ok()) return st;
    static const std::string blank;
    if (create_matrix) {  // empty r_name string signals printing columns (utils.rs:133-146)
        DistRow<T> columns;
        for (const auto& q : in.qry_names) columns.emplace_back(&blank, &q, (T)1);
        emit(columns);
    }
    const uint64_t nq = in.qry_names.size();
    const bool tri = same_files && in.ref_names.size() == nq;
    return distance_blocks(ctx, algo, 16, est, LASH_MODEL_FRAC, std::is_same<T, float>::value, in, same_files,
                           [&](uint64_t row0, uint64_t n_rows, const void* block) {
                               const T* v = static_cast<const T*>(block);
                               for (uint64_t r = 0; r < n_rows; ++r) {
                                   const uint64_t i = row0 + r;
                                   const uint64_t cols = tri ? i + 1 : nq;  // utils.rs:158-160 with idx = list position
                                   DistRow<T> row;
                                   row.reserve(cols);
                                   for (uint64_t j = 0; j < cols; ++j) row.emplace_back(&in.ref_names[i], &in.qry_names[j], v[r * nq + j]);
                                   emit(row);
                               }
                           });
}

// utils.rs:84-94
template <class T, class F>
Status hmh_distance(lash_ctx* ctx, const std::vector<std::string>& reference_names, const std::string& ref_sketch_file,
                    const std::vector<std::string>& query_names, const std::string& query_sketch_file, bool create_matrix,
                    bool same_files, F&& emit) {
    return distance_generic<T>(ctx, LASH_ALGO_HMH, nullptr, reference_names, ref_sketch_file, query_names, query_sketch_file,
                               create_matrix, same_files, emit);
}
// utils.rs:186-197
template <class T, class F>
Status ull_distance(lash_ctx* ctx, const std::vector<std::string>& reference_names, const std::string& ref_sketch_file,
                    const std::vector<std::string>& query_names, const std::string& query_sketch_file, const std::string& estimator,
                    bool create_matrix, bool same_files, F&& emit) {
    return distance_generic<T>(ctx, LASH_ALGO_ULL, &estimator, reference_names, ref_sketch_file, query_names, query_sketch_file,
                               create_matrix, same_files, emit);
}
// utils.rs:290-299
template <class T, class F>
Status hll_distance(lash_ctx* ctx, const std::vector<std::string>& reference_names, const std::string& ref_sketch_file,
                    const std::vector<std::string>& query_names, const std::string& query_sketch_file, bool create_matrix,
                    bool same_files, F&& emit) {
    return distance_generic<T>(ctx, LASH_ALGO_HLL, nullptr, reference_names, ref_sketch_file, query_names, query_sketch_file,
                               create_matrix, same_files, emit);
}

// main.rs:415-423.  T-typed arithmetic: with --fp32 ln / powf run in f32 on the f32-cast frac.
template <class T>
inline T compute_distance(T frac, size_t kmer_length, uint8_t equation) {
    const T k = (T)kmer_length;
    if (equation == 1) return std::fmin(-std::log(frac) / k, (T)1);
    return (T)1 - std::pow(frac, (T)1 / k);
}

// Rust `{:.6}` (exact, round-half-even); appends to out
void append_fixed6(std::string& out, double v);
void append_fixed6(std::string& out, float v);
// the bulk writer's fast path of the same formatting ('\n' after each value; out: 341 bytes per value); returns bytes written
size_t format_fixed6_bulk(const double* v, size_t n, char* out);

// main.rs:429-471: the emit callback that writes the TSV list / the --dm matrix
template <class T>
class PrintDist {
  public:
    PrintDist(std::string& sink, bool create_matrix, size_t kmer_length, uint64_t equation)
        : sink_(sink), create_matrix_(create_matrix), k_(kmer_length), eq_(equation) {}
    void operator()(const DistRow<T>& distance_list) {
        if (create_matrix_ && !distance_list.empty() && std::get<0>(distance_list[0])->empty()) {
            for (const auto& col : distance_list) {
                sink_.push_back('\t');
                sink_ += *std::get<1>(col);
            }
            return;
        }
        size_t i = 0;
        for (const auto& row : distance_list) {
            const std::string& r_name = *std::get<0>(row);
            const std::string& q_name = *std::get<1>(row);
            const T d = (q_name == r_name) ? (T)0 : compute_distance<T>(std::get<2>(row), k_, (uint8_t)eq_);  // main.rs:452-456
            if (!create_matrix_) {
                sink_ += r_name;
                sink_.push_back('\t');
                sink_ += q_name;
                sink_.push_back('\t');
                append_fixed6(sink_, d);
                sink_.push_back('\n');
            } else {
                if (i == 0) {
                    sink_.push_back('\n');
                    sink_ += r_name;
                }
                sink_.push_back('\t');
                append_fixed6(sink_, d);
            }
            ++i;
        }
    }

  private:
    std::string& sink_;
    bool create_matrix_;
    size_t k_;
    uint64_t eq_;
};

// `lash dist` (main.rs:279-613); see lash_host_dist in include/lash_host.h for the arguments
// rank / world: this process writes only its row range, into "<output_file>.part<rank, 4 digits>" when world > 1
// (rank 0's part carries the header); the parts concatenated in rank order are the single-process file.
Status dist_command(lash_ctx* ctx, const std::string& ref_prefix, const std::string& query_prefix, const std::string& output_file,
                    const std::string& estimator, uint64_t model, bool dm, bool fp32, int threads, bool fused, int rank = 0,
                    int world = 1);

}  // namespace lash
