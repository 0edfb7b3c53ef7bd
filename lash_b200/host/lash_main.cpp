// `lash-b200`: the reference's two sub-commands on top of the host layer, same flags and defaults
// (reference src/main.rs:26-177):
//   lash-b200 sketch -f <list file> [-o sketch] [-k 16] [-t threads] [-a hmh|hll|ull] [-p 10] [-s 42]
//   lash-b200 dist   -q <query prefix> -r <reference prefix> [-o dist] [-t threads] [-e fgra|ml] [-m 1|0] [--fp32] [--dm]
// Extras (not in the reference): --device N (GPU index, default 0); --rank R --world W (dist: one process per GPU, each
// writes its row range to <output>.partRRRR); --mirror (dist: take `frac` from the GPU and
// run compute_distance + print_dist on the host exactly as main.rs does, instead of the fused kernel epilogue).
// This is SURVEY.md 8f row 4 (CLI parity); it adds no arithmetic of its own.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "lash_host.hpp"

namespace {

struct Args {
    std::vector<std::pair<std::string, std::string>> kv;  // (canonical long name, value); flags have value "1"
    const std::string* get(const std::string& k) const {
        for (const auto& e : kv)
            if (e.first == k) return &e.second;
        return nullptr;
    }
};

struct Opt {
    char short_name;  // 0: none
    const char* long_name;
    bool takes_value;
};

bool parse(int argc, char** argv, int first, const std::vector<Opt>& opts, Args& out, std::string& err) {
    for (int i = first; i < argc; ++i) {
        const std::string a = argv[i];
        const Opt* hit = nullptr;
        std::string inline_val;
        bool has_inline = false;
        for (const auto& o : opts) {
            if (a == std::string("--") + o.long_name || (o.short_name && a == std::string("-") + o.short_name)) hit = &o;
            const std::string pre = std::string("--") + o.long_name + "=";
            if (a.compare(0, pre.size(), pre) == 0) {
                hit = &o;
                inline_val = a.substr(pre.size());
                has_inline = true;
            }
            if (hit) break;
        }
        if (!hit) {
            err = "error: unexpected argument '" + a + "' found";
            return false;
        }
        if (!hit->takes_value) {
            out.kv.emplace_back(hit->long_name, "1");
        } else if (has_inline) {
            out.kv.emplace_back(hit->long_name, inline_val);
        } else {
            if (i + 1 >= argc) {
                err = "error: a value is required for '--" + std::string(hit->long_name) + "' but none was supplied";
                return false;
            }
            out.kv.emplace_back(hit->long_name, argv[++i]);
        }
    }
    return true;
}

bool to_u64(const std::string& s, uint64_t& v) {
    if (s.empty()) return false;
    char* end = nullptr;
    v = strtoull(s.c_str(), &end, 10);
    return *end == 0 && s[0] != '-';
}

int usage() {
    fprintf(stderr,
            "Fast and Memory Efficient (Meta)genome Sketching via HyperLogLog, HyperMinhash and UltraLogLog (B200)\n\n"
            "Usage: lash-b200 <COMMAND>\n\nCommands:\n"
            "  sketch  Sketches genomes and serializes them, sketches are compressed\n"
            "          -f, --file <list>  -o, --output <prefix=sketch>  -k, --kmer <16>  -t, --threads <n>\n"
            "          -a, --algorithm <hmh|hll|ull = hmh>  -p, --precision <10>  -s, --seed <42>  [--device <0>]\n"
            "  dist    Computes distance between sketches\n"
            "          -q, --query <prefix>  -r, --reference <prefix>  -o, --output_file <dist>  -t, --threads <n>\n"
            "          -e, --estimator <fgra|ml = fgra>  -m, --model <1|0 = 1>  --fp32  --dm  [--device <0>] [--mirror]\n");
    return 2;
}

// per-command help: the reference's own text (clap output quoted in its README), plus this binary's extras
int help_sketch() {
    printf("Sketches genomes and serializes them, sketches are compressed\n\n"
           "Usage: lash-b200 sketch [OPTIONS] --file <file>\n\nOptions:\n"
           "  -f, --file <file>            One file containing list of FASTA/FASTQ files (.gz/.bz2/.zstd supported), one per line. File must be UTF-8.\n"
           "  -o, --output <output>        Input a prefix/name for your output files [default: sketch]\n"
           "  -k, --kmer <kmer_length>     Length of the kmer [default: 16]\n"
           "  -t, --threads <threads>      Number of threads to use, default to all logical cores\n"
           "  -a, --algorithm <algorithm>  Which algorithm to use: HyperMinHash (hmh), UltraLogLog (ull), or HyperLogLog (hll) [default: hmh]\n"
           "  -p, --precision <precision>  Specifiy precision, for ull and hll only. [default: 10]\n"
           "  -s, --seed <seed>            Random seed [default: 42]\n"
           "      --device <device>        GPU index [default: 0]  (lash-b200 only)\n"
           "  -h, --help                   Print help\n");
    return 0;
}
int help_dist() {
    printf("Computes distance between sketches\n\n"
           "Usage: lash-b200 dist [OPTIONS] --query <query> --reference <reference>\n\nOptions:\n"
           "  -q, --query <query>              Prefix to search for query genome files\n"
           "  -r, --reference <reference>      Prefix to search for reference genome files\n"
           "  -o, --output_file <output_file>  Name of output file to write results [default: dist]\n"
           "  -t, --threads <threads>          Number of threads to use, default to all logical cores\n"
           "  -e, --estimator <estimator>      Specify estimator (fgra or ml), for ull only [default: fgra]\n"
           "  -m, --model <model>              Equation used to calculate distance: 1 for poisson model or 0 for binomial model [default: 1]\n"
           "      --fp32                       Distance output in float 32 instead of 64\n"
           "      --dm                         Prints distance matrix\n"
           "      --device <device>            GPU index [default: 0]  (lash-b200 only)\n"
           "      --mirror                     frac from the GPU, compute_distance + print_dist on the host as main.rs does  (lash-b200 only)\n"
           "      --rank <r> --world <w>       one process per GPU: write this rank's row range to <output_file>.partRRRR  (lash-b200 only)\n"
           "      --allow-hll-bias-regime      hll: write pairs whose estimate needs the HLL++ bias tables as distance 1 instead of failing  (lash-b200 only)\n"
           "  -h, --help                       Print help\n");
    return 0;
}
bool wants_help(int argc, char** argv) {
    for (int i = 2; i < argc; ++i)
        if (!strcmp(argv[i], "-h") || !strcmp(argv[i], "--help")) return true;
    return false;
}

int fail(const std::string& m) {
    fprintf(stderr, "%s\n", m.c_str());
    return 1;
}

int run_sketch(int argc, char** argv) {
    Args a;
    std::string err;
    if (!parse(argc, argv, 2,
               {{'f', "file", true}, {'o', "output", true}, {'k', "kmer", true}, {'t', "threads", true}, {'a', "algorithm", true},
                {'p', "precision", true}, {'s', "seed", true}, {0, "device", true}},
               a, err))
        return fail(err);
    if (!a.get("file")) return fail("error: the following required arguments were not provided:\n  --file <file>");
    const std::string output = a.get("output") ? *a.get("output") : "sketch";
    const std::string alg = a.get("algorithm") ? *a.get("algorithm") : "hmh";
    uint64_t k = 16, threads = std::max(1u, std::thread::hardware_concurrency()), precision = 10, seed = 42, device = 0;
    if (a.get("kmer") && !to_u64(*a.get("kmer"), k)) return fail("error: invalid value for '--kmer <kmer_length>'");
    if (a.get("threads") && !to_u64(*a.get("threads"), threads)) return fail("error: invalid value for '--threads <threads>'");
    if (a.get("precision") && !to_u64(*a.get("precision"), precision)) return fail("error: invalid value for '--precision <precision>'");
    if (a.get("seed") && !to_u64(*a.get("seed"), seed)) return fail("error: invalid value for '--seed <seed>'");
    if (a.get("device") && !to_u64(*a.get("device"), device)) return fail("error: invalid value for '--device'");

    // main.rs:200-207: one path per line, blank lines skipped, lines kept verbatim
    std::ifstream in(*a.get("file"));
    if (!in) return fail("cannot open " + *a.get("file"));
    std::vector<std::string> files;
    for (std::string line; std::getline(in, line);) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.find_first_not_of(" \t\r\n\v\f") == std::string::npos) continue;
        files.push_back(line);
    }
    int algo;
    if (alg == "hmh") algo = LASH_ALGO_HMH;
    else if (alg == "hll") algo = LASH_ALGO_HLL;
    else if (alg == "ull") algo = LASH_ALGO_ULL;
    else return fail("Algorithm must be either hmh, ull, or hll");  // main.rs:245

    lash_ctx* ctx = nullptr;
    if (lash_ctx_create((int)device, &ctx) != LASH_OK) return fail(lash_gpu_last_error());
    lash::SketchFilesStats st{};
    lash::Status rc;
    const uint32_t th = (uint32_t)std::max<uint64_t>(threads, 1);
    if (algo == LASH_ALGO_HMH) rc = lash::sketch_files<lash::Hmh>(ctx, std::nullopt, files, (size_t)k, output, th, seed, &st);
    else if (algo == LASH_ALGO_HLL) rc = lash::sketch_files<lash::Hll>(ctx, (uint32_t)precision, files, (size_t)k, output, th, seed, &st);
    else rc = lash::sketch_files<lash::Ull>(ctx, (uint32_t)precision, files, (size_t)k, output, th, seed, &st);
    // the parameter JSON is written whether or not sketching succeeded (main.rs:249-276 precedes `result`)
    const int prc = lash_host_write_parameters(output.c_str(), algo, (int)precision, (int)k, seed);
    lash_ctx_destroy(ctx);
    if (!rc.ok()) return fail(rc.message);
    if (prc != 0) return fail(lash_host_last_error());
    fprintf(stderr, "sketched %zu files: %llu records, %llu bases kept, %.3f s (GPU kernels %.3f ms, %llu pushes)\n", files.size(),
            (unsigned long long)st.n_records, (unsigned long long)st.n_bases_kept, st.seconds_total, st.gpu_kernel_ms,
            (unsigned long long)st.n_pushes);
    return 0;
}

int run_dist(int argc, char** argv) {
    Args a;
    std::string err;
    if (!parse(argc, argv, 2,
               {{'q', "query", true}, {'r', "reference", true}, {'o', "output_file", true}, {'t', "threads", true},
                {'e', "estimator", true}, {'m', "model", true}, {0, "fp32", false}, {0, "dm", false}, {0, "device", true},
                {0, "mirror", false}, {0, "rank", true}, {0, "world", true}, {0, "allow-hll-bias-regime", false}},
               a, err))
        return fail(err);
    if (!a.get("query") || !a.get("reference"))
        return fail("error: the following required arguments were not provided:\n  --query <query>\n  --reference <reference>");
    const std::string output = a.get("output_file") ? *a.get("output_file") : "dist";
    const std::string estimator = a.get("estimator") ? *a.get("estimator") : "fgra";
    uint64_t threads = std::max(1u, std::thread::hardware_concurrency()), model = 1, device = 0, rank = 0, world = 1;
    if (a.get("rank") && !to_u64(*a.get("rank"), rank)) return fail("error: invalid value for '--rank'");
    if (a.get("world") && !to_u64(*a.get("world"), world)) return fail("error: invalid value for '--world'");
    if (a.get("threads") && !to_u64(*a.get("threads"), threads)) return fail("error: invalid value for '--threads <threads>'");
    if (a.get("model") && !to_u64(*a.get("model"), model)) return fail("error: invalid value for '--model <model>'");
    if (a.get("device") && !to_u64(*a.get("device"), device)) return fail("error: invalid value for '--device'");
    lash_ctx* ctx = nullptr;
    if (lash_ctx_create((int)device, &ctx) != LASH_OK) return fail(lash_gpu_last_error());
    const lash::Status rc = lash::dist_command(ctx, *a.get("reference"), *a.get("query"), output, estimator, model, a.get("dm") != nullptr,
                                               a.get("fp32") != nullptr, (int)std::max<uint64_t>(threads, 1), a.get("mirror") == nullptr, (int)rank,
                                               (int)world);
    lash_ctx_destroy(ctx);
    if (!rc.ok()) return fail(rc.message);
    if (rc.code > 0) {
        // LASH_W_HLL_BIAS_REGIME: some pair's HLL estimate lies in (linear-counting threshold, 5m], where the reference's
        // len() subtracts the HLL++ empirical bias (utils.rs:315,358 -> streaming_algorithms).  Those tables cannot be
        // reproduced offline; the cells were written as distance 1.  A silently different file is worse than none.
        if (!a.get("allow-hll-bias-regime")) {
            std::string out_path = output;
            if (world > 1) {
                char suffix[16];
                snprintf(suffix, sizeof(suffix), ".part%04d", (int)rank);
                out_path += suffix;
            }
            remove(out_path.c_str());
            return fail("error: " + rc.message + "; no output written.  These sketches are too small for HLL at this precision in this build "
                        "(use ull or hmh, or a smaller -p); --allow-hll-bias-regime writes such pairs as distance 1.");
        }
        fprintf(stderr, "warning: %s\n", rc.message.c_str());
    }
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) return usage();
    const std::string cmd = argv[1];
    if (cmd == "sketch") return wants_help(argc, argv) ? help_sketch() : run_sketch(argc, argv);
    if (cmd == "dist") return wants_help(argc, argv) ? help_dist() : run_dist(argc, argv);
    if (cmd == "help" || cmd == "-h" || cmd == "--help") {
        if (argc > 2 && !strcmp(argv[2], "sketch")) return help_sketch();
        if (argc > 2 && !strcmp(argv[2], "dist")) return help_dist();
        usage();
        return 0;
    }
    if (cmd == "-V" || cmd == "--version") {
        printf("lash-b200 0.1.4 (B200 hot paths; reference jianshu93/lash 0.1.4)\n");
        return 0;
    }
    return usage();
}
