// sketch_files<S> (reference src/utils.rs:439-581) on top of the C ABI.
//
// Reference shape: files.par_iter().map(|file| { reader; sketch = S::new(p); for record { filter_out_n;
// skip if shorter than k; k-mers -> add_kmer } sketch }).collect(), then S::save of every sketch into
// one zstd stream + the names JSON.  "Each file is processed to completion in its own task."
//
// Here: the same "parallel by sample" -- a pool of host workers takes files one at a time (dynamic,
// like rayon's work stealing) -- but a worker does no hashing: it parses, filters + 2-bit packs
// (pack.cpp) into a pinned chunk and hands full chunks to the GPU (lash_sketch_push) while it fills
// its second chunk.  Several small files share a chunk (one span each); a large record is split
// across chunks with a (k-1)-base overlap so that every k-mer start is produced exactly once.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>

#include "fastx.hpp"
#include "lash_host.hpp"
#include "pack.hpp"
#include "sketch_io.hpp"

namespace lash {

using lashhost::BaseStream;
using lashhost::FastxReader;

namespace {

constexpr uint64_t kDefaultChunk = 16ull << 20;
constexpr size_t kFeed = 1u << 16;  // sequence bytes handed to the packer per call (room is checked per call)

// Pinned staging blocks are kept across calls: page-locking is slow (cudaHostAlloc runs at ~1-3 GB/s, so the
// 16 workers x 2 x 16 MiB of a default call cost more than sketching a few hundred Mbp) and a host that sketches
// batch after batch should pay it once.  lash_host_release_pinned() returns the blocks to the driver.
class PinnedPool {
  public:
    void* get(uint64_t bytes, uint64_t* got) {
        {
            std::lock_guard<std::mutex> g(mu_);
            auto it = free_.lower_bound(bytes);
            if (it != free_.end() && it->first <= 2 * bytes) {
                void* p = it->second;
                *got = it->first;
                pooled_ -= it->first;
                free_.erase(it);
                return p;
            }
        }
        void* p = nullptr;
        if (lash_host_alloc(bytes, &p) != LASH_OK) return nullptr;
        *got = bytes;
        return p;
    }
    void put(void* p, uint64_t bytes) {
        {
            std::lock_guard<std::mutex> g(mu_);
            if (pooled_ + bytes <= kMaxPooled) {
                free_.emplace(bytes, p);
                pooled_ += bytes;
                return;
            }
        }
        lash_host_free(p);
    }
    void clear() {
        std::lock_guard<std::mutex> g(mu_);
        for (auto& kv : free_) lash_host_free(kv.second);
        free_.clear();
        pooled_ = 0;
    }

  private:
    static constexpr uint64_t kMaxPooled = 1ull << 30;
    std::mutex mu_;
    std::multimap<uint64_t, void*> free_;
    uint64_t pooled_ = 0;
};
PinnedPool& pinned_pool() {
    static PinnedPool* pool = new PinnedPool();  // never destroyed: the CUDA runtime may be gone at exit
    return *pool;
}

// serialises access to the (not thread-safe) sketcher handle
struct Gpu {
    lash_sketcher* sk = nullptr;
    std::mutex mu;
    std::atomic<uint64_t> pushes{0};
    std::string err;  // first error
    std::atomic<bool> failed{false};
    void fail(const std::string& m) {
        std::lock_guard<std::mutex> g(mu);
        if (!failed.exchange(true)) err = m;
    }
};

// One pinned chunk being filled: spans of (possibly several) genomes, their record tables
class Chunk {
  public:
    // pinned: staging for lash_sketch_push; !pinned: plain memory for the parse+pack dry run (no GPU involved)
    // text mode (lash_sketch_push_ascii): the chunk holds the raw sequence bytes of the records, one separator byte after
    // every record; the device does filter_out_n + the 2-bit packing.  The host work per base is a memcpy.
    void set_text_mode(bool on) { text_ = on; }
    bool text_mode() const { return text_; }

    bool alloc(uint64_t bytes, bool pinned = true) {
        cap_ = bytes;
        pinned_ = pinned;
        void* p = nullptr;
        if (pinned) {
            p = pinned_pool().get(bytes, &block_);
            if (!p) return false;
        } else if (posix_memalign(&p, 64, bytes) != 0) {
            return false;
        }
        buf_ = static_cast<uint8_t*>(p);
        return true;
    }
    bool allocated() const { return buf_ != nullptr; }
    void release() {
        if (buf_ && pinned_) pinned_pool().put(buf_, block_);
        else if (buf_) free(buf_);
        buf_ = nullptr;
    }
    bool empty() const { return spans_.empty() && tspans_.empty() && !span_open_; }
    uint64_t used() const { return off_; }
    uint64_t cap() const { return cap_; }
    // room (in bases; text mode: bytes) the current span can still take
    uint64_t room() const {
        if (!text_) return bs_.room();
        const uint64_t used = off_ + tw_ + 64;  // separator + 16-byte pad + slack
        return cap_ > used ? cap_ - used : 0;
    }
    bool can_begin_span() const { return off_ + 4096 <= cap_; }

    void begin_span(uint64_t genome) {
        genome_ = genome;
        if (text_) {
            tw_ = 0;
            n_seps_ = 0;
            rec_begin_ = 0;
            span_open_ = true;
            return;
        }
        bs_.attach(buf_ + off_, cap_ - off_);
        rec_first_ = rec_start_.size();
        rec_start_.push_back(0);
        n_rec_ = 0;
        uniform_len_ = 0;
        uniform_ = true;
        rec_begin_ = 0;
        span_open_ = true;
    }
    void begin_record() { rec_begin_ = text_ ? tw_ : bs_.size(); }
    uint64_t append(const uint8_t* s, size_t n) {
        if (!text_) return bs_.append_filtered(s, n);
        uint8_t* dst = buf_ + off_ + tw_;
        memcpy(dst, s, n);
        // a sequence byte equal to the separator would split the record: it is not a base, so any other non-base byte
        // stands for it (filter_out_n deletes both)
        for (uint8_t* q = static_cast<uint8_t*>(memchr(dst, LASH_TEXT_RECORD_SEP, n)); q;
             q = static_cast<uint8_t*>(memchr(q + 1, LASH_TEXT_RECORD_SEP, (size_t)(dst + n - (q + 1)))))
            *q = '\n';
        tw_ += n;
        return 0;  // bases kept: only the device knows
    }
    // ---- text mode, plain FASTA read straight into the pinned chunk (no intermediate buffer, no memcpy) ----------------
    // The file's bytes land where the GPU will read them; the only host work per byte is the scan below: every header
    // line is overwritten in place -- its '>' by the record separator, its text by 'N' (never a base) -- and a sequence
    // byte equal to the separator by a line feed.  State carries across pieces and chunks.
    struct FastaScan {
        bool in_header = false;      // the piece starts inside a header line
        bool at_line_start = true;   // the byte before the piece was '\n' (or the piece starts the file)
        uint64_t n_headers = 0;
    };
    uint8_t* text_ptr() { return buf_ + off_ + tw_; }
    void text_commit_fasta(size_t n, FastaScan& st) {
        uint8_t* p = buf_ + off_ + tw_;
        size_t i = 0;
        while (i < n) {
            if (st.in_header) {
                uint8_t* nl = static_cast<uint8_t*>(memchr(p + i, '\n', n - i));
                const size_t end = nl ? (size_t)(nl - p) : n;
                memset(p + i, 'N', end - i);
                if (!nl) break;
                st.in_header = false;
                i = end + 1;
                continue;
            }
            uint8_t* g = static_cast<uint8_t*>(memchr(p + i, '>', n - i));
            const size_t seq_end = g ? (size_t)(g - p) : n;
            for (uint8_t* q = static_cast<uint8_t*>(memchr(p + i, LASH_TEXT_RECORD_SEP, seq_end - i)); q;
                 q = static_cast<uint8_t*>(memchr(q + 1, LASH_TEXT_RECORD_SEP, (size_t)(p + seq_end - (q + 1)))))
                *q = '\n';
            if (!g) break;
            const bool line_start = seq_end > 0 ? p[seq_end - 1] == '\n' : st.at_line_start;
            if (line_start) {   // a header: the previous record ends here
                p[seq_end] = LASH_TEXT_RECORD_SEP;
                ++n_seps_;
                ++st.n_headers;
                rec_begin_ = tw_ + seq_end + 1;
                st.in_header = true;
            }                   // else: a '>' inside a sequence line is just a byte filter_out_n deletes
            i = seq_end + 1;
        }
        st.at_line_start = n > 0 ? p[n - 1] == '\n' : st.at_line_start;
        tw_ += n;
    }

    // the (k-1)-base overlap of a record that was split across chunks
    void push_carry(const std::vector<uint8_t>& carry) {
        if (text_) {
            if (!carry.empty()) append(carry.data(), carry.size());
        } else {
            for (uint8_t c : carry) bs_.push_base(c);
        }
    }
    uint64_t record_len() const { return bs_.size() - rec_begin_; }
    unsigned record_base(uint64_t i) const { return bs_.base_at(rec_begin_ + i); }
    // utils.rs:460-462: a (filtered) record shorter than k contributes nothing -- it is dropped here
    void end_record(int k) {
        if (text_) {  // records shorter than k are dropped by the device (their starts are all invalid)
            buf_[off_ + tw_++] = LASH_TEXT_RECORD_SEP;
            ++n_seps_;
            return;
        }
        const uint64_t len = record_len();
        if (len < (uint64_t)k) {
            bs_.truncate(rec_begin_);
            return;
        }
        close_record(len);
    }
    // first part of a record that continues in the next chunk; returns the bases to carry over
    void split_record(int k, std::vector<uint8_t>& carry) {
        if (text_) {
            // carry = the last k-1 BASES of this part (the bytes between them are deleted by the filter anyway)
            const uint8_t* b = buf_ + off_;
            carry.clear();
            for (uint64_t pos = tw_; pos > rec_begin_ && carry.size() < (size_t)(k - 1);) {
                const uint8_t c = b[--pos];
                if (c == 'A' || c == 'C' || c == 'G' || c == 'T') carry.push_back(c);
            }
            std::reverse(carry.begin(), carry.end());
            buf_[off_ + tw_++] = LASH_TEXT_RECORD_SEP;
            ++n_seps_;
            return;
        }
        const uint64_t len = record_len();
        const uint64_t c = std::min<uint64_t>(len, (uint64_t)(k - 1));
        carry.resize(c);
        for (uint64_t i = 0; i < c; ++i) carry[i] = (uint8_t)record_base(len - c + i);
        if (len < (uint64_t)k) bs_.truncate(rec_begin_);  // no k-mer starts here: the carry holds all of it
        else close_record(len);
    }
    void end_span() {
        if (text_) {
            span_open_ = false;
            if (tw_ == 0) return;
            lash_text_span sp;
            sp.genome = genome_;
            sp.byte_off = off_;
            sp.n_bytes = tw_;
            // records = separators strictly inside the span + 1; one record needs no boundary bitmask on the device
            const uint8_t* b = buf_ + off_;
            uint64_t inner = n_seps_;
            if (inner && b[0] == LASH_TEXT_RECORD_SEP) --inner;
            if (inner && tw_ > 1 && b[tw_ - 1] == LASH_TEXT_RECORD_SEP) --inner;
            sp.n_rec = (uint32_t)std::min<uint64_t>(inner + 1, 0xffffffffull);
            sp.reserved = 0;
            tspans_.push_back(sp);
            off_ += (tw_ + 15) / 16 * 16 + 16;
            return;
        }
        const uint64_t n = bs_.size();
        if (n == 0) {  // nothing kept: no span
            rec_start_.resize(rec_first_);
            span_open_ = false;
            return;
        }
        lash_span sp;
        sp.genome = genome_;
        sp.byte_off = off_;
        sp.n_bases = n;
        sp.rec_first = rec_first_;
        sp.n_rec = n_rec_;
        sp.rec_len = 0;
        if (n_rec_ <= 1) {
            rec_start_.resize(rec_first_);  // one record: no table needed
            sp.rec_first = 0;
        } else if (uniform_ && uniform_len_ <= 0xffffffffull) {
            // fixed-length reads: boundaries are arithmetic, no 8 B/read table over PCIe (lash_span.rec_len)
            rec_start_.resize(rec_first_);
            sp.rec_first = 0;
            sp.rec_len = (uint32_t)uniform_len_;
        }
        spans_.push_back(sp);
        off_ += bs_.finalize();
        span_open_ = false;
    }
    // hand the chunk to the GPU; returns the ticket (0 = nothing to push)
    bool submit(Gpu& gpu, uint64_t* ticket) {
        *ticket = 0;
        if ((spans_.empty() && tspans_.empty()) || !gpu.sk) {  // nothing to push, or dry run
            if (!spans_.empty() || !tspans_.empty()) gpu.pushes.fetch_add(1);
            reset();
            return true;
        }
        std::lock_guard<std::mutex> g(gpu.mu);
        const int rc = text_ ? lash_sketch_push_ascii(gpu.sk, buf_, off_, tspans_.data(), (uint32_t)tspans_.size(), ticket)
                             : lash_sketch_push(gpu.sk, buf_, off_, spans_.data(), (uint32_t)spans_.size(),
                                                rec_start_.empty() ? nullptr : rec_start_.data(), rec_start_.size(), ticket);
        if (rc != LASH_OK) {
            if (!gpu.failed.exchange(true)) gpu.err = lash_gpu_last_error();
            return false;
        }
        gpu.pushes.fetch_add(1);
        return true;
    }
    void reset() {
        spans_.clear();
        tspans_.clear();
        rec_start_.clear();
        off_ = 0;
        span_open_ = false;
    }

  private:
    void close_record(uint64_t len) {
        // uniform = every record has the same length except possibly the last one (which may be shorter)
        if (n_rec_ == 0) uniform_len_ = len;
        else if (last_len_ != uniform_len_ || len > uniform_len_) uniform_ = false;
        last_len_ = len;
        rec_start_.push_back(bs_.size());
        ++n_rec_;
    }
    uint8_t* buf_ = nullptr;
    uint64_t cap_ = 0, off_ = 0, block_ = 0;
    BaseStream bs_;
    std::vector<lash_span> spans_;
    std::vector<lash_text_span> tspans_;
    uint64_t tw_ = 0;     // text mode: bytes written in the open span
    uint64_t n_seps_ = 0; // text mode: record separators written in the open span
    bool text_ = false;
    std::vector<uint64_t> rec_start_;
    uint64_t genome_ = 0, rec_first_ = 0, rec_begin_ = 0, uniform_len_ = 0, last_len_ = 0;
    uint32_t n_rec_ = 0;
    bool uniform_ = true, span_open_ = false, pinned_ = true;
};

struct Worker {
    Chunk chunk[2];
    uint64_t ticket[2] = {0, 0};
    int cur = 0;
    uint64_t n_records = 0, n_in = 0, n_kept = 0;

    // submit the current chunk and switch to the other one (waiting until its last copy has left the host)
    bool rotate(Gpu& gpu) {
        if (!chunk[cur].submit(gpu, &ticket[cur])) return false;
        cur ^= 1;
        chunk[cur].set_text_mode(chunk[cur ^ 1].text_mode());
        if (!chunk[cur].allocated() && !chunk[cur].alloc(chunk[cur ^ 1].cap(), gpu.sk != nullptr)) {  // second buffer on first need
            gpu.fail(std::string("pinned staging allocation failed: ") + lash_gpu_last_error());
            return false;
        }
        if (ticket[cur]) {
            std::lock_guard<std::mutex> g(gpu.mu);
            if (lash_sketch_wait_copied(gpu.sk, ticket[cur]) != LASH_OK) {
                if (!gpu.failed.exchange(true)) gpu.err = lash_gpu_last_error();
                return false;
            }
            ticket[cur] = 0;
        }
        chunk[cur].reset();
        return true;
    }

    FastxReader rd;  // one reader per worker: its read buffer is allocated once, not per file

    // plain (uncompressed) FASTA in text mode: read() straight into the pinned chunk.  Returns 0 = not applicable (the
    // generic reader takes over), 1 = done, -1 = error.
    int sketch_fasta_direct(Gpu& gpu, const std::string& path, uint64_t genome, int k, std::string& err) {
        static const bool off = getenv("LASH_FASTA_DIRECT") && !strcmp(getenv("LASH_FASTA_DIRECT"), "0");
        if (off) return 0;
        const int fd = ::open(path.c_str(), O_RDONLY | O_CLOEXEC);
        if (fd < 0) return 0;
        uint8_t m[6] = {0, 0, 0, 0, 0, 0};
        struct stat sb;
        const bool plain_fasta = fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && ::pread(fd, m, 6, 0) >= 1 && m[0] == '>';
        if (!plain_fasta) {   // compressed ('>' is no magic byte of gzip / bzip2 / xz / zstd), FASTQ, empty, a pipe ...
            ::close(fd);
            return 0;
        }
        constexpr size_t kPiece = 1u << 20;   // read + scan granularity: the scan finds the piece in the core's L2
        if (!chunk[cur].can_begin_span() && !rotate(gpu)) { ::close(fd); return -1; }
        chunk[cur].begin_span(genome);
        chunk[cur].begin_record();
        Chunk::FastaScan st;
        std::vector<uint8_t> carry;
        int rc = 1;
        for (;;) {
            if (chunk[cur].room() < 4096) {
                // chunk full: close this part of the record (a header needs no overlap), continue in the other chunk
                if (st.in_header) carry.clear();
                else chunk[cur].split_record(k, carry);
                chunk[cur].end_span();
                if (!rotate(gpu)) { rc = -1; break; }
                chunk[cur].begin_span(genome);
                chunk[cur].begin_record();
                chunk[cur].push_carry(carry);
            }
            const size_t want = (size_t)std::min<uint64_t>(chunk[cur].room(), kPiece);
            const ssize_t got = ::read(fd, chunk[cur].text_ptr(), want);
            if (got < 0) {
                if (errno == EINTR) continue;
                err = "Invalid input file " + path + ": read failed: " + strerror(errno);
                rc = -1;
                break;
            }
            if (got == 0) break;
            n_in += (uint64_t)got;
            chunk[cur].text_commit_fasta((size_t)got, st);
        }
        ::close(fd);
        n_records += st.n_headers;
        if (rc == 1) chunk[cur].end_span();
        return rc;
    }

    bool sketch_file(Gpu& gpu, const std::string& path, uint64_t genome, int k, std::string& err) {
        if (chunk[cur].text_mode()) {
            const int rc = sketch_fasta_direct(gpu, path, genome, k, err);
            if (rc != 0) return rc > 0;
        }
        if (!rd.open(path)) {
            err = "Invalid input file " + path + ": " + rd.err();  // utils.rs:453 expect("Invalid input file")
            return false;
        }
        if (!chunk[cur].can_begin_span() && !rotate(gpu)) return false;
        chunk[cur].begin_span(genome);
        std::vector<uint8_t> carry;
        bool seen_record = false, record_open = false;
        for (;;) {
            FastxReader::Ev e = rd.next();
            if (e.type == FastxReader::kEof) break;
            if (e.type == FastxReader::kError) {
                // utils.rs:453 `.expect("Invalid input file")` only covers a file that cannot be opened or recognised;
                // a record that fails to parse later is skipped (`if let Ok(seqrec) = res`, utils.rs:458) and the file
                // keeps the sketch of what came before.  The same here: warn, keep the records read so far.
                if (!seen_record) {
                    err = "Invalid input file " + path + ": " + rd.err();
                    return false;
                }
                fprintf(stderr, "warning: %s: %s -- the rest of the file is ignored (records so far: %llu)\n", path.c_str(), rd.err().c_str(),
                        (unsigned long long)n_records);
                if (record_open) chunk[cur].end_record(k);
                break;
            }
            if (e.type == FastxReader::kBegin) {
                ++n_records;
                seen_record = record_open = true;
                chunk[cur].begin_record();
            } else if (e.type == FastxReader::kEnd) {
                record_open = false;
                chunk[cur].end_record(k);
            } else {
                n_in += e.n;
                for (size_t off = 0; off < e.n;) {
                    const size_t n = std::min(kFeed, e.n - off);
                    if (chunk[cur].room() < n) {
                        // chunk full in the middle of a record: close this part, continue in the other chunk
                        chunk[cur].split_record(k, carry);
                        chunk[cur].end_span();
                        if (!rotate(gpu)) return false;
                        chunk[cur].begin_span(genome);
                        chunk[cur].begin_record();
                        chunk[cur].push_carry(carry);
                        if (chunk[cur].room() < n) {
                            err = "staging chunk too small";
                            return false;
                        }
                    }
                    n_kept += chunk[cur].append(e.p + off, n);
                    off += n;
                }
            }
        }
        chunk[cur].end_span();
        return true;
    }
};

// Staging chunk size when the caller does not choose: about half of a worker's share of the packed input
// (input bytes / 4), between 1 and 16 MiB -- small jobs do not page-lock memory they will never fill.
uint64_t auto_chunk_bytes(const std::vector<std::string>& files, uint32_t n_workers) {
    uint64_t total = 0;
    for (const auto& f : files) {
        struct stat sb;
        if (stat(f.c_str(), &sb) == 0 && S_ISREG(sb.st_mode)) total += (uint64_t)sb.st_size;
        if (total > (1ull << 36)) break;
    }
    const uint64_t per_worker = total / 4 / std::max(1u, n_workers);
    const uint64_t want = ((per_worker / 2 + (1u << 20) - 1) >> 20) << 20;
    return std::min<uint64_t>(std::max<uint64_t>(want, 1u << 20), kDefaultChunk);
}

}  // namespace

void release_pinned() { pinned_pool().clear(); }

// Who filters and packs: 1 = host (SIMD packer, 0.25 B/base over PCIe), 2 = device (raw text, 1 B/base over PCIe, the
// host only copies).  0 = auto: the host packer when the CPU has the SIMD path (measured on the B200 boxes: 4.3 Gbp/s per
// worker packed against 4.8-7.3 as text, but the text moves 2 B/base through host DRAM against 1.25 and both are bound by
// host memory bandwidth from 4 workers on -- packed 41 Gbp/s, text 25 Gbp/s at 16 workers), the device otherwise (the
// scalar packer does 1 Gbp/s per worker).  LASH_INGEST=packed|ascii overrides.
static std::atomic<int> g_ingest_mode{0};
void set_ingest_mode(int mode) { g_ingest_mode.store(mode); }
static bool use_text_ingest(uint32_t n_workers) {
    int mode = g_ingest_mode.load();
    if (const char* e = getenv("LASH_INGEST")) {
        if (!strcmp(e, "ascii") || !strcmp(e, "text")) mode = 2;
        else if (!strcmp(e, "packed")) mode = 1;
    }
    if (mode == 1) return false;
    if (mode == 2) return true;
    (void)n_workers;
    return !lashhost::pack_has_simd();
}

// Parse + filter + pack every file exactly as sketch_files does, but drop the chunks instead of pushing
// them: the host-side ingest ceiling (no GPU, no sketch -- a measurement aid for bench.py).
Status pack_files_dry(const std::vector<std::string>& files, size_t kmer_length, uint32_t threads, uint64_t chunk_bytes,
                      SketchFilesStats* stats) {
    const auto t0 = std::chrono::steady_clock::now();
    if (kmer_length < 1 || kmer_length > 32) return Status{LASH_E_INVALID, "k-mer length must be 1-32"};
    const uint64_t n_files = files.size();
    SketchFilesStats st{};
    if (n_files) {
        if (chunk_bytes == 0) chunk_bytes = kDefaultChunk;
        uint32_t n_workers = threads ? threads : std::max(1u, std::thread::hardware_concurrency());
        n_workers = (uint32_t)std::min<uint64_t>(n_workers, n_files);
        Gpu gpu;  // sk == nullptr: submit() only counts
        std::vector<Worker> workers(n_workers);
        const bool text = use_text_ingest(n_workers);
        for (auto& w : workers) {
            w.chunk[0].set_text_mode(text);
            if (!w.chunk[0].alloc(chunk_bytes, false)) return Status{LASH_E_NOMEM, "staging allocation failed"};
        }
        std::atomic<uint64_t> next_file{0};
        auto run = [&](Worker& w) {
            for (;;) {
                const uint64_t f = next_file.fetch_add(1);
                if (f >= n_files || gpu.failed.load()) break;
                std::string err;
                if (!w.sketch_file(gpu, files[f], f, (int)kmer_length, err)) {
                    gpu.fail(err);
                    break;
                }
            }
            uint64_t t;
            w.chunk[w.cur].submit(gpu, &t);
        };
        std::vector<std::thread> pool;
        for (uint32_t i = 1; i < n_workers; ++i) pool.emplace_back(run, std::ref(workers[i]));
        run(workers[0]);
        for (auto& t : pool) t.join();
        for (auto& w : workers) {
            st.n_records += w.n_records;
            st.n_bases_in += w.n_in;
            st.n_bases_kept += w.n_kept;
            for (auto& c : w.chunk) c.release();
        }
        st.n_pushes = gpu.pushes.load();
        if (gpu.failed.load()) return Status{LASH_HOST_E_FORMAT, gpu.err};
    }
    st.seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats) *stats = st;
    return Status{};
}

Status sketch_files_impl(lash_ctx* ctx, int algo, std::optional<uint32_t> precision, const std::vector<std::string>& files,
                         size_t kmer_length, const std::string* output_name, uint32_t threads, uint64_t seed, uint64_t chunk_bytes,
                         void* regs_out, SketchFilesStats* stats) {
    const auto t0 = std::chrono::steady_clock::now();
    if (!ctx) return Status{LASH_E_INVALID, "sketch_files: NULL ctx"};
    if (kmer_length < 1 || kmer_length > 32) return Status{LASH_E_INVALID, "k-mer length must be 1-32"};  // utils.rs:500-502
    int p = 14;
    if (algo == LASH_ALGO_HLL || algo == LASH_ALGO_ULL) {
        if (!precision) return Status{LASH_E_INVALID, algo == LASH_ALGO_HLL ? "HLL needs precision" : "ULL needs precision"};  // utils.rs:408,423
        p = (int)*precision;
    }
    const size_t rb = lash_sketch_reg_bytes(algo, p);
    if (rb == 0) return Status{LASH_E_INVALID, algo == LASH_ALGO_ULL ? "failed to create ULL" : "bad algorithm / precision"};
    const uint64_t n_files = files.size();
    std::vector<uint8_t> regs_local;
    uint8_t* regs = static_cast<uint8_t*>(regs_out);
    if (!regs) {
        regs_local.resize(rb * n_files);
        regs = regs_local.data();
    }
    SketchFilesStats st{};
    if (n_files) {
        uint32_t n_workers = threads ? threads : std::max(1u, std::thread::hardware_concurrency());
        n_workers = (uint32_t)std::min<uint64_t>(n_workers, n_files);
        if (chunk_bytes == 0) chunk_bytes = auto_chunk_bytes(files, n_workers);
        chunk_bytes = std::max<uint64_t>(chunk_bytes, 1u << 20) / 16 * 16;

        Gpu gpu;
        if (lash_sketch_open(ctx, algo, p, (int)kmer_length, seed, n_files, &gpu.sk) != LASH_OK)
            return Status{LASH_E_CUDA, lash_gpu_last_error()};
        std::vector<Worker> workers(n_workers);
        bool alloc_ok = true;
        const bool text = use_text_ingest(n_workers);
        if (text) chunk_bytes = std::min<uint64_t>(chunk_bytes * 4, kDefaultChunk);  // same bases per chunk, bounded
        for (auto& w : workers) {
            w.chunk[0].set_text_mode(text);
            alloc_ok = alloc_ok && w.chunk[0].alloc(chunk_bytes);  // chunk[1]: on first rotate
        }
        std::atomic<uint64_t> next_file{0};
        const auto t_open = std::chrono::steady_clock::now();
        st.seconds_open = std::chrono::duration<double>(t_open - t0).count();
        if (alloc_ok) {
            auto run = [&](Worker& w) {
                for (;;) {
                    const uint64_t f = next_file.fetch_add(1);
                    if (f >= n_files || gpu.failed.load()) break;
                    std::string err;
                    if (!w.sketch_file(gpu, files[f], f, (int)kmer_length, err)) {
                        if (!err.empty()) gpu.fail(err);
                        break;
                    }
                }
                if (!gpu.failed.load()) {
                    uint64_t t;
                    w.chunk[w.cur].submit(gpu, &t);
                }
            };
            std::vector<std::thread> pool;
            for (uint32_t i = 1; i < n_workers; ++i) pool.emplace_back(run, std::ref(workers[i]));
            run(workers[0]);
            for (auto& t : pool) t.join();
        } else {
            gpu.fail(std::string("pinned staging allocation failed: ") + lash_gpu_last_error());
        }
        const auto t_joined = std::chrono::steady_clock::now();
        st.seconds_workers = std::chrono::duration<double>(t_joined - t_open).count();
        int rc = LASH_OK;
        if (!gpu.failed.load()) rc = lash_sketch_fetch(gpu.sk, 0, n_files, regs);
        else lash_sketch_sync(gpu.sk);  // chunks must not be freed under an in-flight copy
        std::string gerr = rc != LASH_OK ? lash_gpu_last_error() : "";
        uint64_t launches = 0;
        lash_sketch_stats(gpu.sk, &st.gpu_kernel_ms, &launches);
        for (auto& w : workers) {
            st.n_records += w.n_records;
            st.n_bases_in += w.n_in;
            st.n_bases_kept += w.n_kept;
            for (auto& c : w.chunk) c.release();
        }
        st.n_pushes = gpu.pushes.load();
        lash_sketch_close(gpu.sk);
        st.seconds_drain = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_joined).count();
        if (gpu.failed.load()) {
            const bool input = gpu.err.rfind("Invalid input file", 0) == 0;
            return Status{input ? LASH_HOST_E_FORMAT : LASH_E_CUDA, gpu.err};
        }
        if (rc != LASH_OK) return Status{rc, gerr};
    }
    if (output_name) {
        // write sketches (utils.rs:566-575) and names (utils.rs:577-580)
        std::string err;
        if (!lashhost::write_sketches(*output_name + "_sketches.bin", algo, p, regs, n_files, (int)threads, err))
            return Status{LASH_HOST_E_IO, err};
        if (!lashhost::write_file(*output_name + "_files.json", lashhost::json_pretty_string_array(files), err))
            return Status{LASH_HOST_E_IO, err};
    }
    st.seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stats) *stats = st;
    return Status{};
}

}  // namespace lash
