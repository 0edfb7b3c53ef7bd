#include "fastx.hpp"

#include <cstdlib>
#include <cstring>

namespace lashhost {

FastxReader::~FastxReader() { unmap_file(map_, map_len_); }

bool FastxReader::open(const std::string& path, size_t buf_bytes) {
    err_.clear();
    unmap_file(map_, map_len_);
    map_ = nullptr;
    map_len_ = 0;
    pos_ = end_ = 0;
    eof_ = false;
    // uncompressed regular files are parsed in place (page cache mapped read-only, no copy)
    static const bool use_mmap = getenv("LASH_FASTX_MMAP") != nullptr;
    if (use_mmap && map_plain_file(path, &map_, &map_len_)) {
        end_ = map_len_;
        eof_ = true;
    } else {
        src_ = open_source(path, err_);
        if (!src_) {
            state_ = kStFailed;
            return false;
        }
        if (buf_.size() < 4096) buf_.resize(buf_bytes < 4096 ? 4096 : buf_bytes);   // kept across open() calls of one reader
    }
    at_line_start_ = true;
    fa_open_ = false;
    state_ = kStStart;
    return true;
}

bool FastxReader::fill() {
    if (eof_) return false;
    if (pos_ > 0) {
        memmove(buf_.data(), buf_.data() + pos_, end_ - pos_);
        end_ -= pos_;
        pos_ = 0;
    }
    if (end_ == buf_.size()) buf_.resize(buf_.size() * 2);  // one line / one FASTQ record longer than the buffer
    const long r = src_->read(buf_.data() + end_, buf_.size() - end_);
    if (r < 0) {
        err_ = src_->err();
        eof_ = true;
        return false;
    }
    if (r == 0) {
        eof_ = true;
        return false;
    }
    end_ += (size_t)r;
    return true;
}

static inline size_t strip_cr(const uint8_t* b, size_t from, size_t to) { return (to > from && b[to - 1] == '\r') ? to - 1 : to; }

FastxReader::Ev FastxReader::next() {
    for (;;) {
        const uint8_t* b = map_ ? map_ : buf_.data();
        switch (state_) {
            case kStFailed:
                return Ev{kError, nullptr, 0};
            case kStDone:
                return Ev{kEof, nullptr, 0};
            case kStStart: {
                if (pos_ == end_ && !fill()) return fail(err_.empty() ? "Invalid input file: empty file" : err_);
                b = map_ ? map_ : buf_.data();
                if (b[pos_] == '>') state_ = kStFaHeader;
                else if (b[pos_] == '@') state_ = kStFqRecord;
                else return fail("Invalid input file: first byte is neither '>' (FASTA) nor '@' (FASTQ)");
                break;
            }
            case kStFaHeader: {
                const uint8_t* nl = static_cast<const uint8_t*>(memchr(b + pos_, '\n', end_ - pos_));
                size_t line_end, next;
                if (!nl) {
                    if (fill()) continue;
                    if (!err_.empty()) return fail(err_);
                    line_end = next = end_;  // header without a line end at EOF: a record with an empty sequence
                } else {
                    line_end = (size_t)(nl - b);
                    next = line_end + 1;
                }
                const size_t from = pos_ + 1;
                const size_t to = strip_cr(b, from, line_end);
                pos_ = next;
                at_line_start_ = true;
                fa_open_ = true;
                state_ = kStFaSeq;
                return Ev{kBegin, b + from, to - from};
            }
            case kStFaSeq: {
                if (pos_ == end_) {
                    if (fill()) continue;
                    if (!err_.empty()) return fail(err_);
                    state_ = kStDone;
                    fa_open_ = false;
                    return Ev{kEnd, nullptr, 0};
                }
                size_t stop = end_;
                for (size_t i = pos_; i < end_;) {
                    const uint8_t* g = static_cast<const uint8_t*>(memchr(b + i, '>', end_ - i));
                    if (!g) break;
                    const size_t j = (size_t)(g - b);
                    const bool line_start = (j == pos_) ? at_line_start_ : b[j - 1] == '\n';
                    if (line_start) {
                        stop = j;
                        break;
                    }
                    i = j + 1;  // a '>' inside a sequence line is just a byte filter_out_n deletes
                }
                if (stop > pos_) {
                    Ev e{kSeq, b + pos_, stop - pos_};
                    at_line_start_ = b[stop - 1] == '\n';
                    pos_ = stop;
                    return e;
                }
                state_ = kStFaHeader;
                fa_open_ = false;
                return Ev{kEnd, nullptr, 0};
            }
            case kStFqRecord: {
                if (pos_ == end_) {
                    if (fill()) continue;
                    if (!err_.empty()) return fail(err_);
                    state_ = kStDone;
                    return Ev{kEof, nullptr, 0};
                }
                size_t le[4], p = pos_;
                int got = 0;
                while (got < 4) {
                    const uint8_t* nl = static_cast<const uint8_t*>(memchr(b + p, '\n', end_ - p));
                    if (!nl) break;
                    le[got++] = (size_t)(nl - b);
                    p = le[got - 1] + 1;
                }
                if (got < 4) {
                    if (!eof_) {
                        if (!fill() && !err_.empty()) return fail(err_);
                        continue;
                    }
                    if (got == 3 && p < end_) {  // last quality line without a line end
                        le[3] = end_;
                        p = end_;
                    } else {
                        bool blank = true;
                        for (size_t i = pos_; i < end_; ++i) blank = blank && (b[i] == '\n' || b[i] == '\r');
                        if (!blank) return fail("Invalid input file: truncated FASTQ record");
                        state_ = kStDone;
                        return Ev{kEof, nullptr, 0};
                    }
                }
                if (b[pos_] != '@') return fail("Invalid input file: FASTQ record does not start with '@'");
                const size_t id0 = pos_ + 1, id1 = strip_cr(b, id0, le[0]);
                const size_t s0 = le[0] + 1, s1 = strip_cr(b, s0, le[1]);
                if (le[1] + 1 >= end_ || b[le[1] + 1] != '+') return fail("Invalid input file: FASTQ separator line does not start with '+'");
                const size_t q0 = le[2] + 1, q1 = strip_cr(b, q0, le[3]);
                if (q1 - q0 != s1 - s0) return fail("Invalid input file: FASTQ sequence and quality lengths differ");
                fq_seq_off_ = s0;
                fq_seq_len_ = s1 - s0;
                fq_next_ = p;
                state_ = kStFqSeq;
                return Ev{kBegin, b + id0, id1 - id0};
            }
            case kStFqSeq:
                state_ = kStFqEnd;
                return Ev{kSeq, b + fq_seq_off_, fq_seq_len_};
            case kStFqEnd:
                pos_ = fq_next_;
                state_ = kStFqRecord;
                return Ev{kEnd, nullptr, 0};
        }
    }
}

int FastxReader::next_record(std::string& id, std::string& seq) {
    Ev e = next();
    if (e.type == kEof) return 0;
    if (e.type != kBegin) return -1;
    id.assign(reinterpret_cast<const char*>(e.p), e.n);
    seq.clear();
    for (;;) {
        e = next();
        if (e.type == kEnd) return 1;
        if (e.type != kSeq) return -1;
        for (size_t i = 0; i < e.n; ++i)
            if (e.p[i] != '\n' && e.p[i] != '\r') seq.push_back((char)e.p[i]);
    }
}

}  // namespace lashhost
