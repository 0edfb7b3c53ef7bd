// FASTA / FASTQ reader with the behaviour of needletail's parse_fastx_file + reader.next() as the
// reference uses it (src/utils.rs:453-458): format chosen by the first byte ('>' / '@'), FASTA
// sequences may span lines, FASTQ records are the strict four-line form, seq() has line breaks
// removed, an empty or unrecognisable file is an error ("Invalid input file").
//
// The reader works at event level so the sketching path can stream sequence bytes straight into
// the packer without assembling records; next_record() assembles them for everyone else.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "io.hpp"

namespace lashhost {

class FastxReader {
  public:
    enum Event { kBegin, kSeq, kEnd, kEof, kError };
    struct Ev {
        Event type;
        const uint8_t* p;  // kBegin: record id line (without '>'/'@' and line end); kSeq: raw sequence bytes
        size_t n;          //         (FASTA: may contain '\n' / '\r'; they are not bases)
    };

    FastxReader() = default;
    ~FastxReader();
    FastxReader(const FastxReader&) = delete;
    FastxReader& operator=(const FastxReader&) = delete;
    bool open(const std::string& path, size_t buf_bytes = 1u << 18);
    Ev next();
    // 1 record, 0 end of file, -1 error
    int next_record(std::string& id, std::string& seq);
    const std::string& err() const { return err_; }

  private:
    bool fill();  // compacts [pos_, end_) to the front and reads more; false when nothing was added
    Ev fail(const std::string& m) {
        err_ = m;
        state_ = kStFailed;
        return Ev{kError, nullptr, 0};
    }
    enum State { kStStart, kStFaHeader, kStFaSeq, kStFqRecord, kStFqSeq, kStFqEnd, kStDone, kStFailed };
    std::unique_ptr<ByteSource> src_;
    std::vector<uint8_t> buf_;
    const uint8_t* map_ = nullptr;  // whole file mapped (uncompressed input); buf_ / src_ unused then
    size_t map_len_ = 0;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false;
    bool at_line_start_ = true;  // FASTA: the byte before pos_ was '\n' (or pos_ is the start of the data)
    bool fa_open_ = false;       // a FASTA record has begun and kEnd is still owed
    State state_ = kStStart;
    size_t fq_seq_off_ = 0, fq_seq_len_ = 0, fq_next_ = 0;
    std::string err_;
};

}  // namespace lashhost
