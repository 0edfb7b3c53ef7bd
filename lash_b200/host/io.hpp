// Byte sources for the record reader: plain files and gzip / bzip2 / xz / zstd streams, sniffed by
// magic bytes like needletail's parse_fastx_file (reference src/utils.rs:453).  zlib is linked; the
// other three are bound at run time (dlopen) because this image ships their .so without headers --
// the prototypes below are the public, ABI-stable ones of zstd 1.x, bzip2 1.0 and liblzma 5.x.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace lashhost {

class ByteSource {
  public:
    virtual ~ByteSource() = default;
    // up to n bytes into dst; 0 = end of stream; -1 = error (message in err())
    virtual long read(uint8_t* dst, size_t n) = 0;
    const std::string& err() const { return err_; }

  protected:
    std::string err_;
};

// Opens `path`, sniffs the compression, returns a decoding source (nullptr + err on failure).
std::unique_ptr<ByteSource> open_source(const std::string& path, std::string& err);

// Maps an UNCOMPRESSED regular file read-only (false: compressed, empty, not mappable -> use open_source).
bool map_plain_file(const std::string& path, const uint8_t** data, size_t* size);
void unmap_file(const uint8_t* data, size_t size);

// ---- zstd (streaming, run-time bound) --------------------------------------------------------------
struct ZstdApi;
const ZstdApi* zstd_api(std::string& err);  // nullptr when libzstd.so.1 cannot be loaded

// One zstd frame written incrementally (what zstd::stream::Encoder::new(writer, 3) + multithread(n)
// + finish() produce, utils.rs:567-575).
class ZstdFileWriter {
  public:
    ZstdFileWriter() = default;
    ~ZstdFileWriter();
    bool open(const std::string& path, int level, int workers, std::string& err);
    bool write(const void* p, size_t n, std::string& err);
    bool finish(std::string& err);

  private:
    bool drain(int end_op, const void* p, size_t n, std::string& err);
    const ZstdApi* api_ = nullptr;
    void* cctx_ = nullptr;
    int fd_ = -1;
    std::vector<uint8_t> out_;
};

// Whole-file helpers
bool read_file(const std::string& path, std::string& out, std::string& err);
bool write_file(const std::string& path, const std::string& data, std::string& err);

}  // namespace lashhost
