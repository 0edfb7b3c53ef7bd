#include "io.hpp"

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cerrno>
#include <cstring>

namespace lashhost {

// ------------------------------------------------------------------------------------------------
// raw file with a small read-ahead buffer (lets the opener peek at the magic bytes)
// ------------------------------------------------------------------------------------------------
namespace {

class RawFile : public ByteSource {
  public:
    ~RawFile() override {
        if (fd_ >= 0) ::close(fd_);
    }
    bool open(const std::string& path) {
        fd_ = ::open(path.c_str(), O_RDONLY | O_CLOEXEC);
        if (fd_ < 0) {
            err_ = "cannot open " + path + ": " + strerror(errno);
            return false;
        }
#ifdef POSIX_FADV_SEQUENTIAL
        posix_fadvise(fd_, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
        return true;
    }
    // first bytes of the file without consuming them
    size_t peek(uint8_t* dst, size_t n) {
        while (head_.size() < n) {
            uint8_t tmp[64];
            long r = ::read(fd_, tmp, sizeof(tmp));
            if (r <= 0) break;
            head_.insert(head_.end(), tmp, tmp + r);
        }
        const size_t m = head_.size() < n ? head_.size() : n;
        memcpy(dst, head_.data(), m);
        return m;
    }
    long read(uint8_t* dst, size_t n) override {
        if (head_pos_ < head_.size()) {
            const size_t m = std::min(n, head_.size() - head_pos_);
            memcpy(dst, head_.data() + head_pos_, m);
            head_pos_ += m;
            return (long)m;
        }
        for (;;) {
            long r = ::read(fd_, dst, n);
            if (r < 0 && errno == EINTR) continue;
            if (r < 0) err_ = std::string("read failed: ") + strerror(errno);
            return r < 0 ? -1 : r;
        }
    }

  private:
    int fd_ = -1;
    std::vector<uint8_t> head_;
    size_t head_pos_ = 0;
};

constexpr size_t kInBuf = 1 << 18;

// Common shape of the four decoders: pull compressed bytes from `raw_` into in_, produce into dst.
class Decoder : public ByteSource {
  public:
    explicit Decoder(std::unique_ptr<RawFile> raw) : raw_(std::move(raw)), in_(kInBuf) {}

  protected:
    // returns false at raw EOF (avail stays 0) or on error (err_ set)
    bool refill() {
        long r = raw_->read(in_.data(), in_.size());
        if (r < 0) {
            err_ = raw_->err();
            return false;
        }
        in_pos_ = 0;
        in_len_ = (size_t)r;
        if (r == 0) raw_eof_ = true;
        return r > 0;
    }
    std::unique_ptr<RawFile> raw_;
    std::vector<uint8_t> in_;
    size_t in_pos_ = 0, in_len_ = 0;
    bool raw_eof_ = false;
};

// ---- gzip (multi-member, like flate2's MultiGzDecoder that needletail uses) ------------------------
class GzSource : public Decoder {
  public:
    explicit GzSource(std::unique_ptr<RawFile> raw) : Decoder(std::move(raw)) {
        memset(&z_, 0, sizeof(z_));
        ok_ = inflateInit2(&z_, 15 + 32) == Z_OK;
        if (!ok_) err_ = "inflateInit2 failed";
    }
    ~GzSource() override {
        if (ok_) inflateEnd(&z_);
    }
    long read(uint8_t* dst, size_t n) override {
        if (!ok_) return -1;
        if (done_) return 0;
        z_.next_out = dst;
        z_.avail_out = (uInt)std::min<size_t>(n, 1u << 30);
        const uInt want = z_.avail_out;
        while (z_.avail_out == want) {
            if (in_pos_ == in_len_) {
                if (!refill()) {
                    if (!err_.empty()) return -1;
                    if (member_open_) {
                        err_ = "truncated gzip stream";
                        return -1;
                    }
                    done_ = true;
                    break;
                }
            }
            z_.next_in = in_.data() + in_pos_;
            z_.avail_in = (uInt)(in_len_ - in_pos_);
            member_open_ = true;
            const int rc = inflate(&z_, Z_NO_FLUSH);
            in_pos_ = in_len_ - z_.avail_in;
            if (rc == Z_STREAM_END) {
                member_open_ = false;
                inflateReset(&z_);  // next member, if any bytes follow
            } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
                err_ = std::string("gzip: ") + (z_.msg ? z_.msg : "inflate error");
                return -1;
            }
        }
        return (long)(want - z_.avail_out);
    }

  private:
    z_stream z_;
    bool ok_ = false, done_ = false, member_open_ = false;
};

// ---- run-time bound libraries ------------------------------------------------------------------------
void* load_lib(const char* const* names, std::string& err) {
    for (const char* const* n = names; *n; ++n) {
        if (void* h = dlopen(*n, RTLD_NOW | RTLD_LOCAL)) return h;
    }
    err = std::string("cannot load ") + names[0] + ": " + (dlerror() ? dlerror() : "not found");
    return nullptr;
}
template <class F>
bool bind(void* h, const char* name, F& fn, std::string& err) {
    fn = reinterpret_cast<F>(dlsym(h, name));
    if (!fn) err = std::string("symbol missing: ") + name;
    return fn != nullptr;
}

}  // namespace

struct ZstdInBuffer { const void* src; size_t size; size_t pos; };
struct ZstdOutBuffer { void* dst; size_t size; size_t pos; };
struct ZstdApi {
    unsigned (*isError)(size_t);
    const char* (*getErrorName)(size_t);
    void* (*createDCtx)();
    size_t (*freeDCtx)(void*);
    size_t (*decompressStream)(void*, ZstdOutBuffer*, ZstdInBuffer*);
    void* (*createCCtx)();
    size_t (*freeCCtx)(void*);
    size_t (*CCtx_setParameter)(void*, int, int);
    size_t (*compressStream2)(void*, ZstdOutBuffer*, ZstdInBuffer*, int);
};
const ZstdApi* zstd_api(std::string& err) {
    static ZstdApi api;
    static std::string load_err;
    static const bool ok = [] {
        static const char* const names[] = {"libzstd.so.1", "libzstd.so", nullptr};
        void* h = load_lib(names, load_err);
        if (!h) return false;
        return bind(h, "ZSTD_isError", api.isError, load_err) && bind(h, "ZSTD_getErrorName", api.getErrorName, load_err) &&
               bind(h, "ZSTD_createDCtx", api.createDCtx, load_err) && bind(h, "ZSTD_freeDCtx", api.freeDCtx, load_err) &&
               bind(h, "ZSTD_decompressStream", api.decompressStream, load_err) &&
               bind(h, "ZSTD_createCCtx", api.createCCtx, load_err) && bind(h, "ZSTD_freeCCtx", api.freeCCtx, load_err) &&
               bind(h, "ZSTD_CCtx_setParameter", api.CCtx_setParameter, load_err) &&
               bind(h, "ZSTD_compressStream2", api.compressStream2, load_err);
    }();
    if (!ok) err = load_err;
    return ok ? &api : nullptr;
}

namespace {

class ZstdSource : public Decoder {
  public:
    ZstdSource(std::unique_ptr<RawFile> raw, const ZstdApi* api) : Decoder(std::move(raw)), api_(api) { d_ = api_->createDCtx(); }
    ~ZstdSource() override {
        if (d_) api_->freeDCtx(d_);
    }
    long read(uint8_t* dst, size_t n) override {
        if (!d_) {
            err_ = "ZSTD_createDCtx failed";
            return -1;
        }
        ZstdOutBuffer out{dst, n, 0};
        while (out.pos == 0) {
            if (in_pos_ == in_len_ && !raw_eof_) {
                if (!refill() && !err_.empty()) return -1;
            }
            if (in_pos_ == in_len_ && raw_eof_) {
                // flush whatever the decoder still buffers; a frame cut short is an error
                ZstdInBuffer none{nullptr, 0, 0};
                const size_t rc = api_->decompressStream(d_, &out, &none);
                if (api_->isError(rc)) {
                    err_ = std::string("zstd: ") + api_->getErrorName(rc);
                    return -1;
                }
                if (out.pos == 0) {
                    if (last_rc_ != 0) {
                        err_ = "truncated zstd stream";
                        return -1;
                    }
                    return 0;
                }
                break;
            }
            ZstdInBuffer in{in_.data() + in_pos_, in_len_ - in_pos_, 0};
            const size_t rc = api_->decompressStream(d_, &out, &in);
            in_pos_ += in.pos;
            if (api_->isError(rc)) {
                err_ = std::string("zstd: ") + api_->getErrorName(rc);
                return -1;
            }
            last_rc_ = rc;
        }
        return (long)out.pos;
    }

  private:
    const ZstdApi* api_;
    void* d_ = nullptr;
    size_t last_rc_ = 0;
};

// ---- bzip2 ------------------------------------------------------------------------------------------
struct BzStream {
    char* next_in; unsigned avail_in; unsigned total_in_lo32; unsigned total_in_hi32;
    char* next_out; unsigned avail_out; unsigned total_out_lo32; unsigned total_out_hi32;
    void* state; void* (*bzalloc)(void*, int, int); void (*bzfree)(void*, void*); void* opaque;
};
struct BzApi {
    int (*init)(BzStream*, int, int);
    int (*decompress)(BzStream*);
    int (*end)(BzStream*);
};
const BzApi* bz_api(std::string& err) {
    static BzApi api;
    static std::string load_err;
    static const bool ok = [] {
        static const char* const names[] = {"libbz2.so.1.0", "libbz2.so.1", "libbz2.so", nullptr};
        void* h = load_lib(names, load_err);
        if (!h) return false;
        return bind(h, "BZ2_bzDecompressInit", api.init, load_err) && bind(h, "BZ2_bzDecompress", api.decompress, load_err) &&
               bind(h, "BZ2_bzDecompressEnd", api.end, load_err);
    }();
    if (!ok) err = load_err;
    return ok ? &api : nullptr;
}
class BzSource : public Decoder {
  public:
    BzSource(std::unique_ptr<RawFile> raw, const BzApi* api) : Decoder(std::move(raw)), api_(api) {
        memset(&s_, 0, sizeof(s_));
        open_ = api_->init(&s_, 0, 0) == 0;
    }
    ~BzSource() override {
        if (open_) api_->end(&s_);
    }
    long read(uint8_t* dst, size_t n) override {
        if (done_) return 0;
        s_.next_out = reinterpret_cast<char*>(dst);
        s_.avail_out = (unsigned)std::min<size_t>(n, 1u << 30);
        const unsigned want = s_.avail_out;
        while (s_.avail_out == want) {
            if (in_pos_ == in_len_) {
                if (!refill()) {
                    if (!err_.empty()) return -1;
                    if (open_ && mid_stream_) {
                        err_ = "truncated bzip2 stream";
                        return -1;
                    }
                    done_ = true;
                    break;
                }
            }
            if (!open_) {  // another stream follows (pbzip2-style concatenation)
                memset(&s_, 0, sizeof(s_));
                open_ = api_->init(&s_, 0, 0) == 0;
                s_.next_out = reinterpret_cast<char*>(dst);
                s_.avail_out = want;
            }
            s_.next_in = reinterpret_cast<char*>(in_.data() + in_pos_);
            s_.avail_in = (unsigned)(in_len_ - in_pos_);
            mid_stream_ = true;
            const int rc = api_->decompress(&s_);
            in_pos_ = in_len_ - s_.avail_in;
            if (rc == 4 /* BZ_STREAM_END */) {
                api_->end(&s_);
                open_ = false;
                mid_stream_ = false;
                if (s_.avail_out != want) break;  // hand out what this stream produced first
            } else if (rc != 0) {
                err_ = "bzip2: decompress error " + std::to_string(rc);
                return -1;
            }
        }
        return (long)(want - s_.avail_out);
    }

  private:
    const BzApi* api_;
    BzStream s_;
    bool open_ = false, done_ = false, mid_stream_ = false;
};

// ---- xz ---------------------------------------------------------------------------------------------
struct LzmaStream {  // liblzma 5.x lzma_stream (public layout), padded generously
    const uint8_t* next_in; size_t avail_in; uint64_t total_in;
    uint8_t* next_out; size_t avail_out; uint64_t total_out;
    const void* allocator; void* internal;
    void* reserved_ptr[4]; uint64_t reserved_int[2]; size_t reserved_sz[2]; int reserved_enum[2];
    uint64_t pad[8];
};
struct LzmaApi {
    int (*stream_decoder)(LzmaStream*, uint64_t, uint32_t);
    int (*code)(LzmaStream*, int);
    void (*end)(LzmaStream*);
};
const LzmaApi* lzma_api(std::string& err) {
    static LzmaApi api;
    static std::string load_err;
    static const bool ok = [] {
        static const char* const names[] = {"liblzma.so.5", "liblzma.so", nullptr};
        void* h = load_lib(names, load_err);
        if (!h) return false;
        return bind(h, "lzma_stream_decoder", api.stream_decoder, load_err) && bind(h, "lzma_code", api.code, load_err) &&
               bind(h, "lzma_end", api.end, load_err);
    }();
    if (!ok) err = load_err;
    return ok ? &api : nullptr;
}
class XzSource : public Decoder {
  public:
    XzSource(std::unique_ptr<RawFile> raw, const LzmaApi* api) : Decoder(std::move(raw)), api_(api) {
        memset(&s_, 0, sizeof(s_));
        open_ = api_->stream_decoder(&s_, UINT64_MAX, 0x08 /* LZMA_CONCATENATED */) == 0;
    }
    ~XzSource() override {
        if (open_) api_->end(&s_);
    }
    long read(uint8_t* dst, size_t n) override {
        if (!open_) {
            err_ = "lzma_stream_decoder failed";
            return -1;
        }
        if (done_) return 0;
        s_.next_out = dst;
        s_.avail_out = n;
        while (s_.avail_out == n) {
            if (in_pos_ == in_len_ && !raw_eof_) {
                if (!refill() && !err_.empty()) return -1;
            }
            s_.next_in = in_.data() + in_pos_;
            s_.avail_in = in_len_ - in_pos_;
            const int rc = api_->code(&s_, raw_eof_ ? 3 /* LZMA_FINISH */ : 0 /* LZMA_RUN */);
            in_pos_ = in_len_ - s_.avail_in;
            if (rc == 1 /* LZMA_STREAM_END */) {
                done_ = true;
                break;
            }
            if (rc != 0) {
                err_ = "xz: decode error " + std::to_string(rc);
                return -1;
            }
        }
        return (long)(n - s_.avail_out);
    }

  private:
    const LzmaApi* api_;
    LzmaStream s_;
    bool open_ = false, done_ = false;
};

}  // namespace

std::unique_ptr<ByteSource> open_source(const std::string& path, std::string& err) {
    auto raw = std::make_unique<RawFile>();
    if (!raw->open(path)) {
        err = raw->err();
        return nullptr;
    }
    uint8_t m[6] = {0, 0, 0, 0, 0, 0};
    const size_t got = raw->peek(m, 6);
    if (got >= 2 && m[0] == 0x1f && m[1] == 0x8b) return std::make_unique<GzSource>(std::move(raw));
    if (got >= 3 && m[0] == 'B' && m[1] == 'Z' && m[2] == 'h') {
        const BzApi* api = bz_api(err);
        if (!api) return nullptr;
        return std::make_unique<BzSource>(std::move(raw), api);
    }
    if (got >= 6 && m[0] == 0xfd && m[1] == '7' && m[2] == 'z' && m[3] == 'X' && m[4] == 'Z' && m[5] == 0) {
        const LzmaApi* api = lzma_api(err);
        if (!api) return nullptr;
        return std::make_unique<XzSource>(std::move(raw), api);
    }
    if (got >= 4 && m[0] == 0x28 && m[1] == 0xb5 && m[2] == 0x2f && m[3] == 0xfd) {
        const ZstdApi* api = zstd_api(err);
        if (!api) return nullptr;
        return std::make_unique<ZstdSource>(std::move(raw), api);
    }
    return raw;
}

bool map_plain_file(const std::string& path, const uint8_t** data, size_t* size) {
    const int fd = ::open(path.c_str(), O_RDONLY | O_CLOEXEC);
    if (fd < 0) return false;
    struct stat sb;
    uint8_t m[6] = {0, 0, 0, 0, 0, 0};
    bool ok = fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0 && ::pread(fd, m, 6, 0) > 0;
    const bool compressed = (m[0] == 0x1f && m[1] == 0x8b) || (m[0] == 'B' && m[1] == 'Z' && m[2] == 'h') ||
                            (m[0] == 0xfd && m[1] == '7' && m[2] == 'z' && m[3] == 'X' && m[4] == 'Z' && m[5] == 0) ||
                            (m[0] == 0x28 && m[1] == 0xb5 && m[2] == 0x2f && m[3] == 0xfd);
    ok = ok && !compressed;
    void* p = MAP_FAILED;
    if (ok) {
        const size_t len = (size_t)sb.st_size;
        p = mmap(nullptr, len, PROT_READ, MAP_PRIVATE | (len <= (256u << 20) ? MAP_POPULATE : 0), fd, 0);
        if (p != MAP_FAILED) {
            madvise(p, len, MADV_SEQUENTIAL);
            *data = static_cast<const uint8_t*>(p);
            *size = len;
        }
    }
    ::close(fd);
    return ok && p != MAP_FAILED;
}
void unmap_file(const uint8_t* data, size_t size) {
    if (data) munmap(const_cast<uint8_t*>(data), size);
}

// ------------------------------------------------------------------------------------------------
// zstd writer
// ------------------------------------------------------------------------------------------------
ZstdFileWriter::~ZstdFileWriter() {
    if (cctx_ && api_) api_->freeCCtx(cctx_);
    if (fd_ >= 0) ::close(fd_);
}
bool ZstdFileWriter::open(const std::string& path, int level, int workers, std::string& err) {
    api_ = zstd_api(err);
    if (!api_) return false;
    fd_ = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
    if (fd_ < 0) {
        err = "cannot create " + path + ": " + strerror(errno);
        return false;
    }
    cctx_ = api_->createCCtx();
    if (!cctx_) {
        err = "ZSTD_createCCtx failed";
        return false;
    }
    api_->CCtx_setParameter(cctx_, 100 /* ZSTD_c_compressionLevel */, level);
    // Encoder::multithread(threads) (utils.rs:569); a library built without ZSTD_MULTITHREAD rejects it, which is fine
    if (workers > 1) api_->CCtx_setParameter(cctx_, 400 /* ZSTD_c_nbWorkers */, workers);
    out_.resize(1 << 20);
    return true;
}
bool ZstdFileWriter::drain(int end_op, const void* p, size_t n, std::string& err) {
    ZstdInBuffer in{p, n, 0};
    for (;;) {
        ZstdOutBuffer out{out_.data(), out_.size(), 0};
        const size_t rc = api_->compressStream2(cctx_, &out, &in, end_op);
        if (api_->isError(rc)) {
            err = std::string("zstd: ") + api_->getErrorName(rc);
            return false;
        }
        size_t off = 0;
        while (off < out.pos) {
            long w = ::write(fd_, out_.data() + off, out.pos - off);
            if (w < 0 && errno == EINTR) continue;
            if (w < 0) {
                err = std::string("write failed: ") + strerror(errno);
                return false;
            }
            off += (size_t)w;
        }
        if (end_op == 0 /* continue */ ? in.pos == in.size : rc == 0) return true;
    }
}
bool ZstdFileWriter::write(const void* p, size_t n, std::string& err) { return n == 0 || drain(0, p, n, err); }
bool ZstdFileWriter::finish(std::string& err) {
    const bool ok = drain(2 /* ZSTD_e_end */, nullptr, 0, err);
    if (fd_ >= 0) {
        if (::close(fd_) != 0 && ok) {
            err = std::string("close failed: ") + strerror(errno);
            fd_ = -1;
            return false;
        }
        fd_ = -1;
    }
    return ok;
}

bool read_file(const std::string& path, std::string& out, std::string& err) {
    auto src = open_source(path, err);
    if (!src) return false;
    out.clear();
    std::vector<uint8_t> buf(1 << 16);
    for (;;) {
        long r = src->read(buf.data(), buf.size());
        if (r < 0) {
            err = src->err();
            return false;
        }
        if (r == 0) return true;
        out.append(reinterpret_cast<const char*>(buf.data()), (size_t)r);
    }
}
bool write_file(const std::string& path, const std::string& data, std::string& err) {
    int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
    if (fd < 0) {
        err = "cannot create " + path + ": " + strerror(errno);
        return false;
    }
    size_t off = 0;
    while (off < data.size()) {
        long w = ::write(fd, data.data() + off, data.size() - off);
        if (w < 0 && errno == EINTR) continue;
        if (w < 0) {
            err = std::string("write failed: ") + strerror(errno);
            ::close(fd);
            return false;
        }
        off += (size_t)w;
    }
    ::close(fd);
    return true;
}

}  // namespace lashhost
