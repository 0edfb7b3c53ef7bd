#include "sketch_io.hpp"

#include <cmath>
#include <cstring>

#include "../../include/lash_gpu.h"
#include "io.hpp"

namespace lashhost {

size_t reg_bytes(int algo, int p) {
    if (algo == LASH_ALGO_HMH) return 32768;
    return (size_t)1 << p;
}

static void put_u64(std::vector<uint8_t>& b, uint64_t v) {
    for (int i = 0; i < 8; ++i) b.push_back((uint8_t)(v >> (8 * i)));
}
static void put_f64(std::vector<uint8_t>& b, double d) {
    uint64_t v;
    memcpy(&v, &d, 8);
    put_u64(b, v);
}
static uint64_t get_u64(const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

// HyperLogLog::with_p's alpha (streaming_algorithms 0.3.3)
static double hll_alpha(int p) {
    if (p == 4) return 0.673;
    if (p == 5) return 0.697;
    if (p == 6) return 0.709;
    return 0.7213 / (1.0 + 1.079 / (double)(1ull << p));
}

bool write_sketches(const std::string& path, int algo, int p, const void* regs, uint64_t n, int threads, std::string& err) {
    ZstdFileWriter w;
    if (!w.open(path, 3, threads, err)) return false;
    const size_t rb = reg_bytes(algo, p);
    const uint8_t* r = static_cast<const uint8_t*>(regs);
    std::vector<uint8_t> hdr;
    for (uint64_t i = 0; i < n; ++i, r += rb) {
        hdr.clear();
        if (algo == LASH_ALGO_ULL) {
            put_u64(hdr, rb);
        } else if (algo == LASH_ALGO_HLL) {
            // the fields push_hash64 maintains incrementally: zero = #empty registers, sum = sum 2^-m[i]
            // (every term and every partial sum is a dyadic rational that f64 holds exactly for the rho
            // values real inputs reach, so the incremental and the direct sum are the same number)
            uint64_t zero = 0;
            double sum = 0.0;
            for (size_t j = 0; j < rb; ++j) {
                zero += r[j] == 0;
                sum += std::ldexp(1.0, -(int)r[j]);
            }
            put_f64(hdr, hll_alpha(p));
            put_u64(hdr, zero);
            put_f64(hdr, sum);
            hdr.push_back((uint8_t)p);
            put_u64(hdr, rb);
        }
        // HMH: registers are u16 little-endian already (every supported host is LE)
        if (!hdr.empty() && !w.write(hdr.data(), hdr.size(), err)) return false;
        if (!w.write(r, rb, err)) return false;
    }
    return w.finish(err);
}

namespace {
// pull-exact-n helper over a ByteSource
struct Puller {
    ByteSource* s;
    std::string* err;
    bool get(uint8_t* dst, size_t n) {
        size_t off = 0;
        while (off < n) {
            long r = s->read(dst + off, n - off);
            if (r < 0) {
                *err = s->err();
                return false;
            }
            if (r == 0) {
                *err = "failed to fill whole buffer";  // what std::io::Read::read_exact reports (utils.rs:102 `?`)
                return false;
            }
            off += (size_t)r;
        }
        return true;
    }
};
}  // namespace

bool read_sketches(const std::string& path, int algo, int* p_inout, uint64_t n, std::vector<uint8_t>& regs, std::string& err) {
    auto src = open_source(path, err);
    if (!src) return false;
    Puller in{src.get(), &err};
    int p = p_inout ? *p_inout : 0;
    regs.clear();
    for (uint64_t i = 0; i < n; ++i) {
        size_t rb;
        if (algo == LASH_ALGO_HMH) {
            rb = 32768;
        } else {
            uint8_t h[33];
            const size_t hn = algo == LASH_ALGO_ULL ? 8 : 33;
            if (!in.get(h, hn)) return false;
            const uint64_t len = get_u64(h + hn - 8);
            int fp = 0;
            while (fp < 63 && (1ull << fp) < len) ++fp;
            if (len == 0 || (1ull << fp) != len || fp > 26) {
                err = "sketch file: register count " + std::to_string(len) + " is not a power of two";
                return false;
            }
            if (algo == LASH_ALGO_HLL && h[24] != fp) {
                err = "sketch file: HLL precision byte does not match the register count";
                return false;
            }
            if (p == 0) p = fp;
            if (fp != p) {
                err = "sketch file: precision " + std::to_string(fp) + " differs from the expected " + std::to_string(p);
                return false;
            }
            rb = (size_t)len;
        }
        const size_t off = regs.size();
        regs.resize(off + rb);
        if (!in.get(regs.data() + off, rb)) return false;
    }
    if (p_inout && algo != LASH_ALGO_HMH) *p_inout = p;
    return true;
}

// ------------------------------------------------------------------------------------------------
// JSON (only the two shapes lash writes)
// ------------------------------------------------------------------------------------------------
static void json_escape(std::string& o, const std::string& s) {
    o.push_back('"');
    for (unsigned char c : s) {
        switch (c) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\b': o += "\\b"; break;
            case '\f': o += "\\f"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            default:
                if (c < 0x20) {
                    static const char* hex = "0123456789abcdef";
                    o += "\\u00";
                    o.push_back(hex[c >> 4]);
                    o.push_back(hex[c & 15]);
                } else {
                    o.push_back((char)c);
                }
        }
    }
    o.push_back('"');
}

std::string json_pretty_string_array(const std::vector<std::string>& v) {
    if (v.empty()) return "[]";
    std::string o = "[\n";
    for (size_t i = 0; i < v.size(); ++i) {
        o += "  ";
        json_escape(o, v[i]);
        o += i + 1 < v.size() ? ",\n" : "\n";
    }
    o += "]";
    return o;
}
std::string json_pretty_string_map(const std::map<std::string, std::string>& m) {
    if (m.empty()) return "{}";
    std::string o = "{\n";
    size_t i = 0;
    for (const auto& kv : m) {
        o += "  ";
        json_escape(o, kv.first);
        o += ": ";
        json_escape(o, kv.second);
        o += ++i < m.size() ? ",\n" : "\n";
    }
    o += "}";
    return o;
}

namespace {
struct JsonIn {
    const std::string& t;
    size_t i = 0;
    std::string* err;
    void ws() {
        while (i < t.size() && (t[i] == ' ' || t[i] == '\n' || t[i] == '\r' || t[i] == '\t')) ++i;
    }
    bool lit(char c) {
        ws();
        if (i < t.size() && t[i] == c) {
            ++i;
            return true;
        }
        return false;
    }
    static void utf8(std::string& o, uint32_t cp) {
        if (cp < 0x80) o.push_back((char)cp);
        else if (cp < 0x800) { o.push_back((char)(0xc0 | (cp >> 6))); o.push_back((char)(0x80 | (cp & 63))); }
        else if (cp < 0x10000) { o.push_back((char)(0xe0 | (cp >> 12))); o.push_back((char)(0x80 | ((cp >> 6) & 63))); o.push_back((char)(0x80 | (cp & 63))); }
        else { o.push_back((char)(0xf0 | (cp >> 18))); o.push_back((char)(0x80 | ((cp >> 12) & 63))); o.push_back((char)(0x80 | ((cp >> 6) & 63))); o.push_back((char)(0x80 | (cp & 63))); }
    }
    bool hex4(uint32_t& v) {
        if (i + 4 > t.size()) return false;
        v = 0;
        for (int k = 0; k < 4; ++k) {
            const char c = t[i++];
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else return false;
        }
        return true;
    }
    bool str(std::string& o) {
        ws();
        if (i >= t.size() || t[i] != '"') {
            *err = "JSON: string expected at offset " + std::to_string(i);
            return false;
        }
        ++i;
        o.clear();
        while (i < t.size()) {
            const char c = t[i++];
            if (c == '"') return true;
            if (c != '\\') {
                o.push_back(c);
                continue;
            }
            if (i >= t.size()) break;
            const char e = t[i++];
            switch (e) {
                case '"': o.push_back('"'); break;
                case '\\': o.push_back('\\'); break;
                case '/': o.push_back('/'); break;
                case 'b': o.push_back('\b'); break;
                case 'f': o.push_back('\f'); break;
                case 'n': o.push_back('\n'); break;
                case 'r': o.push_back('\r'); break;
                case 't': o.push_back('\t'); break;
                case 'u': {
                    uint32_t cp;
                    if (!hex4(cp)) { *err = "JSON: bad \\u escape"; return false; }
                    if (cp >= 0xd800 && cp < 0xdc00 && i + 6 <= t.size() && t[i] == '\\' && t[i + 1] == 'u') {
                        i += 2;
                        uint32_t lo;
                        if (!hex4(lo)) { *err = "JSON: bad \\u escape"; return false; }
                        cp = 0x10000 + ((cp - 0xd800) << 10) + (lo - 0xdc00);
                    }
                    utf8(o, cp);
                    break;
                }
                default:
                    *err = "JSON: bad escape";
                    return false;
            }
        }
        *err = "JSON: unterminated string";
        return false;
    }
    bool end() {
        ws();
        if (i != t.size()) {
            *err = "JSON: trailing characters";
            return false;
        }
        return true;
    }
};
}  // namespace

bool json_parse_string_array(const std::string& text, std::vector<std::string>& out, std::string& err) {
    JsonIn in{text, 0, &err};
    out.clear();
    if (!in.lit('[')) {
        err = "JSON: '[' expected";
        return false;
    }
    if (in.lit(']')) return in.end();
    for (;;) {
        std::string s;
        if (!in.str(s)) return false;
        out.push_back(std::move(s));
        if (in.lit(',')) continue;
        if (in.lit(']')) return in.end();
        err = "JSON: ',' or ']' expected";
        return false;
    }
}
bool json_parse_string_map(const std::string& text, std::map<std::string, std::string>& out, std::string& err) {
    JsonIn in{text, 0, &err};
    out.clear();
    if (!in.lit('{')) {
        err = "JSON: '{' expected";
        return false;
    }
    if (in.lit('}')) return in.end();
    for (;;) {
        std::string k, v;
        if (!in.str(k)) return false;
        if (!in.lit(':')) {
            err = "JSON: ':' expected";
            return false;
        }
        if (!in.str(v)) return false;
        out[k] = v;
        if (in.lit(',')) continue;
        if (in.lit('}')) return in.end();
        err = "JSON: ',' or '}' expected";
        return false;
    }
}

}  // namespace lashhost
