/*
 * lash_oracle.c -- CPU restatement of jianshu93/lash's two hot paths.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (lash_b200/, include/lash_gpu.h) never links, imports or calls anything in oracle/.
 *
 * PARITY STATUS: "parity unpinned" for everything except XXH3.
 *   The reference (/root/reference/src/{utils,main}.rs) only *calls* five crates.io
 *   dependencies that are not vendored and cannot be built offline (no cargo/rustc):
 *     xxhash-rust 0.8.15, kmerutils 0.0.14, hyperminhash 0.1.4,
 *     streaming_algorithms 0.3.3, ultraloglog 0.1.6        (Cargo.lock)
 *   and the reference has no tests, fixtures or golden vectors.  XXH3 is a frozen public
 *   spec and is pinned here against libxxhash 0.8.2 / python-xxhash (tests/golden/xxh3_kat.json).
 *   The other crates are restated from their published algorithms (hash4j UltraLogLog +
 *   FGRA/ML estimators, Heule et al. HLL++ as implemented by streaming_algorithms,
 *   axiomhq HyperMinHash); each 1-bit convention that could differ is a named switch below.
 *
 * Every function cites the reference call site (file:line under /root/reference) it follows.
 * Floating point: compiled with -ffp-contract=off; all sums run in the reference's order
 * (register index order, one scalar accumulator).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LO_HMH 0
#define LO_HLL 1
#define LO_ULL 2

/* ---- convention switches (SURVEY Appendix A, items marked "recalled") ------------------- */
/* hyperminhash: which half of the 128-bit XXH3 result is `x` (index+lz) and which is `y` (sig). */
#ifndef LO_HMH_X_IS_HIGH64
#define LO_HMH_X_IS_HIGH64 1
#endif
/* hyperminhash: Rust port returns cardinality as f64 (Go original truncates to u64). */
#ifndef LO_HMH_CARD_TRUNC
#define LO_HMH_CARD_TRUNC 0
#endif

/* ======================================================================================= */
/* XXH3 (xxhash-rust 0.8.15 == XXH3 v0.8 spec, default secret)                              */
/* ======================================================================================= */
#define XXH_PRIME32_1 0x9E3779B1U
#define XXH_PRIME32_2 0x85EBCA77U
#define XXH_PRIME32_3 0xC2B2AE3DU
#define XXH_PRIME64_1 0x9E3779B185EBCA87ULL
#define XXH_PRIME64_2 0xC2B2AE3D27D4EB4FULL
#define XXH_PRIME_MX1 0x165667919E3779F9ULL
#define XXH_PRIME_MX2 0x9FB21C651E98DF25ULL

/* first 32 bytes of XXH3's kSecret (public constant of the spec) */
static const uint8_t kSecret32[32] = {
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c,
    0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f,
};

/* the two secret words the short paths need, folded (checked against kSecret32 in lo_selfcheck) */
#define LO_SECRET_X_8_16 0xc73ab174c5ecd5a2ULL
#define LO_SECRET_X_16_24 0xc4f023344dc994acULL
static uint64_t rd64(const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}
static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint32_t bswap32(uint32_t x) {
    return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
}

int lo_selfcheck(void) {
    return (rd64(kSecret32 + 8) ^ rd64(kSecret32 + 16)) == LO_SECRET_X_8_16 &&
           (rd64(kSecret32 + 16) ^ rd64(kSecret32 + 24)) == LO_SECRET_X_16_24;
}

/* xxh3_64_with_seed(&v.to_le_bytes(), seed): the 4..8-byte short path, len = 8.
 * Call sites: utils.rs:412 (HLL), utils.rs:428 (ULL). */
__attribute__((hot)) uint64_t lo_xxh3_64_le64(uint64_t v, uint64_t seed) {
    uint64_t s = seed ^ ((uint64_t)bswap32((uint32_t)seed) << 32);
    uint32_t in1 = (uint32_t)v;         /* bytes 0..3 */
    uint32_t in2 = (uint32_t)(v >> 32); /* bytes len-4..len-1 */
    uint64_t bitflip = LO_SECRET_X_8_16 - s; /* readLE64(kSecret+8) ^ readLE64(kSecret+16) */
    uint64_t in64 = (uint64_t)in2 + ((uint64_t)in1 << 32);
    uint64_t h = in64 ^ bitflip;
    /* XXH3_rrmxmx(h, len=8) */
    h ^= rotl64(h, 49) ^ rotl64(h, 24);
    h *= XXH_PRIME_MX2;
    h ^= (h >> 35) + 8;
    h *= XXH_PRIME_MX2;
    return h ^ (h >> 28);
}

/* xxh3_128_with_seed(&w.to_le_bytes(), seed): the 4..8-byte short path, len = 4.
 * Call site: utils.rs:397 via hyperminhash::Sketch::add_bytes_with_seed. */
void lo_xxh3_128_le32(uint32_t w, uint64_t seed, uint64_t* out_lo, uint64_t* out_hi) {
    uint64_t s = seed ^ ((uint64_t)bswap32((uint32_t)seed) << 32);
    uint32_t in_lo = w, in_hi = w; /* len == 4: both reads cover the same 4 bytes */
    uint64_t in64 = (uint64_t)in_lo + ((uint64_t)in_hi << 32);
    uint64_t bitflip = LO_SECRET_X_16_24 + s; /* readLE64(kSecret+16) ^ readLE64(kSecret+24) */
    uint64_t keyed = in64 ^ bitflip;
    __uint128_t m = (__uint128_t)keyed * (XXH_PRIME64_1 + (4u << 2));
    uint64_t lo = (uint64_t)m, hi = (uint64_t)(m >> 64);
    hi += lo << 1;
    lo ^= hi >> 3;
    lo ^= lo >> 35;
    lo *= XXH_PRIME_MX2;
    lo ^= lo >> 28;
    /* XXH3_avalanche(hi) */
    hi ^= hi >> 37;
    hi *= XXH_PRIME_MX1;
    hi ^= hi >> 32;
    *out_lo = lo;
    *out_hi = hi;
}

/* ======================================================================================= */
/* sequence front end                                                                       */
/* ======================================================================================= */

/* utils.rs:33-41 filter_out_n: keep only uppercase A,C,G,T; everything else is deleted. */
size_t lo_filter_out_n(const uint8_t* seq, size_t n, uint8_t* out) {
    size_t o = 0;
    for (size_t i = 0; i < n; ++i) {
        uint8_t c = seq[i];
        if (c == 'A' || c == 'C' || c == 'T' || c == 'G') out[o++] = c;
    }
    return o;
}

/* utils.rs:57-64 mask_bits */
uint64_t lo_mask_bits(uint64_t v, int k) {
    int b = 2 * k;
    return b == 64 ? v : (v & ((1ULL << b) - 1));
}

/* kmerutils 2-bit alphabet (Sequence::new(&seq, 2), utils.rs:464): A=0 C=1 G=2 T=3 */
static inline uint64_t base2(uint8_t c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        default: return 3; /* 'T' (input already filtered) */
    }
}

typedef void (*lo_kmer_fn)(void* ctx, uint64_t masked);

/* utils.rs:464-499: KmerSeqIterator over one filtered record; canonical = min(fwd, revcomp);
 * masked = mask_bits(value, k).  All three kmerutils word types (Kmer32bit, Kmer16b32bit,
 * Kmer64bit) hold the k-mer with its first base in the most significant 2 bits of the 2k-bit
 * value, so one code path serves k in [1,32]. */
static void for_each_canonical_kmer(const uint8_t* filtered, size_t n, int k, lo_kmer_fn fn, void* ctx) {
    if (n < (size_t)k) return; /* utils.rs:460-462 */
    uint64_t mask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    uint64_t fwd = 0, rc = 0;
    for (size_t i = 0; i < n; ++i) {
        uint64_t b = base2(filtered[i]);
        fwd = ((fwd << 2) | b) & mask;
        rc = (rc >> 2) | ((3 - b) << (2 * (k - 1)));
        if (i + 1 >= (size_t)k) {
            uint64_t canon = fwd < rc ? fwd : rc; /* utils.rs:470,482,494 */
            fn(ctx, lo_mask_bits(canon, k));      /* utils.rs:471-474 */
        }
    }
}

struct kmer_collect {
    uint64_t* out;
    size_t n;
};
static void collect_fn(void* c, uint64_t m) {
    struct kmer_collect* kc = (struct kmer_collect*)c;
    kc->out[kc->n++] = m;
}
/* test helper: canonical masked k-mers of one *filtered* record; returns the count */
size_t lo_canonical_kmers(const uint8_t* filtered, size_t n, int k, uint64_t* out) {
    struct kmer_collect kc = {out, 0};
    for_each_canonical_kmer(filtered, n, k, collect_fn, &kc);
    return kc.n;
}

/* ======================================================================================= */
/* register updates (KmerSketch::add_kmer, utils.rs:395-398, 411-413, 427-429)              */
/* ======================================================================================= */
static inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }

/* streaming_algorithms 0.3.3 HyperLogLog::push_hash64: j = low p bits, rho = clz(x >> p) - p + 1 */
void lo_hll_push_hash64(uint8_t* regs, int p, uint64_t x) {
    uint64_t j = x & ((1ULL << p) - 1);
    uint64_t w = x >> p;
    int rho = clz64(w) - p + 1; /* get_rho(w, 64 - p): w == 0 -> 64 - p + 1 */
    if (regs[j] < rho) regs[j] = (uint8_t)rho;
}

/* ultraloglog 0.1.6 (port of hash4j UltraLogLog) */
static inline uint64_t ull_unpack(uint8_t r) {
    int sh = ((r >> 2) - 2) & 63; /* Java/Rust wrapping shift semantics */
    return (uint64_t)(4 | (r & 3)) << sh;
}
static inline uint8_t ull_pack(uint64_t hp) {
    int nlz = clz64(hp) + 1; /* hp != 0 */
    uint64_t low2 = (nlz >= 64) ? 0 : ((hp << nlz) >> 62);
    return (uint8_t)(((unsigned)(-nlz) << 2) | (unsigned)low2);
}
void lo_ull_add(uint8_t* regs, int p, uint64_t h) {
    uint64_t idx = h >> (64 - p);
    int nlz = clz64(~(~h << p)); /* in [0, 64-p] */
    uint64_t hp = ull_unpack(regs[idx]);
    hp |= 1ULL << (nlz + p - 1);
    regs[idx] = ull_pack(hp);
}
/* UltraLogLog::merge (utils.rs:260-262) */
void lo_ull_merge(const uint8_t* a, const uint8_t* b, uint8_t* out, int p) {
    size_t m = (size_t)1 << p;
    for (size_t i = 0; i < m; ++i) {
        uint64_t hp = ull_unpack(a[i]) | ull_unpack(b[i]);
        out[i] = hp ? ull_pack(hp) : 0;
    }
}

/* hyperminhash 0.1.4 Sketch::add_bytes_with_seed -> add_hash(x, y); P=14, Q=6, R=10 */
void lo_hmh_add_hash(uint16_t* regs, uint64_t x, uint64_t y) {
    uint64_t k = x >> 50;
    int lz = clz64((x << 14) ^ ((1ULL << 14) - 1)) + 1;
    uint16_t sig = (uint16_t)(y & 1023);
    uint16_t reg = (uint16_t)((lz << 10) | sig);
    if (regs[k] < reg) regs[k] = reg;
}

static size_t reg_bytes(int algo, int p) {
    return algo == LO_HMH ? 16384 * 2 : ((size_t)1 << p);
}
size_t lo_reg_bytes(int algo, int p) { return reg_bytes(algo, p); }

struct sk_ctx {
    int algo, p;
    uint64_t seed;
    void* regs;
};
/* KmerSketch::add_kmer for the three impls */
void lo_add_kmer(int algo, int p, void* regs, uint64_t masked, uint64_t seed) {
    if (algo == LO_HMH) { /* utils.rs:395-398: only the low 32 bits are hashed */
        uint64_t lo, hi;
        lo_xxh3_128_le32((uint32_t)masked, seed, &lo, &hi);
#if LO_HMH_X_IS_HIGH64
        lo_hmh_add_hash((uint16_t*)regs, hi, lo);
#else
        lo_hmh_add_hash((uint16_t*)regs, lo, hi);
#endif
    } else if (algo == LO_HLL) { /* utils.rs:411-413 */
        lo_hll_push_hash64((uint8_t*)regs, p, lo_xxh3_64_le64(masked, seed));
    } else { /* utils.rs:427-429 */
        lo_ull_add((uint8_t*)regs, p, lo_xxh3_64_le64(masked, seed));
    }
}
static void sk_fn(void* c, uint64_t masked) {
    struct sk_ctx* s = (struct sk_ctx*)c;
    lo_add_kmer(s->algo, s->p, s->regs, masked, s->seed);
}

/* One FASTA/FASTQ record (raw bytes as needletail would hand them over, newlines already
 * removed) into a sketch: utils.rs:457-503. */
void lo_sketch_add_record(int algo, int p, int k, uint64_t seed, const uint8_t* raw, size_t n, void* regs) {
    uint8_t* f = (uint8_t*)malloc(n ? n : 1); /* utils.rs:459 allocates per record too */
    size_t fn = lo_filter_out_n(raw, n, f);
    struct sk_ctx s = {algo, p, seed, regs};
    for_each_canonical_kmer(f, fn, k, sk_fn, &s);
    free(f);
}

/* sketch_files (utils.rs:439-510): one task per genome ("file"), order-preserving.
 *   seqs:      all records' raw bytes, concatenated
 *   rec_off:   n_rec_total+1 byte offsets into seqs
 *   gen_rec:   n_genomes+1 record indices: genome g owns records [gen_rec[g], gen_rec[g+1])
 *   regs_out:  n_genomes * reg_bytes, zero-filled here
 * threads mirrors rayon's pool size (main.rs:189-192). */
struct sk_job {
    int algo, p, k;
    uint64_t seed;
    const uint8_t* seqs;
    const uint64_t* rec_off;
    const uint64_t* gen_rec;
    uint64_t n_genomes;
    uint8_t* regs_out;
    volatile uint64_t next;
};
static void* sk_worker(void* arg) {
    struct sk_job* j = (struct sk_job*)arg;
    size_t rb = reg_bytes(j->algo, j->p);
    for (;;) {
        uint64_t g = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (g >= j->n_genomes) break;
        void* regs = j->regs_out + g * rb;
        for (uint64_t r = j->gen_rec[g]; r < j->gen_rec[g + 1]; ++r)
            lo_sketch_add_record(j->algo, j->p, j->k, j->seed, j->seqs + j->rec_off[r],
                                 (size_t)(j->rec_off[r + 1] - j->rec_off[r]), regs);
    }
    return NULL;
}
int lo_sketch_genomes(int algo, int p, int k, uint64_t seed, const uint8_t* seqs, const uint64_t* rec_off,
                      const uint64_t* gen_rec, uint64_t n_genomes, void* regs_out, int threads) {
    if (k < 1 || k > 32) return -1; /* utils.rs:500-502 panics */
    if (algo == LO_ULL && (p < 3 || p > 26)) return -2;
    if (algo == LO_HLL && (p < 4 || p > 18)) return -2;
    size_t rb = reg_bytes(algo, p);
    memset(regs_out, 0, rb * n_genomes);
    struct sk_job job = {algo, p, k, seed, seqs, rec_off, gen_rec, n_genomes, (uint8_t*)regs_out, 0};
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, sk_worker, &job);
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th);
    return 0;
}

/* ======================================================================================= */
/* estimators                                                                               */
/* ======================================================================================= */

/* ---- streaming_algorithms HyperLogLog::len() (HLL++), utils.rs:315,358 ------------------ */
static const double HLL_THRESH[15] = {10, 20, 40, 80, 220, 400, 900, 1800, 3100, 6500, 11500, 20000, 50000, 120000, 350000};
static double hll_alpha(int p) {
    if (p == 4) return 0.673;
    if (p == 5) return 0.697;
    if (p == 6) return 0.709;
    return 0.7213 / (1.0 + 1.079 / (double)(1ULL << p));
}
static inline double pow2neg(unsigned r) { /* 2^-r, the crate's bit hack */
    uint64_t bits = (0xFFFFFFFFFFFFFFFFULL - (uint64_t)r) << 54 >> 2;
    double d;
    memcpy(&d, &bits, 8);
    return d;
}
/* status: 0 ok; 1 = estimate falls in the HLL++ bias-table regime (e <= 5m, not linear counting):
 * Google's empirical bias tables are not reproducible offline, the oracle returns NaN there. */
double lo_hll_len(const uint8_t* regs, int p, int* status) {
    size_t m = (size_t)1 << p;
    size_t zero = 0;
    double sum = 0.0;
    for (size_t i = 0; i < m; ++i) { /* union(): zero recount + sequential sum of 2^-m[i] */
        zero += regs[i] == 0;
        sum += pow2neg(regs[i]);
    }
    if (status) *status = 0;
    if (zero > 0) {
        double h = (double)m * log((double)m / (double)zero);
        if (h <= HLL_THRESH[p - 4]) return h;
    }
    double e = hll_alpha(p) * (double)(m * m) / sum;
    if (e <= (double)(5 * m)) {
        if (status) *status = 1;
        return NAN;
    }
    return e;
}

/* ---- ultraloglog FGRA: UltraLogLog::get_distinct_count_estimate (utils.rs:215,266) ------- */
#define ULL_TAU 0.8194911375910897
#define ULL_V 0.6118931496978437
#define ULL_ETA0 4.663135422063788
#define ULL_ETA1 2.1378502137958524
#define ULL_ETA2 2.781144650979996
#define ULL_ETA3 0.9824082545153715

static double g_pow2tau, g_pow2mtau, g_pow4mtau, g_etaX, g_eta23X, g_eta13X, g_eta3012XX, g_phi1, g_pinit;
static double g_ull_reg[256]; /* REGISTER_CONTRIBUTIONS: eta[i&3] * 2^(-tau*(3+(i>>2))) */
static double g_ull_factor[27];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void ull_init(void) {
    g_pow2tau = pow(2.0, ULL_TAU);
    g_pow2mtau = pow(2.0, -ULL_TAU);
    g_pow4mtau = pow(4.0, -ULL_TAU);
    g_etaX = ULL_ETA0 - ULL_ETA1 - ULL_ETA2 + ULL_ETA3;
    g_eta23X = (ULL_ETA2 - ULL_ETA3) / g_etaX;
    g_eta13X = (ULL_ETA1 - ULL_ETA3) / g_etaX;
    g_eta3012XX = (ULL_ETA3 * ULL_ETA0 - ULL_ETA1 * ULL_ETA2) / (g_etaX * g_etaX);
    g_phi1 = ULL_ETA0 / (g_pow2tau * (2.0 * g_pow2tau - 1.0));
    g_pinit = g_etaX * (g_pow4mtau / (2.0 - g_pow2mtau));
    const double eta[4] = {ULL_ETA0, ULL_ETA1, ULL_ETA2, ULL_ETA3};
    for (int i = 0; i < 256; ++i) g_ull_reg[i] = eta[i & 3] * pow(2.0, -ULL_TAU * (double)(3 + (i >> 2)));
    for (int p = 3; p <= 26; ++p) {
        double m = (double)(1ULL << p);
        g_ull_factor[p] = m * pow(m, 1.0 / ULL_TAU) / (1.0 + ULL_V * (1.0 + ULL_TAU) / (2.0 * m));
    }
}
/* tables exported so the GPU library can be fed the *same* doubles (they are data, not code) */
const double* lo_ull_register_contributions(void) {
    pthread_once(&g_once, ull_init);
    return g_ull_reg;
}
double lo_ull_estimation_factor(int p) {
    pthread_once(&g_once, ull_init);
    return g_ull_factor[p];
}

static double psi_prime(double z, double z2) { return (z + g_eta23X) * (z2 + g_eta13X) + g_eta3012XX; }
static double ull_sigma(double z) {
    if (z <= 0.0) return ULL_ETA3;
    if (z >= 1.0) return INFINITY;
    double powZ = z, nextPowZ = z * z, s = 0.0, powTau = g_etaX;
    for (;;) {
        double oldS = s;
        double nn = nextPowZ * nextPowZ;
        s += powTau * (powZ - nextPowZ) * psi_prime(nextPowZ, nn);
        if (!(s > oldS)) return s / z;
        powZ = nextPowZ;
        nextPowZ = nn;
        powTau *= g_pow2tau;
    }
}
static double ull_phi(double z, double zSquare) {
    if (z <= 0.0) return 0.0;
    if (z >= 1.0) return g_phi1;
    double previousPowZ = zSquare, powZ = z, nextPowZ = sqrt(powZ);
    double p = g_pinit / (1.0 + nextPowZ);
    double ps = psi_prime(powZ, previousPowZ);
    double s = nextPowZ * (ps + ps) * p;
    for (;;) {
        previousPowZ = powZ;
        powZ = nextPowZ;
        double oldS = s;
        nextPowZ = sqrt(powZ);
        double nextPs = psi_prime(powZ, previousPowZ);
        p *= g_pow2mtau / (1.0 + nextPowZ);
        s += nextPowZ * ((nextPs + nextPs) - (powZ + nextPowZ) * ps) * p;
        if (!(s > oldS)) return s;
        ps = nextPs;
    }
}
double lo_ull_fgra(const uint8_t* regs, int p) {
    pthread_once(&g_once, ull_init);
    size_t m = (size_t)1 << p;
    int64_t c0 = 0, c4 = 0, c8 = 0, c10 = 0, w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    double sum = 0.0;
    int off = 4 * p + 4;
    for (size_t i = 0; i < m; ++i) {
        int r = regs[i];
        int r2 = r - off;
        if (r2 < 0) {
            if (r2 < -8) c0++;
            if (r2 == -8) c4++;
            if (r2 == -4) c8++;
            if (r2 == -2) c10++;
        } else if (r < 252) {
            sum += g_ull_reg[r2];
        } else {
            if (r == 252) w0++;
            if (r == 253) w1++;
            if (r == 254) w2++;
            if (r == 255) w3++;
        }
    }
    if (c0 > 0 || c4 > 0 || c8 > 0 || c10 > 0) {
        double alpha = (double)((int64_t)m + 3 * (c0 + c4 + c8 + c10));
        double beta = (double)((int64_t)m - c0 - c4);
        double gamma = (double)(4 * c0 + 2 * c4 + 3 * c8 + c10);
        double q = (sqrt(beta * beta + 4.0 * alpha * gamma) - beta) / (2.0 * alpha);
        double rz = q * q;
        double z = rz * rz;
        if (c0 > 0) sum += (double)c0 * ull_sigma(z);
        if (c4 > 0) sum += (double)c4 * (g_pow2mtau * g_etaX) * psi_prime(z, z * z);
        if (c8 > 0) sum += (double)c8 * (z * (g_pow4mtau * (ULL_ETA0 - ULL_ETA1)) + g_pow4mtau * ULL_ETA1);
        if (c10 > 0) sum += (double)c10 * (z * (g_pow4mtau * (ULL_ETA2 - ULL_ETA3)) + g_pow4mtau * ULL_ETA3);
    }
    if (w0 > 0 || w1 > 0 || w2 > 0 || w3 > 0) {
        double c = (double)(w0 + w1 + w2 + w3);
        double alpha = (double)m + 3.0 * c;
        double beta = (double)(w0 + w1 + 2 * (w2 + w3));
        double gamma = (double)((int64_t)m + 2 * w0 + w2 - w3);
        double z = sqrt((sqrt(beta * beta + 4.0 * alpha * gamma) - beta) / (2.0 * alpha));
        double rz = sqrt(z);
        double s = ull_phi(rz, z) * c;
        s += z * (1.0 + rz) * ((double)w0 * ULL_ETA0 + (double)w1 * ULL_ETA1 + (double)w2 * ULL_ETA2 + (double)w3 * ULL_ETA3);
        s += rz * ((double)(w0 + w1) * (z * (g_pow2mtau * (ULL_ETA0 - ULL_ETA2)) + g_pow2mtau * ULL_ETA2) +
                   (double)(w2 + w3) * (z * (g_pow2mtau * (ULL_ETA1 - ULL_ETA3)) + g_pow2mtau * ULL_ETA3));
        sum += s * pow(g_pow2mtau, (double)(65 - p)) / ((1.0 + rz) * (1.0 + z));
    }
    return g_ull_factor[p] * pow(sum, -1.0 / ULL_TAU);
}

/* ---- ultraloglog MaximumLikelihoodEstimator.estimate (utils.rs:216,267) ------------------ */
#define ULL_INV_SQRT_FISHER 0.7608621002725182
#define ULL_ML_BIAS 0.48147376527720065

/* hash4j DistinctCountUtil.solveMaximumLikelihoodEquation (Ertl 2017, Alg. 8) */
static double ldexp_bits(double x, int e) { /* exact x * 2^e for normal x and result */
    return ldexp(x, e);
}
double lo_solve_ml(double a, const int32_t* b, int n, double eps) {
    if (a == 0.0) return INFINITY;
    int kMax = n;
    while (kMax >= 0 && b[kMax] == 0) --kMax;
    if (kMax < 0) return 0.0;
    int kMin = kMax;
    int64_t s1 = b[kMax];
    double s2 = ldexp_bits((double)b[kMax], kMax);
    for (int k = kMax - 1; k >= 0; --k) {
        int32_t t = b[k];
        if (t > 0) {
            s1 += t;
            s2 += ldexp_bits((double)t, k);
            kMin = k;
        }
    }
    double gPrev = 0.0, x;
    if (s2 <= 1.5 * a)
        x = (double)s1 / (0.5 * s2 + a);
    else
        x = log1p(s2 / a) * ((double)s1 / s2);
    double dx = x;
    while (dx > x * eps) {
        int kappa = ilogb(x) + 2;
        int sh = (kMax > kappa ? kMax : kappa) + 1;
        double xp = ldexp_bits(x, -sh);
        double xp2 = xp * xp;
        double h = xp - xp2 / 3.0 + (xp2 * xp2) * (1.0 / 45.0 - xp2 / 472.5);
        for (int k = kappa - 1; k >= kMax; --k) {
            double hp = 1.0 - h;
            h = (xp + h * hp) / (xp + hp);
            xp += xp;
        }
        double g = (double)b[kMax] * h;
        for (int k = kMax - 1; k >= kMin; --k) {
            double hp = 1.0 - h;
            h = (xp + h * hp) / (xp + hp);
            xp += xp;
            g += (double)b[k] * h;
        }
        g += x * a;
        if (gPrev < g && g <= (double)s1)
            dx *= (g - (double)s1) / (gPrev - g);
        else
            dx = 0.0;
        x += dx;
        gPrev = g;
    }
    return x;
}
static uint64_t ull_ml_contribute(int r, int32_t* b, int p) {
    int r2 = r - 4 * p - 4;
    if (r2 < 0) {
        uint64_t ret = 4;
        if (r2 == -2 || r2 == -8) { b[0] += 1; ret -= 2; }
        if (r2 == -2 || r2 == -4) { b[1] += 1; ret -= 1; }
        return ret << (62 - p);
    } else {
        int k = r2 >> 2;
        uint64_t ret = 0xE000000000000000ULL;
        uint64_t y0 = r & 1, y1 = (r >> 1) & 1;
        ret -= y0 << 63;
        ret -= y1 << 62;
        b[k] += (int32_t)y0;
        b[k + 1] += (int32_t)y1;
        b[k + 2] += 1;
        return ret >> (k + p);
    }
}
/* exposes the integer statistics so tests can compare GPU counts exactly */
void lo_ull_ml_stats(const uint8_t* regs, int p, uint64_t* S_out, int32_t* b_out /*[66]*/) {
    size_t m = (size_t)1 << p;
    uint64_t S = 0;
    memset(b_out, 0, 66 * sizeof(int32_t));
    for (size_t i = 0; i < m; ++i) S += ull_ml_contribute(regs[i], b_out, p);
    *S_out = S;
}
double lo_ull_ml(const uint8_t* regs, int p) {
    size_t m = (size_t)1 << p;
    uint64_t S;
    int32_t b[66];
    lo_ull_ml_stats(regs, p, &S, b);
    if (S == 0) return regs[0] == 0 ? 0.0 : INFINITY;
    b[63 - p] += b[64 - p];
    double factor = (double)(m << 1);
    double a = (double)S * factor * 0x1p-64;
    double eps = 1e-3 * ULL_INV_SQRT_FISHER / sqrt((double)m);
    return factor * lo_solve_ml(a, b, 63 - p, eps) / (1.0 + ULL_ML_BIAS / (double)m);
}

/* ---- hyperminhash 0.1.4 (port of axiomhq/hyperminhash), utils.rs:164 --------------------- */
#define HMH_P 14
#define HMH_M 16384
#define HMH_Q 6
#define HMH_R 10
#define HMH_C 0.169919487159739093975315012348
static double hmh_beta(double ez) {
    double zl = log(ez + 1.0);
    return -0.370393911 * ez + 0.070471823 * zl + 0.17393686 * pow(zl, 2.0) + 0.16339839 * pow(zl, 3.0) +
           -0.09237745 * pow(zl, 4.0) + 0.03738027 * pow(zl, 5.0) + -0.005384159 * pow(zl, 6.0) +
           0.00042419 * pow(zl, 7.0);
}
double lo_hmh_cardinality(const uint16_t* regs) {
    double sum = 0.0, ez = 0.0;
    for (int i = 0; i < HMH_M; ++i) {
        unsigned lz = regs[i] >> HMH_R;
        if (lz == 0) ez += 1.0;
        sum += 1.0 / pow(2.0, (double)lz);
    }
    double alpha = 0.7213 / (1.0 + 1.079 / (double)HMH_M);
    double c = alpha * (double)HMH_M * ((double)HMH_M - ez) / (hmh_beta(ez) + sum);
#if LO_HMH_CARD_TRUNC
    c = (double)(uint64_t)c;
#endif
    return c;
}
static double hmh_expected_collision(double n, double m) {
    double x = 0.0;
    for (int i = 1; i <= 64; ++i) {
        for (int j = 1; j <= 1024; ++j) {
            double b1, b2;
            if (i != 64) {
                double den = pow(2.0, (double)(HMH_P + HMH_R + i));
                b1 = (1024.0 + j) / den;
                b2 = (1024.0 + j + 1.0) / den;
            } else {
                double den = pow(2.0, (double)(HMH_P + HMH_R + i - 1));
                b1 = j / den;
                b2 = (j + 1.0) / den;
            }
            double prx = pow(1.0 - b2, n) - pow(1.0 - b1, n);
            double pry = pow(1.0 - b2, m) - pow(1.0 - b1, m);
            x += prx * pry;
        }
    }
    return x * (double)HMH_P + 0.5;
}
double lo_hmh_expected_collisions(double n, double m) {
    if (n < m) { double t = n; n = m; m = t; }
    if (n > pow(2.0, pow(2.0, (double)HMH_Q) + (double)HMH_R)) return 18446744073709551615.0;
    if (n > pow(2.0, (double)(HMH_P + 5))) {
        double d = (4.0 * n / m) / pow((1.0 + n) / m, 2.0);
        return HMH_C * pow(2.0, (double)(HMH_P - HMH_R)) * d + 0.5;
    }
    return hmh_expected_collision(n, m) / (double)HMH_P;
}
void lo_hmh_counts(const uint16_t* a, const uint16_t* b, uint32_t* C_out, uint32_t* N_out) {
    uint32_t C = 0, N = 0;
    for (int i = 0; i < HMH_M; ++i) {
        if (a[i] != 0 && a[i] == b[i]) C++;
        if (a[i] != 0 || b[i] != 0) N++;
    }
    *C_out = C;
    *N_out = N;
}
double lo_hmh_similarity(const uint16_t* a, const uint16_t* b) {
    uint32_t C, N;
    lo_hmh_counts(a, b, &C, &N);
    if (C == 0) return 0.0;
    double n = lo_hmh_cardinality(a), m = lo_hmh_cardinality(b);
    double ec = lo_hmh_expected_collisions(n, m);
    if ((double)C < ec) return 0.0;
    return ((double)C - ec) / (double)N;
}

/* ======================================================================================= */
/* distance (main.rs:415-423) and the three *_distance inner loops                          */
/* ======================================================================================= */
double lo_compute_distance_f64(double frac, int k, int model) {
    double kk = (double)k;
    if (model == 2) return frac; /* oracle-only test aid: expose `frac` itself */
    if (model == 1) return fmin(-log(frac) / kk, 1.0);
    return 1.0 - pow(frac, 1.0 / kk);
}
float lo_compute_distance_f32(float frac, int k, int model) {
    float kk = (float)k;
    if (model == 2) return frac;
    if (model == 1) return fminf(-logf(frac) / kk, 1.0f);
    return 1.0f - powf(frac, 1.0f / kk);
}

/* per-sketch cardinality: utils.rs:213-219 (ULL), :314-316 (HLL); HMH recomputes per pair */
double lo_cardinality(int algo, int p, int estimator, const void* regs, int* status) {
    if (status) *status = 0;
    if (algo == LO_HLL) return lo_hll_len((const uint8_t*)regs, p, status);
    if (algo == LO_ULL) return estimator == 0 ? lo_ull_fgra((const uint8_t*)regs, p) : lo_ull_ml((const uint8_t*)regs, p);
    return lo_hmh_cardinality((const uint16_t*)regs);
}

/* `frac` of one pair = 2s/(1+s): utils.rs:164-167 (HMH), :260-275 (ULL), :355-363 (HLL).
 * card_r / card_q are the precomputed per-sketch cardinalities (ignored for HMH). */
double lo_pair_fraction(int algo, int p, int estimator, const void* r, const void* q, double card_r, double card_q,
                        uint8_t* scratch, int* status) {
    double s;
    if (status) *status = 0;
    if (algo == LO_HMH) {
        s = lo_hmh_similarity((const uint16_t*)q, (const uint16_t*)r);
        if (!(s > 0.0)) s = 0.0; /* f64::max(0.0) */
    } else if (algo == LO_ULL) {
        lo_ull_merge((const uint8_t*)r, (const uint8_t*)q, scratch, p);
        double u = estimator == 0 ? lo_ull_fgra(scratch, p) : lo_ull_ml(scratch, p);
        double sim = (card_r + card_q - u) / u;
        s = sim < 0.0 ? 0.0 : sim; /* NaN propagates, as in utils.rs:274 */
    } else {
        size_t m = (size_t)1 << p;
        const uint8_t* a = (const uint8_t*)r;
        const uint8_t* b = (const uint8_t*)q;
        for (size_t i = 0; i < m; ++i) scratch[i] = a[i] > b[i] ? a[i] : b[i]; /* clone + union */
        double u = lo_hll_len(scratch, p, status);
        double sim = (card_r + card_q - u) / u;
        s = fmax(sim, 0.0); /* f64::max: NaN -> 0.0 */
    }
    return 2.0 * s / (1.0 + s);
}

/* All-vs-all: one task per reference sketch, serial over queries (utils.rs:150,248,342).
 * out is dense row-major [n_ref][n_qry] (f64, or f32 when fp32): final Mash distances
 * (compute_distance applied, main.rs:455).  triangular: only j <= i is computed (utils.rs:158,256,350
 * with idx = array position); other cells are left untouched.  The name-equality -> 0 rule
 * (main.rs:452-453) is the caller's business, exactly as at the GPU C ABI.
 * flags_out (optional, [n_ref][n_qry] bytes): 1 where the HLL bias-table regime was hit. */
struct dist_job {
    int algo, p, k, estimator, model, fp32, triangular;
    const uint8_t* ref;
    uint64_t n_ref;
    const uint8_t* qry;
    uint64_t n_qry;
    const double* card_r;
    const double* card_q;
    void* out;
    uint8_t* flags;
    volatile uint64_t next;
};
static void* dist_worker(void* arg) {
    struct dist_job* j = (struct dist_job*)arg;
    size_t rb = reg_bytes(j->algo, j->p);
    uint8_t* scratch = (uint8_t*)malloc(rb);
    for (;;) {
        uint64_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n_ref) break;
        uint64_t jmax = j->triangular ? (i + 1 < j->n_qry ? i + 1 : j->n_qry) : j->n_qry;
        for (uint64_t q = 0; q < jmax; ++q) {
            int st = 0;
            double frac = lo_pair_fraction(j->algo, j->p, j->estimator, j->ref + i * rb, j->qry + q * rb,
                                           j->card_r ? j->card_r[i] : 0.0, j->card_q ? j->card_q[q] : 0.0, scratch, &st);
            if (j->flags) j->flags[i * j->n_qry + q] = (uint8_t)st;
            if (j->fp32)
                ((float*)j->out)[i * j->n_qry + q] = lo_compute_distance_f32((float)frac, j->k, j->model);
            else
                ((double*)j->out)[i * j->n_qry + q] = lo_compute_distance_f64(frac, j->k, j->model);
        }
    }
    free(scratch);
    return NULL;
}
int lo_dist(int algo, int p, int k, int estimator, int model, int fp32, const void* ref, uint64_t n_ref,
            const void* qry, uint64_t n_qry, int triangular, void* out, uint8_t* flags_out, int threads) {
    if (model != 0 && model != 1 && model != 2) return -1; /* 2 = raw frac (test aid) */
    size_t rb = reg_bytes(algo, p);
    double *cr = NULL, *cq = NULL;
    if (algo != LO_HMH) {
        cr = (double*)malloc(sizeof(double) * (n_ref ? n_ref : 1));
        cq = (double*)malloc(sizeof(double) * (n_qry ? n_qry : 1));
        for (uint64_t i = 0; i < n_ref; ++i) cr[i] = lo_cardinality(algo, p, estimator, (const uint8_t*)ref + i * rb, NULL);
        for (uint64_t i = 0; i < n_qry; ++i) cq[i] = lo_cardinality(algo, p, estimator, (const uint8_t*)qry + i * rb, NULL);
    }
    struct dist_job job = {algo, p, k, estimator, model, fp32, triangular, (const uint8_t*)ref, n_ref,
                           (const uint8_t*)qry, n_qry, cr, cq, out, flags_out, 0};
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, dist_worker, &job);
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(cr);
    free(cq);
    return 0;
}
