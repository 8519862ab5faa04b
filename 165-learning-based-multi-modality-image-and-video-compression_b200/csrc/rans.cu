// rans.cu -- host-side range-ANS byte coder, bitstream-compatible with the reference's `compressai.ans`
// (compressai/cpp_exts/rans/rans_interface.cpp:108-284 on top of third_party/ryg_rans/rans64.h: 64-bit
// state, 32-bit renormalisation words, 16-bit probability precision, 4-bit bypass nibbles for symbols
// outside the CDF's range).  SURVEY.md section 8f row 1: with the symbols / indexes now produced on the
// GPU in microseconds, the coder is what remains of compress(); this version keeps ONE serial rANS state
// per image (byte-identical streams) but takes flat int32 buffers instead of Python lists and codes the
// images of a batch on parallel host threads.
#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace mmc {

constexpr int kPrecision = 16;          // rans_interface.cpp:49
constexpr int kBypassBits = 4;          // rans_interface.cpp:51
constexpr uint32_t kMaxBypass = (1u << kBypassBits) - 1;
constexpr uint64_t kRansL = 1ull << 31; // rans64.h:59

struct CdfTable {
    const int32_t *cdfs;
    int n_cdfs, stride;
    const int32_t *sizes, *offsets;
};

// One coded unit: (start, freq) against 2^bits.  Bypass nibbles are (val, 1) against 2^4, which is exactly
// what Rans64EncPutBits does (rans_interface.cpp:67-85: x = (x << nbits) | val).
struct Token {
    uint32_t start, freq, bits;
};

static inline void put(uint64_t &x, uint32_t *&ptr, const Token &t)
{
    const uint64_t x_max = ((kRansL >> t.bits) << 32) * t.freq;
    if (x >= x_max) {
        *--ptr = (uint32_t)x;
        x >>= 32;
    }
    x = ((x / t.freq) << t.bits) + (x % t.freq) + t.start;
}

// Tokens of one symbol in stream order (rans_interface.cpp:117-171).
static inline int tokens_of(int32_t symbol, int32_t index, const CdfTable &T, Token *out)
{
    const int32_t *cdf = T.cdfs + (size_t)index * T.stride;
    const int32_t max_value = T.sizes[index] - 2;
    int32_t value = symbol - T.offsets[index];
    uint32_t raw = 0;
    if (value < 0) {
        raw = (uint32_t)(-2 * value - 1);
        value = max_value;
    } else if (value >= max_value) {
        raw = (uint32_t)(2 * (value - max_value));
        value = max_value;
    }
    int n = 0;
    out[n++] = Token{(uint32_t)cdf[value], (uint32_t)(cdf[value + 1] - cdf[value]), (uint32_t)kPrecision};
    if (value == max_value) {
        int32_t n_bypass = 0;
        while (((uint64_t)raw >> (n_bypass * kBypassBits)) != 0) ++n_bypass;   // 64-bit shift: raw may need all 8 nibbles
        int32_t v = n_bypass;
        while (v >= (int32_t)kMaxBypass) {
            out[n++] = Token{kMaxBypass, 1, (uint32_t)kBypassBits};
            v -= kMaxBypass;
        }
        out[n++] = Token{(uint32_t)v, 1, (uint32_t)kBypassBits};
        for (int32_t j = 0; j < n_bypass; ++j) out[n++] = Token{(raw >> (j * kBypassBits)) & kMaxBypass, 1, (uint32_t)kBypassBits};
    }
    return n;
}

static int validate(const int32_t *indexes, int64_t n, const CdfTable &T, const char *name)
{
    MMC_CHECK_ARG(T.cdfs && T.sizes && T.offsets && T.n_cdfs >= 1 && T.stride >= 2, "%s: bad CDF table", name);
    for (int i = 0; i < T.n_cdfs; ++i)
        MMC_CHECK_ARG(T.sizes[i] >= 2 && T.sizes[i] <= T.stride, "%s: cdf length %d of row %d outside [2, %d]", name, T.sizes[i], i, T.stride);
    for (int64_t i = 0; i < n; ++i)
        MMC_CHECK_ARG(indexes[i] >= 0 && indexes[i] < T.n_cdfs, "%s: index %d at %lld outside [0, %d)", name, indexes[i], (long long)i, T.n_cdfs);
    return MMC_OK;
}

// ---- fast path: reciprocal-multiply encoder --------------------------------------------------------------------------
// The serial dependency of rANS encoding is x -> x / freq.  The reference divides (rans64.h Rans64EncPut); with one
// precomputed entry per (CDF row, value) the quotient is an exact multiply-high + shift (Granlund-Montgomery, the
// RansEncSymbol idea of ryg_rans): for 2 <= freq <= 2^16 and x < freq * 2^47 <= 2^63,
//     floor(x / freq) == mulhi64(x, ceil(2^(63 + s) / freq)) >> (s - 1),   s = ceil(log2 freq)
// (checked exhaustively over freq with edge and random x in tests/test_host_cpu.py via mmc_rans_selftest).  freq == 1
// uses rcp = 2^64 - 1 (mulhi gives x - 1) and a bias of 2^16 - 1, as ryg does.  The stream is byte-identical.
typedef unsigned __int128 u128;
struct EncEntry {
    uint64_t rcp;
    uint32_t bias;        // start (+ 2^16 - 1 when freq == 1)
    uint32_t freq_shift;  // freq in bits 0..19, shift in bits 24..31
};
static inline uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a * b) >> 64); }

static inline EncEntry make_entry(uint32_t start, uint32_t freq)
{
    EncEntry e;
    if (freq < 2) {
        e.rcp = ~0ull; e.bias = start + (1u << kPrecision) - 1; e.freq_shift = freq;
    } else {
        uint32_t shift = 0;
        while (freq > (1u << shift)) ++shift;
        e.rcp = (uint64_t)((((u128)1 << (shift + 63)) + freq - 1) / freq);
        e.bias = start;
        e.freq_shift = freq | ((shift - 1) << 24);
    }
    return e;
}

// Encoder tables derived from one set of CDFs; cached by content hash (the Python side passes freshly copied tables on
// every call, and building ~100k reciprocals costs a few ms).
struct EncTable {
    uint64_t hash = 0;
    std::vector<EncEntry> entries;       // [row base + value]
    std::vector<uint32_t> row_base;
    std::vector<int32_t> max_value, offset;
    bool ok = true;
};

static uint64_t hash_words(const int32_t *p, size_t n, uint64_t h)
{
    for (size_t i = 0; i < n; ++i) { h ^= (uint32_t)p[i]; h *= 0x100000001b3ull; h ^= h >> 29; }
    return h;
}

static std::shared_ptr<const EncTable> get_enc_table(const CdfTable &T)
{
    static std::mutex mu;
    static std::vector<std::shared_ptr<const EncTable>> cache;
    uint64_t h = 0xcbf29ce484222325ull ^ ((uint64_t)T.n_cdfs << 32) ^ (uint64_t)T.stride;
    h = hash_words(T.sizes, T.n_cdfs, h);
    h = hash_words(T.offsets, T.n_cdfs, h);
    for (int i = 0; i < T.n_cdfs; ++i) h = hash_words(T.cdfs + (size_t)i * T.stride, (size_t)T.sizes[i], h);
    {
        std::lock_guard<std::mutex> g(mu);
        for (auto &t : cache) if (t->hash == h) return t;
    }
    auto t = std::make_shared<EncTable>();
    t->hash = h;
    t->row_base.resize(T.n_cdfs); t->max_value.resize(T.n_cdfs); t->offset.resize(T.n_cdfs);
    size_t total = 0;
    for (int i = 0; i < T.n_cdfs; ++i) { t->row_base[i] = (uint32_t)total; total += (size_t)T.sizes[i] - 1; }
    t->entries.resize(total);
    for (int i = 0; i < T.n_cdfs; ++i) {
        const int32_t *cdf = T.cdfs + (size_t)i * T.stride;
        t->max_value[i] = T.sizes[i] - 2;
        t->offset[i] = T.offsets[i];
        for (int v = 0; v + 1 < T.sizes[i]; ++v) {
            const int64_t freq = (int64_t)cdf[v + 1] - cdf[v];
            if (freq < 1 || freq > (1 << kPrecision) || cdf[v] < 0) { t->ok = false; t->entries[t->row_base[i] + v] = make_entry(0, 1); continue; }
            t->entries[t->row_base[i] + v] = make_entry((uint32_t)cdf[v], (uint32_t)freq);
        }
    }
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() >= 16) cache.erase(cache.begin());
    cache.push_back(t);
    return t;
}

// Output words are produced back to front.  Every coded unit emits at most one word, so n + slack words always hold the
// regular symbols; the (rare) bypass symbols check the remaining room and grow the buffer when they need to.
struct BackBuffer {
    std::unique_ptr<uint32_t[]> buf;     // uninitialised: only the words actually produced are ever read
    size_t cap;
    uint32_t *ptr, *end;
    explicit BackBuffer(size_t words) : buf(new uint32_t[words]), cap(words) { end = buf.get() + cap; ptr = end; }
    void ensure(size_t words)
    {
        if ((size_t)(ptr - buf.get()) >= words) return;
        const size_t used = end - ptr, bigger = cap * 2 + words;
        std::unique_ptr<uint32_t[]> nb(new uint32_t[bigger]);
        memcpy(nb.get() + bigger - used, ptr, used * sizeof(uint32_t));
        buf.swap(nb);
        cap = bigger;
        end = buf.get() + cap;
        ptr = end - used;
    }
    size_t bytes() const { return (size_t)(end - ptr) * sizeof(uint32_t); }
};

// Encodes one stream back to front into `out`; returns false on an index outside the table.
static bool encode_one(const int32_t *symbols, const int32_t *indexes, int64_t n, const CdfTable &T, const EncTable &E, BackBuffer &out)
{
    uint64_t x = kRansL;
    Token tmp[16];
    const EncEntry *entries = E.entries.data();
    const uint32_t *row_base = E.row_base.data();
    const int32_t *max_value = E.max_value.data(), *offset = E.offset.data();
    const uint32_t n_cdfs = (uint32_t)T.n_cdfs;
    uint32_t *ptr = out.ptr;                          // at least 64 words of slack below the n regular words
    constexpr int64_t kAhead = 12;                    // table entries are fetched a few symbols ahead of the serial state chain
    for (int64_t i = n - 1; i >= 0; --i) {          // rANS is LIFO: code the last symbol first
        if (i >= kAhead) {
            const uint32_t pi = (uint32_t)indexes[i - kAhead];
            if (pi < n_cdfs) {
                const int32_t pv = symbols[i - kAhead] - offset[pi];
                if ((uint32_t)pv < (uint32_t)max_value[pi]) __builtin_prefetch(entries + row_base[pi] + (uint32_t)pv);
            }
        }
        const uint32_t idx = (uint32_t)indexes[i];
        if (idx >= n_cdfs) return false;
        const int32_t v = symbols[i] - offset[idx];
        if ((uint32_t)v < (uint32_t)max_value[idx]) {
            const EncEntry e = entries[row_base[idx] + (uint32_t)v];
            const uint32_t freq = e.freq_shift & 0xFFFFFu, shift = e.freq_shift >> 24;
            // renormalise without a data-dependent branch: the word is always stored, the pointer moves only when x >= x_max
            // (x_max = ((L >> 16) << 32) * freq)
            const bool r = x >= ((uint64_t)freq << (31 - kPrecision + 32));
            ptr[-1] = (uint32_t)x;
            ptr -= r;
            x = r ? (x >> 32) : x;
            const uint64_t q = mulhi64(x, e.rcp) >> shift;
            x += e.bias + q * ((1u << kPrecision) - freq);              // (q << 16) + (x - q * freq) + start
        } else {
            // escape symbol + bypass nibbles (rans_interface.cpp:117-171): rare, keeps the division-based path
            out.ptr = ptr;
            out.ensure(16 + 64);
            ptr = out.ptr;
            const int nt = tokens_of(symbols[i], indexes[i], T, tmp);
            for (int t = nt - 1; t >= 0; --t) put(x, ptr, tmp[t]);
        }
    }
    *--ptr = (uint32_t)(x >> 32);                    // Rans64EncFlush: low word first in the stream
    *--ptr = (uint32_t)x;
    out.ptr = ptr;
    return true;
}

static inline uint32_t get_bits(uint64_t &x, const uint32_t *&ptr, const uint32_t *end, uint32_t nbits, bool &ok)
{
    const uint32_t val = (uint32_t)(x & ((1u << nbits) - 1));
    x >>= nbits;
    if (x < kRansL) {
        if (ptr >= end) { ok = false; return val; }
        x = (x << 32) | *ptr++;
    }
    return val;
}

// Decoder side table: for every CDF row the symbol at which the search for a 16-bit cumulative value starts, per 256-wide bucket
// of that value (the reference scans the row linearly from 0, rans_interface.cpp:246-249; a binary search per symbol was the hot
// spot of decode: ~36 ns per symbol).  Cached by content hash like the encoder table.
struct DecTable {
    uint64_t hash = 0;
    std::vector<uint16_t> start;    // [row][256]
};

static std::shared_ptr<const DecTable> get_dec_table(const CdfTable &T)
{
    static std::mutex mu;
    static std::vector<std::shared_ptr<const DecTable>> cache;
    uint64_t h = 0x9ae16a3b2f90404full ^ ((uint64_t)T.n_cdfs << 32) ^ (uint64_t)T.stride;
    h = hash_words(T.sizes, T.n_cdfs, h);
    for (int i = 0; i < T.n_cdfs; ++i) h = hash_words(T.cdfs + (size_t)i * T.stride, (size_t)T.sizes[i], h);
    {
        std::lock_guard<std::mutex> g(mu);
        for (auto &t : cache) if (t->hash == h) return t;
    }
    auto t = std::make_shared<DecTable>();
    t->hash = h;
    t->start.resize((size_t)T.n_cdfs * 256);
    for (int i = 0; i < T.n_cdfs; ++i) {
        const int32_t *cdf = T.cdfs + (size_t)i * T.stride;
        const int32_t len = T.sizes[i];
        for (int b = 0; b < 256; ++b) {
            const int32_t *it = std::upper_bound(cdf, cdf + len, (int32_t)(b << 8));
            const int64_t sidx = (it - cdf) - 1;
            t->start[(size_t)i * 256 + b] = (uint16_t)(sidx < 0 ? 0 : (sidx > 65535 ? 65535 : sidx));
        }
    }
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() >= 16) cache.erase(cache.begin());
    cache.push_back(t);
    return t;
}

static bool decode_one(const uint32_t *ptr, const uint32_t *end, const int32_t *indexes, int64_t n, const CdfTable &T, const DecTable &D, int32_t *out)
{
    if (end - ptr < 2) return false;
    uint64_t x = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32);
    ptr += 2;
    bool ok = true;
    const uint16_t *start_tab = D.start.data();
    for (int64_t i = 0; i < n && ok; ++i) {
        const int32_t idx = indexes[i];
        const int32_t *cdf = T.cdfs + (size_t)idx * T.stride;
        const int32_t len = T.sizes[idx], max_value = len - 2;
        const uint32_t cum = (uint32_t)(x & ((1u << kPrecision) - 1));
        // largest s with cdf[s] <= cum (the row is increasing): start from the bucket's first candidate, walk up
        int32_t s = start_tab[(size_t)idx * 256 + (cum >> 8)];
        if (cdf[s] > (int32_t)cum) return false;                 // cum below the row's first entry: not a stream of these tables
        while (s + 1 < len && cdf[s + 1] <= (int32_t)cum) ++s;
        if (s > max_value) return false;
        const uint32_t start = (uint32_t)cdf[s], freq = (uint32_t)(cdf[s + 1] - cdf[s]);
        x = freq * (x >> kPrecision) + (x & ((1u << kPrecision) - 1)) - start;
        if (x < kRansL) {
            if (ptr >= end) return false;
            x = (x << 32) | *ptr++;
        }
        int32_t value = s;
        if (value == max_value) {
            int32_t val = (int32_t)get_bits(x, ptr, end, kBypassBits, ok);
            int32_t n_bypass = val;
            while (ok && val == (int32_t)kMaxBypass) {
                val = (int32_t)get_bits(x, ptr, end, kBypassBits, ok);
                n_bypass += val;
            }
            if (n_bypass > 8) return false;   // more than 32 raw bits cannot come from an int32 symbol
            uint32_t raw = 0;
            for (int j = 0; j < n_bypass && ok; ++j) raw |= get_bits(x, ptr, end, kBypassBits, ok) << (j * kBypassBits);
            value = (int32_t)(raw >> 1);
            if (raw & 1) value = -value - 1;
            else value += max_value;
        }
        out[i] = value + T.offsets[idx];
    }
    return ok;
}

template <typename F>
static void parallel_for(int n, F f)
{
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::min<unsigned>(hw ? hw : 1, (unsigned)n);
    if (nt <= 1) { for (int i = 0; i < n; ++i) f(i); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([=]() { for (int i = t; i < n; i += nt) f(i); });
    for (auto &t : th) t.join();
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_rans_encode_batch_host(const int32_t *symbols, const int32_t *indexes, int batch, int64_t n, const int32_t *cdfs,
                               int n_cdfs, int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets, uint8_t *out,
                               size_t cap_per_stream, size_t *nbytes)
{
    const char *name = "mmc_rans_encode_batch_host";
    MMC_CHECK_ARG(batch >= 0 && n >= 0 && nbytes, "%s: bad argument", name);
    if (batch == 0) return MMC_OK;
    MMC_CHECK_ARG(symbols && indexes, "%s: NULL buffer", name);
    CdfTable T{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
    int rc = validate(indexes, 0, T, name);          // the table; the indexes are range-checked inside the coding loop
    if (rc) return rc;
    std::shared_ptr<const EncTable> E = get_enc_table(T);
    MMC_CHECK_ARG(E->ok, "%s: CDF rows must be strictly increasing with steps of at most 2^%d", name, kPrecision);
    std::vector<std::unique_ptr<BackBuffer>> streams(batch);
    std::vector<int> good(batch, 1);
    parallel_for(batch, [&](int b) {
        streams[b].reset(new BackBuffer((size_t)n + 128));
        good[b] = encode_one(symbols + (size_t)b * n, indexes + (size_t)b * n, n, T, *E, *streams[b]) ? 1 : 0;
    });
    for (int b = 0; b < batch; ++b)
        MMC_CHECK_ARG(good[b], "%s: stream %d has an index outside [0, %d)", name, b, n_cdfs);
    bool fits = out != nullptr;
    for (int b = 0; b < batch; ++b) {
        nbytes[b] = streams[b]->bytes();
        fits = fits && nbytes[b] <= cap_per_stream;
    }
    if (!fits) {
        set_error("%s: output capacity %zu bytes per stream is too small (sizes returned in nbytes)", name, cap_per_stream);
        return MMC_EINVAL;
    }
    for (int b = 0; b < batch; ++b) memcpy(out + (size_t)b * cap_per_stream, streams[b]->ptr, nbytes[b]);
    return MMC_OK;
}

// Exhaustive-over-freq check of the reciprocal quotient against the division it replaces (CPU test hook).
int64_t mmc_rans_selftest(void)
{
    int64_t bad = 0;
    uint64_t r = 0x9E3779B97F4A7C15ull;
    for (uint32_t freq = 1; freq <= (1u << kPrecision); ++freq) {
        const EncEntry e = make_entry(0, freq);
        const uint64_t x_max = (uint64_t)freq << 47;
        const uint64_t edge[8] = {1, freq, freq + 1ull, x_max - 1, x_max - freq, x_max / 2, kRansL, kRansL + freq - 1};
        for (int k = 0; k < 24; ++k) {
            r ^= r << 13; r ^= r >> 7; r ^= r << 17;
            const uint64_t x = k < 8 ? edge[k] : r % x_max;
            if (x == 0 || x >= x_max) continue;
            const uint64_t q = mulhi64(x, e.rcp) >> (e.freq_shift >> 24);
            const uint64_t got = x + e.bias + q * ((1u << kPrecision) - freq);
            const uint64_t want = ((x / freq) << kPrecision) + (x % freq);
            bad += got != want;
        }
    }
    return bad;
}

int mmc_rans_decode_batch_host(const uint8_t *streams, const size_t *stream_offsets, const size_t *nbytes, const int32_t *indexes,
                               int batch, int64_t n, const int32_t *cdfs, int n_cdfs, int cdf_stride, const int32_t *cdf_sizes,
                               const int32_t *offsets, int32_t *symbols_out)
{
    const char *name = "mmc_rans_decode_batch_host";
    MMC_CHECK_ARG(batch >= 0 && n >= 0, "%s: bad argument", name);
    if (batch == 0) return MMC_OK;
    MMC_CHECK_ARG(streams && stream_offsets && nbytes && indexes && symbols_out, "%s: NULL buffer", name);
    CdfTable T{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
    int rc = validate(indexes, (int64_t)batch * n, T, name);
    if (rc) return rc;
    for (int b = 0; b < batch; ++b) MMC_CHECK_ARG(nbytes[b] % 4 == 0, "%s: stream %d length %zu is not a multiple of 4", name, b, nbytes[b]);
    std::shared_ptr<const DecTable> D = get_dec_table(T);
    std::vector<int> ok(batch, 1);
    parallel_for(batch, [&](int b) {
        std::vector<uint32_t> w(nbytes[b] / 4);     // copy: the byte stream need not be 4-byte aligned
        memcpy(w.data(), streams + stream_offsets[b], nbytes[b]);
        ok[b] = decode_one(w.data(), w.data() + w.size(), indexes + (size_t)b * n, n, T, *D, symbols_out + (size_t)b * n) ? 1 : 0;
    });
    for (int b = 0; b < batch; ++b)
        if (!ok[b]) { set_error("%s: stream %d is truncated or does not match the CDF tables", name, b); return MMC_EINVAL; }
    return MMC_OK;
}

}  // extern "C"
