// rans.cu -- host-side range-ANS byte coder, bitstream-compatible with the reference's `compressai.ans`
// (compressai/cpp_exts/rans/rans_interface.cpp:108-284 on top of third_party/ryg_rans/rans64.h: 64-bit
// state, 32-bit renormalisation words, 16-bit probability precision, 4-bit bypass nibbles for symbols
// outside the CDF's range).  SURVEY.md section 8f row 1: with the symbols / indexes now produced on the
// GPU in microseconds, the coder is what remains of compress(); this version keeps ONE serial rANS state
// per image (byte-identical streams) but takes flat int32 buffers instead of Python lists and codes the
// images of a batch on parallel host threads.
#include <algorithm>
#include <cstring>
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

// Encodes one stream; returns its 32-bit words (front = first word of the stream).
static void encode_one(const int32_t *symbols, const int32_t *indexes, int64_t n, const CdfTable &T, std::vector<uint32_t> &words)
{
    // upper bound on emitted words: one per token (each put renormalises at most once) + 2 for the flush
    size_t ntok = 0;
    Token tmp[16];
    for (int64_t i = 0; i < n; ++i) {
        const int32_t max_value = T.sizes[indexes[i]] - 2;
        const int32_t v = symbols[i] - T.offsets[indexes[i]];
        ntok += (v < 0 || v >= max_value) ? 12 : 1;
    }
    std::vector<uint32_t> buf(ntok + 2);
    uint32_t *end = buf.data() + buf.size(), *ptr = end;
    uint64_t x = kRansL;
    for (int64_t i = n - 1; i >= 0; --i) {          // rANS is LIFO: code the last symbol first
        const int nt = tokens_of(symbols[i], indexes[i], T, tmp);
        for (int t = nt - 1; t >= 0; --t) put(x, ptr, tmp[t]);
    }
    *--ptr = (uint32_t)(x >> 32);                    // Rans64EncFlush: low word first in the stream
    *--ptr = (uint32_t)x;
    words.assign(ptr, end);
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

static bool decode_one(const uint32_t *ptr, const uint32_t *end, const int32_t *indexes, int64_t n, const CdfTable &T, int32_t *out)
{
    if (end - ptr < 2) return false;
    uint64_t x = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32);
    ptr += 2;
    bool ok = true;
    for (int64_t i = 0; i < n && ok; ++i) {
        const int32_t idx = indexes[i];
        const int32_t *cdf = T.cdfs + (size_t)idx * T.stride;
        const int32_t len = T.sizes[idx], max_value = len - 2;
        const uint32_t cum = (uint32_t)(x & ((1u << kPrecision) - 1));
        // first entry > cum (the reference scans linearly, rans_interface.cpp:246-249; the row is increasing)
        const int32_t *it = std::upper_bound(cdf, cdf + len, (int32_t)cum);
        const int32_t s = (int32_t)(it - cdf) - 1;
        if (s < 0 || s > max_value) return false;
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
    int rc = validate(indexes, (int64_t)batch * n, T, name);
    if (rc) return rc;
    std::vector<std::vector<uint32_t>> streams(batch);
    parallel_for(batch, [&](int b) { encode_one(symbols + (size_t)b * n, indexes + (size_t)b * n, n, T, streams[b]); });
    bool fits = out != nullptr;
    for (int b = 0; b < batch; ++b) {
        nbytes[b] = streams[b].size() * sizeof(uint32_t);
        fits = fits && nbytes[b] <= cap_per_stream;
    }
    if (!fits) {
        set_error("%s: output capacity %zu bytes per stream is too small (sizes returned in nbytes)", name, cap_per_stream);
        return MMC_EINVAL;
    }
    for (int b = 0; b < batch; ++b) memcpy(out + (size_t)b * cap_per_stream, streams[b].data(), nbytes[b]);
    return MMC_OK;
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
    std::vector<int> ok(batch, 1);
    parallel_for(batch, [&](int b) {
        std::vector<uint32_t> w(nbytes[b] / 4);     // copy: the byte stream need not be 4-byte aligned
        memcpy(w.data(), streams + stream_offsets[b], nbytes[b]);
        ok[b] = decode_one(w.data(), w.data() + w.size(), indexes + (size_t)b * n, n, T, symbols_out + (size_t)b * n) ? 1 : 0;
    });
    for (int b = 0; b < batch; ++b)
        if (!ok[b]) { set_error("%s: stream %d is truncated or does not match the CDF tables", name, b); return MMC_EINVAL; }
    return MMC_OK;
}

}  // extern "C"
