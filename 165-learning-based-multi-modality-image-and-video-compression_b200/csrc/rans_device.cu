// rans_device.cu -- range-ANS coder ON THE DEVICE ("GPU/parallel rANS").
//
// The reference codes one image as ONE serial rANS stream (compressai/cpp_exts/rans/rans_interface.cpp:108-200: 64-bit state,
// 32-bit words) -- a dependency chain over every symbol of the image that no device can shorten, which is why csrc/rans.cu keeps it
// on host threads for byte-identical streams.  This file is the other half of that row: a container of our own ("lane container",
// magic "MMCL") in which the symbols of an image are dealt round-robin onto S independent rANS lanes, so that the symbols never
// leave the GPU and compress() stops being bound by the host coder.  Same probability model as the reference coder: the model's
// 16-bit quantised CDF tables (entropy_models.py:206-214), symbol -> CDF row through the same `indexes`, the same escape scheme
// for values outside a row's range (4-bit bypass nibbles: rans_interface.cpp:117-171, restated in csrc/rans.cu tokens_of), so the
// rate is the reference's plus the lane headers.  Streams of the two containers are NOT interchangeable; the host coder stays
// the default and the reference-compatible one.
//
// Container (little endian), one per image and latent tensor:
//   u32 magic "MMCL" | u32 n symbols | u32 S lanes | u32 0 | u32 state[S] | u32 words[S] | u16 payload[sum(words)] | pad to 4 bytes
// Symbol i (logical order) is position i / S of lane i % S.  A lane is a classic word-renormalised rANS (32-bit state x in
// [2^16, 2^32), 16-bit words): the encoder walks its tokens LAST to first with, per token (start, freq) against 2^bits,
//       if (x >= ((2^16 >> bits) << 16) * freq) { emit low 16 bits of x; x >>= 16; }       x = ((x / freq) << bits) + x % freq + start
// starting from x = 2^16; state[l] is the final x, payload words are stored in the order the decoder reads them (reverse of
// emission).  The decoder starts from state[l] and per token does  slot = x & (2^bits - 1);  x = freq * (x >> bits) + slot - start;
// if (x < 2^16) x = (x << 16) | next word.  The tests hold a CPU restatement of exactly this text and compare bytes.
//
// Kernels: lanes are threads (32 lanes per warp read symbols / indexes coalesced); the per-lane chain is serial, so the loads of a
// chunk of 32 steps (symbol, index -> CDF entry) are issued ahead of the chain and only the state arithmetic stays on it.
//   table   rans_lane_table:         (start, freq, 32-bit reciprocal of freq) per CDF entry, so the chain has no division
//   pass 1  rans_lane_encode<false>: final state and word count of every lane        (no stores)
//   scan    rans_lane_header:        lane offsets, header, total size, capacity check  (one CTA per image)
//   pass 2  rans_lane_encode<true>:  the same chain again, words written in place
#include "common.cuh"

namespace mmc {

constexpr int kLanePrecision = 16;
constexpr int kLaneBypassBits = 4;
constexpr uint32_t kLaneMaxBypass = (1u << kLaneBypassBits) - 1;
constexpr uint32_t kLaneL = 1u << 16;
constexpr uint32_t kLaneMagic = 0x4C434D4Du;   // "MMCL"
constexpr int kLaneChunk = 32;          // decoder: indexes prefetched per chunk
constexpr int kMaxLanes = 1024;

enum LaneStatus { LANE_OK = 0, LANE_BAD_INDEX = 1, LANE_BAD_CDF = 2, LANE_CAPACITY = 4, LANE_BAD_STREAM = 8 };

struct LaneTables {
    const int32_t *cdfs;
    int n_cdfs, stride;
    const int32_t *sizes, *offsets;
};

template <bool kWrite>
__device__ __forceinline__ void lane_put(uint32_t &x, uint16_t *&ptr, uint32_t &count, uint32_t start, uint32_t freq, uint32_t bits)
{
    const uint64_t x_max = (uint64_t)((kLaneL >> bits) << 16) * freq;
    if ((uint64_t)x >= x_max) {
        if (kWrite) *--ptr = (uint16_t)(x & 0xffffu);
        ++count;
        x >>= 16;
    }
    const uint32_t q = x / freq;
    x = (q << bits) + (x - q * freq) + start;
}

// One 8-byte entry per (CDF row, value): what the chain needs for the 16-bit main token, prepared off the chain.
//   tok = start << 16 | (freq - 1);  rcp = floor(2^32 / freq) (2^32 - 1 for freq == 1): umulhi(x, rcp) is x / freq or one less for every
//   x < 2^32 (x rcp / 2^32 lies in (x / freq - 1, x / freq]), so ONE compare-and-fix makes the quotient exact;  rcp == 0: invalid entry.
struct __align__(8) LaneEntry {
    uint32_t tok, rcp;
};

// the main token on the chain: ~11 dependent instructions, no division
template <bool kWrite>
__device__ __forceinline__ void lane_put_main(uint32_t &x, uint16_t *&ptr, uint32_t &count, uint32_t tok, uint32_t rcp)
{
    const uint32_t start = tok >> 16, freq = (tok & 0xffffu) + 1;
    if ((x >> 16) >= freq) {                       // x >= freq << 16
        if (kWrite) *--ptr = (uint16_t)(x & 0xffffu);
        ++count;
        x >>= 16;
    }
    uint32_t q = __umulhi(x, rcp), r = x - q * freq;
    if (r >= freq) { ++q; r -= freq; }
    x = (q << kLanePrecision) + r + start;
}

__global__ void __launch_bounds__(256) rans_lane_table_kernel(const int32_t *__restrict__ cdfs, const int32_t *__restrict__ sizes, int n_cdfs, int stride,
                                                              LaneEntry *__restrict__ tab)
{
    const int64_t n = (int64_t)n_cdfs * stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / stride), v = (int)(i - (int64_t)row * stride);
        LaneEntry e;
        e.tok = 0; e.rcp = 0;
        const int len = sizes[row];
        if (len >= 2 && len <= stride && v + 1 < len) {
            const int32_t c0 = cdfs[i], c1 = cdfs[i + 1];
            if (c0 >= 0 && c1 > c0 && c1 - c0 <= (1 << kLanePrecision) && c0 < (1 << kLanePrecision)) {
                const uint32_t freq = (uint32_t)(c1 - c0);
                e.tok = ((uint32_t)c0 << 16) | (freq - 1);
                e.rcp = freq == 1 ? 0xffffffffu : (uint32_t)((1ull << 32) / freq);
            }
        }
        tab[i] = e;
    }
}

// all tokens of an escaped symbol, last token first (the forward order is: main token, count nibbles, raw nibbles).  The lane
// state goes in and out BY VALUE: reference parameters of a non-inlined function would pin x / count / ptr to the stack
struct LaneState {
    uint32_t x, count;
    uint16_t *ptr;
};
template <bool kWrite>
__device__ __noinline__ LaneState lane_put_escape(LaneState st, uint32_t raw, uint32_t tok, uint32_t rcp)
{
    uint32_t x = st.x, count = st.count;
    uint16_t *ptr = st.ptr;
    int n_bypass = 0;
    while (n_bypass < 8 && (raw >> (n_bypass * kLaneBypassBits)) != 0) ++n_bypass;
    for (int j = n_bypass - 1; j >= 0; --j) lane_put<kWrite>(x, ptr, count, (raw >> (j * kLaneBypassBits)) & kLaneMaxBypass, 1, kLaneBypassBits);
    // count nibbles forward: 15 while the remainder >= 15, then the remainder
    int full = 0, v = n_bypass;
    while (v >= (int)kLaneMaxBypass) { ++full; v -= kLaneMaxBypass; }
    lane_put<kWrite>(x, ptr, count, (uint32_t)v, 1, kLaneBypassBits);
    for (int j = 0; j < full; ++j) lane_put<kWrite>(x, ptr, count, kLaneMaxBypass, 1, kLaneBypassBits);
    lane_put_main<kWrite>(x, ptr, count, tok, rcp);
    return LaneState{x, count, ptr};
}

constexpr int kLaneSmemRows = 512;      // CDF rows whose length / offset are staged in shared memory (more: read from global memory)
constexpr int kCh = 16;                 // steps per chunk of the software pipeline below

struct LaneChunk {                      // one chunk's operands, ready for the chain
    uint2 ent[kCh];                     // (tok, rcp) table entries
    uint32_t raw[kCh];                  // raw bits of escaped symbols
    uint32_t esc;                       // bit k: symbol k is escaped
};

// positions hi - 1 - k (k = 0 .. kCh - 1) of this lane, clamped to its first position: always a valid address, extras are ignored
__device__ __forceinline__ void lane_load(const int32_t *__restrict__ sym, const int32_t *__restrict__ idx, int lane, int S, int64_t hi,
                                          int32_t (&sv)[kCh], int32_t (&iv)[kCh])
{
#pragma unroll
    for (int k = 0; k < kCh; ++k) {
        int64_t step = hi - 1 - k;
        step = step < 0 ? 0 : step;
        const int64_t i = (int64_t)lane + step * S;
        sv[k] = __ldg(sym + i);
        iv[k] = __ldg(idx + i);
    }
}

// symbol / index -> table entry, branch-free: out-of-range indexes and row lengths are clamped and FLAGGED, never skipped
__device__ __forceinline__ void lane_prep(const LaneTables &T, const LaneEntry *__restrict__ tab, const int32_t *s_max, const int32_t *s_offs, bool in_smem,
                                          const int32_t (&sv)[kCh], const int32_t (&iv)[kCh], LaneChunk &c, int &bad)
{
    c.esc = 0;
#pragma unroll
    for (int k = 0; k < kCh; ++k) {
        const int32_t ix = min(max(iv[k], 0), T.n_cdfs - 1);
        bad |= (ix != iv[k]) ? LANE_BAD_INDEX : 0;
        const int32_t mv = in_smem ? s_max[ix] : __ldg(T.sizes + ix) - 2;
        const int32_t ov = in_smem ? s_offs[ix] : __ldg(T.offsets + ix);
        const int32_t max_value = min(max(mv, 0), T.stride - 2);
        bad |= (max_value != mv) ? LANE_BAD_CDF : 0;
        const int32_t value = sv[k] - ov;
        const bool neg = value < 0, over = value >= max_value;
        // raw = -2 value - 1 (value < 0), 2 (value - max_value) (value >= max_value): unsigned 32-bit forms of rans_interface.cpp:124-131
        const uint32_t raw_neg = (uint32_t)(-(value + 1)) * 2u + 1u, raw_over = (uint32_t)(value - max_value) * 2u;
        c.raw[k] = neg ? raw_neg : (over ? raw_over : 0u);
        c.esc |= (neg || over) ? (1u << k) : 0u;
        const int32_t v = (neg || over) ? max_value : value;
        c.ent[k] = __ldg(reinterpret_cast<const uint2 *>(tab + (size_t)ix * T.stride + v));
    }
}

// Encoder of one lane per thread.  A lone warp per SM issues a DEPENDENT instruction only every ~4-5 cycles and has no other warp
// to hide a memory latency behind, so the loop is a three-deep software pipeline over chunks of kCh steps: while the chain of
// chunk c runs on registers, the table entries of chunk c + 1 are being fetched (their addresses computed from the symbols /
// indexes loaded one iteration earlier) and the symbols / indexes of chunk c + 2 are being loaded.  The preparation of the next
// chunk sits in the SAME basic block as the (branch-free) chain of the current one, so the instruction scheduler fills the
// chain's latency slots with it; loads are consumed one loop iteration after they are issued, so they cannot sink to their uses
// (without this structure every symbol paid its own two memory round trips: measured 820 cycles per symbol, then ~500 with the
// phases merely separated).  Chunks that contain an escaped symbol or are ragged take the per-step path.
template <bool kWrite>
__global__ void __launch_bounds__(32) rans_lane_encode_kernel(const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes, int64_t n, int S,
                                                              LaneTables T, const LaneEntry *__restrict__ tab, uint32_t *__restrict__ states,
                                                              uint32_t *__restrict__ words, const uint32_t *__restrict__ lane_off,
                                                              uint8_t *__restrict__ out, size_t cap, int *__restrict__ status)
{
    __shared__ int32_t s_max[kLaneSmemRows], s_offs[kLaneSmemRows];      // per row: length - 2, offset
    const int b = blockIdx.y, lane = blockIdx.x * 32 + threadIdx.x;
    const bool in_smem = T.n_cdfs <= kLaneSmemRows;
    if (in_smem) {
        for (int i = threadIdx.x; i < T.n_cdfs; i += 32) { s_max[i] = __ldg(T.sizes + i) - 2; s_offs[i] = __ldg(T.offsets + i); }
        __syncwarp();
    }
    if (lane >= S) return;
    const int32_t *sym = symbols + (int64_t)b * n, *idx = indexes + (int64_t)b * n;
    const int64_t steps = lane < n ? (n - lane + S - 1) / S : 0;
    uint32_t x = kLaneL, count = 0;
    uint16_t *ptr = nullptr;
    if (kWrite) {
        if (*status != LANE_OK) return;
        const size_t payload = 16 + (size_t)8 * S;
        ptr = reinterpret_cast<uint16_t *>(out + (size_t)b * cap + payload) + lane_off[(size_t)b * S + lane] + words[(size_t)b * S + lane];
    }
    int bad = 0;
    if (steps > 0) {
        int32_t sv[kCh], iv[kCh];
        LaneChunk cur, nxt;
        int64_t hi = steps;                                  // the current chunk covers steps [hi - kCh, hi) of this lane, last first
        lane_load(sym, idx, lane, S, hi, sv, iv);
        lane_prep(T, tab, s_max, s_offs, in_smem, sv, iv, cur, bad);
        lane_load(sym, idx, lane, S, hi - kCh, sv, iv);
        for (; hi > 0; hi -= kCh) {
            const int m = hi < kCh ? (int)hi : kCh;
            if (cur.esc == 0 && m == kCh) {
                // ---- the common case: one basic block = preparation of chunk c + 1, loads of chunk c + 2, chain of chunk c ----
                lane_prep(T, tab, s_max, s_offs, in_smem, sv, iv, nxt, bad);
                lane_load(sym, idx, lane, S, hi - 2 * kCh, sv, iv);
#pragma unroll
                for (int k = 0; k < kCh; ++k) lane_put_main<kWrite>(x, ptr, count, cur.ent[k].x, cur.ent[k].y);
            } else {
                lane_prep(T, tab, s_max, s_offs, in_smem, sv, iv, nxt, bad);
                lane_load(sym, idx, lane, S, hi - 2 * kCh, sv, iv);
#pragma unroll
                for (int k = 0; k < kCh; ++k) {          // static indices: the chunk stays in registers
                    if (k >= m) continue;
                    if ((cur.esc >> k) & 1u) {
                        const LaneState r = lane_put_escape<kWrite>(LaneState{x, count, ptr}, cur.raw[k], cur.ent[k].x, cur.ent[k].y);
                        x = r.x; count = r.count; ptr = r.ptr;
                    } else {
                        lane_put_main<kWrite>(x, ptr, count, cur.ent[k].x, cur.ent[k].y);
                    }
                }
            }
            cur = nxt;
            if (bad) break;
        }
    }
    if (bad) atomicOr(status, bad);
    if (!kWrite) {
        states[(size_t)b * S + lane] = x;
        words[(size_t)b * S + lane] = count;
    }
}

// one CTA per image: exclusive scan of the lane word counts, header, total size
__global__ void __launch_bounds__(256) rans_lane_header_kernel(int64_t n, int S, const uint32_t *__restrict__ states, const uint32_t *__restrict__ words,
                                                               uint32_t *__restrict__ lane_off, uint8_t *__restrict__ out, size_t cap,
                                                               uint64_t *__restrict__ nbytes, int *__restrict__ status)
{
    __shared__ uint32_t s_off[kMaxLanes + 1];
    const int b = blockIdx.x;
    // word counts into shared memory with parallel loads, then the (short) serial scan there instead of S dependent global loads
    for (int l = threadIdx.x; l < S; l += blockDim.x) s_off[l] = words[(size_t)b * S + l];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int l = 0; l < S; ++l) { const uint32_t w = s_off[l]; s_off[l] = acc; acc += w; }
        s_off[S] = acc;
    }
    __syncthreads();
    const size_t total = (16 + (size_t)8 * S + (size_t)2 * s_off[S] + 3) & ~(size_t)3;
    if (threadIdx.x == 0) {
        nbytes[b] = total;
        if (total > cap) atomicOr(status, LANE_CAPACITY);
    }
    if (total > cap) return;
    uint32_t *hdr = reinterpret_cast<uint32_t *>(out + (size_t)b * cap);
    if (threadIdx.x == 0) { hdr[0] = kLaneMagic; hdr[1] = (uint32_t)n; hdr[2] = (uint32_t)S; hdr[3] = 0; }
    for (int l = threadIdx.x; l < S; l += blockDim.x) {
        hdr[4 + l] = states[(size_t)b * S + l];
        hdr[4 + S + l] = words[(size_t)b * S + l];
        lane_off[(size_t)b * S + l] = s_off[l];
    }
    // zero the padding half-word
    if (threadIdx.x == 0 && (s_off[S] & 1)) reinterpret_cast<uint16_t *>(out + (size_t)b * cap + 16 + (size_t)8 * S)[s_off[S]] = 0;
}

__device__ __forceinline__ uint32_t lane_get_bits(uint32_t &x, const uint16_t *&ptr, const uint16_t *end, bool &ok)
{
    const uint32_t val = x & kLaneMaxBypass;
    x >>= kLaneBypassBits;
    if (x < kLaneL) {
        if (ptr >= end) { ok = false; return val; }
        x = (x << 16) | *ptr++;
    }
    return val;
}

__global__ void __launch_bounds__(32) rans_lane_decode_kernel(const uint8_t *__restrict__ streams, const uint64_t *__restrict__ stream_off,
                                                              const uint64_t *__restrict__ stream_bytes, const int32_t *__restrict__ indexes,
                                                              int64_t n, LaneTables T, int32_t *__restrict__ out, int *__restrict__ status)
{
    const int b = blockIdx.y, lane = blockIdx.x * 32 + threadIdx.x;
    const uint8_t *s = streams + stream_off[b];
    const uint64_t len = stream_bytes[b];
    if (len < 16) { if (lane == 0) atomicOr(status, LANE_BAD_STREAM); return; }
    const uint32_t *hdr = reinterpret_cast<const uint32_t *>(s);
    const int S = (int)hdr[2];
    if (hdr[0] != kLaneMagic || hdr[1] != (uint32_t)n || S < 1 || S > kMaxLanes || len < 16 + (uint64_t)8 * S) {
        if (lane == 0) atomicOr(status, LANE_BAD_STREAM);
        return;
    }
    if (lane >= S) return;
    uint64_t off = 0, all = 0;
    for (int l = 0; l < S; ++l) {
        const uint32_t w = hdr[4 + S + l];
        if (l < lane) off += w;
        all += w;
    }
    if (16 + (uint64_t)8 * S + 2 * all > len) { atomicOr(status, LANE_BAD_STREAM); return; }
    const uint16_t *ptr = reinterpret_cast<const uint16_t *>(s + 16 + (size_t)8 * S) + off;
    const uint16_t *end = ptr + hdr[4 + S + lane];
    uint32_t x = hdr[4 + lane];
    const int32_t *idx = indexes + (int64_t)b * n;
    int32_t *o = out + (int64_t)b * n;
    const int64_t steps = lane < n ? (n - lane + S - 1) / S : 0;
    bool ok = true;
    for (int64_t lo = 0; lo < steps && ok; lo += kLaneChunk) {
        const int m = steps - lo < kLaneChunk ? (int)(steps - lo) : kLaneChunk;
        int32_t iv[kLaneChunk];
#pragma unroll
        for (int k = 0; k < kLaneChunk; ++k) iv[k] = k < m ? __ldg(idx + (int64_t)lane + (lo + k) * S) : 0;
#pragma unroll 1
        for (int k = 0; k < m && ok; ++k) {
            const int32_t ix = iv[k];
            if (ix < 0 || ix >= T.n_cdfs) { atomicOr(status, LANE_BAD_INDEX); ok = false; break; }
            const int32_t *row = T.cdfs + (size_t)ix * T.stride;
            const int32_t lenr = __ldg(T.sizes + ix), max_value = lenr - 2;
            if (max_value < 0 || lenr > T.stride) { atomicOr(status, LANE_BAD_CDF); ok = false; break; }
            const uint32_t cum = x & 0xffffu;
            // largest s with row[s] <= cum: binary search over the increasing row
            int lo_s = 0, hi_s = lenr - 1;
            if (__ldg(row) > (int32_t)cum) { ok = false; break; }
            while (hi_s - lo_s > 1) {
                const int mid = (lo_s + hi_s) >> 1;
                if (__ldg(row + mid) <= (int32_t)cum) lo_s = mid; else hi_s = mid;
            }
            int32_t sidx = lo_s;
            if (sidx > max_value) { ok = false; break; }
            const uint32_t start = (uint32_t)__ldg(row + sidx), freq = (uint32_t)__ldg(row + sidx + 1) - start;
            x = freq * (x >> kLanePrecision) + cum - start;
            if (x < kLaneL) {
                if (ptr >= end) { ok = false; break; }
                x = (x << 16) | *ptr++;
            }
            int32_t value = sidx;
            if (value == max_value) {
                int32_t val = (int32_t)lane_get_bits(x, ptr, end, ok);
                int32_t n_bypass = val;
                while (ok && val == (int32_t)kLaneMaxBypass) {
                    val = (int32_t)lane_get_bits(x, ptr, end, ok);
                    n_bypass += val;
                }
                if (n_bypass > 8) { ok = false; break; }
                uint32_t raw = 0;
                for (int j = 0; j < n_bypass && ok; ++j) raw |= lane_get_bits(x, ptr, end, ok) << (j * kLaneBypassBits);
                value = (int32_t)(raw >> 1);
                if (raw & 1) value = -value - 1;
                else value += max_value;
            }
            o[(int64_t)lane + (lo + k) * S] = value + __ldg(T.offsets + ix);
        }
    }
    if (!ok || ptr != end || x != kLaneL) atomicOr(status, LANE_BAD_STREAM);
}

// Lane count: the smallest power of two that keeps a lane's chain at <= 8192 symbols, within [4, 1024]; small tensors (the hyper
// latents) are then split further, up to 64 lanes, while a lane keeps >= 2048 symbols -- the chain length is the coding time
// (two passes of ~0.15 us per symbol on one thread), the lane count the header overhead (8 bytes per lane).
static int lanes_for(int64_t n)
{
    int s = 4;
    while (s < kMaxLanes && (int64_t)s * 8192 < n) s *= 2;
    while (s < 64 && n / (2 * s) >= 2048) s *= 2;
    return s;
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_rans_lanes_default(int64_t n) { return lanes_for(n); }

int mmc_rans_device_workspace(int batch, int lanes, int n_cdfs, int cdf_stride, size_t *bytes)
{
    MMC_CHECK_ARG(batch >= 0 && lanes >= 1 && lanes <= kMaxLanes && n_cdfs >= 1 && cdf_stride >= 2 && bytes, "mmc_rans_device_workspace: bad argument");
    // encoder table (8 bytes per CDF entry, rebuilt by every call: a few microseconds) + states, words, lane offsets (u32 each)
    *bytes = (size_t)8 * n_cdfs * cdf_stride + (size_t)3 * 4 * (size_t)batch * lanes + 64;
    return MMC_OK;
}

int mmc_rans_encode_device(const int32_t *symbols, const int32_t *indexes, int batch, int64_t n, const int32_t *cdfs, int n_cdfs,
                           int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets, int lanes, uint8_t *out,
                           size_t cap_per_stream, uint64_t *nbytes, void *workspace, int *status, void *stream)
{
    const char *name = "mmc_rans_encode_device";
    MMC_CHECK_ARG(batch >= 0 && n >= 0 && n < (1ll << 32) && lanes >= 1 && lanes <= kMaxLanes, "%s: bad argument (lanes in [1, %d])", name, kMaxLanes);
    MMC_CHECK_ARG(n_cdfs >= 1 && cdf_stride >= 2, "%s: bad CDF table", name);
    MMC_CHECK_ARG(cap_per_stream % 4 == 0 && cap_per_stream >= 16 + (size_t)8 * lanes, "%s: capacity must be a multiple of 4 and hold the header", name);
    if (batch == 0) return MMC_OK;
    MMC_CHECK_ARG((n == 0 || (symbols && indexes)) && cdfs && cdf_sizes && offsets && out && nbytes && workspace && status, "%s: NULL buffer", name);
    MMC_CHECK_ARG(batch <= 65535, "%s: batch <= 65535", name);
    cudaStream_t st = (cudaStream_t)stream;
    MMC_CHECK_ARG(aligned16(workspace), "%s: workspace must be 16-byte aligned", name);
    LaneEntry *tab = (LaneEntry *)workspace;
    uint32_t *states = (uint32_t *)(tab + (size_t)n_cdfs * cdf_stride), *words = states + (size_t)batch * lanes, *lane_off = words + (size_t)batch * lanes;
    LaneTables T{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
    MMC_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
    rans_lane_table_kernel<<<elementwise_grid((int64_t)n_cdfs * cdf_stride, 256), 256, 0, st>>>(cdfs, cdf_sizes, n_cdfs, cdf_stride, tab);
    MMC_CHECK_LAUNCH(name);
    const dim3 grid((unsigned)((lanes + 31) / 32), (unsigned)batch);
    rans_lane_encode_kernel<false><<<grid, 32, 0, st>>>(symbols, indexes, n, lanes, T, tab, states, words, lane_off, out, cap_per_stream, status);
    MMC_CHECK_LAUNCH(name);
    rans_lane_header_kernel<<<batch, 256, 0, st>>>(n, lanes, states, words, lane_off, out, cap_per_stream, nbytes, status);
    MMC_CHECK_LAUNCH(name);
    rans_lane_encode_kernel<true><<<grid, 32, 0, st>>>(symbols, indexes, n, lanes, T, tab, states, words, lane_off, out, cap_per_stream, status);
    MMC_CHECK_LAUNCH(name);
    return MMC_OK;
}

int mmc_rans_decode_device(const uint8_t *streams, const uint64_t *stream_offsets, const uint64_t *stream_bytes, const int32_t *indexes,
                           int batch, int64_t n, int max_lanes, const int32_t *cdfs, int n_cdfs, int cdf_stride, const int32_t *cdf_sizes,
                           const int32_t *offsets, int32_t *symbols_out, int *status, void *stream)
{
    const char *name = "mmc_rans_decode_device";
    MMC_CHECK_ARG(batch >= 0 && n >= 0 && n < (1ll << 32) && max_lanes >= 1 && max_lanes <= kMaxLanes, "%s: bad argument", name);
    MMC_CHECK_ARG(n_cdfs >= 1 && cdf_stride >= 2, "%s: bad CDF table", name);
    if (batch == 0) return MMC_OK;
    MMC_CHECK_ARG(streams && stream_offsets && stream_bytes && (n == 0 || (indexes && symbols_out)) && cdfs && cdf_sizes && offsets && status, "%s: NULL buffer", name);
    MMC_CHECK_ARG(batch <= 65535, "%s: batch <= 65535", name);
    cudaStream_t st = (cudaStream_t)stream;
    LaneTables T{cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets};
    MMC_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
    const dim3 grid((unsigned)((max_lanes + 31) / 32), (unsigned)batch);
    rans_lane_decode_kernel<<<grid, 32, 0, st>>>(streams, stream_offsets, stream_bytes, indexes, n, T, symbols_out, status);
    MMC_CHECK_LAUNCH(name);
    return MMC_OK;
}

}  // extern "C"
