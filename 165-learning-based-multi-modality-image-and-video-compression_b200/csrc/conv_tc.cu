// conv_tc.cu -- tensor-core implicit-GEMM convolution / transposed convolution for sm_100a.
//
// Replaces conv()/deconv() (compressai/models/utils.py:128-146: nn.Conv2d / nn.ConvTranspose2d,
// k in {1,3,5}, stride in {1,2}, padding k/2, output_padding stride-1) fused with what follows them in
// the transform stacks: bias, ReLU / LeakyReLU (compressai/models/google.py:254-269,363-377) and
// GDN / IGDN (compressai/layers/gdn.py:77-92), plus the |y| copy that feeds h_a (google.py:283).
//
// GEMM view: M = output pixels (128 per tile: a TH x TW patch of one image), N = Cout (<= 256 per
// pass), K = taps x Cin walked as (tap, 64-channel chunk) blocks.
//   * A operand: NHWC bf16 activations.  One 4-D TMA box {64 ch, TW, TH, 1} per K block lands in
//     shared memory as 128 rows x 128 B with the 128B swizzle, i.e. directly in the canonical
//     K-major UMMA layout.  Stride-2 convolutions use the tensor map's element strides, borders
//     (padding) are TMA out-of-bounds zero fill -- there is no im2col buffer and no halo logic.
//   * A transposed convolution is split into its stride^2 output phases; every phase is a stride-1
//     correlation over the taps of matching parity (3x3 / 3x2 / 2x3 / 2x2 for k=5), so no zero
//     insertion and no wasted MACs.
//   * B operand: weights pre-packed [tap][Cout][Cin] bf16, one 2-D TMA box {64, Ntile} per K block.
//   * tcgen05.mma (cta_group::1, M=128, N=Ntile, K=16) issued by one thread, fp32 accumulators in
//     TMEM, double buffered so the epilogue of tile i overlaps the main loop of tile i+1.
//   * Epilogue (4 warps, one accumulator row = one pixel per thread): tcgen05.ld -> bias -> act ->
//     [GDN] -> bf16 / fp32 NHWC stores.  GDN is a second tensor-core contraction: the epilogue
//     squares the activations into a swizzled bf16 tile in shared memory, one thread issues
//     D2[128 x C] = X2[128 x C] * gamma^T with gamma resident in shared memory, and the result is
//     combined as x * rsqrt(beta + D2) (sqrt for IGDN).
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2-5 = epilogue.
// Persistent grid: one CTA per SM, tiles strided across CTAs.
#include <cuda.h>

#include "common.cuh"

namespace mmc {

constexpr int kTcThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kMaxTaps = 32;
constexpr int kMaxCout = 1024;
constexpr int kABytes = 128 * 128;  // 128 pixels x 64 bf16

struct Tap {
    int16_t dy, dx;  // offset of the input patch for this tap (in input pixels)
    int32_t brow;    // first row of this tap in the packed weight matrix (tap * Cout)
};

struct TcParams {
    CUtensorMap tmA, tmB, tmG;
    Tap taps[kMaxTaps];
    int phase_begin[5];  // taps of phase p are [phase_begin[p], phase_begin[p+1])
    int n_phases;
    int a_stride;    // element stride of the A box along W and H (conv stride; 1 for deconv phases)
    int out_stride;  // output pixels per grid cell (deconv: stride; conv: 1)
    int B, Gh, Gw;   // per-phase pixel grid
    int Ho, Wo;
    int TH, TW, tiles_y, tiles_x;
    int Cout, Ntile, n_blocks;
    int kchunks;     // ceil(Cin / 64)
    int num_stages, acc_stages;
    int act, gdn, out_f32, out2;
    int gdn_chunk;
    int tiles_per_phase, total_tiles;
    const float *bias;
    const float *beta;
    void *y;
    __nv_bfloat16 *y2;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(addr), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap *tm, uint64_t *bar, void *dst, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *tm, uint64_t *bar, void *dst, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *tm)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, M=128, K=16
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v)
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor: K-major operand, 128B swizzle, rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address          bits [0,14)
    d |= (uint64_t)1 << 16;                     // leading byte offset    bits [16,30)  (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset     bits [32,46)
    d |= (uint64_t)1 << 46;                     // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                     // layout: SWIZZLE_128B
    return d;
}
// Instruction descriptor for kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=n.
__device__ __forceinline__ uint32_t make_idesc(int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ float act_tc(float v, int act)
{
    if (act == MMC_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == MMC_ACT_LEAKY_RELU) return v > 0.0f ? v : 0.01f * v;
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}

struct TileCoord {
    int phase, b, y0, x0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const TcParams &P, int tile)
{
    TileCoord t;
    t.phase = tile / P.tiles_per_phase;
    int r = tile - t.phase * P.tiles_per_phase;
    int nb = r % P.n_blocks; r /= P.n_blocks;
    int tx = r % P.tiles_x;  r /= P.tiles_x;
    int ty = r % P.tiles_y;
    t.b = r / P.tiles_y;
    t.y0 = ty * P.TH;
    t.x0 = tx * P.TW;
    t.n0 = nb * P.Ntile;
    return t;
}

// 16 consecutive output channels of one pixel -> global memory (bf16 or fp32 NHWC) [+ bf16 secondary]
__device__ __forceinline__ void store16(const TcParams &P, int64_t off, const float *v)
{
    if (P.out_f32) {
        float4 *dst = reinterpret_cast<float4 *>((float *)P.y + off);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
        uint4 *dst = reinterpret_cast<uint4 *>((__nv_bfloat16 *)P.y + off);
        dst[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        dst[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
    }
    if (P.out2) {
        float w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = (P.out2 == 1) ? fabsf(v[i]) : v[i];
        uint4 *dst = reinterpret_cast<uint4 *>(P.y2 + off);
        dst[0] = make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
        dst[1] = make_uint4(pack_bf16(w[8], w[9]), pack_bf16(w[10], w[11]), pack_bf16(w[12], w[13]), pack_bf16(w[14], w[15]));
    }
}

__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_kernel(const __grid_constant__ TcParams P)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
    __shared__ uint64_t tmem_full_bar[2], tmem_empty_bar[2], gdn_bar, gload_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float bias_s[kMaxCout], beta_s[256];

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = kABytes + P.Ntile * 128;
    uint8_t *sG = smem + (size_t)P.num_stages * stage_bytes;       // gamma: (Cout/64) tiles of [Cout][64] bf16
    uint8_t *sA2 = sG + (size_t)P.Cout * P.Cout * 2;               // x^2:   (Cout/64) tiles of [128][64] bf16

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.num_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 128); }
        mbar_init(&gdn_bar, 1);
        mbar_init(&gload_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&P.tmA);
        prefetch_tmap(&P.tmB);
        if (P.gdn) prefetch_tmap(&P.tmG);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = threadIdx.x; i < kMaxCout; i += kTcThreads) {
        bias_s[i] = (P.bias && i < P.Cout) ? P.bias[i] : 0.0f;
        if (i < 256) beta_s[i] = (P.gdn && i < P.Cout) ? P.beta[i] : 1.0f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            if (P.gdn) {
                mbar_expect_tx(&gload_bar, (uint32_t)(P.Cout * P.Cout * 2));
                for (int kc = 0; kc < P.Cout / 64; ++kc)
                    tma_load_2d(&P.tmG, &gload_bar, sG + (size_t)kc * P.Cout * 128, kc * 64, 0);
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
                TileCoord t = decode_tile(P, tile);
                const int cx = t.x0 * P.a_stride, cy = t.y0 * P.a_stride;
                for (int tp = P.phase_begin[t.phase]; tp < P.phase_begin[t.phase + 1]; ++tp) {
                    const Tap tap = P.taps[tp];
                    for (int kc = 0; kc < P.kchunks; ++kc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t *a = smem + (size_t)stage * stage_bytes;
                        mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
                        tma_load_4d(&P.tmA, &full_bar[stage], a, kc * 64, cx + tap.dx, cy + tap.dy, t.b);
                        tma_load_2d(&P.tmB, &full_bar[stage], a + kABytes, kc * 64, tap.brow + t.n0);
                        if (++stage == P.num_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(P.Ntile);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
                TileCoord t = decode_tile(P, tile);
                const int as = (P.acc_stages == 2) ? (it & 1) : 0;
                const uint32_t aphase = (P.acc_stages == 2) ? ((it >> 1) & 1) : (it & 1);
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * P.Ntile);
                const int nkb = (P.phase_begin[t.phase + 1] - P.phase_begin[t.phase]) * P.kchunks;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t adesc = make_desc(a_addr), bdesc = make_desc(a_addr + kABytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k)   // 4 x K=16 per 64-channel chunk: +32 B per step inside the swizzle atom
                        tc_mma(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    tc_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs have read it
                    if (++stage == P.num_stages) { stage = 0; phase ^= 1; }
                }
                tc_commit(&tmem_full_bar[as]);      // accumulator complete -> epilogue
            }
        }
    } else {
        // ===================== epilogue warps (2..5) =====================
        const int q = warp & 3;                 // TMEM lane quarter this warp can access
        const int row = q * 32 + lane;          // accumulator row == pixel of the tile
        const int th = row / P.TW, tw = row - th * P.TW;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t norm_col = (uint32_t)(P.acc_stages * P.Ntile);
        uint32_t gdn_phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
            TileCoord t = decode_tile(P, tile);
            const int as = (P.acc_stages == 2) ? (it & 1) : 0;
            const uint32_t aphase = (P.acc_stages == 2) ? ((it >> 1) & 1) : (it & 1);
            const int gy = t.y0 + th, gx = t.x0 + tw;
            const bool valid = gy < P.Gh && gx < P.Gw;
            const int py = t.phase / P.out_stride, px = t.phase - py * P.out_stride;
            const int oy = gy * P.out_stride + py, ox = gx * P.out_stride + px;
            const int64_t pix_off = (((int64_t)t.b * P.Ho + oy) * P.Wo + ox) * P.Cout + t.n0;
            const uint32_t acc_addr = tmem_base + lane_addr + (uint32_t)(as * P.Ntile);

            mbar_wait(&tmem_full_bar[as], aphase);
            tc_fence_after();

            if (!P.gdn) {
                for (int c0 = 0; c0 < P.Ntile; c0 += 16) {
                    float v[16];
                    tmem_ld16(acc_addr + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = act_tc(v[i] + bias_s[t.n0 + c0 + i], P.act);
                    if (valid) store16(P, pix_off + c0, v);
                }
            } else {
                // ---- pass 1: x = acc + bias ; x^2 (bf16) -> swizzled K-major tile in shared memory ----
                for (int c0 = 0; c0 < P.Cout; c0 += 16) {
                    float v[16];
                    tmem_ld16(acc_addr + c0, v);
                    tmem_ld_wait();
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float a = act_tc(v[2 * i] + bias_s[c0 + 2 * i], P.act), b = act_tc(v[2 * i + 1] + bias_s[c0 + 2 * i + 1], P.act);
                        pk[i] = pack_bf16(a * a, b * b);
                    }
                    uint8_t *tile_base = sA2 + (size_t)(c0 >> 6) * kABytes + (size_t)row * 128;
                    const int j0 = (c0 & 63) >> 3;   // 16-byte chunk index inside the 128-byte row
                    *reinterpret_cast<uint4 *>(tile_base + (((j0) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4 *>(tile_base + (((j0 + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
                tc_fence_before();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int g0 = 0; g0 < P.Cout; g0 += P.gdn_chunk) {
                    if (threadIdx.x == 64) {
                        // ---- norm[128 x chunk] = X2[128 x C] * gamma[g0:g0+chunk, :]^T on the tensor cores ----
                        if (it == 0 && g0 == 0) mbar_wait(&gload_bar, 0);
                        tc_fence_after();
                        const uint32_t idesc = make_idesc(P.gdn_chunk);
                        for (int kc = 0; kc < P.Cout / 64; ++kc) {
                            const uint64_t adesc = make_desc(smem_u32(sA2 + (size_t)kc * kABytes));
                            const uint64_t bdesc = make_desc(smem_u32(sG + (size_t)kc * P.Cout * 128 + (size_t)g0 * 128));
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc_mma(tmem_base + norm_col, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kc | k) != 0);
                        }
                        tc_commit(&gdn_bar);
                    }
                    mbar_wait(&gdn_bar, gdn_phase);
                    gdn_phase ^= 1;
                    tc_fence_after();
                    // ---- pass 2: y = x * rsqrt(beta + norm)  (IGDN: * sqrt) ----
                    for (int c0 = g0; c0 < g0 + P.gdn_chunk; c0 += 16) {
                        float v[16], nrm[16];
                        tmem_ld16(acc_addr + c0, v);
                        tmem_ld16(tmem_base + lane_addr + norm_col + (uint32_t)(c0 - g0), nrm);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float x = act_tc(v[i] + bias_s[c0 + i], P.act);
                            float n = beta_s[c0 + i] + nrm[i];
                            v[i] = (P.gdn == MMC_GDN_INVERSE) ? x * sqrtf(n) : x * rsqrtf(n);
                        }
                        if (valid) store16(P, pix_off + c0, v);
                    }
                    // all 128 threads are done with the norm columns (and, on the last chunk, with sA2)
                    tc_fence_before();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 torch layout -> bf16 [tap][Cout][Cin]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_weights_kernel(const float *__restrict__ w, int transposed, int Cin, int Cout,
                                                          int kk, __nv_bfloat16 *__restrict__ out)
{
    int64_t n = (int64_t)kk * Cout * Cin;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int ci = (int)(i % Cin);
        int64_t r = i / Cin;
        int co = (int)(r % Cout);
        int tap = (int)(r / Cout);
        int64_t src = transposed ? (((int64_t)ci * Cout + co) * kk + tap) : (((int64_t)co * Cin + ci) * kk + tap);
        out[i] = __float2bfloat16_rn(__ldg(w + src));
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int encode_map(CUtensorMap *tm, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                      const uint32_t *box, const uint32_t *estr, const char *what)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return MMC_ECUDA; }
    cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = estr[i]; }
    for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r); return MMC_ECUDA; }
    return MMC_OK;
}

static int pick_ntile(int cout)
{
    for (int n = 256; n >= 16; n -= 16)
        if (cout % n == 0) return n;
    return 0;
}

static void pick_tile(int gh, int gw, int a_stride, int *TH, int *TW)
{
    const int cand[][2] = {{8, 16}, {16, 8}, {4, 32}, {32, 4}, {2, 64}, {64, 2}, {1, 128}, {128, 1}};
    int64_t best = -1;
    for (auto &c : cand) {
        if (c[0] * a_stride > 256 || c[1] * a_stride > 256) continue;   // TMA box limit
        int64_t tiles = (int64_t)((gh + c[0] - 1) / c[0]) * ((gw + c[1] - 1) / c[1]);
        if (best < 0 || tiles < best) { best = tiles; *TH = c[0]; *TW = c[1]; }
    }
}

static int tc_validate(const mmc_conv_desc *d, const char *name)
{
    MMC_CHECK_ARG(d != nullptr, "%s: descriptor is NULL", name);
    MMC_CHECK_ARG(d->B >= 0 && d->H >= 1 && d->W >= 1 && d->Cin >= 1 && d->Cout >= 1, "%s: bad shape", name);
    MMC_CHECK_ARG(d->k == 1 || d->k == 3 || d->k == 5, "%s: kernel size %d not in {1,3,5}", name, d->k);
    MMC_CHECK_ARG(d->stride == 1 || d->stride == 2, "%s: stride %d not in {1,2}", name, d->stride);
    // Cin is walked in 64-channel TMA boxes; a ragged last box is zero-filled by TMA on both operands.
    // The global row pitch (Cin * 2 B) must be a multiple of 16 B for the tensor maps.
    MMC_UNSUPPORTED(d->Cin % 8 != 0 || d->Cin < 32, "%s: tensor-core path needs Cin %% 8 == 0 and Cin >= 32 (got %d); use the direct kernel", name, d->Cin);
    MMC_UNSUPPORTED(d->Cout > kMaxCout, "%s: Cout=%d exceeds %d", name, d->Cout, kMaxCout);
    MMC_UNSUPPORTED(d->Cout % 16 != 0, "%s: tensor-core path needs Cout %% 16 == 0 (got %d); use the direct kernel", name, d->Cout);
    return MMC_OK;
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_conv_pack_weights(const mmc_conv_desc *d, const float *w, void *w_packed, size_t *bytes, void *stream)
{
    int rc = tc_validate(d, "mmc_conv_pack_weights");
    if (rc) return rc;
    size_t need = (size_t)d->k * d->k * d->Cout * d->Cin * sizeof(__nv_bfloat16);
    if (bytes) *bytes = need;
    if (!w_packed) return MMC_OK;
    MMC_CHECK_ARG(w != nullptr, "mmc_conv_pack_weights: w is NULL");
    int64_t n = (int64_t)d->k * d->k * d->Cout * d->Cin;
    pack_weights_kernel<<<elementwise_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(w, d->transposed, d->Cin, d->Cout, d->k * d->k,
                                                                                  (__nv_bfloat16 *)w_packed);
    MMC_CHECK_LAUNCH("mmc_conv_pack_weights");
    return MMC_OK;
}

int mmc_conv_forward_tc(const mmc_conv_desc *d, const void *x, const void *w_packed, const float *bias,
                        const float *beta_eff, const void *gamma_eff_bf16, void *y, void *y2, void *stream)
{
    const char *name = "mmc_conv_forward_tc";
    int rc = tc_validate(d, name);
    if (rc) return rc;
    MMC_CHECK_ARG(d->in_dtype == MMC_BF16 && d->in_layout == MMC_NHWC, "%s: input must be NHWC bf16", name);
    MMC_CHECK_ARG(d->out_layout == MMC_NHWC, "%s: output must be NHWC", name);
    MMC_CHECK_ARG(d->act >= 0 && d->act <= MMC_ACT_LEAKY_RELU, "%s: bad act", name);
    MMC_CHECK_ARG(d->gdn >= 0 && d->gdn <= MMC_GDN_INVERSE, "%s: bad gdn mode", name);
    MMC_CHECK_ARG(d->out2_bf16 >= 0 && d->out2_bf16 <= 2, "%s: bad out2_bf16", name);
    MMC_CHECK_ARG(d->gdn == MMC_GDN_NONE || (beta_eff && gamma_eff_bf16), "%s: GDN needs beta/gamma", name);
    MMC_CHECK_ARG(!d->out2_bf16 || y2, "%s: out2_bf16 set but y2 is NULL", name);
    MMC_UNSUPPORTED(d->gdn != MMC_GDN_NONE && (d->Cout > 192 || d->Cout % 64 != 0),
                    "%s: fused GDN supports Cout in {64,128,192} (got %d)", name, d->Cout);
    if (d->B == 0) return MMC_OK;
    MMC_CHECK_ARG(x && w_packed && y, "%s: NULL buffer", name);
    MMC_CHECK_ARG(aligned16(x) && aligned16(w_packed) && aligned16(y) && (!y2 || aligned16(y2)), "%s: buffers must be 16-byte aligned", name);

    TcParams P;
    memset(&P, 0, sizeof(P));
    const int k = d->k, s = d->stride, pad = k / 2, kk = k * k;
    int Ho, Wo;
    mmc_conv_out_size(d, &Ho, &Wo);
    P.Ho = Ho; P.Wo = Wo; P.B = d->B; P.Cout = d->Cout;
    P.kchunks = (d->Cin + 63) / 64;
    P.act = d->act; P.gdn = d->gdn; P.out_f32 = (d->out_dtype == MMC_F32); P.out2 = d->out2_bf16;
    P.bias = bias; P.beta = beta_eff; P.y = y; P.y2 = (__nv_bfloat16 *)y2;

    // ---- phases and taps ----
    int nt = 0;
    if (!d->transposed) {
        P.n_phases = 1; P.a_stride = s; P.out_stride = 1; P.Gh = Ho; P.Gw = Wo;
        P.phase_begin[0] = 0;
        for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx) P.taps[nt++] = Tap{(int16_t)(ky - pad), (int16_t)(kx - pad), (ky * k + kx) * d->Cout};
        P.phase_begin[1] = nt;
    } else {
        // oy = iy*s - pad + ky  =>  for output phase py: ky == (py + pad) mod s, iy = qy + (py + pad - ky)/s
        P.n_phases = s * s; P.a_stride = 1; P.out_stride = s; P.Gh = d->H; P.Gw = d->W;
        for (int ph = 0; ph < s * s; ++ph) {
            const int py = ph / s, px = ph % s;
            P.phase_begin[ph] = nt;
            for (int ky = 0; ky < k; ++ky) {
                if (((py + pad - ky) % s) != 0) continue;
                for (int kx = 0; kx < k; ++kx) {
                    if (((px + pad - kx) % s) != 0) continue;
                    P.taps[nt++] = Tap{(int16_t)((py + pad - ky) / s), (int16_t)((px + pad - kx) / s), (ky * k + kx) * d->Cout};
                }
            }
        }
        P.phase_begin[s * s] = nt;
    }
    (void)kk;

    // ---- tiling ----
    P.Ntile = (d->gdn != MMC_GDN_NONE) ? d->Cout : pick_ntile(d->Cout);
    MMC_UNSUPPORTED(P.Ntile == 0 || P.Ntile > 256, "%s: no valid N tile for Cout=%d", name, d->Cout);
    P.n_blocks = d->Cout / P.Ntile;
    pick_tile(P.Gh, P.Gw, P.a_stride, &P.TH, &P.TW);
    P.tiles_y = (P.Gh + P.TH - 1) / P.TH;
    P.tiles_x = (P.Gw + P.TW - 1) / P.TW;
    int64_t tpp = (int64_t)d->B * P.tiles_y * P.tiles_x * P.n_blocks;
    MMC_CHECK_ARG(tpp * P.n_phases < (1ll << 31), "%s: too many tiles", name);
    P.tiles_per_phase = (int)tpp;
    P.total_tiles = (int)(tpp * P.n_phases);
    P.gdn_chunk = 0;
    if (d->gdn != MMC_GDN_NONE) P.gdn_chunk = (2 * P.Ntile + P.Ntile <= 512) ? P.Ntile : P.Ntile / 2;
    P.acc_stages = (2 * P.Ntile + P.gdn_chunk <= 512) ? 2 : 1;
    MMC_UNSUPPORTED(P.acc_stages * P.Ntile + P.gdn_chunk > 512 || (P.gdn_chunk % 16) != 0, "%s: TMEM budget exceeded", name);

    // ---- shared memory budget ----
    const size_t stage_bytes = kABytes + (size_t)P.Ntile * 128;
    size_t fixed = 1024;  // alignment slack
    if (d->gdn != MMC_GDN_NONE) fixed += (size_t)d->Cout * d->Cout * 2 + (size_t)(d->Cout / 64) * kABytes;
    // dynamic shared memory available next to the kernel's static allocation (227 KB per CTA on sm_100)
    static size_t budget = 0;
    if (budget == 0) {
        cudaFuncAttributes fa;
        MMC_CHECK_CUDA(cudaFuncGetAttributes(&fa, conv_tc_kernel));
        size_t avail = 227 * 1024 - fa.sharedSizeBytes;
        MMC_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)avail));
        budget = avail;
    }
    MMC_UNSUPPORTED(fixed + 2 * stage_bytes > budget, "%s: shared memory budget exceeded", name);
    int stages = (int)((budget - fixed) / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    P.num_stages = stages;
    const size_t smem = fixed + (size_t)stages * stage_bytes;

    // ---- tensor maps ----
    {
        uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
        uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
        uint32_t box[4] = {64, (uint32_t)(P.TW * P.a_stride), (uint32_t)(P.TH * P.a_stride), 1};
        uint32_t es[4] = {1, (uint32_t)P.a_stride, (uint32_t)P.a_stride, 1};
        rc = encode_map(&P.tmA, x, 4, dims, str, box, es, "activations");
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)d->Cin, (uint64_t)k * k * d->Cout};
        uint64_t str[1] = {(uint64_t)d->Cin * 2};
        uint32_t box[2] = {64, (uint32_t)P.Ntile};
        uint32_t es[2] = {1, 1};
        rc = encode_map(&P.tmB, w_packed, 2, dims, str, box, es, "weights");
        if (rc) return rc;
    }
    if (d->gdn != MMC_GDN_NONE) {
        MMC_CHECK_ARG(aligned16(gamma_eff_bf16), "%s: gamma must be 16-byte aligned", name);
        uint64_t dims[2] = {(uint64_t)d->Cout, (uint64_t)d->Cout};
        uint64_t str[1] = {(uint64_t)d->Cout * 2};
        uint32_t box[2] = {64, (uint32_t)d->Cout};
        uint32_t es[2] = {1, 1};
        rc = encode_map(&P.tmG, gamma_eff_bf16, 2, dims, str, box, es, "gamma");
        if (rc) return rc;
    }

    int grid = P.total_tiles < kNumSMs ? P.total_tiles : kNumSMs;
    conv_tc_kernel<<<grid, kTcThreads, smem, (cudaStream_t)stream>>>(P);
    MMC_CHECK_LAUNCH(name);
    return MMC_OK;
}

}  // extern "C"
