// wgrad_tc.cu -- weight gradient of conv() / deconv() on tensor cores (tcgen05), plus the planar staging it reads.
//
// Backward of compressai/models/utils.py:128-146 with respect to the weights (training step of the two-branch codec,
// examples/train.py:239-253).  Both layer types reduce to one form.  Let S be the tensor on the low-resolution side of the
// layer (conv: the output gradient dy; deconv: the input x) and L the one on the high-resolution side (conv: the input x;
// deconv: the output gradient dy), then
//     dW[cs][cl][ky][kx] = sum_{b, qy, qx}  S[b, cs, qy, qx] * L[b, cl, qy*s + ky - pad, qx*s + kx - pad]
// which is exactly torch's weight layout for both nn.Conv2d (Cout, Cin, k, k) and nn.ConvTranspose2d (Cin, Cout, k, k).
// The same kernel with k = 1 gives the GDN gamma gradient  dgamma[i][j] = sum_pixels t_i * x_j^2  (SURVEY.md Appendix E).
//
// GEMM view per tap: M = cs (128 per tile), N = cl (<= 256 per tile), K = pixels, walked as 64-pixel patches (TH x TW).
// Operands are read straight from the NHWC bf16 activations / gradients that the forward and input-gradient kernels use:
// [pixel][channel] with channels contiguous is the MN-major UMMA operand layout, so no transposed copy exists.  One TMA
// box {64 ch, TW, TH, 1} per 64-channel atom per K block lands as 64 pixel rows x 128 B with the 128 B swizzle
// (canonical MN-major SW128 layout: 8-row groups 1024 B apart = SBO, atoms 8192 B apart = LBO); the layer stride is the
// tensor map's element stride on L, borders and ragged channel tiles are TMA out-of-bounds zero fill -- the same
// addressing as the forward kernel's activation loads.
// Work item = (tap, M tile, N tile, K split); fp32 accumulators in TMEM (double buffered), split-K partial sums are
// reduced with fp32 atomics into a [tap][Cs][Cl] workspace; mmc_wgrad_finalize re-lays it out as torch's (cs, cl, k, k),
// applies a scale (GDN: -1/2) and an optional mask (MaskedConv2d, layers/layers.py:75-78).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace mmc {

constexpr int kWgThreads = 192;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (one TMEM lane quarter each)
constexpr int kWgMaxStages = 8;

struct WgParams {
    CUtensorMap tmS, tmL;
    int Cs, Cl, k, stride, pad;
    int Hs, Ws, B;
    int m_tiles, n_tiles, Ntile, n_atoms;
    int TH, TW, tiles_y, tiles_x;   // K block = TH x TW = 64 pixels of the S grid
    int kblocks;         // B * tiles_y * tiles_x
    int splits, kb_per_split;
    int total_items;
    int num_stages;
    int debug;           // profiling / bring-up aid (env MMC_WG_DEBUG): bit 0 no TMA loads, bit 1 no MMAs, bit 2 no TMEM loads
    float *ws;           // [k*k][Cs][Cl] fp32, zero-initialised by the caller
};

struct WgItem {
    int tap, mt, nt, kb0, kb1;
};

__device__ __forceinline__ WgItem wg_item(const WgParams &P, int item)
{
    WgItem it;
    int split = item % P.splits;
    int r = item / P.splits;
    it.nt = r % P.n_tiles; r /= P.n_tiles;
    it.mt = r % P.m_tiles;
    it.tap = r / P.m_tiles;
    it.kb0 = split * P.kb_per_split;
    it.kb1 = min(it.kb0 + P.kb_per_split, P.kblocks);
    return it;
}

constexpr int kAtomBytes = 64 * 128;   // 64 pixels x 64 channels bf16

// MN-major operand, 128B swizzle: 64-element (128 B) rows along M/N, 8 K-rows per 1024 B group (SBO), atoms LBO apart
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(kAtomBytes >> 4) << 16;     // leading byte offset: next 64-channel atom
    d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset: next 8 pixels
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgParams P)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[kWgMaxStages], empty_bar[kWgMaxStages], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_s;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_bytes = 2 * kAtomBytes, b_bytes = P.n_atoms * kAtomBytes, stage_bytes = a_bytes + b_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.num_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&P.tmS);
        prefetch_tmap(&P.tmL);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
            const WgItem it = wg_item(P, item);
            const int ky = it.tap / P.k, kx = it.tap - ky * P.k;
            int tx = it.kb0 % P.tiles_x;
            int r = it.kb0 / P.tiles_x;
            int ty = r % P.tiles_y;
            int b = r / P.tiles_y;
            for (int kb = it.kb0; kb < it.kb1; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t *a = smem + (size_t)stage * stage_bytes;
                if (elect_one()) {
                    if (P.debug & 1) {
                        mbar_arrive(&full_bar[stage]);
                    } else {
                        mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
                        const int x0 = tx * P.TW, y0 = ty * P.TH;
                        tma_load_4d(&P.tmS, &full_bar[stage], a, it.mt * 128, x0, y0, b);
                        tma_load_4d(&P.tmS, &full_bar[stage], a + kAtomBytes, it.mt * 128 + 64, x0, y0, b);
                        for (int at = 0; at < P.n_atoms; ++at)
                            tma_load_4d(&P.tmL, &full_bar[stage], a + a_bytes + at * kAtomBytes, it.nt * P.Ntile + at * 64,
                                        x0 * P.stride + kx - P.pad, y0 * P.stride + ky - P.pad, b);
                    }
                }
                __syncwarp();
                if (++stage == P.num_stages) { stage = 0; phase ^= 1; }
                if (++tx == P.tiles_x) { tx = 0; if (++ty == P.tiles_y) { ty = 0; ++b; } }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc(P.Ntile) | (1u << 15) | (1u << 16);   // A and B are MN-major
        int stage = 0;
        uint32_t phase = 0;
        int n_it = 0;
        for (int item = blockIdx.x; item < P.total_items; item += gridDim.x, ++n_it) {
            const WgItem it = wg_item(P, item);
            const int as = n_it & 1;
            const uint32_t aphase = (n_it >> 1) & 1;
            mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * P.Ntile);
            for (int kb = it.kb0; kb < it.kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint64_t adesc = make_desc_mn(a_addr), bdesc = make_desc_mn(a_addr + a_bytes);
                if (elect_one()) {
                    if (!(P.debug & 2)) {
                        // K = 16 pixels per step = 16 rows of 128 B: +2048 B (= +128 in descriptor units)
                        tc_mma(d_tmem, adesc, bdesc, idesc, kb != it.kb0);
                        tc_mma(d_tmem, adesc + 128, bdesc + 128, idesc, 1);
                        tc_mma(d_tmem, adesc + 256, bdesc + 256, idesc, 1);
                        tc_mma(d_tmem, adesc + 384, bdesc + 384, idesc, 1);
                    }
                    tc_commit(&empty_bar[stage]);
                    if (kb == it.kb1 - 1) tc_commit(&tmem_full_bar[as]);
                }
                __syncwarp();
                if (++stage == P.num_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: TMEM -> fp32 atomics into the [tap][Cs][Cl] workspace =====================
        const int q = warp & 3;                       // TMEM lane quarter of this warp
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        int n_it = 0;
        for (int item = blockIdx.x; item < P.total_items; item += gridDim.x, ++n_it) {
            const WgItem it = wg_item(P, item);
            const int as = n_it & 1;
            const uint32_t aphase = (n_it >> 1) & 1;
            mbar_wait(&tmem_full_bar[as], aphase);
            tc_fence_after();
            const int m = it.mt * 128 + row;
            float *dst = P.ws + ((size_t)it.tap * P.Cs + m) * P.Cl + it.nt * P.Ntile;
            const int n_valid = min(P.Ntile, P.Cl - it.nt * P.Ntile);
            for (int c0 = 0; c0 < P.Ntile; c0 += 16) {
                float v[16];
                if (P.debug & 4) continue;
                tmem_ld16(tmem_base + lane_addr + (uint32_t)(as * P.Ntile + c0), v);
                tmem_ld_wait();
                if (m < P.Cs) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < n_valid) atomicAdd(dst + c0 + i, v[i]);
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

// Narrow tensors (image / reconstruction gradient, <= 8 channels): gather the k x k neighbourhood of every low-resolution
// position into one NHWC row [q][tap][8 ch], so that the weight gradient of the edge layers is ONE k = 1 GEMM with N = k*k*8
// columns instead of k*k GEMMs with a 16-wide N tile (the MMA costs the same ~115 cycles for N = 16 as for N = 208).
__global__ void __launch_bounds__(256) im2col8_kernel(const uint4 *__restrict__ x, int H, int W, int k, int stride, int Hs, int Ws, int64_t n,
                                                     uint4 *__restrict__ out)
{
    const int kk = k * k, pad = k / 2;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const int tap = (int)(i % kk);
        int64_t q = i / kk;
        const int qx = (int)(q % Ws); q /= Ws;
        const int qy = (int)(q % Hs);
        const int64_t b = q / Hs;
        const int iy = qy * stride + tap / k - pad, ix = qx * stride + tap % k - pad;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + (b * H + iy) * (int64_t)W + ix);
        out[i] = v;
    }
}

__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const float *__restrict__ ws, int taps, int Cs, int Cl, float scale,
                                                            const float *__restrict__ mask, int accumulate, float *__restrict__ dw)
{
    const int64_t n = (int64_t)taps * Cs * Cl, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // i indexes the OUTPUT (cs, cl, tap) so that writes are coalesced; the workspace read is strided
        const int tap = (int)(i % taps);
        const int64_t r = i / taps;
        float v = scale * ws[(size_t)tap * Cs * Cl + r];
        if (mask) v *= mask[i];
        dw[i] = accumulate ? dw[i] + v : v;
    }
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_wgrad_tc(const void *s_nhwc, const void *l_nhwc, int64_t B, int Cs, int Cl, int Hs, int Ws, int Hl, int Wl, int k, int stride,
                 float *workspace, void *stream)
{
    const char *name = "mmc_wgrad_tc";
    MMC_CHECK_ARG(B >= 0 && Cs >= 1 && Cl >= 1 && Hs >= 1 && Ws >= 1 && Hl >= 1 && Wl >= 1, "%s: bad shape", name);
    MMC_CHECK_ARG(k == 1 || k == 3 || k == 5, "%s: kernel size %d not in {1,3,5}", name, k);
    MMC_CHECK_ARG(stride == 1 || stride == 2, "%s: stride %d not in {1,2}", name, stride);
    MMC_UNSUPPORTED(Cs % 8 != 0 || Cl % 8 != 0, "%s: channel counts must be multiples of 8 (got %d, %d); pad narrow tensors to 8 channels", name, Cs, Cl);
    if (B == 0) return MMC_OK;
    MMC_CHECK_ARG(s_nhwc && l_nhwc && workspace && aligned16(s_nhwc) && aligned16(l_nhwc), "%s: NULL or unaligned buffer", name);
    WgParams P;
    memset(&P, 0, sizeof(P));
    P.Cs = Cs; P.Cl = Cl; P.k = k; P.stride = stride; P.pad = k / 2;
    P.Hs = Hs; P.Ws = Ws; P.B = (int)B;
    P.ws = workspace;
    P.m_tiles = (Cs + 127) / 128;
    // N tile: the smallest multiple of 16 that covers Cl in the fewest (<= 256-wide) tiles
    P.n_tiles = (Cl + 255) / 256;
    P.Ntile = (((Cl + P.n_tiles - 1) / P.n_tiles) + 15) & ~15;
    P.n_atoms = (P.Ntile + 63) / 64;
    {
        // 64-pixel K patch: the shape that wastes the fewest out-of-range pixels on this grid
        const int cand[][2] = {{8, 8}, {4, 16}, {16, 4}, {2, 32}, {32, 2}, {1, 64}, {64, 1}};
        int64_t best = -1;
        for (auto &c : cand) {
            if (c[0] * stride > 256 || c[1] * stride > 256) continue;
            const int64_t t = (int64_t)((Hs + c[0] - 1) / c[0]) * ((Ws + c[1] - 1) / c[1]);
            if (best < 0 || t < best) { best = t; P.TH = c[0]; P.TW = c[1]; }
        }
    }
    P.tiles_y = (Hs + P.TH - 1) / P.TH;
    P.tiles_x = (Ws + P.TW - 1) / P.TW;
    const int64_t kblocks = B * P.tiles_y * P.tiles_x;
    MMC_CHECK_ARG(kblocks < (1ll << 31), "%s: too many K blocks", name);
    P.kblocks = (int)kblocks;
    const int tiles = k * k * P.m_tiles * P.n_tiles;
    // split K so that the grid covers the machine about twice -- but every split pays 128 x Ntile fp32 atomics in its epilogue, so a
    // split must carry enough K blocks to amortise them (measured: a 192x192 3x3 layer on 4 x 32x48 pixels took 1.6 ms with 17
    // splits of 6 K blocks, all of it atomics); every split non-empty
    // (Tried: choosing the split count that minimises rounds x K blocks per item -- fewer, longer items in one or two full rounds.
    // Measured SLOWER, tran_conv1 1.56 -> 2.16 ms: the kernel is bound by L2 bandwidth, not by the per-SM pipeline -- every tap and
    // every M tile re-reads its operand patches, 21 GB through L2 for a 256x256 3x3 layer at 16 x 256x384 in 0.96 ms -- and more,
    // shorter items spread those reads better.  The lever is operand reuse across taps, not the schedule.)
    int splits = (2 * kNumSMs + tiles - 1) / tiles;
    const int max_splits = (P.kblocks + 31) / 32;
    if (splits > max_splits) splits = max_splits;
    if (splits > P.kblocks) splits = P.kblocks;
    if (splits < 1) splits = 1;
    P.kb_per_split = (P.kblocks + splits - 1) / splits;
    P.splits = (P.kblocks + P.kb_per_split - 1) / P.kb_per_split;
    P.total_items = tiles * P.splits;
    {
        uint64_t dims[4] = {(uint64_t)Cs, (uint64_t)Ws, (uint64_t)Hs, (uint64_t)B};
        uint64_t str[3] = {(uint64_t)Cs * 2, (uint64_t)Ws * Cs * 2, (uint64_t)Hs * Ws * Cs * 2};
        uint32_t box[4] = {64, (uint32_t)P.TW, (uint32_t)P.TH, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        int rc = encode_map(&P.tmS, s_nhwc, 4, dims, str, box, es, "wgrad S");
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)Cl, (uint64_t)Wl, (uint64_t)Hl, (uint64_t)B};
        uint64_t str[3] = {(uint64_t)Cl * 2, (uint64_t)Wl * Cl * 2, (uint64_t)Hl * Wl * Cl * 2};
        uint32_t box[4] = {64, (uint32_t)(P.TW * stride), (uint32_t)(P.TH * stride), 1};
        uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        int rc = encode_map(&P.tmL, l_nhwc, 4, dims, str, box, es, "wgrad L");
        if (rc) return rc;
    }
    static PerDevice<size_t> budget_dev;
    size_t budget = budget_dev.cur().load(std::memory_order_relaxed);
    if (budget == 0) {
        cudaFuncAttributes fa;
        MMC_CHECK_CUDA(cudaFuncGetAttributes(&fa, wgrad_tc_kernel));
        size_t avail = 227 * 1024 - fa.sharedSizeBytes;
        MMC_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)avail));
        budget = avail;
        budget_dev.cur().store(avail, std::memory_order_relaxed);
    }
    const size_t stage_bytes = (size_t)(2 + P.n_atoms) * kAtomBytes;
    int stages = (int)((budget - 1024) / stage_bytes);
    if (stages > kWgMaxStages) stages = kWgMaxStages;
    MMC_UNSUPPORTED(stages < 2, "%s: shared memory budget exceeded", name);
    P.num_stages = stages;
    if (const char *g = getenv("MMC_WG_DEBUG")) P.debug = atoi(g);
    const int grid = P.total_items < kNumSMs ? P.total_items : kNumSMs;
    wgrad_tc_kernel<<<grid, kWgThreads, 1024 + (size_t)stages * stage_bytes, (cudaStream_t)stream>>>(P);
    MMC_CHECK_LAUNCH(name);
    return MMC_OK;
}

int mmc_im2col8(const void *x_nhwc8, int64_t B, int H, int W, int k, int stride, int Hs, int Ws, void *out, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && Hs >= 1 && Ws >= 1 && (k == 1 || k == 3 || k == 5) && (stride == 1 || stride == 2), "mmc_im2col8: bad argument");
    if (B == 0) return MMC_OK;
    MMC_CHECK_ARG(x_nhwc8 && out && aligned16(x_nhwc8) && aligned16(out), "mmc_im2col8: NULL or unaligned buffer");
    const int64_t n = B * Hs * Ws * k * k;
    im2col8_kernel<<<elementwise_grid(n, 256), 256, 0, (cudaStream_t)stream>>>((const uint4 *)x_nhwc8, H, W, k, stride, Hs, Ws, n, (uint4 *)out);
    MMC_CHECK_LAUNCH("mmc_im2col8");
    return MMC_OK;
}

int mmc_wgrad_finalize(const float *workspace, int k, int Cs, int Cl, float scale, const float *mask, int accumulate, float *dw, void *stream)
{
    MMC_CHECK_ARG(workspace && dw && k >= 1 && Cs >= 1 && Cl >= 1, "mmc_wgrad_finalize: bad argument");
    const int64_t n = (int64_t)k * k * Cs * Cl;
    wgrad_finalize_kernel<<<elementwise_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(workspace, k * k, Cs, Cl, scale, mask, accumulate, dw);
    MMC_CHECK_LAUNCH("mmc_wgrad_finalize");
    return MMC_OK;
}

}  // extern "C"
