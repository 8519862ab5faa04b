// attention.cu -- the kernels of the RGB-T master codec that are not convolutions (compressai/models/master.py):
//   * token side of Spatial_aligner (master.py:484-742): LayerNorm (optionally fused with the residual add in front of it), GELU, and
//     the windowed multi-head cross-attention itself;
//   * Channel_aligner tail (master.py:193-210): global average of a head (two-pass, fixed order), the same average obtained by
//     linearity from border-corrected channel sums of the trunk output (mmc_conv3x3_mean: no convolution at all), and the
//     per-sample affine gamma * guide + beta.
//
// The Linear layers of the attention block are per-token, so they run as 1x1 tensor-core convolutions on the un-shifted,
// un-partitioned (B, H, W, C) token grid; this file holds what is left.  Everything here is HBM / L2-bound at a few bytes per element
// and the token grids are small (<= 128 x 192 tokens of 96 channels per 512 x 768 master image), so the design goal is one pass over
// the data with coalesced vector accesses, not arithmetic throughput.
//
//   window_attention_kernel: one warp per (window, head).  The cyclic shift (torch.roll by -shift, master.py:664-668), the
//   window partition (master.py:431-443) and their inverses are index arithmetic on the loads / stores; the relative-position
//   bias (master.py:512-522, 549-552) is looked up from the (2 ws - 1)^2 x heads table by coordinate difference and the
//   0 / -100 shift mask (master.py:625-643) is recomputed from the three row / column bands instead of being read.
#include "common.cuh"

namespace mmc {

constexpr int kMaxWin = 16;   // tokens per window (ws <= 4)
constexpr int kHeadDim = 32;  // dim 96 / 3 heads (master.py:718)

__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const __nv_bfloat16 *__restrict__ x, const __nv_bfloat16 *__restrict__ delta,
                                                             const float *__restrict__ w, const float *__restrict__ b, int64_t rows, int C,
                                                             float eps, __nv_bfloat16 *__restrict__ sum_out, __nv_bfloat16 *__restrict__ y)
{
    // one warp per token; C <= 32 * 8
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < rows; r += nwarps) {
        float v[8];
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            v[i] = 0.0f;
            if (c < C) {
                float t = __bfloat162float(x[r * C + c]);
                if (delta) {
                    // the residual stream is bf16: round the sum as the unfused add would, so both consumers see one value
                    const __nv_bfloat16 tb = __float2bfloat16(t + __bfloat162float(delta[r * C + c]));
                    if (sum_out) sum_out[r * C + c] = tb;
                    t = __bfloat162float(tb);
                }
                v[i] = t;
                s += t;
            }
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            if (c < C) q += (v[i] - mean) * (v[i] - mean);
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);   // biased variance, as F.layer_norm
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            if (c < C) y[r * C + c] = __float2bfloat16((v[i] - mean) * rstd * w[c] + b[c]);
        }
    }
}

__global__ void __launch_bounds__(256) gelu_bf16_kernel(const __nv_bfloat162 *__restrict__ x, int64_t n2, __nv_bfloat162 *__restrict__ y)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 v = __bfloat1622float2(x[i]);
        // nn.GELU() default: exact erf form (master.py:464)
        y[i] = __floats2bfloat162_rn(0.5f * v.x * (1.0f + erff(v.x * 0.70710678118654752f)), 0.5f * v.y * (1.0f + erff(v.y * 0.70710678118654752f)));
    }
}

__device__ __forceinline__ int shift_band(int s, int extent, int ws, int shift) { return s < extent - ws ? 0 : (s < extent - shift ? 1 : 2); }

__global__ void __launch_bounds__(128) window_attention_kernel(const __nv_bfloat16 *__restrict__ q, const __nv_bfloat16 *__restrict__ kv,
                                                               const float *__restrict__ table, int B, int H, int W, int heads, int ws, int shift,
                                                               float scale, __nv_bfloat16 *__restrict__ out)
{
    __shared__ float sQ[4][kMaxWin][kHeadDim + 1], sK[4][kMaxWin][kHeadDim + 1], sV[4][kMaxWin][kHeadDim + 1], sP[4][kMaxWin][kMaxWin + 1];
    __shared__ int sPix[4][kMaxWin];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int C = heads * kHeadDim, N = ws * ws, nwx = W / ws, nwy = H / ws;
    const int64_t items = (int64_t)B * nwy * nwx * heads;
    for (int64_t it = (int64_t)blockIdx.x * 4 + wib; it < items; it += (int64_t)gridDim.x * 4) {
        const int h = (int)(it % heads);
        int64_t t = it / heads;
        const int wx = (int)(t % nwx); t /= nwx;
        const int wy = (int)(t % nwy);
        const int b = (int)(t / nwy);
        // window token n = (ty, tx) sits at shifted coordinates (wy ws + ty, wx ws + tx) = original coordinates rolled by +shift
        for (int n = 0; n < N; ++n) {
            int oy = wy * ws + n / ws + shift, ox = wx * ws + n % ws + shift;
            if (oy >= H) oy -= H;
            if (ox >= W) ox -= W;
            const int64_t pix = ((int64_t)b * H + oy) * W + ox;
            if (lane == 0) sPix[wib][n] = (int)(pix - (int64_t)b * H * W);
            sQ[wib][n][lane] = __bfloat162float(q[pix * C + h * kHeadDim + lane]) * scale;
            sK[wib][n][lane] = __bfloat162float(kv[pix * 2 * C + h * kHeadDim + lane]);
            sV[wib][n][lane] = __bfloat162float(kv[pix * 2 * C + C + h * kHeadDim + lane]);
        }
        __syncwarp();
        // scores: lane -> query i = lane / 2, keys j in [8 (lane & 1), +8)
        {
            const int i = lane >> 1, j0 = (lane & 1) * 8;
            float sc[8];
            float mx = -INFINITY;
            if (i < N) {
                const int ty = i / ws, tx = i % ws;
                const int ri = shift ? 3 * shift_band(wy * ws + ty, H, ws, shift) + shift_band(wx * ws + tx, W, ws, shift) : 0;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int j = j0 + jj;
                    float a = -INFINITY;
                    if (j < N) {
                        a = 0.0f;
#pragma unroll
                        for (int d = 0; d < kHeadDim; ++d) a = fmaf(sQ[wib][i][d], sK[wib][j][d], a);
                        const int uy = j / ws, ux = j % ws;
                        a += table[((ty - uy + ws - 1) * (2 * ws - 1) + (tx - ux + ws - 1)) * heads + h];
                        if (shift) {
                            const int rj = 3 * shift_band(wy * ws + uy, H, ws, shift) + shift_band(wx * ws + ux, W, ws, shift);
                            if (rj != ri) a -= 100.0f;
                        }
                    }
                    sc[jj] = a;
                    mx = fmaxf(mx, a);
                }
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            float sum = 0.0f;
            if (i < N) {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    sc[jj] = (j0 + jj < N) ? __expf(sc[jj] - mx) : 0.0f;
                    sum += sc[jj];
                }
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            if (i < N) {
                const float inv = 1.0f / sum;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) sP[wib][i][j0 + jj] = sc[jj] * inv;
            }
        }
        __syncwarp();
        // context: lane = channel d of the head
        for (int i = 0; i < N; ++i) {
            float a = 0.0f;
            for (int j = 0; j < N; ++j) a = fmaf(sP[wib][i][j], sV[wib][j][lane], a);
            out[((int64_t)b * H * W + sPix[wib][i]) * C + h * kHeadDim + lane] = __float2bfloat16(a);
        }
        __syncwarp();
    }
}

// ---- Channel_aligner tail (master.py:193-210): global average of an fp32 NHWC map, and guide * gamma + beta --------------
// Two passes, both with a fixed summation order, so a sample's result does not depend on the batch it is part of (images are
// independent units): pass 1 -- one CTA per (32-channel group, sample, row split) sums its rows; pass 2 adds the splits.
__global__ void __launch_bounds__(256) channel_mean_partial_kernel(const float *__restrict__ x, int64_t HW, int C, int splits, float *__restrict__ part)
{
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane, b = blockIdx.y, sp = blockIdx.z;
    const int64_t per = (HW + splits - 1) / splits, r0 = sp * per, r1 = (r0 + per < HW) ? r0 + per : HW;
    float s = 0.0f;
    if (c < C)
        for (int64_t r = r0 + ty; r < r1; r += 8) s += x[((int64_t)b * HW + r) * C + c];
    red[ty][lane] = s;
    __syncthreads();
    if (ty == 0 && c < C) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][lane];
        part[((int64_t)b * splits + sp) * C + c] = t;
    }
}

__global__ void __launch_bounds__(256) channel_mean_final_kernel(const float *__restrict__ part, int64_t HW, int C, int splits, int64_t n, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (b, c)
    if (i >= n) return;
    const int64_t b = i / C;
    const int c = (int)(i % C);
    float t = 0.0f;
    for (int sp = 0; sp < splits; ++sp) t += part[(b * splits + sp) * C + c];
    out[i] = t / (float)HW;
}

__global__ void __launch_bounds__(256) channel_affine_kernel(const __nv_bfloat16 *__restrict__ x, const float *__restrict__ gamma,
                                                             const float *__restrict__ beta, int64_t HW, int C, int64_t n,
                                                             __nv_bfloat16 *__restrict__ y)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t b = i / (HW * C);
        y[i] = __float2bfloat16(fmaf(gamma[b * C + c], __bfloat162float(x[i]), beta[b * C + c]));
    }
}

// 8 channels per thread (C % 8 == 0, 16-byte aligned buffers)
__global__ void __launch_bounds__(256) channel_affine_v8_kernel(const uint4 *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                int64_t HW, int C8, int64_t n8, uint4 *__restrict__ y)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        const int64_t b = i / (HW * C8);
        const float4 *g = reinterpret_cast<const float4 *>(gamma + (b * C8 + c8) * 8), *bt = reinterpret_cast<const float4 *>(beta + (b * C8 + c8) * 8);
        const float4 g0 = g[0], g1 = g[1], b0 = bt[0], b1 = bt[1];
        const uint4 q = x[i];
        const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&q);
        const float2 v0 = __bfloat1622float2(p[0]), v1 = __bfloat1622float2(p[1]), v2 = __bfloat1622float2(p[2]), v3 = __bfloat1622float2(p[3]);
        uint4 o;
        __nv_bfloat162 *po = reinterpret_cast<__nv_bfloat162 *>(&o);
        po[0] = __floats2bfloat162_rn(fmaf(g0.x, v0.x, b0.x), fmaf(g0.y, v0.y, b0.y));
        po[1] = __floats2bfloat162_rn(fmaf(g0.z, v1.x, b0.z), fmaf(g0.w, v1.y, b0.w));
        po[2] = __floats2bfloat162_rn(fmaf(g1.x, v2.x, b1.x), fmaf(g1.y, v2.y, b1.y));
        po[3] = __floats2bfloat162_rn(fmaf(g1.z, v3.x, b1.z), fmaf(g1.w, v3.y, b1.w));
        y[i] = o;
    }
}

// ---- spatial mean of a 3x3 / stride-1 / padding-1 convolution WITHOUT computing the convolution -------------------------------
// Channel_aligner only uses conv5 / conv6 through AdaptiveAvgPool2d(1) (master.py:193-194,205-206).  The mean over output positions
// of sum_tap W[tap] . t[p + tap] is, by linearity, sum_tap W[tap] . S_tap / HW with S_tap = the sum of t over the positions tap can
// reach = total - excluded border row - excluded border column + their corner.  So one pass over t (9 channel vectors per sample:
// total, first / last row, first / last column, four corners) and a 9 x C x O matrix-vector product replace two 256 -> 64 convolutions
// on the full-resolution map and their 2 x 200 MB fp32 outputs.  Fixed summation orders throughout (batch-independent results).
__global__ void __launch_bounds__(256) border_total_partial_kernel(const uint4 *__restrict__ t, int64_t HW, int C8, int splits, float *__restrict__ part)
{
    __shared__ float red[8][32][9];
    const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c8 = blockIdx.x * 32 + lane, b = blockIdx.y, sp = blockIdx.z;
    const int64_t per = (HW + splits - 1) / splits, r0 = sp * per, r1 = (r0 + per < HW) ? r0 + per : HW;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c8 < C8) {
        const uint4 *src = t + (int64_t)b * HW * C8 + c8;
        int64_t r = r0 + ty;
        for (; r + 24 < r1; r += 32) {          // four independent 16-byte loads in flight per thread
            const uint4 q4[4] = {src[r * C8], src[(r + 8) * C8], src[(r + 16) * C8], src[(r + 24) * C8]};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&q4[u]);
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 v = __bfloat1622float2(p[j]); acc[2 * j] += v.x; acc[2 * j + 1] += v.y; }
            }
        }
        for (; r < r1; r += 8) {
            const uint4 q = src[r * C8];
            const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 v = __bfloat1622float2(p[j]); acc[2 * j] += v.x; acc[2 * j + 1] += v.y; }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) red[ty][lane][k] = acc[k];
    __syncthreads();
    if (ty == 0 && c8 < C8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += red[i][lane][k];
            part[((int64_t)b * splits + sp) * (C8 * 8) + c8 * 8 + k] = s;
        }
    }
}

// sums[b][0..8][c]: total, row 0, row H-1, column 0, column W-1, corners (0,0) (0,W-1) (H-1,0) (H-1,W-1)
// edges: one CTA per (32-channel group, sample, edge), 8 position lanes, fixed-order reduction
__global__ void __launch_bounds__(256) border_edges_kernel(const __nv_bfloat16 *__restrict__ t, int H, int W, int C, float *__restrict__ sums)
{
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane, b = blockIdx.y, e = blockIdx.z;
    const int len = (e < 2) ? W : H;
    const int64_t first = (e == 1) ? (int64_t)(H - 1) * W : (e == 3 ? W - 1 : 0), step = (e < 2) ? 1 : W;
    const __nv_bfloat16 *img = t + (int64_t)b * H * W * C + c;
    float s = 0.0f;
    if (c < C)
        for (int i = ty; i < len; i += 8) s += __bfloat162float(img[(first + (int64_t)i * step) * C]);
    red[ty][lane] = s;
    __syncthreads();
    if (ty == 0 && c < C) {
        float v = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) v += red[i][lane];
        sums[((int64_t)b * 9 + 1 + e) * C + c] = v;
    }
}

// totals (sum of the row splits) and the four corner pixels; one thread per (b, c)
__global__ void __launch_bounds__(256) border_sums_kernel(const __nv_bfloat16 *__restrict__ t, const float *__restrict__ part, int H, int W, int C,
                                                          int splits, int64_t n, float *__restrict__ sums)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t b = i / C;
    const int c = (int)(i % C);
    const __nv_bfloat16 *img = t + b * H * W * C + c;
    float total = 0.0f;
    for (int sp = 0; sp < splits; ++sp) total += part[(b * splits + sp) * C + c];
    float *o = sums + b * 9 * C + c;
    o[0] = total;
    o[5 * C] = __bfloat162float(img[0]);
    o[6 * C] = __bfloat162float(img[(int64_t)(W - 1) * C]);
    o[7 * C] = __bfloat162float(img[(int64_t)(H - 1) * W * C]);
    o[8 * C] = __bfloat162float(img[((int64_t)(H - 1) * W + W - 1) * C]);
}

// out[b][o] = bias[o] + (1 / HW) sum_{c, ky, kx} w[o][c][ky][kx] * S(ky - 1, kx - 1)[b][c]; one warp per (b, o), lanes over c
__global__ void __launch_bounds__(256) conv3x3_mean_kernel(const float *__restrict__ sums, const float *__restrict__ w, const float *__restrict__ bias,
                                                           int C, int O, float inv_hw, int64_t n, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= n) return;
    const int64_t b = item / O;
    const int o = (int)(item % O);
    const float *s = sums + b * 9 * C;
    float acc = 0.0f;
    for (int c = lane; c < C; c += 32) {
        const float T = s[c], r0 = s[C + c], rl = s[2 * C + c], c0 = s[3 * C + c], cl = s[4 * C + c];
        const float k00 = s[5 * C + c], k0l = s[6 * C + c], kl0 = s[7 * C + c], kll = s[8 * C + c];
        const float *wk = w + ((int64_t)o * C + c) * 9;
        // tap offset -1 cannot reach the LAST row / column of t, offset +1 cannot reach the FIRST
        const float rex[3] = {rl, 0.0f, r0}, cex[3] = {cl, 0.0f, c0};
        const float cor[3][3] = {{kll, 0.0f, kl0}, {0.0f, 0.0f, 0.0f}, {k0l, 0.0f, k00}};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) acc = fmaf(wk[ky * 3 + kx], T - rex[ky] - cex[kx] + cor[ky][kx], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[item] = (bias ? bias[o] : 0.0f) + acc * inv_hw;
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_layernorm_bf16(const void *x, const void *delta, const float *weight, const float *bias, int64_t rows, int C, float eps,
                       void *sum_out, void *y, void *stream)
{
    MMC_CHECK_ARG(rows >= 0 && C >= 1 && C <= 256, "mmc_layernorm_bf16: needs rows >= 0 and 1 <= C <= 256 (got %d)", C);
    MMC_CHECK_ARG(!sum_out || delta, "mmc_layernorm_bf16: sum_out without delta");
    if (rows == 0) return MMC_OK;
    MMC_CHECK_ARG(x && weight && bias && y, "mmc_layernorm_bf16: NULL buffer");
    layernorm_bf16_kernel<<<elementwise_grid(rows, 8), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)x, (const __nv_bfloat16 *)delta, weight, bias, rows, C, eps, (__nv_bfloat16 *)sum_out, (__nv_bfloat16 *)y);
    MMC_CHECK_LAUNCH("mmc_layernorm_bf16");
    return MMC_OK;
}

int mmc_gelu_bf16(const void *x, int64_t n, void *y, void *stream)
{
    MMC_CHECK_ARG(n >= 0 && n % 2 == 0, "mmc_gelu_bf16: n must be even and >= 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && y, "mmc_gelu_bf16: NULL buffer");
    gelu_bf16_kernel<<<elementwise_grid(n / 2, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat162 *)x, n / 2, (__nv_bfloat162 *)y);
    MMC_CHECK_LAUNCH("mmc_gelu_bf16");
    return MMC_OK;
}

int mmc_window_attention(const void *q, const void *kv, const float *bias_table, int B, int H, int W, int heads, int head_dim,
                         int window, int shift, float scale, void *out, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && heads >= 1, "mmc_window_attention: bad shape");
    MMC_UNSUPPORTED(head_dim != kHeadDim, "mmc_window_attention: head_dim must be %d (got %d)", kHeadDim, head_dim);
    MMC_UNSUPPORTED(window < 1 || window > 4, "mmc_window_attention: window size must be in 1..4 (got %d)", window);
    MMC_CHECK_ARG(H % window == 0 && W % window == 0, "mmc_window_attention: the %dx%d token grid is not a multiple of the window size %d", H, W, window);
    MMC_CHECK_ARG(shift >= 0 && shift < window, "mmc_window_attention: shift must be in [0, window)");
    MMC_CHECK_ARG((int64_t)H * W < (1ll << 31), "mmc_window_attention: token grid too large");
    if (B == 0) return MMC_OK;
    MMC_CHECK_ARG(q && kv && bias_table && out, "mmc_window_attention: NULL buffer");
    const int64_t items = (int64_t)B * (H / window) * (W / window) * heads;
    window_attention_kernel<<<elementwise_grid(items, 4), 128, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)q, (const __nv_bfloat16 *)kv, bias_table, B, H, W, heads, window, shift, scale, (__nv_bfloat16 *)out);
    MMC_CHECK_LAUNCH("mmc_window_attention");
    return MMC_OK;
}

static int channel_mean_splits(int64_t HW) { int64_t s = (HW + 511) / 512; return (int)(s < 1 ? 1 : (s > 128 ? 128 : s)); }

int mmc_channel_mean_workspace(int B, int64_t HW, int C, size_t *bytes)
{
    MMC_CHECK_ARG(B >= 0 && HW >= 1 && C >= 1 && bytes, "mmc_channel_mean_workspace: bad argument");
    *bytes = (size_t)B * channel_mean_splits(HW) * C * sizeof(float);
    return MMC_OK;
}

int mmc_channel_mean(const float *x, int B, int64_t HW, int C, void *workspace, float *out, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && B <= 65535 && HW >= 1 && C >= 1, "mmc_channel_mean: bad shape");
    if (B == 0) return MMC_OK;
    MMC_CHECK_ARG(x && out && workspace, "mmc_channel_mean: NULL buffer");
    const int splits = channel_mean_splits(HW);
    channel_mean_partial_kernel<<<dim3((C + 31) / 32, B, splits), 256, 0, (cudaStream_t)stream>>>(x, HW, C, splits, (float *)workspace);
    MMC_CHECK_LAUNCH("mmc_channel_mean");
    const int64_t n = (int64_t)B * C;
    channel_mean_final_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float *)workspace, HW, C, splits, n, out);
    MMC_CHECK_LAUNCH("mmc_channel_mean");
    return MMC_OK;
}

int mmc_channel_affine_bf16(const void *x, const float *gamma, const float *beta, int B, int64_t HW, int C, void *y, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && HW >= 1 && C >= 1, "mmc_channel_affine_bf16: bad shape");
    if (B == 0) return MMC_OK;
    MMC_CHECK_ARG(x && gamma && beta && y, "mmc_channel_affine_bf16: NULL buffer");
    const int64_t n = (int64_t)B * HW * C;
    if (C % 8 == 0 && aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta)) {
        channel_affine_v8_kernel<<<elementwise_grid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4 *)x, gamma, beta, HW, C / 8, n / 8, (uint4 *)y);
        MMC_CHECK_LAUNCH("mmc_channel_affine_bf16");
        return MMC_OK;
    }
    channel_affine_kernel<<<elementwise_grid(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x, gamma, beta, HW, C, n, (__nv_bfloat16 *)y);
    MMC_CHECK_LAUNCH("mmc_channel_affine_bf16");
    return MMC_OK;
}

int mmc_conv3x3_mean_workspace(int B, int H, int W, int C, size_t *bytes)
{
    MMC_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && C >= 1 && bytes, "mmc_conv3x3_mean_workspace: bad argument");
    *bytes = ((size_t)B * channel_mean_splits((int64_t)H * W) * C + (size_t)B * 9 * C) * sizeof(float);
    return MMC_OK;
}

int mmc_conv3x3_mean(const void *t, int B, int H, int W, int C, const float *weight, const float *bias, int O, void *workspace, float *out, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && B <= 65535 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0 && O >= 1, "mmc_conv3x3_mean: bad shape (C must be a multiple of 8)");
    if (B == 0) return MMC_OK;
    MMC_CHECK_ARG(t && weight && workspace && out && aligned16(t), "mmc_conv3x3_mean: NULL or unaligned buffer");
    const int64_t HW = (int64_t)H * W;
    const int splits = channel_mean_splits(HW);
    float *part = (float *)workspace, *sums = part + (size_t)B * splits * C;
    cudaStream_t st = (cudaStream_t)stream;
    border_total_partial_kernel<<<dim3((C / 8 + 31) / 32, B, splits), 256, 0, st>>>((const uint4 *)t, HW, C / 8, splits, part);
    MMC_CHECK_LAUNCH("mmc_conv3x3_mean");
    const int64_t n = (int64_t)B * C;
    border_sums_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const __nv_bfloat16 *)t, part, H, W, C, splits, n, sums);
    MMC_CHECK_LAUNCH("mmc_conv3x3_mean");
    border_edges_kernel<<<dim3((C + 31) / 32, B, 4), 256, 0, st>>>((const __nv_bfloat16 *)t, H, W, C, sums);
    MMC_CHECK_LAUNCH("mmc_conv3x3_mean");
    const int64_t items = (int64_t)B * O;
    conv3x3_mean_kernel<<<(unsigned)((items + 7) / 8), 256, 0, st>>>(sums, weight, bias, C, O, 1.0f / (float)HW, items, out);
    MMC_CHECK_LAUNCH("mmc_conv3x3_mean");
    return MMC_OK;
}

}  // extern "C"
