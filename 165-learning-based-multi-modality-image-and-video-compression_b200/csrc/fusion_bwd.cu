// fusion_bwd.cu -- backward passes of the non-convolution steps of the cross-modality fusion layers:
//   * ESA gate (compressai/models/google.py:1445-1459): 7x7 / stride-3 max-pool (forward variant that records the arg-max, and its
//     adjoint), adjoint of the bilinear upsampling, backward of x * sigmoid(c);
//   * token side of Spatial_aligner (compressai/models/master.py:463-568, 572-706): LayerNorm (with the residual add in front of
//     it), GELU and the windowed multi-head cross-attention.
// With these the training step of the fusion models keeps every activation-sized pass on libmmcodec kernels (the convolutions
// and Linear layers already were: conv_tc.cu / wgrad_tc.cu).  All of them are HBM / L2-bound passes over NHWC bf16 maps; every
// adjoint is written as a GATHER (one thread owns one gradient element and sums its contributions in fp32 in a fixed order), so
// there are no atomics on activation-sized tensors and results do not depend on scheduling.  Parameter gradients (LayerNorm
// weight / bias, the relative-position bias table) are reduced per CTA in shared memory and flushed with one atomic per entry.
#include "common.cuh"

namespace mmc {

// ---------------------------------------------------------------------------------------------------------------------------
// max-pool: forward with arg-max (window-local index ky * k + kx, first maximum in scan order; a NaN takes over and the last
// NaN wins -- the selection rule of torch's max_pool2d kernels), and the adjoint as a gather over the <= ceil(k / stride)^2
// windows that cover an input pixel
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_idx_kernel(const __nv_bfloat16 *__restrict__ x, int H, int W, int C, int k, int stride, int Ho, int Wo,
                                                          int64_t n, __nv_bfloat16 *__restrict__ y, uint8_t *__restrict__ idx)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho);
        const int64_t b = t / Ho;
        float m = -INFINITY;
        int best = 0;
        for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx) {
                const float v = __bfloat162float(x[((b * H + oy * stride + ky) * W + ox * stride + kx) * C + c]);
                if (v > m || v != v) { m = v; best = ky * k + kx; }
            }
        y[i] = __float2bfloat16(m);
        idx[i] = (uint8_t)best;
    }
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const __nv_bfloat16 *__restrict__ gy, const uint8_t *__restrict__ idx, int H, int W, int C, int k,
                                                          int stride, int Ho, int Wo, int64_t n, __nv_bfloat16 *__restrict__ dx)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int ix = (int)(t % W); t /= W;
        const int iy = (int)(t % H);
        const int64_t b = t / H;
        // windows oy with oy * stride <= iy < oy * stride + k
        const int oy_hi = min(iy / stride, Ho - 1), ox_hi = min(ix / stride, Wo - 1);
        const int oy_lo = max(0, (iy - k + stride) / stride), ox_lo = max(0, (ix - k + stride) / stride);
        float acc = 0.0f;
        for (int oy = oy_lo; oy <= oy_hi; ++oy)
            for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                const int64_t o = ((b * Ho + oy) * Wo + ox) * C + c;
                const int want = (iy - oy * stride) * k + (ix - ox * stride);
                if ((int)idx[o] == want) acc += __bfloat162float(gy[o]);
            }
        dx[i] = __float2bfloat16(acc);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// adjoint of F.interpolate(mode="bilinear", align_corners=False): small[sy][sx] collects weight(oy, sy) * weight(ox, sx) * g[oy][ox]
// over the destination pixels whose two source neighbours include (sy, sx); the weights come from the SAME source-coordinate
// function as the forward kernel (esa.cu), so forward and adjoint agree to the last bit of the interpolation weights
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src_b(int dst, float scale, int in, int *i0, int *i1, float *l1)
{
    float s = ((float)dst + 0.5f) * scale - 0.5f;
    s = s < 0.0f ? 0.0f : s;
    const int a = (int)s;
    *i0 = a;
    *i1 = a + (a < in - 1 ? 1 : 0);
    *l1 = s - (float)a;
}
__device__ __forceinline__ float bilinear_weight(int dst, float scale, int in, int src)
{
    int i0, i1;
    float l1;
    bilinear_src_b(dst, scale, in, &i0, &i1, &l1);
    return (i0 == src ? 1.0f - l1 : 0.0f) + (i1 == src ? l1 : 0.0f);
}
// destination range that can touch source index `src` (conservative by one on both sides; the weight is exact)
__device__ __forceinline__ void bilinear_dst_range(int src, float scale, int out, int *lo, int *hi)
{
    const float inv = 1.0f / scale;
    int a = (int)floorf(((float)src - 0.5f) * inv - 0.5f) - 1;
    int b = (int)ceilf(((float)src + 1.5f) * inv - 0.5f) + 1;
    *lo = a < 0 ? 0 : a;
    *hi = b > out - 1 ? out - 1 : b;
}

__global__ void __launch_bounds__(256) upsample_bwd_v8_kernel(const uint4 *__restrict__ g, int hs, int ws, int H, int W, int C8, float sy, float sx,
                                                              int64_t n8, uint4 *__restrict__ dsmall)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8);
        int64_t t = i / C8;
        const int px = (int)(t % ws); t /= ws;
        const int py = (int)(t % hs);
        const int64_t b = t / hs;
        int y_lo, y_hi, x_lo, x_hi;
        bilinear_dst_range(py, sy, H, &y_lo, &y_hi);
        bilinear_dst_range(px, sx, W, &x_lo, &x_hi);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int oy = y_lo; oy <= y_hi; ++oy) {
            const float wy = bilinear_weight(oy, sy, hs, py);
            if (wy == 0.0f) continue;
            const uint4 *row = g + ((b * H + oy) * (int64_t)W) * C8 + c;
            for (int ox = x_lo; ox <= x_hi; ++ox) {
                const float w = wy * bilinear_weight(ox, sx, ws, px);
                if (w == 0.0f) continue;
                const uint4 q = row[(int64_t)ox * C8];
                const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&q);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 v = __bfloat1622float2(p[j]);
                    acc[2 * j] = fmaf(w, v.x, acc[2 * j]);
                    acc[2 * j + 1] = fmaf(w, v.y, acc[2 * j + 1]);
                }
            }
        }
        uint4 o;
        __nv_bfloat162 *po = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) po[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
        dsmall[i] = o;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// y = x * sigmoid(c):  dx = g * s,  dc = g * x * s * (1 - s)
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sigmoid_gate_bwd_kernel(const uint4 *__restrict__ g, const uint4 *__restrict__ x, const uint4 *__restrict__ c,
                                                               int64_t n8, uint4 *__restrict__ dx, uint4 *__restrict__ dc)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 gv = g[i], xv = x[i], cv = c[i];
        uint4 o1, o2;
        const __nv_bfloat162 *gp = reinterpret_cast<const __nv_bfloat162 *>(&gv), *xp = reinterpret_cast<const __nv_bfloat162 *>(&xv);
        const __nv_bfloat162 *cp = reinterpret_cast<const __nv_bfloat162 *>(&cv);
        __nv_bfloat162 *p1 = reinterpret_cast<__nv_bfloat162 *>(&o1), *p2 = reinterpret_cast<__nv_bfloat162 *>(&o2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 gg = __bfloat1622float2(gp[j]), xx = __bfloat1622float2(xp[j]), cc = __bfloat1622float2(cp[j]);
            const float s0 = 1.0f / (1.0f + __expf(-cc.x)), s1 = 1.0f / (1.0f + __expf(-cc.y));
            p1[j] = __floats2bfloat162_rn(gg.x * s0, gg.y * s1);
            p2[j] = __floats2bfloat162_rn(gg.x * xx.x * s0 * (1.0f - s0), gg.y * xx.y * s1 * (1.0f - s1));
        }
        dx[i] = o1;
        dc[i] = o2;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// GELU (exact erf form, master.py:464):  d/dx [x Phi(x)] = Phi(x) + x phi(x)
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_grad(float v)
{
    const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * v * v);
    return cdf + v * pdf;
}
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const __nv_bfloat162 *__restrict__ g, const __nv_bfloat162 *__restrict__ x, int64_t n2,
                                                       __nv_bfloat162 *__restrict__ dx)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 v = __bfloat1622float2(x[i]), gg = __bfloat1622float2(g[i]);
        dx[i] = __floats2bfloat162_rn(gg.x * gelu_grad(v.x), gg.y * gelu_grad(v.y));
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// LayerNorm backward, one warp per token (C <= 256).  v = the normalised row (x, or the bf16 sum x + delta the forward kernel
// wrote), y = (v - mean) rstd w + b:   dv = rstd (g w - mean_c(g w) - xhat mean_c(g w xhat))  [+ g_sum: the gradient that reaches
// the same row through the residual stream];  dw = sum_rows g xhat,  db = sum_rows g.
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const __nv_bfloat16 *__restrict__ g, const __nv_bfloat16 *__restrict__ v_in,
                                                            const __nv_bfloat16 *__restrict__ g_sum, const float *__restrict__ w, int64_t rows, int C,
                                                            float eps, __nv_bfloat16 *__restrict__ dv, float *__restrict__ dw, float *__restrict__ db)
{
    __shared__ float red_w[8][256], red_b[8][256];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    float aw[8], ab[8], wv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        aw[i] = ab[i] = 0.0f;
        const int c = lane + 32 * i;
        wv[i] = c < C ? w[c] : 0.0f;
    }
    for (int64_t r = warp; r < rows; r += nwarps) {
        float v[8], gg[8];
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            v[i] = gg[i] = 0.0f;
            if (c < C) {
                v[i] = __bfloat162float(v_in[r * C + c]);
                gg[i] = __bfloat162float(g[r * C + c]);
                s += v[i];
            }
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (lane + 32 * i < C) q += (v[i] - mean) * (v[i] - mean);
        const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
        float m1 = 0.0f, m2 = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (lane + 32 * i < C) {
                v[i] = (v[i] - mean) * rstd;          // xhat
                const float gw = gg[i] * wv[i];
                m1 += gw;
                m2 += gw * v[i];
                aw[i] += gg[i] * v[i];
                ab[i] += gg[i];
            }
        }
        m1 = warp_sum(m1) / (float)C;
        m2 = warp_sum(m2) / (float)C;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                float d = rstd * (gg[i] * wv[i] - m1 - v[i] * m2);
                if (g_sum) d += __bfloat162float(g_sum[r * C + c]);
                dv[r * C + c] = __float2bfloat16(d);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { red_w[wib][lane + 32 * i] = aw[i]; red_b[wib][lane + 32 * i] = ab[i]; }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sw = 0.0f, sb = 0.0f;
        for (int k = 0; k < nw; ++k) { sw += red_w[k][c]; sb += red_b[k][c]; }
        atomicAdd(dw + c, sw);
        atomicAdd(db + c, sb);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// windowed multi-head cross-attention backward (forward: attention.cu window_attention_kernel; master.py:535-568).  One warp per
// (window, head): reloads q (scaled), k, v, recomputes the probabilities P exactly as the forward kernel does, then
//   dV = P^T dO,   dP = dO V^T,   dS = P o (dP - rowsum(dP o P)),   dQ = scale dS K,   dK = dS^T (scale Q),
//   dTable[rel(i, j)][h] += dS[i][j]   (the 0 / -100 shift mask is a constant).
// Every (pixel, head) slice of dq / dkv belongs to exactly one warp: plain stores.  The bias-table gradient is reduced per CTA in
// shared memory and flushed with one atomic per entry.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int kMaxWinB = 16;
constexpr int kHeadDimB = 32;
constexpr int kMaxTable = 49 * 8;   // (2 ws - 1)^2 entries x heads, ws <= 4, heads <= 8

__device__ __forceinline__ int shift_band_b(int s, int extent, int ws, int shift) { return s < extent - ws ? 0 : (s < extent - shift ? 1 : 2); }

__global__ void __launch_bounds__(128) window_attention_bwd_kernel(const __nv_bfloat16 *__restrict__ q, const __nv_bfloat16 *__restrict__ kv,
                                                                   const float *__restrict__ table, const __nv_bfloat16 *__restrict__ dout, int B, int H,
                                                                   int W, int heads, int ws, int shift, float scale, __nv_bfloat16 *__restrict__ dq,
                                                                   __nv_bfloat16 *__restrict__ dkv, float *__restrict__ dtable)
{
    __shared__ float sQ[4][kMaxWinB][kHeadDimB + 1], sK[4][kMaxWinB][kHeadDimB + 1], sV[4][kMaxWinB][kHeadDimB + 1], sO[4][kMaxWinB][kHeadDimB + 1];
    __shared__ float sP[4][kMaxWinB][kMaxWinB + 1], sS[4][kMaxWinB][kMaxWinB + 1];
    __shared__ int sPix[4][kMaxWinB];
    __shared__ float sT[kMaxTable];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int C = heads * kHeadDimB, N = ws * ws, nwx = W / ws, nwy = H / ws, tw = 2 * ws - 1;
    for (int i = threadIdx.x; i < kMaxTable; i += blockDim.x) sT[i] = 0.0f;
    __syncthreads();
    const int64_t items = (int64_t)B * nwy * nwx * heads;
    for (int64_t it = (int64_t)blockIdx.x * 4 + wib; it < items; it += (int64_t)gridDim.x * 4) {
        const int h = (int)(it % heads);
        int64_t t = it / heads;
        const int wx = (int)(t % nwx); t /= nwx;
        const int wy = (int)(t % nwy);
        const int b = (int)(t / nwy);
        for (int n = 0; n < N; ++n) {
            int oy = wy * ws + n / ws + shift, ox = wx * ws + n % ws + shift;
            if (oy >= H) oy -= H;
            if (ox >= W) ox -= W;
            const int64_t pix = ((int64_t)b * H + oy) * W + ox;
            if (lane == 0) sPix[wib][n] = (int)(pix - (int64_t)b * H * W);
            sQ[wib][n][lane] = __bfloat162float(q[pix * C + h * kHeadDimB + lane]) * scale;
            sK[wib][n][lane] = __bfloat162float(kv[pix * 2 * C + h * kHeadDimB + lane]);
            sV[wib][n][lane] = __bfloat162float(kv[pix * 2 * C + C + h * kHeadDimB + lane]);
            sO[wib][n][lane] = __bfloat162float(dout[pix * C + h * kHeadDimB + lane]);
        }
        __syncwarp();
        // lane -> query i = lane / 2, keys j in [8 (lane & 1), +8): probabilities (as the forward kernel), then dP and dS
        {
            const int i = lane >> 1, j0 = (lane & 1) * 8;
            float sc[8], dp[8];
            float mx = -INFINITY;
            if (i < N) {
                const int ty = i / ws, tx = i % ws;
                const int ri = shift ? 3 * shift_band_b(wy * ws + ty, H, ws, shift) + shift_band_b(wx * ws + tx, W, ws, shift) : 0;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int j = j0 + jj;
                    float a = -INFINITY, d = 0.0f;
                    if (j < N) {
                        a = 0.0f;
#pragma unroll
                        for (int e = 0; e < kHeadDimB; ++e) {
                            a = fmaf(sQ[wib][i][e], sK[wib][j][e], a);
                            d = fmaf(sO[wib][i][e], sV[wib][j][e], d);
                        }
                        const int uy = j / ws, ux = j % ws;
                        a += table[((ty - uy + ws - 1) * tw + (tx - ux + ws - 1)) * heads + h];
                        if (shift) {
                            const int rj = 3 * shift_band_b(wy * ws + uy, H, ws, shift) + shift_band_b(wx * ws + ux, W, ws, shift);
                            if (rj != ri) a -= 100.0f;
                        }
                    }
                    sc[jj] = a;
                    dp[jj] = d;
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
            float dot = 0.0f;
            if (i < N) {
                const float inv = 1.0f / sum;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) { sc[jj] *= inv; dot += sc[jj] * dp[jj]; }
            }
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            if (i < N) {
                const int ty = i / ws, tx = i % ws;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int j = j0 + jj;
                    const float ds = sc[jj] * (dp[jj] - dot);
                    sP[wib][i][j] = sc[jj];
                    sS[wib][i][j] = ds;
                    if (j < N) {
                        const int uy = j / ws, ux = j % ws;
                        atomicAdd(&sT[((ty - uy + ws - 1) * tw + (tx - ux + ws - 1)) * heads + h], ds);
                    }
                }
            }
        }
        __syncwarp();
        // lane = channel d of the head
        for (int j = 0; j < N; ++j) {
            float a_v = 0.0f, a_k = 0.0f, a_q = 0.0f;
            for (int i = 0; i < N; ++i) {
                a_v = fmaf(sP[wib][i][j], sO[wib][i][lane], a_v);     // dV[j] = sum_i P[i][j] dO[i]
                a_k = fmaf(sS[wib][i][j], sQ[wib][i][lane], a_k);     // dK[j] = sum_i dS[i][j] (scale q[i])
                a_q = fmaf(sS[wib][j][i], sK[wib][i][lane], a_q);     // dQ[j] = scale sum_i dS[j][i] K[i]
            }
            const int64_t pix = (int64_t)b * H * W + sPix[wib][j];
            dq[pix * C + h * kHeadDimB + lane] = __float2bfloat16(a_q * scale);
            dkv[pix * 2 * C + h * kHeadDimB + lane] = __float2bfloat16(a_k);
            dkv[pix * 2 * C + C + h * kHeadDimB + lane] = __float2bfloat16(a_v);
        }
        __syncwarp();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tw * tw * heads; i += blockDim.x)
        if (sT[i] != 0.0f) atomicAdd(dtable + i, sT[i]);
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_maxpool_nhwc_bf16_idx(const void *x, int B, int H, int W, int C, int k, int stride, void *y, void *idx, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && C >= 1, "mmc_maxpool_nhwc_bf16_idx: bad shape");
    MMC_CHECK_ARG(k >= 1 && k <= 15 && stride >= 1 && H >= k && W >= k, "mmc_maxpool_nhwc_bf16_idx: window %d does not fit the %dx%d map (k <= 15)", k, H, W);
    const int Ho = (H - k) / stride + 1, Wo = (W - k) / stride + 1;
    const int64_t n = (int64_t)B * Ho * Wo * C;
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && y && idx, "mmc_maxpool_nhwc_bf16_idx: NULL buffer");
    maxpool_idx_kernel<<<elementwise_grid(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x, H, W, C, k, stride, Ho, Wo, n,
                                                                                  (__nv_bfloat16 *)y, (uint8_t *)idx);
    MMC_CHECK_LAUNCH("mmc_maxpool_nhwc_bf16_idx");
    return MMC_OK;
}

int mmc_maxpool_nhwc_bf16_bwd(const void *gy, const void *idx, int B, int H, int W, int C, int k, int stride, void *dx, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && C >= 1, "mmc_maxpool_nhwc_bf16_bwd: bad shape");
    MMC_CHECK_ARG(k >= 1 && k <= 15 && stride >= 1 && H >= k && W >= k, "mmc_maxpool_nhwc_bf16_bwd: window %d does not fit the %dx%d map (k <= 15)", k, H, W);
    const int Ho = (H - k) / stride + 1, Wo = (W - k) / stride + 1;
    const int64_t n = (int64_t)B * H * W * C;
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(gy && idx && dx, "mmc_maxpool_nhwc_bf16_bwd: NULL buffer");
    maxpool_bwd_kernel<<<elementwise_grid(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)gy, (const uint8_t *)idx, H, W, C, k, stride, Ho, Wo,
                                                                                  n, (__nv_bfloat16 *)dx);
    MMC_CHECK_LAUNCH("mmc_maxpool_nhwc_bf16_bwd");
    return MMC_OK;
}

int mmc_upsample_bilinear_bwd_bf16(const void *g, int B, int H, int W, int C, int hs, int ws, void *dsmall, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && hs >= 1 && ws >= 1 && H >= 1 && W >= 1 && C >= 8, "mmc_upsample_bilinear_bwd_bf16: bad shape");
    MMC_UNSUPPORTED(C % 8 != 0, "mmc_upsample_bilinear_bwd_bf16: C must be a multiple of 8 (got %d)", C);
    const int64_t n8 = (int64_t)B * hs * ws * (C / 8);
    if (n8 == 0) return MMC_OK;
    MMC_CHECK_ARG(g && dsmall && aligned16(g) && aligned16(dsmall), "mmc_upsample_bilinear_bwd_bf16: NULL or unaligned buffer");
    upsample_bwd_v8_kernel<<<elementwise_grid(n8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4 *)g, hs, ws, H, W, C / 8, (float)hs / (float)H,
                                                                                      (float)ws / (float)W, n8, (uint4 *)dsmall);
    MMC_CHECK_LAUNCH("mmc_upsample_bilinear_bwd_bf16");
    return MMC_OK;
}

int mmc_sigmoid_gate_bwd_bf16(const void *g, const void *x, const void *gate, int64_t n, void *dx, void *dgate, void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_sigmoid_gate_bwd_bf16: n < 0");
    MMC_UNSUPPORTED(n % 8 != 0, "mmc_sigmoid_gate_bwd_bf16: n must be a multiple of 8");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(g && x && gate && dx && dgate && aligned16(g) && aligned16(x) && aligned16(gate) && aligned16(dx) && aligned16(dgate),
                  "mmc_sigmoid_gate_bwd_bf16: NULL or unaligned buffer");
    sigmoid_gate_bwd_kernel<<<elementwise_grid(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4 *)g, (const uint4 *)x, (const uint4 *)gate, n / 8,
                                                                                          (uint4 *)dx, (uint4 *)dgate);
    MMC_CHECK_LAUNCH("mmc_sigmoid_gate_bwd_bf16");
    return MMC_OK;
}

int mmc_gelu_bwd_bf16(const void *g, const void *x, int64_t n, void *dx, void *stream)
{
    MMC_CHECK_ARG(n >= 0 && n % 2 == 0, "mmc_gelu_bwd_bf16: n must be even");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(g && x && dx, "mmc_gelu_bwd_bf16: NULL buffer");
    gelu_bwd_kernel<<<elementwise_grid(n / 2, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat162 *)g, (const __nv_bfloat162 *)x, n / 2,
                                                                                  (__nv_bfloat162 *)dx);
    MMC_CHECK_LAUNCH("mmc_gelu_bwd_bf16");
    return MMC_OK;
}

int mmc_layernorm_bwd_bf16(const void *g, const void *v, const void *g_sum, const float *weight, int64_t rows, int C, float eps, void *dv,
                           float *dweight, float *dbias, void *stream)
{
    MMC_CHECK_ARG(rows >= 0 && C >= 1, "mmc_layernorm_bwd_bf16: bad shape");
    MMC_UNSUPPORTED(C > 256, "mmc_layernorm_bwd_bf16: C <= 256 (got %d)", C);
    if (rows == 0) return MMC_OK;
    MMC_CHECK_ARG(g && v && weight && dv && dweight && dbias, "mmc_layernorm_bwd_bf16: NULL buffer");
    // few CTAs: every CTA ends with 2 C atomics; a B200 needs ~4 resident CTAs per SM to cover the load latency of this pass
    int64_t ctas = (rows + 7) / 8;
    if (ctas > 4 * kNumSMs) ctas = 4 * kNumSMs;
    layernorm_bwd_kernel<<<(int)ctas, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)g, (const __nv_bfloat16 *)v, (const __nv_bfloat16 *)g_sum, weight,
                                                                    rows, C, eps, (__nv_bfloat16 *)dv, dweight, dbias);
    MMC_CHECK_LAUNCH("mmc_layernorm_bwd_bf16");
    return MMC_OK;
}

int mmc_window_attention_bwd(const void *q, const void *kv, const float *bias_table, const void *dout, int B, int H, int W, int heads, int head_dim,
                             int window, int shift, float scale, void *dq, void *dkv, float *dtable, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && heads >= 1 && window >= 1, "mmc_window_attention_bwd: bad shape");
    MMC_UNSUPPORTED(head_dim != kHeadDimB || window * window > kMaxWinB || heads > 8,
                    "mmc_window_attention_bwd: supports head_dim 32, windows of <= 16 tokens, <= 8 heads (got %d, %d, %d)", head_dim, window * window, heads);
    MMC_CHECK_ARG(H % window == 0 && W % window == 0, "mmc_window_attention_bwd: the token grid must be a multiple of the window");
    MMC_CHECK_ARG(shift >= 0 && shift < window, "mmc_window_attention_bwd: shift must be in [0, window)");
    const int64_t items = (int64_t)B * (H / window) * (W / window) * heads;
    if (items == 0) return MMC_OK;
    MMC_CHECK_ARG(q && kv && bias_table && dout && dq && dkv && dtable, "mmc_window_attention_bwd: NULL buffer");
    int64_t ctas = (items + 3) / 4;
    if (ctas > 8 * kNumSMs) ctas = 8 * kNumSMs;
    window_attention_bwd_kernel<<<(int)ctas, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)q, (const __nv_bfloat16 *)kv, bias_table,
                                                                           (const __nv_bfloat16 *)dout, B, H, W, heads, window, shift, scale,
                                                                           (__nv_bfloat16 *)dq, (__nv_bfloat16 *)dkv, dtable);
    MMC_CHECK_LAUNCH("mmc_window_attention_bwd");
    return MMC_OK;
}

}  // extern "C"
