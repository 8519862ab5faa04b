// esa.cu -- the non-convolution steps of the enhanced-spatial-attention gate (ESA.forward, compressai/models/google.py:1445-1459)
// on NHWC bf16 maps: the 7x7 / stride-3 max-pool, the bilinear upsampling fused with the `c3 + cf` add, and the sigmoid gate
// `x * sigmoid(c4)`.  All three are HBM-bound elementwise / small-window passes; each replaces two or three library launches and
// the intermediate tensors between them.
#include "common.cuh"

namespace mmc {

__global__ void __launch_bounds__(256) maxpool_nhwc_kernel(const __nv_bfloat162 *__restrict__ x, int H, int W, int C2, int k, int stride,
                                                           int Ho, int Wo, int64_t n, __nv_bfloat162 *__restrict__ y)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C2);
        int64_t t = i / C2;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho);
        const int64_t b = t / Ho;
        float m0 = -INFINITY, m1 = -INFINITY;
        bool nan0 = false, nan1 = false;
        for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx) {
                const float2 v = __bfloat1622float2(x[((b * H + oy * stride + ky) * W + ox * stride + kx) * C2 + c]);
                nan0 |= (v.x != v.x); nan1 |= (v.y != v.y);      // max_pool2d propagates NaN
                m0 = fmaxf(m0, v.x); m1 = fmaxf(m1, v.y);
            }
        y[i] = __floats2bfloat162_rn(nan0 ? NAN : m0, nan1 ? NAN : m1);
    }
}

// F.interpolate(mode="bilinear", align_corners=False): src = max(0, (dst + 0.5) * in / out - 0.5), neighbours clamped at the edge
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in, int *i0, int *i1, float *l1)
{
    float s = ((float)dst + 0.5f) * scale - 0.5f;
    s = s < 0.0f ? 0.0f : s;
    const int a = (int)s;
    *i0 = a;
    *i1 = a + (a < in - 1 ? 1 : 0);
    *l1 = s - (float)a;
}

__global__ void __launch_bounds__(256) upsample_add_kernel(const __nv_bfloat162 *__restrict__ small, const __nv_bfloat162 *__restrict__ add,
                                                           int hs, int ws, int H, int W, int C2, float sy, float sx, int64_t n,
                                                           __nv_bfloat162 *__restrict__ y)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C2);
        int64_t t = i / C2;
        const int ox = (int)(t % W); t /= W;
        const int oy = (int)(t % H);
        const int64_t b = t / H;
        int y0, y1, x0, x1;
        float ly, lx;
        bilinear_src(oy, sy, hs, &y0, &y1, &ly);
        bilinear_src(ox, sx, ws, &x0, &x1, &lx);
        const __nv_bfloat162 *base = small + b * hs * ws * C2 + c;
        const float2 v00 = __bfloat1622float2(base[((int64_t)y0 * ws + x0) * C2]), v01 = __bfloat1622float2(base[((int64_t)y0 * ws + x1) * C2]);
        const float2 v10 = __bfloat1622float2(base[((int64_t)y1 * ws + x0) * C2]), v11 = __bfloat1622float2(base[((int64_t)y1 * ws + x1) * C2]);
        const float2 a = __bfloat1622float2(add[i]);
        const float hy = 1.0f - ly, hx = 1.0f - lx;
        y[i] = __floats2bfloat162_rn(hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x) + a.x,
                                     hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y) + a.y);
    }
}

// 8 channels (one 16-byte vector) per thread: used whenever C % 8 == 0
__global__ void __launch_bounds__(256) upsample_add_v8_kernel(const uint4 *__restrict__ small, const uint4 *__restrict__ add, int hs, int ws, int H, int W,
                                                              int C8, float sy, float sx, int64_t n, uint4 *__restrict__ y)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8);
        int64_t t = i / C8;
        const int ox = (int)(t % W); t /= W;
        const int oy = (int)(t % H);
        const int64_t b = t / H;
        int y0, y1, x0, x1;
        float ly, lx;
        bilinear_src(oy, sy, hs, &y0, &y1, &ly);
        bilinear_src(ox, sx, ws, &x0, &x1, &lx);
        const uint4 *base = small + b * hs * ws * C8 + c;
        const uint4 q00 = base[((int64_t)y0 * ws + x0) * C8], q01 = base[((int64_t)y0 * ws + x1) * C8];
        const uint4 q10 = base[((int64_t)y1 * ws + x0) * C8], q11 = base[((int64_t)y1 * ws + x1) * C8];
        const uint4 qa = add[i];
        uint4 o;
        const __nv_bfloat162 *p00 = reinterpret_cast<const __nv_bfloat162 *>(&q00), *p01 = reinterpret_cast<const __nv_bfloat162 *>(&q01);
        const __nv_bfloat162 *p10 = reinterpret_cast<const __nv_bfloat162 *>(&q10), *p11 = reinterpret_cast<const __nv_bfloat162 *>(&q11);
        const __nv_bfloat162 *pa = reinterpret_cast<const __nv_bfloat162 *>(&qa);
        __nv_bfloat162 *po = reinterpret_cast<__nv_bfloat162 *>(&o);
        const float hy = 1.0f - ly, hx = 1.0f - lx;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 v00 = __bfloat1622float2(p00[j]), v01 = __bfloat1622float2(p01[j]), v10 = __bfloat1622float2(p10[j]), v11 = __bfloat1622float2(p11[j]);
            const float2 a = __bfloat1622float2(pa[j]);
            po[j] = __floats2bfloat162_rn(hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x) + a.x,
                                          hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y) + a.y);
        }
        y[i] = o;
    }
}

__global__ void __launch_bounds__(256) sigmoid_gate_kernel(const uint4 *__restrict__ x, const uint4 *__restrict__ g, int64_t n8, uint4 *__restrict__ y)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 xv = x[i], gv = g[i], o;
        const __nv_bfloat162 *xp = reinterpret_cast<const __nv_bfloat162 *>(&xv), *gp = reinterpret_cast<const __nv_bfloat162 *>(&gv);
        __nv_bfloat162 *op = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 a = __bfloat1622float2(xp[j]), s = __bfloat1622float2(gp[j]);
            op[j] = __floats2bfloat162_rn(a.x / (1.0f + __expf(-s.x)), a.y / (1.0f + __expf(-s.y)));
        }
        y[i] = o;
    }
}

__global__ void sigmoid_gate_tail_kernel(const __nv_bfloat16 *__restrict__ x, const __nv_bfloat16 *__restrict__ g, int n, __nv_bfloat16 *__restrict__ y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __float2bfloat16(__bfloat162float(x[i]) / (1.0f + __expf(-__bfloat162float(g[i]))));
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_maxpool_nhwc_bf16(const void *x, int B, int H, int W, int C, int k, int stride, void *y, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && C >= 2 && C % 2 == 0, "mmc_maxpool_nhwc_bf16: bad shape (C must be even)");
    MMC_CHECK_ARG(k >= 1 && stride >= 1 && H >= k && W >= k, "mmc_maxpool_nhwc_bf16: window %d does not fit the %dx%d map", k, H, W);
    const int Ho = (H - k) / stride + 1, Wo = (W - k) / stride + 1;
    const int64_t n = (int64_t)B * Ho * Wo * (C / 2);
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && y, "mmc_maxpool_nhwc_bf16: NULL buffer");
    maxpool_nhwc_kernel<<<elementwise_grid(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat162 *)x, H, W, C / 2, k, stride, Ho, Wo, n,
                                                                                   (__nv_bfloat162 *)y);
    MMC_CHECK_LAUNCH("mmc_maxpool_nhwc_bf16");
    return MMC_OK;
}

int mmc_upsample_bilinear_add_bf16(const void *small, int B, int hs, int ws, int C, const void *add, int H, int W, void *y, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && hs >= 1 && ws >= 1 && H >= 1 && W >= 1 && C >= 2 && C % 2 == 0, "mmc_upsample_bilinear_add_bf16: bad shape (C must be even)");
    const int64_t n = (int64_t)B * H * W * (C / 2);
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(small && add && y, "mmc_upsample_bilinear_add_bf16: NULL buffer");
    if (C % 8 == 0 && aligned16(small) && aligned16(add) && aligned16(y)) {
        const int64_t n8 = (int64_t)B * H * W * (C / 8);
        upsample_add_v8_kernel<<<elementwise_grid(n8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4 *)small, (const uint4 *)add, hs, ws, H, W, C / 8,
                                                                                          (float)hs / (float)H, (float)ws / (float)W, n8, (uint4 *)y);
        MMC_CHECK_LAUNCH("mmc_upsample_bilinear_add_bf16");
        return MMC_OK;
    }
    upsample_add_kernel<<<elementwise_grid(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat162 *)small, (const __nv_bfloat162 *)add, hs, ws, H, W,
                                                                                   C / 2, (float)hs / (float)H, (float)ws / (float)W, n, (__nv_bfloat162 *)y);
    MMC_CHECK_LAUNCH("mmc_upsample_bilinear_add_bf16");
    return MMC_OK;
}

int mmc_sigmoid_gate_bf16(const void *x, const void *gate, int64_t n, void *y, void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_sigmoid_gate_bf16: n < 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && gate && y && aligned16(x) && aligned16(gate) && aligned16(y), "mmc_sigmoid_gate_bf16: NULL or unaligned buffer");
    const int64_t n8 = n / 8;
    if (n8 > 0) {
        sigmoid_gate_kernel<<<elementwise_grid(n8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4 *)x, (const uint4 *)gate, n8, (uint4 *)y);
        MMC_CHECK_LAUNCH("mmc_sigmoid_gate_bf16");
    }
    const int tail = (int)(n - 8 * n8);
    if (tail > 0) {
        sigmoid_gate_tail_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x + 8 * n8, (const __nv_bfloat16 *)gate + 8 * n8, tail,
                                                                    (__nv_bfloat16 *)y + 8 * n8);
        MMC_CHECK_LAUNCH("mmc_sigmoid_gate_bf16");
    }
    return MMC_OK;
}

}  // extern "C"
