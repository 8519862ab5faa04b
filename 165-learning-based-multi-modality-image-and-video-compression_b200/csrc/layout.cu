// layout.cu -- layout / dtype conversion at the API edges (NCHW fp32 <-> NHWC bf16/fp32).
// The reference keeps everything NCHW fp32 (torch default); the transform kernels stream NHWC bf16.
#include "common.cuh"

namespace mmc {

// y = x * rsqrt(norm) (GDN) or x * sqrt(norm) (IGDN) on fp32 maps: the elementwise half of the fp32-mode GDN, whose norm
// contraction runs on the tensor cores (three-term split, see split_bf16x3_kernel)
__global__ void __launch_bounds__(256) gdn_apply_kernel(const float *__restrict__ x, const float *__restrict__ norm, int inverse, int64_t n4,
                                                        float *__restrict__ y)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 a = ldg_stream(reinterpret_cast<const float4 *>(x) + i);
        const float4 b = ldg_stream(reinterpret_cast<const float4 *>(norm) + i);
        float4 o;
        if (inverse) o = make_float4(a.x * sqrtf(b.x), a.y * sqrtf(b.y), a.z * sqrtf(b.z), a.w * sqrtf(b.w));
        else o = make_float4(a.x * rsqrtf(b.x), a.y * rsqrtf(b.y), a.z * rsqrtf(b.z), a.w * rsqrtf(b.w));
        reinterpret_cast<float4 *>(y)[i] = o;
    }
}

// [B][C][HW] -> [B][HW][C] through a 32x33 shared tile; both sides coalesced.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) transpose_c_hw_kernel(const TIn *__restrict__ x, int C, int64_t HW, bool to_nhwc,
                                                            TOut *__restrict__ y)
{
    __shared__ float tile[32][33];
    const int64_t b = blockIdx.z;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;  // 32 x 8
    const TIn *xb = x + b * C * HW;
    TOut *yb = y + b * C * HW;
    if (to_nhwc) {
        for (int r = ty; r < 32; r += 8) {   // rows = channels, cols = pixels
            int c = c0 + r; int64_t p = p0 + tx;
            tile[r][tx] = (c < C && p < HW) ? (float)xb[(int64_t)c * HW + p] : 0.0f;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {   // rows = pixels, cols = channels
            int64_t p = p0 + r; int c = c0 + tx;
            if (p < HW && c < C) yb[p * C + c] = (TOut)tile[tx][r];
        }
    } else {
        for (int r = ty; r < 32; r += 8) {   // rows = pixels, cols = channels
            int64_t p = p0 + r; int c = c0 + tx;
            tile[r][tx] = (p < HW && c < C) ? (float)xb[p * C + c] : 0.0f;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {   // rows = channels, cols = pixels
            int c = c0 + r; int64_t p = p0 + tx;
            if (c < C && p < HW) yb[(int64_t)c * HW + p] = (TOut)tile[tx][r];
        }
    }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float *__restrict__ x, int64_t n, __nv_bfloat16 *__restrict__ y)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = __float2bfloat16_rn(x[i]);
}

// fp32 precision mode (mmcodec.precision("fp32")): x = hi + lo + O(2^-17 x) with hi = bf16(x), lo = bf16(x - hi).  One pixel row of C
// fp32 channels becomes 3 C bf16 channels [hi | lo | hi]; against weights laid out [w_hi | w_hi | w_lo] along the input channels
// the ordinary bf16 tensor-core convolution then accumulates x_hi w_hi + x_lo w_hi + x_hi w_lo in fp32: every term of the product
// down to 2^-16 relative (~1e-5 on a layer output, against 4e-3 for plain bf16 operands and ~4e-4 for a single TF32 pass).
template <bool kSquare>
__global__ void __launch_bounds__(256) split_bf16x3_kernel(const float *__restrict__ x, int64_t pixels, int C, __nv_bfloat16 *__restrict__ y)
{
    const int cq = C >> 2;                                  // float4 groups per pixel
    const int64_t n = pixels * cq;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t p = i / cq;
        const int q = (int)(i - p * cq);
        const float4 v = ldg_stream(reinterpret_cast<const float4 *>(x) + i);
        // kSquare: the operand of the GDN norm contraction, x^2 (fp32 mode of layers/gdn.py:77-92)
        const float f[4] = {kSquare ? v.x * v.x : v.x, kSquare ? v.y * v.y : v.y, kSquare ? v.z * v.z : v.z, kSquare ? v.w * v.w : v.w};
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            hi[k] = __float2bfloat16_rn(f[k]);
            lo[k] = __float2bfloat16_rn(f[k] - __bfloat162float(hi[k]));
        }
        __nv_bfloat16 *row = y + p * 3 * C + q * 4;
        *reinterpret_cast<uint2 *>(row) = *reinterpret_cast<const uint2 *>(hi);
        *reinterpret_cast<uint2 *>(row + C) = *reinterpret_cast<const uint2 *>(lo);
        *reinterpret_cast<uint2 *>(row + 2 * C) = *reinterpret_cast<const uint2 *>(hi);
    }
}

template <typename TIn, typename TOut>
static int launch_transpose(const TIn *x, int64_t B, int C, int64_t HW, bool to_nhwc, TOut *y, cudaStream_t st, const char *name)
{
    MMC_CHECK_ARG(B >= 0 && C >= 1 && HW >= 0 && B <= 65535, "%s: bad shape", name);
    if (B * HW == 0) return MMC_OK;
    MMC_CHECK_ARG(x && y, "%s: NULL buffer", name);
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B);
    transpose_c_hw_kernel<TIn, TOut><<<grid, 256, 0, st>>>(x, C, HW, to_nhwc, y);
    MMC_CHECK_LAUNCH(name);
    return MMC_OK;
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_nchw_f32_to_nhwc_bf16(const float *x, int64_t B, int C, int64_t HW, void *y, void *stream)
{
    return launch_transpose<float, __nv_bfloat16>(x, B, C, HW, true, (__nv_bfloat16 *)y, (cudaStream_t)stream, "mmc_nchw_f32_to_nhwc_bf16");
}

int mmc_nhwc_bf16_to_nchw_f32(const void *x, int64_t B, int C, int64_t HW, float *y, void *stream)
{
    return launch_transpose<__nv_bfloat16, float>((const __nv_bfloat16 *)x, B, C, HW, false, y, (cudaStream_t)stream, "mmc_nhwc_bf16_to_nchw_f32");
}

int mmc_nhwc_f32_to_nchw_f32(const float *x, int64_t B, int C, int64_t HW, float *y, void *stream)
{
    return launch_transpose<float, float>(x, B, C, HW, false, y, (cudaStream_t)stream, "mmc_nhwc_f32_to_nchw_f32");
}

int mmc_split_f32_bf16x3(const float *x, int64_t pixels, int C, int square, void *y, void *stream)
{
    MMC_CHECK_ARG(pixels >= 0 && C >= 4 && C % 4 == 0, "mmc_split_f32_bf16x3: C must be a positive multiple of 4");
    if (pixels == 0) return MMC_OK;
    MMC_CHECK_ARG(x && y && aligned16(x) && (reinterpret_cast<uintptr_t>(y) & 7u) == 0, "mmc_split_f32_bf16x3: NULL or unaligned buffer");
    const int grid = elementwise_grid(pixels * (C >> 2), 256);
    if (square) split_bf16x3_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, pixels, C, (__nv_bfloat16 *)y);
    else split_bf16x3_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, pixels, C, (__nv_bfloat16 *)y);
    MMC_CHECK_LAUNCH("mmc_split_f32_bf16x3");
    return MMC_OK;
}

int mmc_gdn_apply_f32(const float *x, const float *norm, int inverse, int64_t n, float *y, void *stream)
{
    MMC_CHECK_ARG(n >= 0 && n % 4 == 0, "mmc_gdn_apply_f32: n must be a non-negative multiple of 4");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && norm && y && aligned16(x) && aligned16(norm) && aligned16(y), "mmc_gdn_apply_f32: NULL or unaligned buffer");
    gdn_apply_kernel<<<elementwise_grid(n >> 2, 256), 256, 0, (cudaStream_t)stream>>>(x, norm, inverse, n >> 2, y);
    MMC_CHECK_LAUNCH("mmc_gdn_apply_f32");
    return MMC_OK;
}

int mmc_f32_to_bf16(const float *x, int64_t n, void *y, void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_f32_to_bf16: n < 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && y, "mmc_f32_to_bf16: NULL buffer");
    f32_to_bf16_kernel<<<elementwise_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, (__nv_bfloat16 *)y);
    MMC_CHECK_LAUNCH("mmc_f32_to_bf16");
    return MMC_OK;
}

}  // extern "C"
