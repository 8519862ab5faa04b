// layout.cu -- layout / dtype conversion at the API edges (NCHW fp32 <-> NHWC bf16/fp32).
// The reference keeps everything NCHW fp32 (torch default); the transform kernels stream NHWC bf16.
#include "common.cuh"

namespace mmc {

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
