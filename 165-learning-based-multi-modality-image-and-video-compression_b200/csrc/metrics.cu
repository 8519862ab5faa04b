// metrics.cu -- per-image rate / distortion of one forward pass, reduced on the device: what the reference's evaluation loop keeps
// of a forward (compressai/utils/eval_model/__main__t.py:151-173: bpp = sum(log(likelihoods)) / (-ln 2 * pixels), psnr from
// F.mse_loss(x, x_hat)).  With these the host only reads two floats per image instead of the reconstruction and the likelihood
// tensors.  Both kernels are single HBM passes: grid = (chunks, images), one atomicAdd per CTA into the image's accumulator.
#include "common.cuh"

namespace mmc {

__device__ __forceinline__ float block_sum_256(float v, float *red)
{
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
    if (threadIdx.x < 8) t = red[threadIdx.x];
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;   // valid in thread 0
}

__global__ void __launch_bounds__(256) image_bits_kernel(const float *__restrict__ lik, int64_t n, float scale, float *__restrict__ out)
{
    __shared__ float red[8];
    const float *p = lik + (int64_t)blockIdx.y * n;
    float acc = 0.0f;
    const int64_t n4 = n / 4, step = (int64_t)gridDim.x * blockDim.x;
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
        const float4 *p4 = reinterpret_cast<const float4 *>(p);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
            const float4 v = p4[i];
            acc += (__log2f(v.x) + __log2f(v.y)) + (__log2f(v.z) + __log2f(v.w));
        }
        for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) acc += __log2f(p[i]);
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) acc += __log2f(p[i]);
    }
    const float t = block_sum_256(acc, red);
    if (threadIdx.x == 0) atomicAdd(out + blockIdx.y, -t * scale);
}

__global__ void __launch_bounds__(256) image_sse_kernel(const float *__restrict__ a, const float *__restrict__ b, int64_t n, float scale,
                                                        float *__restrict__ out)
{
    __shared__ float red[8];
    const float *pa = a + (int64_t)blockIdx.y * n, *pb = b + (int64_t)blockIdx.y * n;
    float acc = 0.0f;
    const int64_t n4 = n / 4, step = (int64_t)gridDim.x * blockDim.x;
    if (((reinterpret_cast<uintptr_t>(pa) | reinterpret_cast<uintptr_t>(pb)) & 15u) == 0) {
        const float4 *a4 = reinterpret_cast<const float4 *>(pa), *b4 = reinterpret_cast<const float4 *>(pb);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
            const float4 u = a4[i], v = b4[i];
            const float d0 = u.x - v.x, d1 = u.y - v.y, d2 = u.z - v.z, d3 = u.w - v.w;
            acc += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
        for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) { const float d = pa[i] - pb[i]; acc += d * d; }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) { const float d = pa[i] - pb[i]; acc += d * d; }
    }
    const float t = block_sum_256(acc, red);
    if (threadIdx.x == 0) atomicAdd(out + blockIdx.y, t * scale);
}

// 8-bit image samples -> fp32 in [0, 1]: torchvision's ToTensor (`img.to(float32).div(255)`, the reference's loader,
// eval_model/__main__t.py:94-101) moved behind the host->device copy, so that the copy carries one byte per sample instead of four.
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t *__restrict__ x, int64_t n, float *__restrict__ y)
{
    const int64_t n4 = n / 4, step = (int64_t)gridDim.x * blockDim.x;
    const uchar4 *x4 = reinterpret_cast<const uchar4 *>(x);
    float4 *y4 = reinterpret_cast<float4 *>(y);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
        const uchar4 v = x4[i];
        y4[i] = make_float4(__fdiv_rn((float)v.x, 255.0f), __fdiv_rn((float)v.y, 255.0f), __fdiv_rn((float)v.z, 255.0f), __fdiv_rn((float)v.w, 255.0f));
    }
    for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) y[i] = __fdiv_rn((float)x[i], 255.0f);
}

static int chunks_for(int B, int64_t n)
{
    int64_t want = (n / 4 + 255) / 256 / 8;                 // ~8 float4 per thread
    int64_t cap = ((int64_t)kNumSMs * 8 + B - 1) / B;       // fill the machine across the batch
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_image_bits(const float *likelihood, int B, int64_t n_per_image, float scale, float *out, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && B <= 65535 && n_per_image >= 0, "mmc_image_bits: bad shape");
    if (B == 0 || n_per_image == 0) return MMC_OK;
    MMC_CHECK_ARG(likelihood && out, "mmc_image_bits: NULL buffer");
    image_bits_kernel<<<dim3(chunks_for(B, n_per_image), B), 256, 0, (cudaStream_t)stream>>>(likelihood, n_per_image, scale, out);
    MMC_CHECK_LAUNCH("mmc_image_bits");
    return MMC_OK;
}

int mmc_image_sse(const float *a, const float *b, int B, int64_t n_per_image, float scale, float *out, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && B <= 65535 && n_per_image >= 0, "mmc_image_sse: bad shape");
    if (B == 0 || n_per_image == 0) return MMC_OK;
    MMC_CHECK_ARG(a && b && out, "mmc_image_sse: NULL buffer");
    image_sse_kernel<<<dim3(chunks_for(B, n_per_image), B), 256, 0, (cudaStream_t)stream>>>(a, b, n_per_image, scale, out);
    MMC_CHECK_LAUNCH("mmc_image_sse");
    return MMC_OK;
}

int mmc_u8_to_f32(const void *x, int64_t n, float *y, void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_u8_to_f32: n < 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && y && (reinterpret_cast<uintptr_t>(x) & 3u) == 0 && aligned16(y), "mmc_u8_to_f32: NULL or unaligned buffer");
    u8_to_f32_kernel<<<elementwise_grid(n / 4 > 0 ? n / 4 : 1, 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t *)x, n, y);
    MMC_CHECK_LAUNCH("mmc_u8_to_f32");
    return MMC_OK;
}

}  // extern "C"
