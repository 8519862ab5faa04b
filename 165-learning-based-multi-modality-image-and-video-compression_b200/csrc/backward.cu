// backward.cu -- elementwise / reduction kernels of the training step (HBM-bound), next to the tensor-core input- and
// weight-gradient kernels (conv_tc.cu with the adjoint descriptor, wgrad_tc.cu).
//
// Backward of (file:line relative to /root/reference/CompressAI; formulas: SURVEY.md Appendix E, checked in tests/ against float64
// autograd of the same ops):
//   ReLU / LeakyReLU after conv()/deconv()        compressai/models/google.py:254-269,363-377
//   bias of nn.Conv2d / nn.ConvTranspose2d        compressai/models/utils.py:128-146
//   GDN / IGDN                                    compressai/layers/gdn.py:77-92
//   NonNegativeParametrizer + LowerBound          compressai/ops/parametrizers.py:61-64, ops/bound_ops.py:40-42
//   GaussianConditional._likelihood / forward     compressai/entropy_models/entropy_models.py:692-731
// Activations and gradients are NHWC bf16 (the transform kernels' format); parameter gradients are fp32.
#include "common.cuh"

namespace mmc {

constexpr int kBwBlock = 256;

__device__ __forceinline__ void unpack8(const uint4 &v, float *f)
{
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float *f)
{
    uint4 v;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
}

// g_in = g_out * act'(y): ReLU -> [y > 0], LeakyReLU(0.01) -> (y > 0 ? 1 : 0.01).  8 bf16 per thread per step.
__global__ void __launch_bounds__(kBwBlock) act_bwd_kernel(const uint4 *__restrict__ g, const uint4 *__restrict__ y, int act, int64_t n8,
                                                          uint4 *__restrict__ out)
{
    const float neg = (act == MMC_ACT_LEAKY_RELU) ? 0.01f : 0.0f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        float gv[8], yv[8];
        unpack8(g[i], gv);
        unpack8(y[i], yv);
#pragma unroll
        for (int k = 0; k < 8; ++k) gv[k] = yv[k] > 0.0f ? gv[k] : gv[k] * neg;
        out[i] = pack8(gv);
    }
}

// out[c] (+)= scale * sum over rows of g[row][c]   (bias gradient; GDN beta gradient).  g is [rows][C] bf16, C % 8 == 0.
// Each thread owns 8 adjacent channels and walks rows; partial sums go through shared memory, one atomic per CTA column.
__global__ void __launch_bounds__(kBwBlock) colsum_kernel(const uint4 *__restrict__ g, int64_t rows, int C8, float scale, float *__restrict__ out)
{
    extern __shared__ float part[];          // [rows_per_block_threads][C8 * 8] folded below
    const int tc = threadIdx.x % C8, tr = threadIdx.x / C8, rpb = kBwBlock / C8;   // C8 <= 256
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (tr < rpb) {
        const int64_t step = (int64_t)gridDim.x * rpb;
        int64_t r = (int64_t)blockIdx.x * rpb + tr;
        for (; r + 3 * step < rows; r += 4 * step) {      // four independent 16-byte loads in flight per thread
            const uint4 q0 = g[r * C8 + tc], q1 = g[(r + step) * C8 + tc], q2 = g[(r + 2 * step) * C8 + tc], q3 = g[(r + 3 * step) * C8 + tc];
            float v0[8], v1[8], v2[8], v3[8];
            unpack8(q0, v0); unpack8(q1, v1); unpack8(q2, v2); unpack8(q3, v3);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += (v0[k] + v1[k]) + (v2[k] + v3[k]);
        }
        for (; r < rows; r += step) {
            float v[8];
            unpack8(g[r * C8 + tc], v);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += v[k];
        }
    }
    float *mine = part + (size_t)threadIdx.x * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) mine[k] = acc[k];
    __syncthreads();
    if (tr == 0) {
        for (int j = 1; j < rpb; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += part[(size_t)(j * C8 + tc) * 8 + k];
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(out + tc * 8 + k, scale * acc[k]);
    }
}

__global__ void __launch_bounds__(kBwBlock) square_kernel(const uint4 *__restrict__ x, int64_t n8, uint4 *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        float v[8];
        unpack8(x[i], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] *= v[k];
        out[i] = pack8(v);
    }
}

// GDN backward, stage 1:  t = g * x * n^(-3/2)  (IGDN: g * x * n^(-1/2)),  n = beta' + gamma' x^2 (fp32, from the 1x1 conv)
__global__ void __launch_bounds__(kBwBlock) gdn_bwd_t_kernel(const uint4 *__restrict__ g, const uint4 *__restrict__ x, const float4 *__restrict__ n,
                                                            int inverse, int64_t n8, uint4 *__restrict__ t)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        float gv[8], xv[8], nv[8];
        unpack8(g[i], gv);
        unpack8(x[i], xv);
        *reinterpret_cast<float4 *>(nv) = n[2 * i];
        *reinterpret_cast<float4 *>(nv + 4) = n[2 * i + 1];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float r = rsqrtf(nv[k]);
            gv[k] = gv[k] * xv[k] * (inverse ? r : r * r * r);
        }
        t[i] = pack8(gv);
    }
}

// GDN backward, stage 2:  dx = g * n^(-1/2) - x * u   (IGDN: g * n^(1/2) + x * u),  u = gamma'^T t (fp32, from the 1x1 conv)
__global__ void __launch_bounds__(kBwBlock) gdn_bwd_dx_kernel(const uint4 *__restrict__ g, const uint4 *__restrict__ x, const float4 *__restrict__ n,
                                                             const float4 *__restrict__ u, int inverse, int64_t n8, uint4 *__restrict__ dx)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        float gv[8], xv[8], nv[8], uv[8];
        unpack8(g[i], gv);
        unpack8(x[i], xv);
        *reinterpret_cast<float4 *>(nv) = n[2 * i];
        *reinterpret_cast<float4 *>(nv + 4) = n[2 * i + 1];
        *reinterpret_cast<float4 *>(uv) = u[2 * i];
        *reinterpret_cast<float4 *>(uv + 4) = u[2 * i + 1];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float r = rsqrtf(nv[k]);
            gv[k] = inverse ? fmaf(xv[k], uv[k], gv[k] * (nv[k] * r)) : fmaf(-xv[k], uv[k], gv[k] * r);
        }
        dx[i] = pack8(gv);
    }
}

// p' = max(p, b)^2 - pedestal  =>  d = dp' * 2 max(p, b);  dp = d * [(p >= b) | (d < 0)]   (parametrizers.py:61-64, bound_ops.py:40-42)
__global__ void __launch_bounds__(kBwBlock) reparam_bwd_kernel(const float *__restrict__ p, const float *__restrict__ dp_eff, float bound, int64_t n,
                                                              float *__restrict__ dp)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float pv = p[i];
        const float d = dp_eff[i] * 2.0f * fmaxf(pv, bound);
        dp[i] = (pv >= bound || d < 0.0f) ? d : 0.0f;
    }
}

// |x| (fp32) -> bf16: the h_a input of ScaleHyperprior (models/google.py:283) on the training path; backward g * sign(x)
__global__ void __launch_bounds__(kBwBlock) abs_bf16_kernel(const float *__restrict__ x, int64_t n, __nv_bfloat16 *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __float2bfloat16_rn(fabsf(x[i]));
}
__global__ void __launch_bounds__(kBwBlock) abs_bwd_kernel(const __nv_bfloat16 *__restrict__ g, const float *__restrict__ x, int64_t n, float *__restrict__ dx)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float xv = x[i], gv = __bfloat162float(g[i]);
        dx[i] = xv > 0.0f ? gv : (xv < 0.0f ? -gv : 0.0f);
    }
}

__device__ __forceinline__ float phi_f(float t) { return 0.3989422804014327f * expf(-0.5f * t * t); }
__device__ __forceinline__ float Phi_f(float t) { return 0.5f * erfcf(-0.7071067811865476f * t); }

// GaussianConditional likelihood backward (entropy_models.py:692-731).  g = dL/d(likelihood).
__global__ void __launch_bounds__(kBwBlock) gc_bwd_kernel(const float *__restrict__ x, const float *__restrict__ scales, const float *__restrict__ means,
                                                         const float *__restrict__ noise, const float *__restrict__ g, float scale_bound,
                                                         float lik_bound, int64_t n, float *__restrict__ dx, float *__restrict__ dscales,
                                                         float *__restrict__ dmeans)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float mv = means ? means[i] : 0.0f;
        const float v = noise ? x[i] + noise[i] : rintf(x[i] - mv) + mv;
        const float d = v - mv, a_ = fabsf(d);
        const float sv = scales[i], s = fmaxf(sv, scale_bound), inv_s = 1.0f / s;
        const float a = (0.5f - a_) * inv_s, b = (-0.5f - a_) * inv_s;
        const float p_raw = Phi_f(a) - Phi_f(b);
        const float gv = g[i];
        const float gm = (lik_bound <= 0.0f || p_raw >= lik_bound || gv < 0.0f) ? gv : 0.0f;
        const float pa = phi_f(a), pb = phi_f(b);
        const float sgn = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
        // noise mode: x_hat = x + u, so d(x_hat)/dx = 1; dequantize mode: round() has zero gradient
        const float dxv = noise ? gm * (-(pa - pb) * inv_s) * sgn : 0.0f;
        const float ds = gm * (-(a * pa - b * pb) * inv_s);
        if (dx) dx[i] = dxv;
        if (dmeans) dmeans[i] = -dxv;
        dscales[i] = (sv >= scale_bound || ds < 0.0f) ? ds : 0.0f;
    }
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_act_bwd(const void *grad_out, const void *y, int act, int64_t n, void *grad_in, void *stream)
{
    MMC_CHECK_ARG(n >= 0 && n % 8 == 0, "mmc_act_bwd: n must be a multiple of 8");
    MMC_CHECK_ARG(act == MMC_ACT_RELU || act == MMC_ACT_LEAKY_RELU, "mmc_act_bwd: act must be ReLU or LeakyReLU");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(grad_out && y && grad_in && aligned16(grad_out) && aligned16(y) && aligned16(grad_in), "mmc_act_bwd: NULL or unaligned buffer");
    act_bwd_kernel<<<elementwise_grid(n / 8, kBwBlock), kBwBlock, 0, (cudaStream_t)stream>>>((const uint4 *)grad_out, (const uint4 *)y, act, n / 8,
                                                                                            (uint4 *)grad_in);
    MMC_CHECK_LAUNCH("mmc_act_bwd");
    return MMC_OK;
}

int mmc_colsum_bf16(const void *g, int64_t rows, int C, float scale, float *out, void *stream)
{
    MMC_CHECK_ARG(rows >= 0 && C >= 8 && C % 8 == 0 && C <= 2048, "mmc_colsum_bf16: C must be a multiple of 8 in [8, 2048]");
    if (rows == 0) return MMC_OK;
    MMC_CHECK_ARG(g && out && aligned16(g), "mmc_colsum_bf16: NULL or unaligned buffer");
    const int C8 = C / 8;
    const int rpb = kBwBlock / C8;
    int64_t want = (rows + (int64_t)rpb * 64 - 1) / ((int64_t)rpb * 64);      // ~64 rows per thread
    int grid = (int)(want < 1 ? 1 : (want > kNumSMs * 4 ? kNumSMs * 4 : want));
    colsum_kernel<<<grid, kBwBlock, kBwBlock * 8 * sizeof(float), (cudaStream_t)stream>>>((const uint4 *)g, rows, C8, scale, out);
    MMC_CHECK_LAUNCH("mmc_colsum_bf16");
    return MMC_OK;
}

int mmc_square_bf16(const void *x, int64_t n, void *out, void *stream)
{
    MMC_CHECK_ARG(n >= 0 && n % 8 == 0, "mmc_square_bf16: n must be a multiple of 8");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && out && aligned16(x) && aligned16(out), "mmc_square_bf16: NULL or unaligned buffer");
    square_kernel<<<elementwise_grid(n / 8, kBwBlock), kBwBlock, 0, (cudaStream_t)stream>>>((const uint4 *)x, n / 8, (uint4 *)out);
    MMC_CHECK_LAUNCH("mmc_square_bf16");
    return MMC_OK;
}

int mmc_gdn_bwd_t(const void *grad_out, const void *x, const float *norm, int inverse, int64_t n, void *t, void *stream)
{
    MMC_CHECK_ARG(n >= 0 && n % 8 == 0, "mmc_gdn_bwd_t: n must be a multiple of 8");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(grad_out && x && norm && t && aligned16(grad_out) && aligned16(x) && aligned16(norm) && aligned16(t), "mmc_gdn_bwd_t: NULL or unaligned buffer");
    gdn_bwd_t_kernel<<<elementwise_grid(n / 8, kBwBlock), kBwBlock, 0, (cudaStream_t)stream>>>((const uint4 *)grad_out, (const uint4 *)x,
                                                                                              (const float4 *)norm, inverse, n / 8, (uint4 *)t);
    MMC_CHECK_LAUNCH("mmc_gdn_bwd_t");
    return MMC_OK;
}

int mmc_gdn_bwd_dx(const void *grad_out, const void *x, const float *norm, const float *u, int inverse, int64_t n, void *dx, void *stream)
{
    MMC_CHECK_ARG(n >= 0 && n % 8 == 0, "mmc_gdn_bwd_dx: n must be a multiple of 8");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(grad_out && x && norm && u && dx && aligned16(grad_out) && aligned16(x) && aligned16(norm) && aligned16(u) && aligned16(dx),
                  "mmc_gdn_bwd_dx: NULL or unaligned buffer");
    gdn_bwd_dx_kernel<<<elementwise_grid(n / 8, kBwBlock), kBwBlock, 0, (cudaStream_t)stream>>>((const uint4 *)grad_out, (const uint4 *)x,
                                                                                               (const float4 *)norm, (const float4 *)u, inverse, n / 8,
                                                                                               (uint4 *)dx);
    MMC_CHECK_LAUNCH("mmc_gdn_bwd_dx");
    return MMC_OK;
}

int mmc_reparam_bwd(const float *p, const float *dp_eff, float bound, int64_t n, float *dp, void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_reparam_bwd: n < 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(p && dp_eff && dp, "mmc_reparam_bwd: NULL buffer");
    reparam_bwd_kernel<<<elementwise_grid(n, kBwBlock), kBwBlock, 0, (cudaStream_t)stream>>>(p, dp_eff, bound, n, dp);
    MMC_CHECK_LAUNCH("mmc_reparam_bwd");
    return MMC_OK;
}

int mmc_abs_to_bf16(const float *x, int64_t n, void *out, void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_abs_to_bf16: n < 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && out, "mmc_abs_to_bf16: NULL buffer");
    abs_bf16_kernel<<<elementwise_grid(n, kBwBlock), kBwBlock, 0, (cudaStream_t)stream>>>(x, n, (__nv_bfloat16 *)out);
    MMC_CHECK_LAUNCH("mmc_abs_to_bf16");
    return MMC_OK;
}

int mmc_abs_bwd(const void *grad_out, const float *x, int64_t n, float *dx, void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_abs_bwd: n < 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(grad_out && x && dx, "mmc_abs_bwd: NULL buffer");
    abs_bwd_kernel<<<elementwise_grid(n, kBwBlock), kBwBlock, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)grad_out, x, n, dx);
    MMC_CHECK_LAUNCH("mmc_abs_bwd");
    return MMC_OK;
}

int mmc_gc_backward(const float *x, const float *scales, const float *means, const float *noise, const float *grad_lik, float scale_bound,
                    float likelihood_bound, int64_t n, float *dx, float *dscales, float *dmeans, void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_gc_backward: n < 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && scales && grad_lik && dscales, "mmc_gc_backward: NULL buffer");
    MMC_CHECK_ARG(!dmeans || means, "mmc_gc_backward: dmeans requested without means");
    gc_bwd_kernel<<<elementwise_grid(n, kBwBlock), kBwBlock, 0, (cudaStream_t)stream>>>(x, scales, means, noise, grad_lik, scale_bound, likelihood_bound,
                                                                                       n, dx, dscales, dmeans);
    MMC_CHECK_LAUNCH("mmc_gc_backward");
    return MMC_OK;
}

}  // extern "C"
