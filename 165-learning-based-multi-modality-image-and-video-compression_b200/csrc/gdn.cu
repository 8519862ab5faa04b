// gdn.cu -- GDN / IGDN as a stand-alone op (drop-in for compressai.layers.GDN.forward,
// compressai/layers/gdn.py:77-92) and the beta/gamma re-parametrisation pre-pass
// (compressai/ops/parametrizers.py:61-64 with LowerBound, compressai/ops/bound_ops.py:36-37).
//
// In the transform stacks GDN is fused into the producing convolution's epilogue (conv_tc.cu,
// conv_direct.cu); this file serves user code that instantiates GDN on its own.  The op is a
// C x C contraction over x^2 per pixel: with fp32 I/O its arithmetic intensity is C/4 FLOP/B,
// i.e. HBM-bound on B200 for every C the models use, so a shared-memory tiled fp32 kernel is the
// right shape (no tensor cores needed to reach the memory roofline for C <= 192).
#include "common.cuh"

namespace mmc {

__global__ void gdn_reparam_kernel(const float *__restrict__ beta, const float *__restrict__ gamma, int C,
                                   float beta_bound, float gamma_bound, float pedestal, float *__restrict__ beta_eff,
                                   float *__restrict__ gamma_eff, __nv_bfloat16 *__restrict__ gamma_bf16)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) {
        float b = lower_bound_f(beta[i], beta_bound);
        beta_eff[i] = __fsub_rn(__fmul_rn(b, b), pedestal);
    }
    if (i < C * C) {
        float g = lower_bound_f(gamma[i], gamma_bound);
        float v = __fsub_rn(__fmul_rn(g, g), pedestal);
        if (gamma_eff) gamma_eff[i] = v;
        if (gamma_bf16) gamma_bf16[i] = __float2bfloat16_rn(v);
    }
}

// Tile: 64 pixels x 64 output channels per CTA, K (= input channel j) in chunks of 16.
// x element (b, c, p) lives at b*C*HW + c*sc + p*sp  (NCHW: sc=HW, sp=1; NHWC: sc=1, sp=C).
constexpr int GP = 64, GN = 64, GK = 16;

__global__ void __launch_bounds__(256) gdn_forward_kernel(const float *__restrict__ x, const float *__restrict__ beta,
                                                         const float *__restrict__ gamma, int inverse, int C,
                                                         int64_t HW, int64_t sc, int64_t sp, int64_t npix,
                                                         float *__restrict__ y)
{
    __shared__ float sX[GK][GP];   // squared inputs
    __shared__ float sG[GK][GN];   // gamma[i][j] transposed to [j][i]
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;          // tx: channel group, ty: pixel group
    const int64_t pix0 = (int64_t)blockIdx.x * GP;
    const int i0 = blockIdx.y * GN;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;

    for (int j0 = 0; j0 < C; j0 += GK) {
        for (int e = tid; e < GK * GP; e += 256) {
            int p, j;
            if (sp == 1) { p = e % GP; j = e / GP; } else { j = e % GK; p = e / GK; }
            int64_t pid = pix0 + p;
            float v = 0.0f;
            if (pid < npix && j0 + j < C) {
                int64_t b = pid / HW, q = pid - b * HW;
                v = __ldg(x + b * C * HW + (int64_t)(j0 + j) * sc + q * sp);
            }
            sX[j][p] = v * v;
        }
        for (int e = tid; e < GK * GN; e += 256) {
            int j = e % GK, i = e / GK;
            float v = 0.0f;
            if (i0 + i < C && j0 + j < C) v = __ldg(gamma + (int64_t)(i0 + i) * C + j0 + j);
            sG[j][i] = v;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < GK; ++j) {
            float a[4], g[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { a[q] = sX[j][ty * 4 + q]; g[q] = sG[j][tx * 4 + q]; }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[q][r] = fmaf(a[q], g[r], acc[q][r]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int64_t pid = pix0 + ty * 4 + q;
        if (pid >= npix) continue;
        int64_t b = pid / HW, pq = pid - b * HW;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int i = i0 + tx * 4 + r;
            if (i >= C) continue;
            int64_t idx = b * C * HW + (int64_t)i * sc + pq * sp;
            float norm = acc[q][r] + __ldg(beta + i);
            float xv = __ldg(x + idx);
            y[idx] = inverse ? xv * sqrtf(norm) : xv * rsqrtf(norm);
        }
    }
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_gdn_reparam(const float *beta, const float *gamma, int C, float beta_bound, float gamma_bound, float pedestal,
                    float *beta_eff, float *gamma_eff, void *gamma_eff_bf16, void *stream)
{
    MMC_CHECK_ARG(C >= 1 && C <= 4096, "mmc_gdn_reparam: C=%d out of range", C);
    MMC_CHECK_ARG(beta && gamma && beta_eff && (gamma_eff || gamma_eff_bf16), "mmc_gdn_reparam: NULL buffer");
    int n = C * C;
    gdn_reparam_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(beta, gamma, C, beta_bound, gamma_bound, pedestal,
                                                                        beta_eff, gamma_eff, (__nv_bfloat16 *)gamma_eff_bf16);
    MMC_CHECK_LAUNCH("mmc_gdn_reparam");
    return MMC_OK;
}

int mmc_gdn_forward(const float *x, const float *beta_eff, const float *gamma_eff, int inverse, int64_t B, int C,
                    int64_t HW, int layout, float *y, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && C >= 1 && HW >= 0, "mmc_gdn_forward: bad shape");
    MMC_CHECK_ARG(layout == MMC_NCHW || layout == MMC_NHWC, "mmc_gdn_forward: bad layout");
    int64_t npix = B * HW;
    if (npix == 0) return MMC_OK;
    MMC_CHECK_ARG(x && beta_eff && gamma_eff && y, "mmc_gdn_forward: NULL buffer");
    int64_t sc = (layout == MMC_NCHW) ? HW : 1, sp = (layout == MMC_NCHW) ? 1 : C;
    dim3 grid((unsigned)((npix + GP - 1) / GP), (unsigned)((C + GN - 1) / GN));
    gdn_forward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, beta_eff, gamma_eff, inverse, C, HW, sc, sp, npix, y);
    MMC_CHECK_LAUNCH("mmc_gdn_forward");
    return MMC_OK;
}

}  // extern "C"
