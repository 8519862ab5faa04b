// conv_direct.cu -- general CUDA-core convolution / transposed convolution (fp32 accumulate).
//
// Replaces conv()/deconv() of compressai/models/utils.py:128-146 (nn.Conv2d / nn.ConvTranspose2d
// with padding=k/2, output_padding=stride-1) plus the activation / GDN that follows it in the
// transform stacks (compressai/models/google.py:143-161,254-269; compressai/layers/gdn.py:77-92)
// for ANY channel count, layout and dtype.  It is the kernel for the layers the tensor-core path
// cannot take (Cin or Cout not a multiple of 64/16: the 3-channel image edges, 1-channel depth
// maps, unit-test sized layers) and the on-device cross-check of conv_tc.cu.
//
// Shape of the computation: a CTA owns PX output pixels x all Cout channels.  K = taps x Cin is
// walked in chunks of 16 input channels per tap, both operands staged in shared memory
// ([16][PX] activations, [16][TN] weights), each thread holding a TPX x TPN register tile.
// A transposed convolution is split into its stride*stride output phases (grid.y) so that only
// the taps that actually contribute to a phase are visited.  With GDN the CTA parks the
// (bias+act) tile in shared memory and applies the C x C normalisation before the store.
#include "common.cuh"

namespace mmc {

constexpr int kKC = 16;          // input channels per K chunk
constexpr int kThreads = 256;

struct DirectParams {
    mmc_conv_desc d;
    int Ho, Wo;
    int pad;
    int64_t npix;     // pixels per phase (conv: B*Ho*Wo; deconv: B*H*W)
    const void *x;
    const float *w;
    const float *bias;
    const float *beta;
    const float *gamma;
    void *y;
    __nv_bfloat16 *y2;
};

__device__ __forceinline__ float act_apply(float v, int act)
{
    switch (act) {
    case MMC_ACT_RELU: return fmaxf(v, 0.0f);
    case MMC_ACT_LEAKY_RELU: return v > 0.0f ? v : 0.01f * v;
    case MMC_ACT_ABS: return fabsf(v);
    case MMC_ACT_QRELU8: return fminf(fmaxf(v, 0.0f), 255.0f);
    default: return v;
    }
}

__device__ __forceinline__ float load_in(const DirectParams &P, int b, int iy, int ix, int ci)
{
    const mmc_conv_desc &d = P.d;
    if (iy < 0 || iy >= d.H || ix < 0 || ix >= d.W || ci >= d.Cin) return 0.0f;
    int64_t idx = (d.in_layout == MMC_NCHW) ? (((int64_t)b * d.Cin + ci) * d.H + iy) * d.W + ix
                                            : (((int64_t)b * d.H + iy) * d.W + ix) * d.Cin + ci;
    if (d.in_dtype == MMC_F32) return __ldg((const float *)P.x + idx);
    return __bfloat162float(((const __nv_bfloat16 *)P.x)[idx]);
}

__device__ __forceinline__ void store_out(const DirectParams &P, int b, int oy, int ox, int co, float v)
{
    const mmc_conv_desc &d = P.d;
    int64_t idx = (d.out_layout == MMC_NCHW) ? (((int64_t)b * d.Cout + co) * P.Ho + oy) * P.Wo + ox
                                             : (((int64_t)b * P.Ho + oy) * P.Wo + ox) * d.Cout + co;
    if (d.out_dtype == MMC_F32) ((float *)P.y)[idx] = v;
    else ((__nv_bfloat16 *)P.y)[idx] = __float2bfloat16_rn(v);
    if (P.y2) P.y2[(((int64_t)b * P.Ho + oy) * P.Wo + ox) * d.Cout + co] = __float2bfloat16_rn(d.out2_bf16 == 1 ? fabsf(v) : v);
}

template <int PX, int TN, int TPX, int TPN>
__global__ void __launch_bounds__(kThreads) conv_direct_kernel(DirectParams P)
{
    static_assert((PX / TPX) * (TN / TPN) == kThreads, "tile/thread mismatch");
    extern __shared__ float smem[];
    float *sA = smem;                       // [kKC][PX]
    float *sW = sA + kKC * PX;              // [kKC][TN]
    int *pb = (int *)(sW + kKC * TN);       // [PX] batch index (-1 = out of range)
    int *py = pb + PX;                      // [PX] output y
    int *px = py + PX;                      // [PX] output x
    float *sOut = (float *)(px + PX);       // [PX][Cout+1] only when GDN is fused

    const mmc_conv_desc &d = P.d;
    const int tid = threadIdx.x;
    const int s = d.stride, k = d.k, kk = k * k;
    const int phase = blockIdx.y;                 // deconv: (phase_y, phase_x); conv: 0
    const int ph_y = d.transposed ? phase / s : 0;
    const int ph_x = d.transposed ? phase % s : 0;
    const int gW = d.transposed ? d.W : P.Wo;     // pixel grid walked by this phase
    const int gH = d.transposed ? d.H : P.Ho;

    for (int p = tid; p < PX; p += kThreads) {
        int64_t pid = (int64_t)blockIdx.x * PX + p;
        if (pid < P.npix) {
            int gx = (int)(pid % gW);
            int64_t t = pid / gW;
            int gy = (int)(t % gH);
            pb[p] = (int)(t / gH);
            py[p] = d.transposed ? gy * s + ph_y : gy;
            px[p] = d.transposed ? gx * s + ph_x : gx;
        } else {
            pb[p] = -1; py[p] = 0; px[p] = 0;
        }
    }
    __syncthreads();

    const int tx = tid % (TN / TPN);   // channel group
    const int ty = tid / (TN / TPN);   // pixel group
    const int ldo = d.Cout + 1;

    for (int co0 = 0; co0 < d.Cout; co0 += TN) {
        float acc[TPX][TPN];
#pragma unroll
        for (int i = 0; i < TPX; ++i)
#pragma unroll
            for (int j = 0; j < TPN; ++j) acc[i][j] = 0.0f;

        for (int tap = 0; tap < kk; ++tap) {
            const int ky = tap / k, kx = tap % k;
            if (d.transposed) {
                // oy = iy*s - pad + ky  =>  ky == (oy + pad) mod s for this phase
                if (((ph_y + P.pad - ky) % s) != 0 || ((ph_x + P.pad - kx) % s) != 0) continue;
            }
            for (int ci0 = 0; ci0 < d.Cin; ci0 += kKC) {
                // ---- stage activations [kKC][PX] ----
                for (int e = tid; e < kKC * PX; e += kThreads) {
                    int p, c;
                    if (d.in_layout == MMC_NHWC) { c = e % kKC; p = e / kKC; }   // channel fastest in memory
                    else                         { p = e % PX;  c = e / PX;  }   // pixel fastest in memory
                    float v = 0.0f;
                    int b = pb[p];
                    if (b >= 0) {
                        int iy, ix;
                        if (d.transposed) {
                            iy = (py[p] + P.pad - ky) / s;   // exact by construction of the phase
                            ix = (px[p] + P.pad - kx) / s;
                            if (py[p] + P.pad - ky < 0) iy = -1;
                            if (px[p] + P.pad - kx < 0) ix = -1;
                        } else {
                            iy = py[p] * s - P.pad + ky;
                            ix = px[p] * s - P.pad + kx;
                        }
                        v = load_in(P, b, iy, ix, ci0 + c);
                    }
                    sA[c * PX + p] = v;
                }
                // ---- stage weights [kKC][TN] ----
                for (int e = tid; e < kKC * TN; e += kThreads) {
                    int c = e % kKC, n = e / kKC;
                    int ci = ci0 + c, co = co0 + n;
                    float v = 0.0f;
                    if (ci < d.Cin && co < d.Cout) {
                        int64_t wi = d.transposed ? (((int64_t)ci * d.Cout + co) * kk + tap)
                                                  : (((int64_t)co * d.Cin + ci) * kk + tap);
                        v = __ldg(P.w + wi);
                    }
                    sW[c * TN + n] = v;
                }
                __syncthreads();
#pragma unroll
                for (int c = 0; c < kKC; ++c) {
                    float a[TPX], wv[TPN];
#pragma unroll
                    for (int i = 0; i < TPX; ++i) a[i] = sA[c * PX + ty * TPX + i];
#pragma unroll
                    for (int j = 0; j < TPN; ++j) wv[j] = sW[c * TN + tx * TPN + j];
#pragma unroll
                    for (int i = 0; i < TPX; ++i)
#pragma unroll
                        for (int j = 0; j < TPN; ++j) acc[i][j] = fmaf(a[i], wv[j], acc[i][j]);
                }
                __syncthreads();
            }
        }
        // ---- epilogue for this channel chunk ----
#pragma unroll
        for (int i = 0; i < TPX; ++i) {
            int p = ty * TPX + i;
            int b = pb[p];
#pragma unroll
            for (int j = 0; j < TPN; ++j) {
                int co = co0 + tx * TPN + j;
                if (co >= d.Cout) continue;
                float v = acc[i][j] + (P.bias ? __ldg(P.bias + co) : 0.0f);
                v = act_apply(v, d.act);
                if (d.gdn != MMC_GDN_NONE) sOut[p * ldo + co] = v;
                else if (b >= 0) store_out(P, b, py[p], px[p], co, v);
            }
        }
    }
    if (d.gdn == MMC_GDN_NONE) return;
    __syncthreads();
    // ---- fused GDN: norm_i = beta_i + sum_j gamma_ij * v_j^2 (layers/gdn.py:83-90) ----
    // thread <-> (channel i, group of 4 pixels): each gamma element is reused for 4 pixels
    const int C = d.Cout;
    const int groups = PX / 4;
    for (int e = tid; e < C * groups; e += kThreads) {
        int i = e % C, g = e / C;
        float n0 = __ldg(P.beta + i), n1 = n0, n2 = n0, n3 = n0;
        const float *grow = P.gamma + (int64_t)i * C;
        const float *o0 = sOut + (g * 4 + 0) * ldo, *o1 = o0 + ldo, *o2 = o1 + ldo, *o3 = o2 + ldo;
        for (int j = 0; j < C; ++j) {
            float gv = __ldg(grow + j);
            n0 = fmaf(gv, o0[j] * o0[j], n0);
            n1 = fmaf(gv, o1[j] * o1[j], n1);
            n2 = fmaf(gv, o2[j] * o2[j], n2);
            n3 = fmaf(gv, o3[j] * o3[j], n3);
        }
        float nn[4] = {n0, n1, n2, n3};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int p = g * 4 + q;
            if (pb[p] < 0) continue;
            float v = sOut[p * ldo + i];
            float r = (d.gdn == MMC_GDN_INVERSE) ? v * sqrtf(nn[q]) : v * rsqrtf(nn[q]);
            // write after all reads of column i are done for this thread; other threads read other
            // (p, j) pairs of sOut, so results go straight to global memory, not back to smem
            store_out(P, pb[p], py[p], px[p], i, r);
        }
    }
}

static int validate_desc(const mmc_conv_desc *d, const char *name)
{
    MMC_CHECK_ARG(d != nullptr, "%s: descriptor is NULL", name);
    MMC_CHECK_ARG(d->B >= 0 && d->H >= 1 && d->W >= 1 && d->Cin >= 1 && d->Cout >= 1, "%s: bad shape", name);
    MMC_CHECK_ARG(d->k == 1 || d->k == 3 || d->k == 5, "%s: kernel size %d not in {1,3,5}", name, d->k);
    MMC_CHECK_ARG(d->stride == 1 || d->stride == 2, "%s: stride %d not in {1,2}", name, d->stride);
    MMC_CHECK_ARG(d->in_dtype == MMC_F32 || d->in_dtype == MMC_BF16, "%s: bad in_dtype", name);
    MMC_CHECK_ARG(d->out_dtype == MMC_F32 || d->out_dtype == MMC_BF16, "%s: bad out_dtype", name);
    MMC_CHECK_ARG(d->in_layout == MMC_NCHW || d->in_layout == MMC_NHWC || d->in_layout == MMC_NHWC_PAD8, "%s: bad in_layout", name);
    MMC_CHECK_ARG(d->out_layout == MMC_NCHW || d->out_layout == MMC_NHWC, "%s: bad out_layout", name);
    MMC_CHECK_ARG(d->act >= 0 && d->act <= MMC_ACT_QRELU8, "%s: bad act", name);
    MMC_CHECK_ARG(d->gdn >= 0 && d->gdn <= MMC_GDN_INVERSE, "%s: bad gdn mode", name);
    return MMC_OK;
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_conv_out_size(const mmc_conv_desc *d, int *Ho, int *Wo)
{
    int rc = validate_desc(d, "mmc_conv_out_size");
    if (rc) return rc;
    int pad = d->k / 2;
    if (d->transposed) {
        *Ho = (d->H - 1) * d->stride - 2 * pad + d->k + (d->stride - 1);
        *Wo = (d->W - 1) * d->stride - 2 * pad + d->k + (d->stride - 1);
    } else {
        *Ho = (d->H + 2 * pad - d->k) / d->stride + 1;
        *Wo = (d->W + 2 * pad - d->k) / d->stride + 1;
    }
    return MMC_OK;
}

int mmc_conv_forward_direct(const mmc_conv_desc *d, const void *x, const float *w, const float *bias,
                            const float *beta_eff, const float *gamma_eff, void *y, void *y2, void *stream)
{
    int rc = validate_desc(d, "mmc_conv_forward_direct");
    if (rc) return rc;
    MMC_CHECK_ARG(d->in_layout != MMC_NHWC_PAD8, "mmc_conv_forward_direct: NHWC_PAD8 input is a tensor-core staging layout");
    MMC_CHECK_ARG(d->gdn == MMC_GDN_NONE || (beta_eff && gamma_eff), "mmc_conv_forward_direct: GDN needs beta/gamma");
    MMC_CHECK_ARG(d->out2_bf16 >= 0 && d->out2_bf16 <= 2, "mmc_conv_forward_direct: bad out2_bf16");
    MMC_CHECK_ARG(!d->out2_bf16 || y2, "mmc_conv_forward_direct: out2_bf16 set but y2 is NULL");
    if (d->B == 0) return MMC_OK;
    MMC_CHECK_ARG(x && w && y, "mmc_conv_forward_direct: NULL buffer");
    DirectParams P;
    P.d = *d;
    mmc_conv_out_size(d, &P.Ho, &P.Wo);
    P.pad = d->k / 2;
    P.x = x; P.w = w; P.bias = bias; P.beta = beta_eff; P.gamma = gamma_eff; P.y = y;
    P.y2 = d->out2_bf16 ? (__nv_bfloat16 *)y2 : nullptr;
    int phases = d->transposed ? d->stride * d->stride : 1;
    P.npix = d->transposed ? (int64_t)d->B * d->H * d->W : (int64_t)d->B * P.Ho * P.Wo;
    cudaStream_t st = (cudaStream_t)stream;
    const bool narrow = d->Cout <= 4 && d->gdn == MMC_GDN_NONE;
    if (narrow) {
        constexpr int PX = 256, TN = 4;
        size_t smem = sizeof(float) * (kKC * PX + kKC * TN) + sizeof(int) * 3 * PX;
        dim3 grid((unsigned)((P.npix + PX - 1) / PX), phases);
        conv_direct_kernel<PX, TN, 1, 4><<<grid, kThreads, smem, st>>>(P);
    } else {
        constexpr int PX = 64, TN = 64;
        size_t smem = sizeof(float) * (kKC * PX + kKC * TN) + sizeof(int) * 3 * PX;
        if (d->gdn != MMC_GDN_NONE) smem += sizeof(float) * PX * (d->Cout + 1);
        MMC_UNSUPPORTED(smem > 200 * 1024, "mmc_conv_forward_direct: Cout=%d too wide for fused GDN", d->Cout);
        auto kern = conv_direct_kernel<PX, TN, 4, 4>;
        if (smem > 48 * 1024) MMC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)((P.npix + PX - 1) / PX), phases);
        kern<<<grid, kThreads, smem, st>>>(P);
    }
    MMC_CHECK_LAUNCH("mmc_conv_forward_direct");
    return MMC_OK;
}

}  // extern "C"
