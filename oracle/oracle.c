/*
 * oracle.c -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the CUDA kernels in
 * 165-learning-based-multi-modality-image-and-video-compression_b200/csrc.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path never links or calls it.
 *
 * Parity status: PINNED.  Every function below is checked against outputs of the reference
 * itself (imported from /root/reference/CompressAI in the authoring container by
 * tests/golden/gen_golden.py; vectors committed under tests/golden/) and against the
 * reference's own known-answer tests (tests/test_ops.py:104-106, tests/test_layers.py:145-159,
 * tests/test_entropy_models.py:74-87).  See tests/test_oracle_golden.py.
 *
 * Citations are relative to /root/reference/CompressAI.  The convolution arithmetic itself
 * lives in a third-party dependency of the reference (PyTorch 2.11.0 nn.Conv2d /
 * nn.ConvTranspose2d, call sites compressai/models/utils.py:128-146); its published
 * definition (cross-correlation, zero padding; transposed conv as the gradient of conv) is
 * restated here with double accumulation.
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Indexing convention shared with the CUDA library: a tensor is viewed as [outer][C][inner]
 * (NCHW: outer=N, inner=H*W; channels-last: outer=N*H*W, inner=1).
 * means_mode: 0 = none, 1 = same shape as x, 2 = one value per channel (C floats).
 * ---------------------------------------------------------------------------------------- */
static inline float mean_at(const float *means, int mode, int64_t i, int64_t C, int64_t inner)
{
    if (mode == 0) return 0.0f;
    if (mode == 1) return means[i];
    return means[(i / inner) % C];
}

/* EntropyModel.quantize(mode="symbols"): compressai/entropy_models/entropy_models.py:169-182
 * (x - mean) in fp32, round-half-even (torch.round), then int32 cast. */
ORC_API void orc_quantize_symbols(const float *x, const float *means, int mode, int64_t outer,
                                  int64_t C, int64_t inner, int32_t *out)
{
    int64_t n = outer * C * inner;
    for (int64_t i = 0; i < n; ++i) {
        volatile float d = x[i];
        if (mode) d = d - mean_at(means, mode, i, C, inner);
        out[i] = (int32_t)rintf(d);
    }
}

/* EntropyModel.quantize(mode="dequantize"): entropy_models.py:169-178. round(x-m)+m */
ORC_API void orc_quantize_dequantize(const float *x, const float *means, int mode, int64_t outer,
                                     int64_t C, int64_t inner, float *out)
{
    int64_t n = outer * C * inner;
    for (int64_t i = 0; i < n; ++i) {
        float m = mean_at(means, mode, i, C, inner);
        volatile float d = x[i];
        if (mode) d = d - m;
        volatile float r = rintf(d);
        if (mode) r = r + m;
        out[i] = r;
    }
}

/* EntropyModel.quantize(mode="noise"): entropy_models.py:163-167; the uniform noise tensor is an
 * input so that both sides consume identical noise. */
ORC_API void orc_quantize_noise(const float *x, const float *noise, int64_t n, float *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = x[i] + noise[i];
}

/* EntropyModel.dequantize: entropy_models.py:190-199. int32 -> float (+ mean) */
ORC_API void orc_dequantize(const int32_t *sym, const float *means, int mode, int64_t outer,
                            int64_t C, int64_t inner, float *out)
{
    int64_t n = outer * C * inner;
    for (int64_t i = 0; i < n; ++i) {
        volatile float v = (float)sym[i];
        if (mode) v = v + mean_at(means, mode, i, C, inner);
        out[i] = v;
    }
}

/* LowerBound forward: compressai/ops/bound_ops.py:36-37 (torch.max propagates NaN). */
static inline float lower_bound_f(float x, float b) { return (x != x) ? x : (x > b ? x : b); }

ORC_API void orc_lower_bound(const float *x, float bound, int64_t n, float *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = lower_bound_f(x[i], bound);
}

/* LowerBound backward: bound_ops.py:40-42.  g * ((x >= b) | (g < 0)) */
ORC_API void orc_lower_bound_bwd(const float *x, const float *g, float bound, int64_t n, float *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = ((x[i] >= bound) || (g[i] < 0.0f)) ? g[i] : 0.0f;
}

/* GaussianConditional.build_indexes: entropy_models.py:735-740.
 * idx = (levels-1) - #{i < levels-1 : max(s, bound) <= table[i]} */
ORC_API void orc_build_indexes(const float *scales, const float *table, int levels, float bound,
                               int64_t n, int32_t *out)
{
    for (int64_t i = 0; i < n; ++i) {
        float s = lower_bound_f(scales[i], bound);
        int32_t idx = levels - 1;
        for (int k = 0; k < levels - 1; ++k) idx -= (s <= table[k]) ? 1 : 0;
        out[i] = idx;
    }
}

/* EntropyBottleneck._build_indexes: entropy_models.py:542-553. idx[n][c][...] = c */
ORC_API void orc_channel_indexes(int64_t outer, int64_t C, int64_t inner, int32_t *out)
{
    int64_t n = outer * C * inner;
    for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)((i / inner) % C);
}

/* ------------------------------------------------------------------------------------------
 * EntropyBottleneck with the default filters (3,3,3,3): entropy_models.py:361-379.
 * Raw parameter blocks, all [C][...] contiguous:
 *   m0 [C][3][1]  m1..m3 [C][3][3]  m4 [C][1][3]
 *   b0..b3 [C][3][1]  b4 [C][1][1]
 *   f0..f3 [C][3][1]
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const float *m[5];
    const float *b[5];
    const float *f[4];
} orc_eb_params;

static inline float softplus_f(float x)
{ /* torch F.softplus(beta=1, threshold=20) */
    return x > 20.0f ? x : log1pf(expf(x));
}

static inline float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

/* _logits_cumulative for one scalar of channel c: entropy_models.py:457-477 */
static float eb_logits_one(const orc_eb_params *p, int64_t c, float v)
{
    float h[3], u[3];
    int fin = 1;
    h[0] = v;
    for (int k = 0; k < 5; ++k) {
        int fout = (k == 4) ? 1 : 3;
        const float *M = p->m[k] + c * fout * fin;
        const float *B = p->b[k] + c * fout;
        for (int o = 0; o < fout; ++o) {
            float acc = 0.0f;
            for (int i = 0; i < fin; ++i) acc += softplus_f(M[o * fin + i]) * h[i];
            u[o] = acc + B[o];
        }
        if (k < 4) {
            const float *F = p->f[k] + c * fout;
            for (int o = 0; o < fout; ++o) u[o] = u[o] + tanhf(F[o]) * tanhf(u[o]);
        }
        for (int o = 0; o < fout; ++o) h[o] = u[o];
        fin = fout;
    }
    return h[0];
}

ORC_API void orc_eb_logits_cumulative(const orc_eb_params *p, const float *x, int64_t outer,
                                      int64_t C, int64_t inner, float *out)
{
    int64_t n = outer * C * inner;
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) out[i] = eb_logits_one(p, (i / inner) % C, x[i]);
}

/* _likelihood: entropy_models.py:480-492 */
static float eb_likelihood_one(const orc_eb_params *p, int64_t c, float v)
{
    float lower = eb_logits_one(p, c, v - 0.5f);
    float upper = eb_logits_one(p, c, v + 0.5f);
    float s = lower + upper;
    float sign = (s > 0.0f) ? -1.0f : ((s < 0.0f) ? 1.0f : 0.0f);
    return fabsf(sigmoid_f(sign * upper) - sigmoid_f(sign * lower));
}

/* EntropyBottleneck.forward: entropy_models.py:495-540.
 * training=0: x_hat = round(x - median) + median ; training=1: x_hat = x + noise.
 * medians = quantiles[:,0,1] (entropy_models.py:388-390), passed as C floats. */
ORC_API void orc_eb_forward(const orc_eb_params *p, const float *x, const float *medians,
                            const float *noise, int training, float lik_bound, int64_t outer,
                            int64_t C, int64_t inner, float *x_hat, float *lik)
{
    int64_t n = outer * C * inner;
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) {
        int64_t c = (i / inner) % C;
        float v;
        if (training) {
            v = x[i] + noise[i];
        } else {
            volatile float d = x[i] - medians[c];
            volatile float r = rintf(d);
            v = r + medians[c];
        }
        x_hat[i] = v;
        float l = eb_likelihood_one(p, c, v);
        lik[i] = (lik_bound > 0.0f) ? lower_bound_f(l, lik_bound) : l;
    }
}

/* GaussianConditional._standardized_cumulative: entropy_models.py:629-635 */
static inline float std_cumulative_f(float t)
{
    const float c = (float)(-0.70710678118654752440); /* float(-(2**-0.5)) */
    return 0.5f * erfcf(c * t);
}

/* GaussianConditional._likelihood + forward: entropy_models.py:692-731 */
ORC_API void orc_gc_forward(const float *x, const float *scales, const float *means,
                            const float *noise, int training, float scale_bound, float lik_bound,
                            int64_t n, float *x_hat, float *lik)
{
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) {
        float m = means ? means[i] : 0.0f;
        float v;
        if (training) {
            v = x[i] + noise[i];
        } else {
            volatile float d = x[i];
            if (means) d = d - m;
            volatile float r = rintf(d);
            if (means) r = r + m;
            v = r;
        }
        x_hat[i] = v;
        volatile float val = v;
        if (means) val = val - m;
        float a = fabsf(val);
        float s = lower_bound_f(scales[i], scale_bound);
        float upper = std_cumulative_f((0.5f - a) / s);
        float lower = std_cumulative_f((-0.5f - a) / s);
        float l = upper - lower;
        lik[i] = (lik_bound > 0.0f) ? lower_bound_f(l, lik_bound) : l;
    }
}

/* bpp numerator: sum(log(lik)) / (-ln 2); examples/train.py:74-77, utils/eval_model/__main__t.py:197-200 */
ORC_API double orc_bits(const float *lik, int64_t n)
{
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) acc += log((double)lik[i]);
    return acc / -log(2.0);
}

/* ------------------------------------------------------------------------------------------
 * GDN: compressai/layers/gdn.py:77-92 with NonNegativeParametrizer (ops/parametrizers.py:61-64)
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_gdn_reparam(const float *beta, const float *gamma, int C, float beta_bound,
                             float gamma_bound, float pedestal, float *beta_eff, float *gamma_eff)
{
    for (int i = 0; i < C; ++i) {
        float b = lower_bound_f(beta[i], beta_bound);
        beta_eff[i] = b * b - pedestal;
    }
    for (int i = 0; i < C * C; ++i) {
        float g = lower_bound_f(gamma[i], gamma_bound);
        gamma_eff[i] = g * g - pedestal;
    }
}

/* x: [B][C][HW] (NCHW).  norm_i = beta_i + sum_j gamma[i][j] * x_j^2 ; y = x * rsqrt(norm) | x * sqrt(norm) */
ORC_API void orc_gdn_forward(const float *x, const float *beta_eff, const float *gamma_eff,
                             int inverse, int64_t B, int C, int64_t HW, float *y)
{
#pragma omp parallel for collapse(2)
    for (int64_t b = 0; b < B; ++b)
        for (int64_t p = 0; p < HW; ++p) {
            const float *xp = x + b * C * HW + p;
            float *yp = y + b * C * HW + p;
            for (int i = 0; i < C; ++i) {
                double norm = beta_eff[i];
                for (int j = 0; j < C; ++j) {
                    double xj = xp[(int64_t)j * HW];
                    norm += (double)gamma_eff[i * C + j] * xj * xj;
                }
                double s = inverse ? sqrt(norm) : 1.0 / sqrt(norm);
                yp[(int64_t)i * HW] = (float)(xp[(int64_t)i * HW] * s);
            }
        }
}

/* ------------------------------------------------------------------------------------------
 * conv(): nn.Conv2d(k, stride, padding=k//2)  -- compressai/models/utils.py:128-135
 * x [B][Cin][H][W], w [Cout][Cin][k][k], bias [Cout] -> y [B][Cout][Ho][Wo]
 * Ho = (H + 2*pad - k)/stride + 1
 * act: 0 none, 1 ReLU, 2 LeakyReLU(0.01), 3 abs   (the activations between transform layers,
 * models/google.py:254-269,363-377; abs is the torch.abs(y) at models/google.py:283)
 * ---------------------------------------------------------------------------------------- */
static inline float act_f(float v, int act)
{
    switch (act) {
    case 1: return v > 0.0f ? v : 0.0f;
    case 2: return v > 0.0f ? v : 0.01f * v;
    case 3: return fabsf(v);
    default: return v;
    }
}

ORC_API void orc_conv2d(const float *x, const float *w, const float *bias, int64_t B, int Cin,
                        int H, int W, int Cout, int k, int stride, int pad, int act, float *y)
{
    int Ho = (H + 2 * pad - k) / stride + 1;
    int Wo = (W + 2 * pad - k) / stride + 1;
#pragma omp parallel for collapse(2)
    for (int64_t b = 0; b < B; ++b)
        for (int co = 0; co < Cout; ++co)
            for (int oy = 0; oy < Ho; ++oy)
                for (int ox = 0; ox < Wo; ++ox) {
                    double acc = bias ? bias[co] : 0.0;
                    for (int ci = 0; ci < Cin; ++ci)
                        for (int ky = 0; ky < k; ++ky) {
                            int iy = oy * stride - pad + ky;
                            if (iy < 0 || iy >= H) continue;
                            for (int kx = 0; kx < k; ++kx) {
                                int ix = ox * stride - pad + kx;
                                if (ix < 0 || ix >= W) continue;
                                acc += (double)x[((b * Cin + ci) * H + iy) * W + ix] *
                                       (double)w[((co * (int64_t)Cin + ci) * k + ky) * k + kx];
                            }
                        }
                    y[((b * Cout + co) * Ho + oy) * (int64_t)Wo + ox] = act_f((float)acc, act);
                }
}

/* deconv(): nn.ConvTranspose2d(k, stride, padding=k//2, output_padding=stride-1)
 *   -- compressai/models/utils.py:138-146
 * x [B][Cin][H][W], w [Cin][Cout][k][k] -> y [B][Cout][Ho][Wo], Ho = (H-1)*stride - 2*pad + k + outpad
 * y[oy][ox] += x[iy][ix] * w[ky][kx]  with  oy = iy*stride - pad + ky */
ORC_API void orc_conv_transpose2d(const float *x, const float *w, const float *bias, int64_t B,
                                  int Cin, int H, int W, int Cout, int k, int stride, int pad,
                                  int outpad, int act, float *y)
{
    int Ho = (H - 1) * stride - 2 * pad + k + outpad;
    int Wo = (W - 1) * stride - 2 * pad + k + outpad;
#pragma omp parallel for collapse(2)
    for (int64_t b = 0; b < B; ++b)
        for (int co = 0; co < Cout; ++co)
            for (int oy = 0; oy < Ho; ++oy)
                for (int ox = 0; ox < Wo; ++ox) {
                    double acc = bias ? bias[co] : 0.0;
                    for (int ky = 0; ky < k; ++ky) {
                        int ty = oy + pad - ky;
                        if (ty < 0 || ty % stride) continue;
                        int iy = ty / stride;
                        if (iy >= H) continue;
                        for (int kx = 0; kx < k; ++kx) {
                            int tx = ox + pad - kx;
                            if (tx < 0 || tx % stride) continue;
                            int ix = tx / stride;
                            if (ix >= W) continue;
                            for (int ci = 0; ci < Cin; ++ci)
                                acc += (double)x[((b * Cin + ci) * H + iy) * W + ix] *
                                       (double)w[((ci * (int64_t)Cout + co) * k + ky) * k + kx];
                        }
                    }
                    y[((b * Cout + co) * Ho + oy) * (int64_t)Wo + ox] = act_f((float)acc, act);
                }
}

/* ------------------------------------------------------------------------------------------
 * pmf_to_quantized_cdf: compressai/cpp_exts/ops/ops.cpp:40-109 (integer algorithm, bit-exact).
 * Returns 0 on success, -1 on a negative / non-finite entry, -2 when all entries round to 0,
 * -3 when no frequency can be stolen.  cdf has n+1 entries.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf)
{
    for (int i = 0; i < n; ++i)
        if (pmf[i] < 0.0f || !isfinite(pmf[i])) return -1;
    cdf[0] = 0;
    for (int i = 0; i < n; ++i) cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision));
    uint32_t total = 0;
    for (int i = 0; i <= n; ++i) total += cdf[i];
    if (total == 0) return -2;
    for (int i = 0; i <= n; ++i)
        cdf[i] = (uint32_t)((((uint64_t)1 << precision) * cdf[i]) / total);
    for (int i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
    cdf[n] = 1u << precision;
    for (int i = 0; i < n; ++i) {
        if (cdf[i] == cdf[i + 1]) {
            uint32_t best_freq = ~0u;
            int best_steal = -1;
            for (int j = 0; j < n; ++j) {
                uint32_t freq = cdf[j + 1] - cdf[j];
                if (freq > 1 && freq < best_freq) {
                    best_freq = freq;
                    best_steal = j;
                }
            }
            if (best_steal < 0) return -3;
            if (best_steal < i) {
                for (int j = best_steal + 1; j <= i; ++j) cdf[j]--;
            } else {
                for (int j = i + 1; j <= best_steal; ++j) cdf[j]++;
            }
        }
    }
    return 0;
}
