// entropy.cu -- entropy-model stage of the codec hot path (HBM-bound elementwise kernels).
//
// Replaces (file:line relative to /root/reference/CompressAI):
//   EntropyModel.quantize / dequantize            compressai/entropy_models/entropy_models.py:157-199
//   EntropyBottleneck.forward / _likelihood       entropy_models.py:457-540
//   GaussianConditional.forward / _likelihood     entropy_models.py:692-731
//   GaussianConditional.build_indexes             entropy_models.py:735-740
//   EntropyBottleneck._build_indexes              entropy_models.py:542-553
//   LowerBound                                    compressai/ops/bound_ops.py:36-42
//   bpp reduction                                 examples/train.py:74-77
//
// Design: every kernel is one pass over its operands with 128-bit loads/stores (4 fp32 per thread
// per step), a grid that is a multiple of the 148 SMs, and streaming (L1::no_allocate) loads since
// nothing is re-read.  Per-channel parameters live in registers (EB) or shared memory (scale
// table).  Optional bit counts are reduced warp-shuffle -> smem -> one atomicAdd per CTA.
#include "common.cuh"

namespace mmc {

constexpr int kBlock = 256;

// -------------------------------------------------------------------------------------------
// channel lookup for the [outer][C][inner] view
// -------------------------------------------------------------------------------------------
struct ChanIndex {
    uint32_t C, inner;
    __device__ __forceinline__ uint32_t operator()(int64_t i) const
    {
        // n < 2^40 in practice; do the division in 64 bit only when needed
        if (inner == 1) return (uint32_t)(i % C);
        return (uint32_t)((i / inner) % C);
    }
};

template <int kMode>
__device__ __forceinline__ float mean_of(const float *__restrict__ means, int64_t i, const ChanIndex &ci)
{
    if (kMode == MMC_MEANS_NONE) return 0.0f;
    if (kMode == MMC_MEANS_FULL) return __ldg(means + i);
    return __ldg(means + ci(i));
}

// op: 0 = symbols (int32), 1 = dequantize (float), 2 = dequantize symbols (int32 in -> float)
template <int kMode, int kOp, bool kVec>
__global__ void __launch_bounds__(kBlock) quantize_kernel(const void *__restrict__ xin,
                                                          const float *__restrict__ means, ChanIndex ci,
                                                          int64_t n, void *__restrict__ outp)
{
    const float *x = (const float *)xin;
    const int32_t *xs = (const int32_t *)xin;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    auto one = [&](int64_t i, float xv, int32_t sv) -> uint32_t {
        float m = mean_of<kMode>(means, i, ci);
        if (kOp == 2) {
            float v = (float)sv;
            if (kMode != MMC_MEANS_NONE) v = __fadd_rn(v, m);
            return __float_as_uint(v);
        }
        float d = (kMode != MMC_MEANS_NONE) ? __fsub_rn(xv, m) : xv;
        float r = rintf(d);  // round-half-even == torch.round
        if (kOp == 0) return (uint32_t)(int32_t)r;  // cvt.rzi of an integral value
        if (kMode != MMC_MEANS_NONE) r = __fadd_rn(r, m);
        return __float_as_uint(r);
    };
    if (kVec) {
        int64_t n4 = n >> 2;
        for (int64_t v = tid; v < n4; v += stride) {
            uint4 in = __ldg(reinterpret_cast<const uint4 *>(xin) + v);
            uint4 o;
            int64_t i = v << 2;
            o.x = one(i + 0, __uint_as_float(in.x), (int32_t)in.x);
            o.y = one(i + 1, __uint_as_float(in.y), (int32_t)in.y);
            o.z = one(i + 2, __uint_as_float(in.z), (int32_t)in.z);
            o.w = one(i + 3, __uint_as_float(in.w), (int32_t)in.w);
            reinterpret_cast<uint4 *>(outp)[v] = o;
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride)
            ((uint32_t *)outp)[i] = one(i, kOp == 2 ? 0.f : x[i], kOp == 2 ? xs[i] : 0);
    } else {
        for (int64_t i = tid; i < n; i += stride)
            ((uint32_t *)outp)[i] = one(i, kOp == 2 ? 0.f : x[i], kOp == 2 ? xs[i] : 0);
    }
}

template <int kOp>
static int launch_quantize(const void *x, const float *means, int mode, int64_t outer, int64_t C, int64_t inner,
                           void *out, cudaStream_t st, const char *name)
{
    MMC_CHECK_ARG(outer >= 0 && C >= 1 && inner >= 1, "%s: bad shape outer=%lld C=%lld inner=%lld", name,
                  (long long)outer, (long long)C, (long long)inner);
    MMC_CHECK_ARG(mode >= 0 && mode <= 2, "%s: bad means_mode %d", name, mode);
    MMC_CHECK_ARG(mode == MMC_MEANS_NONE || means != nullptr, "%s: means is NULL", name);
    MMC_CHECK_ARG(C < (1ll << 31) && inner < (1ll << 31), "%s: C/inner too large", name);
    int64_t n = outer * C * inner;
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && out, "%s: NULL buffer", name);
    bool vec = aligned16(x) && aligned16(out) && (mode != MMC_MEANS_FULL || aligned16(means));
    ChanIndex ci{(uint32_t)C, (uint32_t)inner};
    int grid = elementwise_grid((n + 3) / 4, kBlock);
#define MMC_Q(M, V) quantize_kernel<M, kOp, V><<<grid, kBlock, 0, st>>>(x, means, ci, n, out)
    if (vec) {
        if (mode == 0) MMC_Q(0, true); else if (mode == 1) MMC_Q(1, true); else MMC_Q(2, true);
    } else {
        if (mode == 0) MMC_Q(0, false); else if (mode == 1) MMC_Q(1, false); else MMC_Q(2, false);
    }
#undef MMC_Q
    MMC_CHECK_LAUNCH(name);
    return MMC_OK;
}

// -------------------------------------------------------------------------------------------
// small elementwise ops
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) add_kernel(const float *__restrict__ a, const float *__restrict__ b,
                                                     int64_t n, float *__restrict__ out)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __fadd_rn(a[i], b[i]);
}

__global__ void __launch_bounds__(kBlock) lower_bound_kernel(const float *__restrict__ x, float b, int64_t n,
                                                             float *__restrict__ out)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = lower_bound_f(x[i], b);
}

__global__ void __launch_bounds__(kBlock) lower_bound_bwd_kernel(const float *__restrict__ x, const float *__restrict__ g,
                                                                 float b, int64_t n, float *__restrict__ out)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = ((x[i] >= b) || (g[i] < 0.0f)) ? g[i] : 0.0f;
}

__global__ void __launch_bounds__(kBlock) channel_indexes_kernel(ChanIndex ci, int64_t n, int32_t *__restrict__ out)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (int32_t)ci(i);
}

// build_indexes: the sorted table sits in shared memory; idx = (L-1) - #{k < L-1 : s <= t_k}.
// The table is sorted ascending (validated by the reference constructor, entropy_models.py:599),
// so the count equals the lower_bound position of s in the first L-1 entries (a log2 guess corrected against the table, see
// below).  NaN compares false everywhere -> L-1, as in the reference.
template <bool kVec>
__global__ void __launch_bounds__(kBlock) build_indexes_kernel(const float *__restrict__ scales,
                                                               const float *__restrict__ table, int levels,
                                                               float bound, int64_t n, int32_t *__restrict__ out)
{
    // One copy of the table per shared-memory bank (entry k of lane l at word 32 k + l): the eight probes of the search are
    // data-dependent, and with a single copy the 32 lanes of a warp hit random banks (~4-way conflicts: the kernel ran at 39 % of
    // the HBM peak, LDS-bound); with a private bank per lane every probe is one conflict-free wavefront.
    __shared__ float tab[256 * 32];
    const int nrep = (levels - 1 > 0 ? levels - 1 : 0);
    for (int k = threadIdx.x; k < nrep * 32; k += blockDim.x) tab[k] = table[k >> 5];
    __syncthreads();
    const int L1 = levels - 1;
    const int lane = threadIdx.x & 31;
    // The reference's table is geometric (exp(linspace(log lo, log hi, L)), models/google.py:32-34), so log2(s) lands within one
    // entry of the answer; the exact position is then fixed up against the TABLE ITSELF (never recomputed: bit-exact for any
    // sorted table, a non-geometric one just walks further).  ~15 instructions per element instead of an 8-probe binary
    // search (~55: the kernel was issue-bound at 39 % of the HBM peak).
    float l0 = 0.0f, inv = 0.0f;
    if (L1 >= 2 && table[0] > 0.0f && table[L1 - 1] > table[0]) {
        l0 = __log2f(table[0]);
        inv = (float)(L1 - 1) / (__log2f(table[L1 - 1]) - l0);
    }
    auto one = [&](float sv) -> int32_t {
        float s = lower_bound_f(sv, bound);
        if (s != s) return L1;
        // first position p in [0, L1] with tab[p] >= s  (count of entries < s)
        int p = __float2int_ru((__log2f(s) - l0) * inv);
        p = min(max(p, 0), L1);
        while (p > 0 && !(tab[((p - 1) << 5) + lane] < s)) --p;
        while (p < L1 && tab[(p << 5) + lane] < s) ++p;
        return p;
    };
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (kVec) {
        // four independent 16-byte loads in flight per thread before the (dependent, ~250-cycle) searches start: with one load per
        // iteration the kernel was latency-bound (28 KB in flight per SM)
        int64_t n4 = n >> 2;
        int64_t v = tid;
        for (; v + 3 * stride < n4; v += 4 * stride) {
            float4 s[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) s[u] = ldg_stream(reinterpret_cast<const float4 *>(scales) + v + u * stride);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                reinterpret_cast<int4 *>(out)[v + u * stride] = make_int4(one(s[u].x), one(s[u].y), one(s[u].z), one(s[u].w));
        }
        for (; v < n4; v += stride) {
            float4 s = ldg_stream(reinterpret_cast<const float4 *>(scales) + v);
            int4 o = make_int4(one(s.x), one(s.y), one(s.z), one(s.w));
            reinterpret_cast<int4 *>(out)[v] = o;
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride) out[i] = one(scales[i]);
    } else {
        for (int64_t i = tid; i < n; i += stride) out[i] = one(scales[i]);
    }
}

// -------------------------------------------------------------------------------------------
// EntropyBottleneck
// -------------------------------------------------------------------------------------------
struct EbRegs {
    float W0[3], W1[9], W2[9], W3[9], W4[3];
    float b0[3], b1[3], b2[3], b3[3], b4;
    float A0[3], A1[3], A2[3], A3[3];
    float median;
};

__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

__device__ __forceinline__ void eb_load(const mmc_eb_params &p, uint32_t c, EbRegs &r)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r.W0[i] = softplus_f(__ldg(p.matrix[0] + c * 3 + i));
        r.W4[i] = softplus_f(__ldg(p.matrix[4] + c * 3 + i));
        r.b0[i] = __ldg(p.bias[0] + c * 3 + i);
        r.b1[i] = __ldg(p.bias[1] + c * 3 + i);
        r.b2[i] = __ldg(p.bias[2] + c * 3 + i);
        r.b3[i] = __ldg(p.bias[3] + c * 3 + i);
        r.A0[i] = tanhf(__ldg(p.factor[0] + c * 3 + i));
        r.A1[i] = tanhf(__ldg(p.factor[1] + c * 3 + i));
        r.A2[i] = tanhf(__ldg(p.factor[2] + c * 3 + i));
        r.A3[i] = tanhf(__ldg(p.factor[3] + c * 3 + i));
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        r.W1[i] = softplus_f(__ldg(p.matrix[1] + c * 9 + i));
        r.W2[i] = softplus_f(__ldg(p.matrix[2] + c * 9 + i));
        r.W3[i] = softplus_f(__ldg(p.matrix[3] + c * 9 + i));
    }
    r.b4 = __ldg(p.bias[4] + c);
    r.median = p.medians ? __ldg(p.medians + c) : 0.0f;
}

__device__ __forceinline__ void eb_layer3(const float *W, const float *b, const float *A, float *h)
{
    float u[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        float acc = W[o * 3 + 0] * h[0];
        acc = fmaf(W[o * 3 + 1], h[1], acc);
        acc = fmaf(W[o * 3 + 2], h[2], acc);
        acc += b[o];
        u[o] = fmaf(A[o], tanhf(acc), acc);
    }
    h[0] = u[0]; h[1] = u[1]; h[2] = u[2];
}

// _logits_cumulative for one scalar (entropy_models.py:457-477)
__device__ __forceinline__ float eb_logits(const EbRegs &r, float v)
{
    float h[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        float acc = fmaf(r.W0[o], v, r.b0[o]);
        h[o] = fmaf(r.A0[o], tanhf(acc), acc);
    }
    eb_layer3(r.W1, r.b1, r.A1, h);
    eb_layer3(r.W2, r.b2, r.A2, h);
    eb_layer3(r.W3, r.b3, r.A3, h);
    float acc = r.W4[0] * h[0];
    acc = fmaf(r.W4[1], h[1], acc);
    acc = fmaf(r.W4[2], h[2], acc);
    return acc + r.b4;
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// _likelihood (entropy_models.py:480-492)
__device__ __forceinline__ float eb_likelihood(const EbRegs &r, float v)
{
    float lower = eb_logits(r, v - 0.5f);
    float upper = eb_logits(r, v + 0.5f);
    float s = lower + upper;
    float sign = (s > 0.0f) ? -1.0f : ((s < 0.0f) ? 1.0f : 0.0f);
    return fabsf(sigmoid_f(sign * upper) - sigmoid_f(sign * lower));
}

// mode 0: forward (x_hat, likelihood[, bf16 copy][, bits]); mode 1: logits only.
// Thread <-> channel binding:
//   kChannelsLast : total thread count is a multiple of C, element i = tid + k*T has channel tid % C
//   otherwise     : blockIdx.y = channel; threads walk the outer x inner elements of that channel
template <bool kChannelsLast, int kModeOp>
__global__ void __launch_bounds__(kBlock) eb_kernel(const float *__restrict__ x, const float *__restrict__ noise,
                                                    mmc_eb_params p, float lik_bound, int64_t outer, uint32_t C,
                                                    int64_t inner, float *__restrict__ x_hat,
                                                    __nv_bfloat16 *__restrict__ x_hat_bf16, float *__restrict__ lik,
                                                    float *__restrict__ bits)
{
    __shared__ float red[32];
    EbRegs r;
    float bit_acc = 0.0f;
    auto body = [&](int64_t i) {
        float xv = x[i];
        if (kModeOp == 1) {
            lik[i] = eb_logits(r, xv);
            return;
        }
        float v;
        if (noise) {
            v = __fadd_rn(xv, noise[i]);
        } else {
            v = __fadd_rn(rintf(__fsub_rn(xv, r.median)), r.median);
        }
        float l = eb_likelihood(r, v);
        if (lik_bound > 0.0f) l = lower_bound_f(l, lik_bound);
        x_hat[i] = v;
        if (x_hat_bf16) x_hat_bf16[i] = __float2bfloat16_rn(v);
        lik[i] = l;
        if (bits) bit_acc -= log2f(l);
    };
    if (kChannelsLast) {
        int64_t n = outer * C;
        int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        int64_t T = (int64_t)gridDim.x * blockDim.x;  // multiple of C by construction
        eb_load(p, (uint32_t)(tid % C), r);
        for (int64_t i = tid; i < n; i += T) body(i);
    } else {
        uint32_t c = blockIdx.y;
        eb_load(p, c, r);
        int64_t per_chan = outer * inner;
        int64_t T = (int64_t)gridDim.x * blockDim.x;
        for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < per_chan; j += T) {
            int64_t o = j / inner, in = j - o * inner;
            body((o * C + c) * inner + in);
        }
    }
    if (kModeOp == 0 && bits) block_atomic_add(bit_acc, red, bits);
}

// ---- backward of the noise-mode likelihood (training) -------------------------------------
// Per element: both logits forward, then each branch again keeping h_k / tanh(u_k) and back-propagating
// delta(u_4) = +-g m sign(q) s sigma'(s logit) down to the input.  The 58 parameter gradients of the thread's channel
// accumulate in registers; dM = dW * sigmoid(M) = dW * (1 - exp(-W)), da = dA * (1 - A^2) are applied once at the end.
struct EbGrads {
    float W0[3], W1[9], W2[9], W3[9], W4[3];
    float b0[3], b1[3], b2[3], b3[3], b4;
    float A0[3], A1[3], A2[3], A3[3];
};

__device__ __forceinline__ void eb_bwd_layer3(const float *W, const float *A, const float *h_in, const float *th, const float *dh_out,
                                              float *dW, float *db, float *dA, float *dh_in)
{
    float du[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        dA[o] += dh_out[o] * th[o];
        du[o] = dh_out[o] * fmaf(A[o], 1.0f - th[o] * th[o], 1.0f);
        db[o] += du[o];
#pragma unroll
        for (int j = 0; j < 3; ++j) dW[o * 3 + j] = fmaf(du[o], h_in[j], dW[o * 3 + j]);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) dh_in[j] = W[j] * du[0] + W[3 + j] * du[1] + W[6 + j] * du[2];
}

// one branch: forward from v keeping intermediates, backward from du4; returns d(logit)/dv * du4
__device__ __forceinline__ float eb_branch_bwd(const EbRegs &r, float v, float du4, EbGrads &G)
{
    float h1[3], h2[3], h3[3], h4[3], t0[3], t1[3], t2[3], t3[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        const float u = fmaf(r.W0[o], v, r.b0[o]);
        t0[o] = tanhf(u);
        h1[o] = fmaf(r.A0[o], t0[o], u);
    }
    auto fwd3 = [](const float *W, const float *b, const float *A, const float *hin, float *t, float *hout) {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            float u = W[o * 3] * hin[0];
            u = fmaf(W[o * 3 + 1], hin[1], u);
            u = fmaf(W[o * 3 + 2], hin[2], u);
            u += b[o];
            t[o] = tanhf(u);
            hout[o] = fmaf(A[o], t[o], u);
        }
    };
    fwd3(r.W1, r.b1, r.A1, h1, t1, h2);
    fwd3(r.W2, r.b2, r.A2, h2, t2, h3);
    fwd3(r.W3, r.b3, r.A3, h3, t3, h4);
    // layer 4 (linear)
    float dh4[3], dh3[3], dh2[3], dh1[3];
    G.b4 += du4;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        G.W4[j] = fmaf(du4, h4[j], G.W4[j]);
        dh4[j] = r.W4[j] * du4;
    }
    eb_bwd_layer3(r.W3, r.A3, h3, t3, dh4, G.W3, G.b3, G.A3, dh3);
    eb_bwd_layer3(r.W2, r.A2, h2, t2, dh3, G.W2, G.b2, G.A2, dh2);
    eb_bwd_layer3(r.W1, r.A1, h1, t1, dh2, G.W1, G.b1, G.A1, dh1);
    float dv = 0.0f;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        G.A0[o] += dh1[o] * t0[o];
        const float du = dh1[o] * fmaf(r.A0[o], 1.0f - t0[o] * t0[o], 1.0f);
        G.b0[o] += du;
        G.W0[o] = fmaf(du, v, G.W0[o]);
        dv = fmaf(r.W0[o], du, dv);
    }
    return dv;
}

template <bool kChannelsLast>
__global__ void __launch_bounds__(kBlock) eb_bwd_kernel(const float *__restrict__ x, const float *__restrict__ noise, const float *__restrict__ g,
                                                        mmc_eb_params p, float lik_bound, int64_t outer, uint32_t C, int64_t inner,
                                                        float *__restrict__ dx, float *__restrict__ dparams /* [C][58] */)
{
    EbRegs r;
    EbGrads G;
    float *Gf = reinterpret_cast<float *>(&G);
#pragma unroll
    for (int i = 0; i < 58; ++i) Gf[i] = 0.0f;
    bool any = false;
    auto body = [&](int64_t i) {
        const float v = x[i] + noise[i];
        const float lo = eb_logits(r, v - 0.5f), up = eb_logits(r, v + 0.5f);
        const float sum = lo + up;
        const float s = (sum > 0.0f) ? -1.0f : ((sum < 0.0f) ? 1.0f : 0.0f);
        const float su = sigmoid_f(s * up), sl = sigmoid_f(s * lo);
        const float q = su - sl, p_raw = fabsf(q);
        const float gv = g[i];
        const float gm = (lik_bound <= 0.0f || p_raw >= lik_bound || gv < 0.0f) ? gv : 0.0f;
        const float sq = q > 0.0f ? 1.0f : (q < 0.0f ? -1.0f : 0.0f);
        const float du_up = gm * sq * s * su * (1.0f - su);
        const float du_lo = -gm * sq * s * sl * (1.0f - sl);
        float d = eb_branch_bwd(r, v + 0.5f, du_up, G);
        d += eb_branch_bwd(r, v - 0.5f, du_lo, G);
        dx[i] = d;
        any = true;
    };
    uint32_t c;
    if (kChannelsLast) {
        int64_t n = outer * C;
        int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        int64_t T = (int64_t)gridDim.x * blockDim.x;  // multiple of C by construction
        c = (uint32_t)(tid % C);
        eb_load(p, c, r);
        for (int64_t i = tid; i < n; i += T) body(i);
    } else {
        c = blockIdx.y;
        eb_load(p, c, r);
        int64_t per_chan = outer * inner;
        int64_t T = (int64_t)gridDim.x * blockDim.x;
        for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < per_chan; j += T) {
            int64_t o = j / inner, in = j - o * inner;
            body((o * C + c) * inner + in);
        }
    }
    // chain through softplus / tanh of the raw parameters
    const float *Wf = reinterpret_cast<const float *>(&r);       // W0..W4 (33), b0..b4 (13), A0..A3 (12): same order as EbGrads
#pragma unroll
    for (int i = 0; i < 33; ++i) Gf[i] *= (1.0f - expf(-Wf[i]));
#pragma unroll
    for (int i = 46; i < 58; ++i) Gf[i] *= (1.0f - Wf[i] * Wf[i]);
    float *dst = dparams + (size_t)c * 58;
    if (kChannelsLast) {
        if (any)
#pragma unroll
            for (int i = 0; i < 58; ++i) atomicAdd(dst + i, Gf[i]);
    } else {
        // the whole block works on channel c: warp-reduce, one atomic per warp
#pragma unroll
        for (int i = 0; i < 58; ++i) {
            const float sum = warp_sum(Gf[i]);
            if ((threadIdx.x & 31) == 0) atomicAdd(dst + i, sum);
        }
    }
}

// ---- eval-mode fast path: likelihood table per (channel, integer symbol) ---------------------------------
// In eval mode x_hat = rint(x - median) + median, so the likelihood is a function of (channel, symbol) only -- the
// same observation update() uses to tabulate the CDFs (entropy_models.py:422-432).  build: one thread per entry.
__global__ void __launch_bounds__(kBlock) eb_build_lut_kernel(mmc_eb_params p, float lik_bound, uint32_t C, int half_width,
                                                              float *__restrict__ lut)
{
    const int width = 2 * half_width + 1;
    int64_t n = (int64_t)C * width;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = (uint32_t)(i / width);
    const int k = (int)(i - (int64_t)c * width) - half_width;
    EbRegs r;
    eb_load(p, c, r);
    float l = eb_likelihood(r, __fadd_rn((float)k, r.median));
    if (lik_bound > 0.0f) l = lower_bound_f(l, lik_bound);
    lut[i] = l;
}

__device__ __noinline__ float eb_likelihood_slow(const mmc_eb_params &p, uint32_t c, float v, float lik_bound)
{
    EbRegs r;
    eb_load(p, c, r);
    float l = eb_likelihood(r, v);
    return lik_bound > 0.0f ? lower_bound_f(l, lik_bound) : l;
}

template <bool kVec>
__global__ void __launch_bounds__(kBlock) eb_lut_kernel(const float *__restrict__ x, mmc_eb_params p, const float *__restrict__ lut,
                                                        int half_width, float lik_bound, ChanIndex ci, int64_t n,
                                                        float *__restrict__ x_hat, __nv_bfloat16 *__restrict__ x_hat_bf16,
                                                        float *__restrict__ lik, float *__restrict__ bits)
{
    __shared__ float red[32];
    float bit_acc = 0.0f;
    const int width = 2 * half_width + 1;
    auto one = [&](int64_t i, float xv, float &v, float &l) {
        const uint32_t c = ci(i);
        const float med = __ldg(p.medians + c);
        const float kf = rintf(__fsub_rn(xv, med));
        v = __fadd_rn(kf, med);
        if (fabsf(kf) <= (float)half_width) l = __ldg(lut + (int64_t)c * width + ((int)kf + half_width));
        else l = eb_likelihood_slow(p, c, v, lik_bound);          // outside the table (also NaN): direct evaluation
        if (bits) bit_acc -= log2f(l);
    };
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (kVec) {
        int64_t n4 = n >> 2;
        for (int64_t q = tid; q < n4; q += stride) {
            float4 xv = ldg_stream(reinterpret_cast<const float4 *>(x) + q);
            float4 v, l;
            one(4 * q + 0, xv.x, v.x, l.x);
            one(4 * q + 1, xv.y, v.y, l.y);
            one(4 * q + 2, xv.z, v.z, l.z);
            one(4 * q + 3, xv.w, v.w, l.w);
            reinterpret_cast<float4 *>(x_hat)[q] = v;
            reinterpret_cast<float4 *>(lik)[q] = l;
            if (x_hat_bf16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                reinterpret_cast<uint2 *>(x_hat_bf16)[q] = make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
            }
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride) {
            float v, l;
            one(i, x[i], v, l);
            x_hat[i] = v; lik[i] = l;
            if (x_hat_bf16) x_hat_bf16[i] = __float2bfloat16_rn(v);
        }
    } else {
        for (int64_t i = tid; i < n; i += stride) {
            float v, l;
            one(i, x[i], v, l);
            x_hat[i] = v; lik[i] = l;
            if (x_hat_bf16) x_hat_bf16[i] = __float2bfloat16_rn(v);
        }
    }
    if (bits) block_atomic_add(bit_acc, red, bits);
}

static int64_t gcd64(int64_t a, int64_t b) { while (b) { int64_t t = a % b; a = b; b = t; } return a; }

static int launch_eb(int op, const float *x, const float *noise, const mmc_eb_params *params, float lik_bound,
                     int64_t outer, int64_t C, int64_t inner, float *x_hat, void *x_hat_bf16, float *lik,
                     float *bits, cudaStream_t st, const char *name)
{
    MMC_CHECK_ARG(params != nullptr, "%s: params is NULL", name);
    MMC_CHECK_ARG(outer >= 0 && C >= 1 && inner >= 1 && C <= 65535, "%s: bad shape outer=%lld C=%lld inner=%lld", name,
                  (long long)outer, (long long)C, (long long)inner);
    for (int k = 0; k < 5; ++k) MMC_CHECK_ARG(params->matrix[k] && params->bias[k], "%s: NULL parameter block", name);
    for (int k = 0; k < 4; ++k) MMC_CHECK_ARG(params->factor[k], "%s: NULL factor block", name);
    int64_t n = outer * C * inner;
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && lik && (op == 1 || x_hat), "%s: NULL buffer", name);
    if (op == 0) MMC_CHECK_ARG(noise != nullptr || params->medians != nullptr, "%s: medians is NULL", name);
    __nv_bfloat16 *xb = (__nv_bfloat16 *)x_hat_bf16;
    if (inner == 1) {
        // total threads T must be a multiple of C: T = unit * m with unit = lcm(kBlock, C)
        int64_t unit_blocks = C / gcd64(C, kBlock);
        int64_t want = elementwise_grid(n, kBlock, 4);
        int64_t m = (want + unit_blocks - 1) / unit_blocks;
        if (m < 1) m = 1;
        int grid = (int)(m * unit_blocks);
        if (op == 0)
            eb_kernel<true, 0><<<grid, kBlock, 0, st>>>(x, noise, *params, lik_bound, outer, (uint32_t)C, inner, x_hat, xb, lik, bits);
        else
            eb_kernel<true, 1><<<grid, kBlock, 0, st>>>(x, noise, *params, lik_bound, outer, (uint32_t)C, inner, x_hat, xb, lik, bits);
    } else {
        int64_t per_chan = outer * inner;
        int64_t gx = (per_chan + kBlock * 4 - 1) / (kBlock * 4);
        int64_t cap = ((int64_t)kNumSMs * 8 + C - 1) / C;
        if (gx > cap) gx = cap;
        if (gx < 1) gx = 1;
        dim3 grid((unsigned)gx, (unsigned)C);
        if (op == 0)
            eb_kernel<false, 0><<<grid, kBlock, 0, st>>>(x, noise, *params, lik_bound, outer, (uint32_t)C, inner, x_hat, xb, lik, bits);
        else
            eb_kernel<false, 1><<<grid, kBlock, 0, st>>>(x, noise, *params, lik_bound, outer, (uint32_t)C, inner, x_hat, xb, lik, bits);
    }
    MMC_CHECK_LAUNCH(name);
    return MMC_OK;
}

static int launch_eb_bwd(const float *x, const float *noise, const float *g, const mmc_eb_params *params, float lik_bound, int64_t outer,
                         int64_t C, int64_t inner, float *dx, float *dparams, cudaStream_t st)
{
    const char *name = "mmc_eb_backward";
    MMC_CHECK_ARG(params != nullptr, "%s: params is NULL", name);
    MMC_CHECK_ARG(outer >= 0 && C >= 1 && inner >= 1 && C <= 65535, "%s: bad shape", name);
    for (int k = 0; k < 5; ++k) MMC_CHECK_ARG(params->matrix[k] && params->bias[k], "%s: NULL parameter block", name);
    for (int k = 0; k < 4; ++k) MMC_CHECK_ARG(params->factor[k], "%s: NULL factor block", name);
    int64_t n = outer * C * inner;
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && noise && g && dx && dparams, "%s: NULL buffer (the backward exists for the noise-mode forward only)", name);
    if (inner == 1) {
        int64_t unit_blocks = C / gcd64(C, kBlock);
        int64_t want = (n + (int64_t)kBlock * 16 - 1) / ((int64_t)kBlock * 16);     // ~16 elements per thread: few atomics
        int64_t m = (want + unit_blocks - 1) / unit_blocks;
        if (m < 1) m = 1;
        eb_bwd_kernel<true><<<(int)(m * unit_blocks), kBlock, 0, st>>>(x, noise, g, *params, lik_bound, outer, (uint32_t)C, inner, dx, dparams);
    } else {
        int64_t per_chan = outer * inner;
        int64_t gx = (per_chan + kBlock * 8 - 1) / (kBlock * 8);
        int64_t cap = ((int64_t)kNumSMs * 8 + C - 1) / C;
        if (gx > cap) gx = cap;
        if (gx < 1) gx = 1;
        eb_bwd_kernel<false><<<dim3((unsigned)gx, (unsigned)C), kBlock, 0, st>>>(x, noise, g, *params, lik_bound, outer, (uint32_t)C, inner, dx, dparams);
    }
    MMC_CHECK_LAUNCH(name);
    return MMC_OK;
}

// -------------------------------------------------------------------------------------------
// GaussianConditional
// -------------------------------------------------------------------------------------------
// erfc with fractional error < 1.2e-7 everywhere (Chebyshev fit of Numerical Recipes' erfcc: t * exp(-z^2 + P(t)),
// t = 1 / (1 + z/2)): 10 FMAs, one MUFU.RCP and one MUFU.EX2 instead of the ~45-instruction erfcf().  The likelihood
// is a difference of two such values of magnitude <= 1/2, so this sits inside the cancellation noise (4 ulp of 1/2)
// that the reference's own fp32 result carries; the kernel then streams at HBM speed instead of being ALU-bound.
// MUFU.RCP alone (1 ulp): __frcp_rn() adds a Newton step and a CALL to a denormal slow path, which made the likelihood kernel
// ALU-bound; every use below feeds an approximation whose own error is larger (operands are >= 0.11 resp. >= 1)
__device__ __forceinline__ float rcp_fast(float v)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float erfc_fast(float x)
{
    const float z = fabsf(x);
    const float t = rcp_fast(fmaf(0.5f, z, 1.0f));
    float p = 0.17087277f;
    p = fmaf(p, t, -0.82215223f);
    p = fmaf(p, t, 1.48851587f);
    p = fmaf(p, t, -1.13520398f);
    p = fmaf(p, t, 0.27886807f);
    p = fmaf(p, t, -0.18628806f);
    p = fmaf(p, t, 0.09678418f);
    p = fmaf(p, t, 0.37409196f);
    p = fmaf(p, t, 1.00002368f);
    p = fmaf(p, t, -1.26551223f);
    const float r = t * __expf(fmaf(-z, z, p));
    return x >= 0.0f ? r : 2.0f - r;
}

__device__ __forceinline__ float std_cumulative(float t)
{
    const float c = -0.70710678118654752440f;  // float(-(2**-0.5)), entropy_models.py:631
    return 0.5f * erfc_fast(c * t);
}

template <bool kMeans, bool kNoise>
__device__ __forceinline__ void gc_one(float xv, float sv, float mv, float nv, float scale_bound, float lik_bound,
                                       float &v, float &l)
{
    if (kNoise) {
        v = __fadd_rn(xv, nv);
    } else if (kMeans) {
        v = __fadd_rn(rintf(__fsub_rn(xv, mv)), mv);
    } else {
        v = rintf(xv);
    }
    float val = kMeans ? __fsub_rn(v, mv) : v;
    float a = fabsf(val);
    float s = lower_bound_f(sv, scale_bound);
    const float inv_s = rcp_fast(s);             // one reciprocal (MUFU) for both CDF arguments
    if (s >= 8.0f) {
        // Wide Gaussians: Phi(u) - Phi(l) is a difference of two numbers near 1/2 (the reference's fp32 result carries
        // ~2.4e-7 of cancellation noise there).  Integrate the density over the bin instead:
        //   h phi(m) [1 + h^2 (m^2 - 1)/24 + h^4 (m^4 - 6 m^2 + 3)/1920],  m = -a/s (bin centre), h = 1/s <= 1/8
        // (truncation < 1e-6 relative for every m whose likelihood is above the 1e-9 floor).
        const float m = -a * inv_s, m2 = m * m, h2 = inv_s * inv_s;
        const float phi = 0.3989422804014327f * __expf(-0.5f * m2);
        const float c2 = (m2 - 1.0f) * (1.0f / 24.0f);
        const float c4 = fmaf(m2, m2 - 6.0f, 3.0f) * (1.0f / 1920.0f);
        l = inv_s * phi * fmaf(h2, fmaf(h2, c4, c2), 1.0f);
    } else {
        float upper = std_cumulative((0.5f - a) * inv_s);
        float lower = std_cumulative((-0.5f - a) * inv_s);
        l = upper - lower;
    }
    if (lik_bound > 0.0f) l = lower_bound_f(l, lik_bound);
}

template <bool kMeans, bool kNoise, bool kVec>
__global__ void __launch_bounds__(kBlock) gc_kernel(const float *__restrict__ x, const float *__restrict__ scales,
                                                    const float *__restrict__ means, const float *__restrict__ noise,
                                                    float scale_bound, float lik_bound, int64_t n,
                                                    float *__restrict__ x_hat, __nv_bfloat16 *__restrict__ x_hat_bf16,
                                                    float *__restrict__ lik, float *__restrict__ bits)
{
    __shared__ float red[32];
    float bit_acc = 0.0f;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    auto scalar = [&](int64_t i) {
        float v, l;
        gc_one<kMeans, kNoise>(x[i], scales[i], kMeans ? means[i] : 0.f, kNoise ? noise[i] : 0.f, scale_bound, lik_bound, v, l);
        x_hat[i] = v;
        if (x_hat_bf16) x_hat_bf16[i] = __float2bfloat16_rn(v);
        lik[i] = l;
        if (bits) bit_acc -= __log2f(l);
    };
    if (kVec) {
        int64_t n4 = n >> 2;
        for (int64_t q = tid; q < n4; q += stride) {
            float4 xv = ldg_stream(reinterpret_cast<const float4 *>(x) + q);
            float4 sv = ldg_stream(reinterpret_cast<const float4 *>(scales) + q);
            float4 mv = make_float4(0, 0, 0, 0), nv = make_float4(0, 0, 0, 0);
            if (kMeans) mv = ldg_stream(reinterpret_cast<const float4 *>(means) + q);
            if (kNoise) nv = ldg_stream(reinterpret_cast<const float4 *>(noise) + q);
            float4 v, l;
            gc_one<kMeans, kNoise>(xv.x, sv.x, mv.x, nv.x, scale_bound, lik_bound, v.x, l.x);
            gc_one<kMeans, kNoise>(xv.y, sv.y, mv.y, nv.y, scale_bound, lik_bound, v.y, l.y);
            gc_one<kMeans, kNoise>(xv.z, sv.z, mv.z, nv.z, scale_bound, lik_bound, v.z, l.z);
            gc_one<kMeans, kNoise>(xv.w, sv.w, mv.w, nv.w, scale_bound, lik_bound, v.w, l.w);
            reinterpret_cast<float4 *>(x_hat)[q] = v;
            reinterpret_cast<float4 *>(lik)[q] = l;
            if (x_hat_bf16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                uint2 pk = make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
                reinterpret_cast<uint2 *>(x_hat_bf16)[q] = pk;
            }
            if (bits) bit_acc -= (__log2f(l.x) + __log2f(l.y)) + (__log2f(l.z) + __log2f(l.w));
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride) scalar(i);
    } else {
        for (int64_t i = tid; i < n; i += stride) scalar(i);
    }
    if (bits) block_atomic_add(bit_acc, red, bits);
}

template <bool kVec>
__global__ void __launch_bounds__(kBlock) bits_kernel(const float *__restrict__ lik, int64_t n, float *__restrict__ bits)
{
    __shared__ float red[32];
    float acc = 0.0f;
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (kVec) {
        int64_t n4 = n >> 2;
        for (int64_t q = tid; q < n4; q += stride) {
            float4 l = ldg_stream(reinterpret_cast<const float4 *>(lik) + q);
            acc -= (log2f(l.x) + log2f(l.y)) + (log2f(l.z) + log2f(l.w));
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride) acc -= log2f(lik[i]);
    } else {
        for (int64_t i = tid; i < n; i += stride) acc -= log2f(lik[i]);
    }
    block_atomic_add(acc, red, bits);
}

// -------------------------------------------------------------------------------------------
// pmf_to_quantized_cdf on the device (compressai/cpp_exts/ops/ops.cpp:40-109 via entropy_models.py:206-214): one CTA per
// table row.  Integer algorithm, bit-exact with the reference: round(pmf * 2^p) frequencies (the tail mass is the last
// symbol), renormalised prefix sums, then every zero-frequency symbol steals one count from the cheapest donor (first index
// among equals).  The scan over symbols is inherently sequential; the donor search and the shifts run across the CTA.
// -------------------------------------------------------------------------------------------
constexpr int kCdfThreads = 256;

__global__ void __launch_bounds__(kCdfThreads) pmf_to_cdf_kernel(const float *__restrict__ pmf, int64_t pmf_pitch, const float *__restrict__ tail_mass,
                                                                 const int32_t *__restrict__ pmf_length, int max_len, int precision,
                                                                 int32_t *__restrict__ cdf_out, int32_t *__restrict__ status)
{
    extern __shared__ uint32_t c[];                 // [max_len + 2]
    __shared__ unsigned long long red_key[kCdfThreads / 32];
    __shared__ uint32_t total_s;
    __shared__ int bad_s;
    const int row = blockIdx.x, tid = threadIdx.x;
    const int n = pmf_length[row] + 1;              // symbols incl. the tail-mass symbol
    const float scale = (float)(1u << precision);
    if (tid == 0) { total_s = 0; bad_s = 0; c[0] = 0; }
    __syncthreads();
    uint32_t part = 0;
    for (int i = tid; i < n; i += kCdfThreads) {
        const float p = (i < n - 1) ? pmf[row * pmf_pitch + i] : tail_mass[row];
        if (!(p >= 0.0f) || isinf(p)) bad_s = 1;
        const uint32_t f = (uint32_t)roundf(p * scale);
        c[i + 1] = f;
        part += f;
    }
    atomicAdd(&total_s, part);
    __syncthreads();
    int32_t *out = cdf_out + (int64_t)row * (max_len + 2);
    if (bad_s || total_s == 0) {
        if (tid == 0) status[row] = MMC_EDOMAIN;
        for (int i = tid; i < max_len + 2; i += kCdfThreads) out[i] = 0;
        return;
    }
    if (tid == 0) {
        const uint64_t one = (uint64_t)1 << precision;
        const uint32_t total = total_s;
        uint32_t run = 0;
        for (int i = 0; i <= n; ++i) {
            run += (uint32_t)((one * c[i]) / total);
            c[i] = run;
        }
        c[n] = 1u << precision;
    }
    __syncthreads();
    int fail = 0;
    for (int i = 0; i < n; ++i) {
        if (c[i] != c[i + 1]) continue;               // uniform: everybody reads the same shared values
        // cheapest donor: smallest frequency > 1, first index among equals  ->  min over (freq << 32 | index)
        unsigned long long best = ~0ull;
        for (int j = tid; j < n; j += kCdfThreads) {
            const uint32_t f = c[j + 1] - c[j];
            if (f > 1) {
                const unsigned long long key = ((unsigned long long)f << 32) | (uint32_t)j;
                if (key < best) best = key;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            if (other < best) best = other;
        }
        if ((tid & 31) == 0) red_key[tid >> 5] = best;
        __syncthreads();
        best = red_key[0];
#pragma unroll
        for (int w = 1; w < kCdfThreads / 32; ++w)
            if (red_key[w] < best) best = red_key[w];
        if (best == ~0ull) { fail = 1; break; }
        const int donor = (int)(uint32_t)best;
        if (donor < i) { for (int j = donor + 1 + tid; j <= i; j += kCdfThreads) c[j]--; }
        else           { for (int j = i + 1 + tid; j <= donor; j += kCdfThreads) c[j]++; }
        __syncthreads();
    }
    if (tid == 0) status[row] = fail ? MMC_EDOMAIN : MMC_OK;
    for (int i = tid; i < max_len + 2; i += kCdfThreads) out[i] = (i <= n) ? (int32_t)c[i] : 0;
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_quantize_symbols(const float *x, const float *means, int means_mode, int64_t outer, int64_t C, int64_t inner,
                         int32_t *out, void *stream)
{
    return launch_quantize<0>(x, means, means_mode, outer, C, inner, out, (cudaStream_t)stream, "mmc_quantize_symbols");
}

int mmc_quantize_dequantize(const float *x, const float *means, int means_mode, int64_t outer, int64_t C, int64_t inner,
                            float *out, void *stream)
{
    return launch_quantize<1>(x, means, means_mode, outer, C, inner, out, (cudaStream_t)stream, "mmc_quantize_dequantize");
}

int mmc_dequantize(const int32_t *symbols, const float *means, int means_mode, int64_t outer, int64_t C, int64_t inner,
                   float *out, void *stream)
{
    return launch_quantize<2>(symbols, means, means_mode, outer, C, inner, out, (cudaStream_t)stream, "mmc_dequantize");
}

int mmc_quantize_noise(const float *x, const float *noise, int64_t n, float *out, void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_quantize_noise: n < 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && noise && out, "mmc_quantize_noise: NULL buffer");
    add_kernel<<<elementwise_grid(n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(x, noise, n, out);
    MMC_CHECK_LAUNCH("mmc_quantize_noise");
    return MMC_OK;
}

int mmc_lower_bound(const float *x, float bound, int64_t n, float *out, void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_lower_bound: n < 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && out, "mmc_lower_bound: NULL buffer");
    lower_bound_kernel<<<elementwise_grid(n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(x, bound, n, out);
    MMC_CHECK_LAUNCH("mmc_lower_bound");
    return MMC_OK;
}

int mmc_lower_bound_bwd(const float *x, const float *grad_out, float bound, int64_t n, float *grad_in, void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_lower_bound_bwd: n < 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && grad_out && grad_in, "mmc_lower_bound_bwd: NULL buffer");
    lower_bound_bwd_kernel<<<elementwise_grid(n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(x, grad_out, bound, n, grad_in);
    MMC_CHECK_LAUNCH("mmc_lower_bound_bwd");
    return MMC_OK;
}

int mmc_build_indexes(const float *scales, const float *table, int levels, float bound, int64_t n, int32_t *out,
                      void *stream)
{
    MMC_CHECK_ARG(levels >= 1 && levels <= 256, "mmc_build_indexes: levels=%d outside [1,256]", levels);
    MMC_CHECK_ARG(n >= 0, "mmc_build_indexes: n < 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(scales && table && out, "mmc_build_indexes: NULL buffer");
    int grid = elementwise_grid((n + 15) / 16, kBlock);      // 4 float4 per thread and iteration
    if (aligned16(scales) && aligned16(out))
        build_indexes_kernel<true><<<grid, kBlock, 0, (cudaStream_t)stream>>>(scales, table, levels, bound, n, out);
    else
        build_indexes_kernel<false><<<grid, kBlock, 0, (cudaStream_t)stream>>>(scales, table, levels, bound, n, out);
    MMC_CHECK_LAUNCH("mmc_build_indexes");
    return MMC_OK;
}

int mmc_channel_indexes(int64_t outer, int64_t C, int64_t inner, int32_t *out, void *stream)
{
    MMC_CHECK_ARG(outer >= 0 && C >= 1 && inner >= 1 && C < (1ll << 31) && inner < (1ll << 31),
                  "mmc_channel_indexes: bad shape");
    int64_t n = outer * C * inner;
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(out, "mmc_channel_indexes: NULL buffer");
    channel_indexes_kernel<<<elementwise_grid(n, kBlock), kBlock, 0, (cudaStream_t)stream>>>(
        ChanIndex{(uint32_t)C, (uint32_t)inner}, n, out);
    MMC_CHECK_LAUNCH("mmc_channel_indexes");
    return MMC_OK;
}

int mmc_pmf_to_quantized_cdf(const float *pmf, int64_t pmf_pitch, const float *tail_mass, const int32_t *pmf_length, int rows, int max_len,
                             int precision, int32_t *cdf, int32_t *status, void *stream)
{
    const char *name = "mmc_pmf_to_quantized_cdf";
    MMC_CHECK_ARG(rows >= 0 && max_len >= 1 && precision >= 1 && precision <= 30 && pmf_pitch >= max_len, "%s: bad argument", name);
    if (rows == 0) return MMC_OK;
    MMC_CHECK_ARG(pmf && tail_mass && pmf_length && cdf && status, "%s: NULL buffer", name);
    const size_t smem = (size_t)(max_len + 2) * sizeof(uint32_t);
    MMC_UNSUPPORTED(smem > 200 * 1024, "%s: rows longer than %d symbols are not supported", name, 200 * 1024 / 4 - 2);
    static PerDevice<int> attr_dev;
    if (!attr_dev.cur().load(std::memory_order_relaxed)) {
        MMC_CHECK_CUDA(cudaFuncSetAttribute(pmf_to_cdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_dev.cur().store(1, std::memory_order_relaxed);
    }
    pmf_to_cdf_kernel<<<rows, kCdfThreads, smem, (cudaStream_t)stream>>>(pmf, pmf_pitch, tail_mass, pmf_length, max_len, precision, cdf, status);
    MMC_CHECK_LAUNCH(name);
    return MMC_OK;
}

int mmc_eb_backward(const float *x, const float *noise, const float *grad_lik, const mmc_eb_params *params, float likelihood_bound,
                    int64_t outer, int64_t C, int64_t inner, float *dx, float *dparams, void *stream)
{
    return launch_eb_bwd(x, noise, grad_lik, params, likelihood_bound, outer, C, inner, dx, dparams, (cudaStream_t)stream);
}

int mmc_eb_forward(const float *x, const float *noise, const mmc_eb_params *params, float likelihood_bound,
                   int64_t outer, int64_t C, int64_t inner, float *x_hat, void *x_hat_bf16, float *likelihood,
                   float *bits, void *stream)
{
    return launch_eb(0, x, noise, params, likelihood_bound, outer, C, inner, x_hat, x_hat_bf16, likelihood, bits,
                     (cudaStream_t)stream, "mmc_eb_forward");
}

int mmc_eb_build_lut(const mmc_eb_params *params, float likelihood_bound, int64_t C, int half_width, float *lut, void *stream)
{
    MMC_CHECK_ARG(params && lut && C >= 1 && C <= 65535 && half_width >= 0 && half_width <= 4096, "mmc_eb_build_lut: bad argument");
    for (int k = 0; k < 5; ++k) MMC_CHECK_ARG(params->matrix[k] && params->bias[k], "mmc_eb_build_lut: NULL parameter block");
    for (int k = 0; k < 4; ++k) MMC_CHECK_ARG(params->factor[k], "mmc_eb_build_lut: NULL factor block");
    MMC_CHECK_ARG(params->medians, "mmc_eb_build_lut: medians is NULL");
    int64_t n = C * (2 * half_width + 1);
    eb_build_lut_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(*params, likelihood_bound, (uint32_t)C,
                                                                                              half_width, lut);
    MMC_CHECK_LAUNCH("mmc_eb_build_lut");
    return MMC_OK;
}

int mmc_eb_forward_lut(const float *x, const mmc_eb_params *params, const float *lut, int half_width, float likelihood_bound,
                       int64_t outer, int64_t C, int64_t inner, float *x_hat, void *x_hat_bf16, float *likelihood, float *bits,
                       void *stream)
{
    const char *name = "mmc_eb_forward_lut";
    MMC_CHECK_ARG(params && lut && half_width >= 0, "%s: bad argument", name);
    MMC_CHECK_ARG(outer >= 0 && C >= 1 && inner >= 1 && C <= 65535 && inner < (1ll << 31), "%s: bad shape", name);
    for (int k = 0; k < 5; ++k) MMC_CHECK_ARG(params->matrix[k] && params->bias[k], "%s: NULL parameter block", name);
    for (int k = 0; k < 4; ++k) MMC_CHECK_ARG(params->factor[k], "%s: NULL factor block", name);
    MMC_CHECK_ARG(params->medians, "%s: medians is NULL", name);
    int64_t n = outer * C * inner;
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && x_hat && likelihood, "%s: NULL buffer", name);
    ChanIndex ci{(uint32_t)C, (uint32_t)inner};
    __nv_bfloat16 *xb = (__nv_bfloat16 *)x_hat_bf16;
    bool vec = aligned16(x) && aligned16(x_hat) && aligned16(likelihood) && (!xb || (reinterpret_cast<uintptr_t>(xb) & 7u) == 0);
    int grid = elementwise_grid((n + 3) / 4, kBlock);
    if (vec)
        eb_lut_kernel<true><<<grid, kBlock, 0, (cudaStream_t)stream>>>(x, *params, lut, half_width, likelihood_bound, ci, n, x_hat, xb, likelihood, bits);
    else
        eb_lut_kernel<false><<<grid, kBlock, 0, (cudaStream_t)stream>>>(x, *params, lut, half_width, likelihood_bound, ci, n, x_hat, xb, likelihood, bits);
    MMC_CHECK_LAUNCH(name);
    return MMC_OK;
}

int mmc_eb_logits_cumulative(const float *x, const mmc_eb_params *params, int64_t outer, int64_t C, int64_t inner,
                             float *logits, void *stream)
{
    return launch_eb(1, x, nullptr, params, 0.0f, outer, C, inner, nullptr, nullptr, logits, nullptr,
                     (cudaStream_t)stream, "mmc_eb_logits_cumulative");
}

int mmc_gc_forward(const float *x, const float *scales, const float *means, const float *noise, float scale_bound,
                   float likelihood_bound, int64_t n, float *x_hat, void *x_hat_bf16, float *likelihood, float *bits,
                   void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_gc_forward: n < 0");
    MMC_CHECK_ARG(scale_bound > 0.0f, "mmc_gc_forward: scale_bound must be > 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(x && scales && x_hat && likelihood, "mmc_gc_forward: NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16 *xb = (__nv_bfloat16 *)x_hat_bf16;
    bool vec = aligned16(x) && aligned16(scales) && aligned16(x_hat) && aligned16(likelihood) &&
               (!means || aligned16(means)) && (!noise || aligned16(noise)) && (!xb || (reinterpret_cast<uintptr_t>(xb) & 7u) == 0);
    int grid = elementwise_grid((n + 3) / 4, kBlock);
#define MMC_GC(M, N, V) gc_kernel<M, N, V><<<grid, kBlock, 0, st>>>(x, scales, means, noise, scale_bound, likelihood_bound, n, x_hat, xb, likelihood, bits)
    if (vec) {
        if (means) { if (noise) MMC_GC(true, true, true); else MMC_GC(true, false, true); }
        else       { if (noise) MMC_GC(false, true, true); else MMC_GC(false, false, true); }
    } else {
        if (means) { if (noise) MMC_GC(true, true, false); else MMC_GC(true, false, false); }
        else       { if (noise) MMC_GC(false, true, false); else MMC_GC(false, false, false); }
    }
#undef MMC_GC
    MMC_CHECK_LAUNCH("mmc_gc_forward");
    return MMC_OK;
}

int mmc_bits(const float *likelihood, int64_t n, float *bits, void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_bits: n < 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(likelihood && bits, "mmc_bits: NULL buffer");
    int grid = elementwise_grid((n + 3) / 4, kBlock, 4);
    if (aligned16(likelihood))
        bits_kernel<true><<<grid, kBlock, 0, (cudaStream_t)stream>>>(likelihood, n, bits);
    else
        bits_kernel<false><<<grid, kBlock, 0, (cudaStream_t)stream>>>(likelihood, n, bits);
    MMC_CHECK_LAUNCH("mmc_bits");
    return MMC_OK;
}

}  // extern "C"
