/*
 * mmcodec.h -- C ABI of libmmcodec.so, the B200 (sm_100a) implementation of the learned-codec
 * hot path of SZU-AdvTech-2022/165 (a CompressAI 1.2.0.dev0 fork).
 *
 * The reference has no FFI for this path: the path is Python modules calling stock torch ops
 * (SURVEY.md section 8b).  Each entry point below therefore replaces one reference *function*;
 * the citation (file:line, relative to /root/reference/CompressAI) is the interface it stands in
 * for.  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer owned by the caller unless the name ends in _host.
 *    The library borrows buffers for the duration of the enqueued work (stream ordered) and
 *    never allocates persistent device memory.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *  - Return value: MMC_OK or a negative error code; mmc_last_error() gives the thread-local
 *    message.  Nothing aborts, nothing falls back to the CPU.
 *  - Elementwise entropy-stage tensors are viewed as [outer][C][inner]:
 *      NCHW contiguous  : outer = N,       inner = H*W
 *      channels-last    : outer = N*H*W,   inner = 1
 *  - Activations between transform layers are NHWC bf16 ("channels-last"); fp32 appears at the
 *    API edges (image in, latents / likelihoods / reconstruction out).
 */
#ifndef MMCODEC_H
#define MMCODEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMC_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define MMC_API __attribute__((visibility("default")))
#else
#define MMC_API
#endif

enum mmc_status {
    MMC_OK = 0,
    MMC_EINVAL = -1,       /* bad shape / alignment / argument -> ValueError on the Python side */
    MMC_ECUDA = -2,        /* CUDA launch or driver error      -> RuntimeError */
    MMC_EUNSUPPORTED = -3, /* valid in the reference, not implemented here (fails loudly) */
    MMC_EDOMAIN = -4       /* pmf_to_quantized_cdf domain error -> ValueError (ops.cpp:46-52) */
};

enum mmc_means_mode { MMC_MEANS_NONE = 0, MMC_MEANS_FULL = 1, MMC_MEANS_PER_CHANNEL = 2 };

enum mmc_dtype { MMC_F32 = 0, MMC_BF16 = 1 };
enum mmc_layout {
    MMC_NCHW = 0,
    MMC_NHWC = 1,
    MMC_NHWC_PAD8 = 2 /* bf16 [B][Hp][Wp][8]: image with <= 8 channels, zero-padded by k/2 pixels and to 8
                         channels (see mmc_conv_pad8_size / mmc_pad_nchw_to_nhwc8); tensor-core input of the
                         image-edge convolutions (g_a.0: 3 -> N, depth branch: 1 -> N) */
};

/* activation fused after bias (and before GDN where both are given) */
enum mmc_act {
    MMC_ACT_NONE = 0,
    MMC_ACT_RELU = 1,       /* nn.ReLU        models/google.py:256,264 */
    MMC_ACT_LEAKY_RELU = 2, /* nn.LeakyReLU() models/google.py:365,373 (slope 0.01) */
    MMC_ACT_ABS = 3,        /* torch.abs(y)   models/google.py:283 (only for the secondary output) */
    MMC_ACT_QRELU8 = 4      /* QReLU(bit_depth=8) forward: clamp(v, 0, 255)   layers/layers.py:268-277, used by the
                               ssf2020 scale hyper-decoder, models/video/google.py:128-148 */
};

enum mmc_gdn_mode { MMC_GDN_NONE = 0, MMC_GDN_FORWARD = 1, MMC_GDN_INVERSE = 2 };

MMC_API int mmc_version(void);
MMC_API const char *mmc_last_error(void);
/* Number of kernels this library has launched on the calling thread since the last reset
 * (bench.py's "gpu_launches"). */
MMC_API int64_t mmc_launch_count(void);
MMC_API void mmc_reset_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Entropy stage (HBM-bound elementwise kernels)
 * ------------------------------------------------------------------------------------------- */

/* EntropyModel.quantize(mode="symbols")   compressai/entropy_models/entropy_models.py:157-182
 * out = int32(rint(x - mean)).  Bit-exact w.r.t. the reference for |x - mean| < 2^31. */
MMC_API int mmc_quantize_symbols(const float *x, const float *means, int means_mode, int64_t outer, int64_t C,
                         int64_t inner, int32_t *out, void *stream);

/* EntropyModel.quantize(mode="dequantize")   entropy_models.py:169-178.  out = rint(x - mean) + mean */
MMC_API int mmc_quantize_dequantize(const float *x, const float *means, int means_mode, int64_t outer, int64_t C,
                            int64_t inner, float *out, void *stream);

/* EntropyModel.quantize(mode="noise")   entropy_models.py:163-167.  out = x + noise; the U(-1/2,1/2)
 * tensor is drawn by the caller (torch generator) so both implementations consume the same noise. */
MMC_API int mmc_quantize_noise(const float *x, const float *noise, int64_t n, float *out, void *stream);

/* EntropyModel.dequantize   entropy_models.py:190-199.  out = float(sym) + mean */
MMC_API int mmc_dequantize(const int32_t *symbols, const float *means, int means_mode, int64_t outer, int64_t C,
                   int64_t inner, float *out, void *stream);

/* LowerBound forward / backward   compressai/ops/bound_ops.py:36-42 */
MMC_API int mmc_lower_bound(const float *x, float bound, int64_t n, float *out, void *stream);
MMC_API int mmc_lower_bound_bwd(const float *x, const float *grad_out, float bound, int64_t n, float *grad_in,
                        void *stream);

/* GaussianConditional.build_indexes   entropy_models.py:735-740
 * out = (levels-1) - #{i < levels-1 : max(scale, bound) <= table[i]}.  `table` is the fp32
 * scale_table BUFFER (never regenerated); levels <= 256. */
MMC_API int mmc_build_indexes(const float *scales, const float *table, int levels, float bound, int64_t n,
                      int32_t *out, void *stream);

/* EntropyBottleneck._build_indexes   entropy_models.py:542-553.  out[o][c][i] = c */
MMC_API int mmc_channel_indexes(int64_t outer, int64_t C, int64_t inner, int32_t *out, void *stream);

/* Raw EntropyBottleneck parameters for the default filters (3,3,3,3)   entropy_models.py:361-379.
 * matrix[0] [C][3][1], matrix[1..3] [C][3][3], matrix[4] [C][1][3]; bias[0..3] [C][3][1],
 * bias[4] [C][1][1]; factor[0..3] [C][3][1]; medians = quantiles[:,0,1] gathered to [C]. */
typedef struct mmc_eb_params {
    const float *matrix[5];
    const float *bias[5];
    const float *factor[4];
    const float *medians;
} mmc_eb_params;

/* EntropyBottleneck.forward   entropy_models.py:495-540 (+ _logits_cumulative :457-477,
 * _likelihood :480-492, LowerBound(1e-9) :527).  noise == NULL: eval (x_hat = rint(x-med)+med);
 * noise != NULL: training (x_hat = x + noise).  Optional outputs (may be NULL):
 *   x_hat_bf16 : bf16 copy of x_hat (feeds the synthesis transform),
 *   bits       : one float, += -sum(log2(likelihood)) over this call (atomicAdd; caller zeroes). */
MMC_API int mmc_eb_forward(const float *x, const float *noise, const mmc_eb_params *params, float likelihood_bound,
                   int64_t outer, int64_t C, int64_t inner, float *x_hat, void *x_hat_bf16,
                   float *likelihood, float *bits, void *stream);

/* Eval-mode fast path of EntropyBottleneck.forward.  In eval mode x_hat = rint(x - median) + median, so the likelihood
 * depends only on (channel, integer symbol) -- the observation update() uses for the CDF tables (entropy_models.py:422-432).
 * mmc_eb_build_lut tabulates bound(likelihood(k + median_c)) for |k| <= half_width into lut[C][2*half_width+1] (caller-owned
 * device buffer, rebuilt when the parameters change); mmc_eb_forward_lut is mmc_eb_forward(noise = NULL) reading that table
 * (symbols outside it are evaluated directly), which turns the SFU-bound kernel into an HBM-bound one. */
MMC_API int mmc_eb_build_lut(const mmc_eb_params *params, float likelihood_bound, int64_t C, int half_width, float *lut,
                             void *stream);
MMC_API int mmc_eb_forward_lut(const float *x, const mmc_eb_params *params, const float *lut, int half_width,
                               float likelihood_bound, int64_t outer, int64_t C, int64_t inner, float *x_hat,
                               void *x_hat_bf16, float *likelihood, float *bits, void *stream);

/* EntropyBottleneck._logits_cumulative   entropy_models.py:457-477 (used by update() and loss()) */
MMC_API int mmc_eb_logits_cumulative(const float *x, const mmc_eb_params *params, int64_t outer, int64_t C,
                             int64_t inner, float *logits, void *stream);

/* GaussianConditional.forward   entropy_models.py:715-731 (+ _likelihood :692-709).
 * means may be NULL; noise as above.  scales/means are fp32 with the same shape as x.
 * Optional outputs as for mmc_eb_forward. */
MMC_API int mmc_gc_forward(const float *x, const float *scales, const float *means, const float *noise,
                   float scale_bound, float likelihood_bound, int64_t n, float *x_hat, void *x_hat_bf16,
                   float *likelihood, float *bits, void *stream);

/* sum(log2(likelihood)) * -1 accumulated into *bits (atomicAdd); examples/train.py:74-77 */
MMC_API int mmc_bits(const float *likelihood, int64_t n, float *bits, void *stream);

/* pmf_to_quantized_cdf   compressai/cpp_exts/ops/ops.cpp:40-109.  HOST function (called from
 * update() once per model).  cdf_host has n+1 entries.  MMC_EDOMAIN on negative / non-finite. */
MMC_API int mmc_pmf_to_quantized_cdf_host(const float *pmf_host, int n, int precision, uint32_t *cdf_host);

/* rANS byte coder, bitstream-compatible with compressai.ans.RansEncoder.encode_with_indexes /
 * RansDecoder.decode_with_indexes   compressai/cpp_exts/rans/rans_interface.cpp:108-284 (one serial 64-bit rANS
 * state per image, 16-bit precision, 4-bit bypass nibbles).  HOST functions on flat int32 buffers; the images of a
 * batch are coded on parallel host threads.
 *   symbols / indexes : [batch][n] int32 (host)      cdfs : [n_cdfs][cdf_stride] int32, row i valid for cdf_sizes[i]
 *   encode: stream b is written at out + b * cap_per_stream, its length to nbytes[b]; MMC_EINVAL (with the required
 *           lengths in nbytes) when cap_per_stream is too small or out is NULL.
 *   decode: stream b is nbytes[b] bytes at streams + stream_offsets[b]. */
MMC_API int mmc_rans_encode_batch_host(const int32_t *symbols, const int32_t *indexes, int batch, int64_t n,
                                       const int32_t *cdfs, int n_cdfs, int cdf_stride, const int32_t *cdf_sizes,
                                       const int32_t *offsets, uint8_t *out, size_t cap_per_stream, size_t *nbytes);
/* Test hook: checks the encoder's reciprocal-multiply state update against the division of the reference coder
 * (third_party/ryg_rans/rans64.h Rans64EncPut) for every frequency 1..2^16 on edge and pseudo-random states; returns the
 * number of mismatches (0). */
MMC_API int64_t mmc_rans_selftest(void);

MMC_API int mmc_rans_decode_batch_host(const uint8_t *streams, const size_t *stream_offsets, const size_t *nbytes,
                                       const int32_t *indexes, int batch, int64_t n, const int32_t *cdfs, int n_cdfs,
                                       int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                                       int32_t *symbols_out);

/* ---------------------------------------------------------------------------------------------
 * GDN   compressai/layers/gdn.py:77-92, compressai/ops/parametrizers.py:47-64
 * ------------------------------------------------------------------------------------------- */

/* beta_eff = max(beta, beta_bound)^2 - pedestal ; gamma_eff likewise ([C][C], row i = output
 * channel).  gamma_eff_bf16 (optional) is the same matrix in bf16 for the fused conv epilogue. */
MMC_API int mmc_gdn_reparam(const float *beta, const float *gamma, int C, float beta_bound, float gamma_bound,
                    float pedestal, float *beta_eff, float *gamma_eff, void *gamma_eff_bf16, void *stream);

/* Stand-alone GDN.forward on fp32 data: y = x * rsqrt(beta_i + sum_j gamma_ij x_j^2) (inverse: sqrt).
 * layout: MMC_NCHW ([B][C][HW]) or MMC_NHWC ([B*HW][C]). */
MMC_API int mmc_gdn_forward(const float *x, const float *beta_eff, const float *gamma_eff, int inverse, int64_t B,
                    int C, int64_t HW, int layout, float *y, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Transforms: conv() / deconv()   compressai/models/utils.py:128-146
 * ------------------------------------------------------------------------------------------- */
typedef struct mmc_conv_desc {
    int transposed;      /* 0: nn.Conv2d(k, stride, padding=k/2); 1: nn.ConvTranspose2d(k, stride,
                            padding=k/2, output_padding=stride-1) */
    int B, H, W;         /* input batch and spatial size */
    int Cin, Cout;
    int k;               /* 1, 3 or 5 */
    int stride;          /* 1 or 2 */
    int in_dtype;        /* mmc_dtype */
    int in_layout;       /* mmc_layout */
    int out_dtype;
    int out_layout;
    int act;             /* mmc_act applied to (acc + bias) */
    int gdn;             /* mmc_gdn_mode applied after act; needs beta_eff / gamma_eff */
    int out2_bf16;       /* secondary NHWC bf16 output y2: 0 none, 1 = |out| (torch.abs(y) feeding h_a,
                            models/google.py:283), 2 = out (y feeding h_a in the mean-scale model, :381),
                            3 = the activations BEFORE the fused GDN / IGDN (saved for the backward pass) */
} mmc_conv_desc;

/* Output spatial size for a descriptor (conv: ceil(H/stride); deconv: H*stride). */
MMC_API int mmc_conv_out_size(const mmc_conv_desc *d, int *Ho, int *Wo);

/* Pack fp32 weights (Conv2d [Cout][Cin][k][k] or ConvTranspose2d [Cin][Cout][k][k]) into the bf16
 * tap-major layout [k*k][Cout][Cin_pad] the tensor-core kernels stream with TMA.
 * Returns the packed size in bytes through *bytes when w_packed is NULL. */
MMC_API int mmc_conv_pack_weights(const mmc_conv_desc *d, const float *w, void *w_packed, size_t *bytes,
                          void *stream);

/* Direct (CUDA-core, fp32 accumulate) convolution for any descriptor; reads fp32 weights in the
 * torch layout.  Used for the 3-channel image-edge layers and as the on-device cross-check of the
 * tensor-core path. */
MMC_API int mmc_conv_forward_direct(const mmc_conv_desc *d, const void *x, const float *w, const float *bias,
                            const float *beta_eff, const float *gamma_eff, void *y, void *y2, void *stream);

/* Tensor-core implicit GEMM (TMA -> smem -> tcgen05.mma -> TMEM -> fused epilogue).
 * Three shape classes (chosen from the descriptor, same choice in mmc_conv_pack_weights):
 *   - NHWC bf16 input, Cin % 8 == 0 (>= 32), Cout % 16 == 0 (<= 1024; split into N tiles <= 256), NHWC output;
 *   - NHWC_PAD8 input (Cin <= 8) forward conv, Cout % 16 == 0, NHWC output;
 *   - transposed, stride 2, Cout <= 4 (the reconstruction layer): planar fp32 NCHW output.
 * Fused GDN needs Cout in {64, 128, 192}.  Anything else returns MMC_EUNSUPPORTED. */
MMC_API int mmc_conv_forward_tc(const mmc_conv_desc *d, const void *x, const void *w_packed, const float *bias,
                        const float *beta_eff, const void *gamma_eff_bf16, void *y, void *y2, void *stream);

/* The same with the input channels coming from TWO NHWC bf16 tensors (channels [0, cin1) from x1, [cin1, Cin) from x2): the
 * torch.cat of two feature maps in front of a convolution (compressai/models/google.py:1153,1161,... `tran_conv*`,
 * `pic2_g_a_conv2..4`, `pic2_g_s_conv2..4`, `entropy_parameters`) is never materialised -- the K loop takes its channel boxes
 * from the first tensor map, then from the second.  cin1 % 64 == 0, (Cin - cin1) % 8 == 0; weights packed for the full Cin. */
MMC_API int mmc_conv_forward_tc2(const mmc_conv_desc *d, const void *x1, int cin1, const void *x2, const void *w_packed,
                                 const float *bias, const float *beta_eff, const void *gamma_eff_bf16, void *y, void *y2,
                                 void *stream);

/* Staging for image-edge convolutions on the tensor cores (Cin <= 8): the fp32 NCHW image is copied once
 * into a zero-padded 8-channel NHWC bf16 buffer [B][Hp][Wp][8]; the conv then reads each 5x5 (3x3) window row
 * as one 128-byte TMA box.  mmc_conv_pad8_size gives (Hp, Wp) for a descriptor with in_layout NHWC_PAD8. */
MMC_API int mmc_conv_pad8_size(const mmc_conv_desc *d, int *Hp, int *Wp);
MMC_API int mmc_pad_nchw_to_nhwc8(const float *x, int64_t B, int C, int H, int W, int pad, int Hp, int Wp, void *out,
                                  void *stream);

/* Layout / dtype conversion at the API edges. */
MMC_API int mmc_nchw_f32_to_nhwc_bf16(const float *x, int64_t B, int C, int64_t HW, void *y, void *stream);
MMC_API int mmc_nhwc_bf16_to_nchw_f32(const void *x, int64_t B, int C, int64_t HW, float *y, void *stream);
MMC_API int mmc_nhwc_f32_to_nchw_f32(const float *x, int64_t B, int C, int64_t HW, float *y, void *stream);
MMC_API int mmc_f32_to_bf16(const float *x, int64_t n, void *y, void *stream);
/* fp32 precision mode: [pixels][C] fp32 -> [pixels][3 C] bf16 = [hi | lo | hi] with hi = bf16(x), lo = bf16(x - hi).  Convolved
 * (mmc_conv_forward_tc) with weights whose input channels are laid out [w_hi | w_hi | w_lo] this evaluates the fp32 layer
 * (compressai/models/utils.py:128-146 on fp32 tensors) to ~1e-5 relative on the bf16 tensor cores.  C % 4 == 0;
 * square != 0 splits x^2 instead (the operand of the GDN norm contraction). */
MMC_API int mmc_split_f32_bf16x3(const float *x, int64_t pixels, int C, int square, void *y, void *stream);
/* fp32-mode GDN / IGDN, elementwise half (layers/gdn.py:88-92): y = x * rsqrt(norm) (inverse: x * sqrt(norm)); norm = beta + gamma x^2
 * comes from a 1x1 mmc_conv_forward_tc over mmc_split_f32_bf16x3(x, square = 1).  n % 4 == 0. */
MMC_API int mmc_gdn_apply_f32(const float *x, const float *norm, int inverse, int64_t n, float *y, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Scale-space flow prediction of the ssf2020 video codec (HBM-bound stencil / gather kernels, planar fp32)
 * ------------------------------------------------------------------------------------------- */

/* ScaleSpaceFlow.gaussian_volume   compressai/models/video/google.py:331-355 (+ gaussian_blur / gaussian_kernel2d,
 * compressai/models/utils.py:155-189): x [planes][H][W] -> volume [planes][num_levels + 1][H][W].  `kernel1d_host` is the
 * HOST copy of gaussian_kernel1d(ksize, sigma) (computed by the caller exactly as the reference does; ksize odd, <= 33);
 * H and W must be multiples of 2^(num_levels - 1).  Workspace size from mmc_gaussian_volume_workspace. */
MMC_API int mmc_gaussian_volume_workspace(int64_t planes, int H, int W, size_t *bytes);
MMC_API int mmc_gaussian_volume(const float *x, int64_t planes, int H, int W, const float *kernel1d_host, int ksize,
                                int num_levels, void *workspace, size_t ws_bytes, float *volume, void *stream);

/* ScaleSpaceFlow.warp_volume / forward_prediction   models/video/google.py:357-382: trilinear F.grid_sample of the
 * volume [N][C][D][H][W] at (base grid + flow, scale_field), border padding, align_corners=False.  `motion_info` is the
 * motion decoder's output [N][3][H][W] (flow x, flow y, scale field); base_x [W] / base_y [H] are the rows of
 * F.affine_grid(identity, align_corners=False) (utils.py:192-195).  Writes x_pred [N][C][H][W] and, when x_res is
 * given, the residual x_cur - x_pred (google.py:262) in the same pass. */
MMC_API int mmc_scale_space_warp(const float *volume, const float *motion_info, const float *base_x, const float *base_y,
                                 int64_t N, int C, int D, int H, int W, const float *x_cur, float *x_pred, float *x_res,
                                 void *stream);

/* Colour transforms of the video pipeline (compressai/transforms/functional.py:26-137), planar fp32:
 *   mmc_color_convert: rgb2ycbcr (to_ycbcr = 1) / ycbcr2rgb (0), ITU-R BT.709, x and y [N][3][HW];
 *   mmc_avg_pool2: F.avg_pool2d(x, 2, 2) per plane (yuv_444_to_420 chroma), input planes x_plane_stride apart;
 *   mmc_upsample2x_bilinear: F.interpolate(scale_factor=2, mode="bilinear", align_corners=False) per plane
 *   (yuv_420_to_444 chroma; also a step of the Gaussian volume), output planes y_plane_stride apart. */
MMC_API int mmc_color_convert(const float *x, int64_t N, int64_t HW, int to_ycbcr, float *y, void *stream);
MMC_API int mmc_avg_pool2(const float *x, int64_t x_plane_stride, int64_t planes, int H, int W, float *y, void *stream);
MMC_API int mmc_upsample2x_bilinear(const float *x, int64_t planes, int H, int W, float *y, int64_t y_plane_stride, void *stream);

/* out = a + b (x_rec = x_pred + x_res_hat, models/video/google.py:271) */
MMC_API int mmc_add(const float *a, const float *b, int64_t n, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Token-side kernels of Spatial_aligner's window cross-attention (compressai/models/master.py:484-742).  The block's
 * Linear layers are per-token and run as 1x1 mmc_conv_forward_tc calls on the (B, H, W, C) bf16 token grid; these are the rest.
 *
 * mmc_layernorm_bf16: y = LayerNorm(x [+ delta]) over the last dimension (nn.LayerNorm(dim), master.py:606,613; biased
 *   variance, fp32 statistics, fp32 affine).  With `delta` the residual add in front of the norm is fused (master.py:699) and,
 *   if `sum_out` is given, the bf16 sum is written too.  rows x C bf16 in / out, C <= 256.
 * mmc_gelu_bf16: nn.GELU() (erf form, master.py:464) on n bf16 values (n even).
 * mmc_window_attention: softmax(scale * Q K^T + relative-position bias [+ shift mask]) V inside ws x ws windows of the token
 *   grid cyclically shifted by `shift` (WindowAttention.forward master.py:535-568 inside SwinTransformerBlock.forward
 *   master.py:652-697: roll, window_partition, attention, window_reverse, roll back -- all as index arithmetic).
 *   q: (B, H, W, heads*head_dim) bf16; kv: (B, H, W, 2*heads*head_dim) bf16, keys then values (qkv2's output layout);
 *   bias_table: ((2 ws - 1)^2, heads) fp32 = relative_position_bias_table; out like q.  head_dim must be 32, ws <= 4. */
/* ---------------------------------------------------------------------------------------------
 * Per-image rate / distortion of a forward pass, reduced on the device -- the numbers the reference's evaluation loop keeps
 * (compressai/utils/eval_model/__main__t.py:151-173).  out[b] += scale * (-sum over image b of log2(likelihood)), resp.
 * out[b] += scale * sum over image b of (a - b)^2; image b is the b-th block of n_per_image consecutive floats (any memory format
 * in which a sample is contiguous).  `out` must be zeroed by the caller (several likelihood tensors accumulate into it):
 * bpp = mmc_image_bits(..., scale = 1 / pixels), mse = mmc_image_sse(..., scale = 1 / n_per_image). */
/* y = x / 255 for 8-bit image samples (torchvision ToTensor, the reference's image loader eval_model/__main__t.py:94-101), bit-exact
 * with `img.to(torch.float32).div(255)`: lets the host->device copy carry the decoded 8-bit image instead of its fp32 expansion. */
MMC_API int mmc_u8_to_f32(const void *x, int64_t n, float *y, void *stream);
MMC_API int mmc_image_bits(const float *likelihood, int B, int64_t n_per_image, float scale, float *out, void *stream);
MMC_API int mmc_image_sse(const float *a, const float *b, int B, int64_t n_per_image, float scale, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Non-convolution steps of the ESA gate (ESA.forward, models/google.py:1445-1459) on NHWC bf16 maps.
 * mmc_maxpool_nhwc_bf16: F.max_pool2d(kernel_size=k, stride=stride), no padding; y is (B, (H-k)/stride+1, (W-k)/stride+1, C).
 * mmc_upsample_bilinear_add_bf16: F.interpolate(small, (H, W), mode="bilinear", align_corners=False) + add  (c3 + cf).
 * mmc_sigmoid_gate_bf16: y = x * sigmoid(gate)  (the `x * m` of google.py:1458-1459), n elements. */
MMC_API int mmc_maxpool_nhwc_bf16(const void *x, int B, int H, int W, int C, int k, int stride, void *y, void *stream);
MMC_API int mmc_upsample_bilinear_add_bf16(const void *small, int B, int hs, int ws, int C, const void *add, int H, int W, void *y, void *stream);
MMC_API int mmc_sigmoid_gate_bf16(const void *x, const void *gate, int64_t n, void *y, void *stream);

/* Channel_aligner tail (master.py:193-210): out[b][c] = mean over the HW positions of an fp32 NHWC map (AdaptiveAvgPool2d(1);
 * two passes over row splits, fixed summation order per sample, independent of B; `workspace` of
 * mmc_channel_mean_workspace bytes), and y = gamma[b][c] * x + beta[b][c] on a bf16 NHWC map. */
MMC_API int mmc_channel_mean_workspace(int B, int64_t HW, int C, size_t *bytes);
MMC_API int mmc_channel_mean(const float *x, int B, int64_t HW, int C, void *workspace, float *out, void *stream);
MMC_API int mmc_channel_affine_bf16(const void *x, const float *gamma, const float *beta, int B, int64_t HW, int C, void *y, void *stream);
/* mean over all output positions of conv3x3(t) (stride 1, padding 1, weight (O, C, 3, 3) fp32, optional bias) WITHOUT computing the
 * convolution: by linearity it is a 9 x C x O contraction of border-corrected channel sums of t (one pass over t).  Replaces
 * `avgpool(conv5(out4))` / `avgpool(conv6(out9))` of Channel_aligner.forward (master.py:193-194, 205-206).  t: (B, H, W, C) bf16 NHWC,
 * C % 8 == 0; out: (B, O) fp32; workspace of mmc_conv3x3_mean_workspace bytes.  Batch-independent summation order. */
MMC_API int mmc_conv3x3_mean_workspace(int B, int H, int W, int C, size_t *bytes);
MMC_API int mmc_conv3x3_mean(const void *t, int B, int H, int W, int C, const float *weight, const float *bias, int O, void *workspace,
                             float *out, void *stream);
MMC_API int mmc_layernorm_bf16(const void *x, const void *delta, const float *weight, const float *bias, int64_t rows, int C, float eps,
                               void *sum_out, void *y, void *stream);
MMC_API int mmc_gelu_bf16(const void *x, int64_t n, void *y, void *stream);
MMC_API int mmc_window_attention(const void *q, const void *kv, const float *bias_table, int B, int H, int W, int heads, int head_dim,
                                 int window, int shift, float scale, void *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Backward of the transforms (training step, examples/train.py:239-253): weight gradient on tensor cores.
 * The input gradient of conv() is deconv() with the same weight tensor and vice versa, i.e. mmc_conv_forward_tc with the
 * adjoint descriptor -- no separate entry point.
 * ------------------------------------------------------------------------------------------- */

/* dW[cs][cl][ky][kx] = sum_{b,qy,qx} S[b,qy,qx,cs] * L[b,qy*stride+ky-k/2,qx*stride+kx-k/2,cl]  (fp32 accumulate),
 * S and L NHWC bf16 with channel counts that are multiples of 8.
 * nn.Conv2d (models/utils.py:128-135): S = grad_output, L = input; nn.ConvTranspose2d (utils.py:138-146): S = input,
 * L = grad_output.  Also the GDN gamma gradient with k = 1 (layers/gdn.py:77-92).  Partial sums are ADDED into
 * workspace ([k*k][Cs][Cl] fp32, zeroed by the caller); mmc_wgrad_finalize writes scale * workspace (* mask) in torch's
 * (Cs, Cl, k, k) weight-gradient layout. */
MMC_API int mmc_wgrad_tc(const void *s_nhwc, const void *l_nhwc, int64_t B, int Cs, int Cl, int Hs, int Ws, int Hl, int Wl,
                         int k, int stride, float *workspace, void *stream);
/* Edge layers (image input / reconstruction gradient with <= 8 channels): out[b][qy][qx][tap][8] = x[b][qy*stride + ky - k/2]
 * [qx*stride + kx - k/2][0..7] (zero outside), NHWC bf16 with 8 channels in, k*k*8 channels out; the weight gradient is then one
 * mmc_wgrad_tc call with k = 1 and Cl = k*k*8. */
MMC_API int mmc_im2col8(const void *x_nhwc8, int64_t B, int H, int W, int k, int stride, int Hs, int Ws, void *out, void *stream);
MMC_API int mmc_wgrad_finalize(const float *workspace, int k, int Cs, int Cl, float scale, const float *mask, int accumulate,
                               float *dw, void *stream);

/* ---- elementwise / reduction kernels of the training step (NHWC bf16 activations and gradients, n % 8 == 0) ---- */

/* grad_in = grad_out * act'(y) for the fused ReLU / LeakyReLU(0.01) epilogues (models/google.py:254-269,363-377) */
MMC_API int mmc_act_bwd(const void *grad_out, const void *y, int act, int64_t n, void *grad_in, void *stream);
/* out[c] += scale * sum_rows g[row][c]: bias gradient of conv()/deconv(); GDN beta gradient (scale = -+1/2).
 * out is fp32 [C] and must be initialised by the caller. */
MMC_API int mmc_colsum_bf16(const void *g, int64_t rows, int C, float scale, float *out, void *stream);
MMC_API int mmc_square_bf16(const void *x, int64_t n, void *out, void *stream);
/* GDN / IGDN backward (layers/gdn.py:77-92) around two 1x1 tensor-core contractions:
 *   norm = beta' + gamma' x^2        (mmc_conv_forward_tc, k = 1, fp32 out)
 *   t    = g x norm^(-3/2)           (IGDN: g x norm^(-1/2))                          mmc_gdn_bwd_t
 *   u    = gamma'^T t                (mmc_conv_forward_tc, k = 1, fp32 out)
 *   dx   = g norm^(-1/2) - x u       (IGDN: g norm^(1/2) + x u)                       mmc_gdn_bwd_dx
 *   dbeta' = -+1/2 colsum(t),  dgamma' = -+1/2 t^T x^2  (mmc_wgrad_tc with k = 1) */
MMC_API int mmc_gdn_bwd_t(const void *grad_out, const void *x, const float *norm, int inverse, int64_t n, void *t, void *stream);
MMC_API int mmc_gdn_bwd_dx(const void *grad_out, const void *x, const float *norm, const float *u, int inverse, int64_t n,
                           void *dx, void *stream);
/* NonNegativeParametrizer backward (ops/parametrizers.py:61-64 + ops/bound_ops.py:40-42):
 * d = dp_eff * 2 max(p, bound); dp = d * [(p >= bound) | (d < 0)] */
MMC_API int mmc_reparam_bwd(const float *p, const float *dp_eff, float bound, int64_t n, float *dp, void *stream);

/* torch.abs(y) feeding h_a (models/google.py:283) on the training path: out = bf16(|x|); backward dx = g * sign(x) */
MMC_API int mmc_abs_to_bf16(const float *x, int64_t n, void *out, void *stream);
MMC_API int mmc_abs_bwd(const void *grad_out, const float *x, int64_t n, float *dx, void *stream);

/* GaussianConditional.forward backward (entropy_models.py:692-731): gradients of the bounded likelihood w.r.t. the
 * input (noise mode only; round() has zero gradient), the scales (through the scale LowerBound) and the means.
 * dx / dmeans may be NULL. */
MMC_API int mmc_gc_backward(const float *x, const float *scales, const float *means, const float *noise, const float *grad_lik,
                            float scale_bound, float likelihood_bound, int64_t n, float *dx, float *dscales, float *dmeans,
                            void *stream);
/* EntropyModel._pmf_to_cdf (entropy_models.py:206-214) with compressai._CXX.pmf_to_quantized_cdf (cpp_exts/ops/ops.cpp:40-109)
 * for every row of a table at once, on the device: row r codes pmf[r][0 .. pmf_length[r]) followed by tail_mass[r];
 * cdf is int32 [rows][max_len + 2] (unused tail zero-filled), status[r] = MMC_OK or MMC_EDOMAIN (negative / non-finite / all-zero
 * pmf, or no symbol can donate frequency).  Bit-exact with the reference's integer algorithm. */
MMC_API int mmc_pmf_to_quantized_cdf(const float *pmf, int64_t pmf_pitch, const float *tail_mass, const int32_t *pmf_length, int rows,
                                     int max_len, int precision, int32_t *cdf, int32_t *status, void *stream);

/* EntropyBottleneck.forward backward, noise mode (entropy_models.py:457-540): dx and the packed parameter gradients
 * dparams [C][58] fp32 (ADDED into; order: _matrix0..4 (3,9,9,9,3), _bias0..4 (3,3,3,3,1), _factor0..3 (3 each)),
 * already chained through softplus / tanh of the raw parameters. */
MMC_API int mmc_eb_backward(const float *x, const float *noise, const float *grad_lik, const mmc_eb_params *params,
                            float likelihood_bound, int64_t outer, int64_t C, int64_t inner, float *dx, float *dparams,
                            void *stream);

/* ---------------------------------------------------------------------------------------------
 * Backward passes of the non-convolution steps of the fusion layers (csrc/fusion_bwd.cu), so that the training step of the
 * RGB + depth / RGB-T models (examples/train.py:208-260) keeps every activation-sized pass on these kernels.  Activations and
 * their gradients are NHWC bf16; every adjoint is a gather with fp32 accumulation (no atomics on activations).
 *
 * ESA gate (ESA.forward, models/google.py:1445-1459):
 * mmc_maxpool_nhwc_bf16_idx: mmc_maxpool_nhwc_bf16 that also records the arg-max of every window as the window-local index
 *   ky * k + kx in a uint8 map shaped like y (first maximum in scan order, NaN takes over: the rule of F.max_pool2d); k <= 15.
 * mmc_maxpool_nhwc_bf16_bwd: dx[b][iy][ix][c] = sum of gy over the windows whose arg-max is (iy, ix).
 * mmc_upsample_bilinear_bwd_bf16: adjoint of F.interpolate(small, (H, W), "bilinear", align_corners=False): g is (B, H, W, C),
 *   dsmall (B, hs, ws, C); C % 8 == 0.  (`c3 + cf` passes g unchanged to cf.)
 * mmc_sigmoid_gate_bwd_bf16: y = x * sigmoid(gate): dx = g * s, dgate = g * x * s * (1 - s); n % 8 == 0.
 *
 * Token side of Spatial_aligner (models/master.py:463-568, 572-706):
 * mmc_gelu_bwd_bf16: dx = g * (Phi(x) + x phi(x)) for nn.GELU() (erf form); n even.
 * mmc_layernorm_bwd_bf16: v = the rows that were normalised (x, or the bf16 sum x + delta written by mmc_layernorm_bf16);
 *   dv = LayerNorm backward of g [+ g_sum, the gradient reaching v through the residual stream; may be NULL];
 *   dweight[C] += sum_rows g * xhat, dbias[C] += sum_rows g (fp32, ADDED into: the caller zeroes or accumulates).  C <= 256.
 * mmc_window_attention_bwd: gradients of mmc_window_attention w.r.t. q (dq like q), kv (dkv like kv) and the relative-position
 *   bias table (dtable ((2 ws - 1)^2, heads) fp32, ADDED into).  Same layout rules and limits as the forward call. */
MMC_API int mmc_maxpool_nhwc_bf16_idx(const void *x, int B, int H, int W, int C, int k, int stride, void *y, void *idx, void *stream);
MMC_API int mmc_maxpool_nhwc_bf16_bwd(const void *gy, const void *idx, int B, int H, int W, int C, int k, int stride, void *dx, void *stream);
MMC_API int mmc_upsample_bilinear_bwd_bf16(const void *g, int B, int H, int W, int C, int hs, int ws, void *dsmall, void *stream);
MMC_API int mmc_sigmoid_gate_bwd_bf16(const void *g, const void *x, const void *gate, int64_t n, void *dx, void *dgate, void *stream);
MMC_API int mmc_gelu_bwd_bf16(const void *g, const void *x, int64_t n, void *dx, void *stream);
MMC_API int mmc_layernorm_bwd_bf16(const void *g, const void *v, const void *g_sum, const float *weight, int64_t rows, int C, float eps,
                                   void *dv, float *dweight, float *dbias, void *stream);
MMC_API int mmc_window_attention_bwd(const void *q, const void *kv, const float *bias_table, const void *dout, int B, int H, int W,
                                     int heads, int head_dim, int window, int shift, float scale, void *dq, void *dkv, float *dtable,
                                     void *stream);

/* ---------------------------------------------------------------------------------------------
 * Range-ANS coder on the device (csrc/rans_device.cu).  The reference's stream is one serial rANS
 * chain per image (cpp_exts/rans/rans_interface.cpp:108-200) and stays on host threads (mmc_rans_encode_batch_host, byte-identical).
 * These entry points code the SAME symbols against the SAME quantised CDF tables / indexes / escape scheme into a container of
 * our own in which an image's symbols are dealt round-robin onto `lanes` independent 32-bit-state / 16-bit-word rANS lanes
 * (layout in the file header; CPU restatement oracle/lane_rans.py), so symbols and indexes never leave the GPU.
 * All pointers are DEVICE pointers; calls are asynchronous on `stream`.
 *   symbols / indexes: int32 [batch][n] in the reference's flattening order; cdfs int32 [n_cdfs][cdf_stride], cdf_sizes / offsets
 *   int32 [n_cdfs] (EntropyModel._quantized_cdf / _cdf_length / _offset, entropy_models.py:216-235).
 *   out: [batch][cap_per_stream] bytes (cap % 4 == 0); nbytes[b] = container size (written even if it does not fit);
 *   status: one int, 0 = ok, bit 0 index out of range, bit 1 bad CDF row, bit 2 capacity too small, bit 3 malformed stream.
 *   workspace: mmc_rans_device_workspace bytes, 16-byte aligned (holds the per-call encoder table: start / freq / exact 64-bit
 *   reciprocal per CDF entry).  mmc_rans_lanes_default(n): lane count used when the caller has no preference (the smallest power
 *   of two in [4, 1024] with at most 8192 symbols per lane: header overhead 8 bytes per lane).
 * Decoding: streams = concatenated containers, stream_offsets[b] (multiples of 4) / stream_bytes[b] uint64; max_lanes >= the
 * largest lane count in the batch; symbols_out int32 [batch][n]. */
MMC_API int mmc_rans_lanes_default(int64_t n);
MMC_API int mmc_rans_device_workspace(int batch, int lanes, int n_cdfs, int cdf_stride, size_t *bytes);
MMC_API int mmc_rans_encode_device(const int32_t *symbols, const int32_t *indexes, int batch, int64_t n, const int32_t *cdfs, int n_cdfs,
                                   int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets, int lanes, uint8_t *out,
                                   size_t cap_per_stream, uint64_t *nbytes, void *workspace, int *status, void *stream);
MMC_API int mmc_rans_decode_device(const uint8_t *streams, const uint64_t *stream_offsets, const uint64_t *stream_bytes,
                                   const int32_t *indexes, int batch, int64_t n, int max_lanes, const int32_t *cdfs, int n_cdfs,
                                   int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets, int32_t *symbols_out, int *status,
                                   void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MMCODEC_H */
