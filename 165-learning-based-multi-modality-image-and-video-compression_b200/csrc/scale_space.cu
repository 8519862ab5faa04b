// scale_space.cu -- scale-space-flow prediction of the ssf2020 video codec (HBM-bound gather / stencil kernels).
//
// Replaces ScaleSpaceFlow.gaussian_volume / warp_volume / forward_prediction
// (compressai/models/video/google.py:331-382) and the helpers they call (compressai/models/utils.py:155-195):
//   volume level 0 = x_ref, level 1 = blur(x_ref), level i>=2 = upsample2x^(i-1)(blur(avgpool2(level source)))
//   x_pred[n,c,h,w] = trilinear sample of volume[n,c,:,:,:] at (w + flow_x*W/2, h + flow_y*H/2, scale) with border
//   padding, align_corners=False (F.grid_sample on a 5-D input).
// Everything is planar fp32 ([planes][H][W]), exactly the reference's layout: these tensors have 3 channels.
#include "common.cuh"

namespace mmc {

constexpr int kBlurTile = 32;
constexpr int kMaxBlurTaps = 33;

struct BlurTaps {
    float w[kMaxBlurTaps];
};

// Depthwise Gaussian blur with replicate padding (utils.py:173-189).  The reference convolves with the outer product
// k k^T; the separable evaluation (rows then columns inside one shared-memory tile) differs from it only by fp32
// summation order.
__global__ void __launch_bounds__(256) gaussian_blur_kernel(const float *__restrict__ x, int H, int W, BlurTaps taps, int ksize,
                                                           float *__restrict__ y, int64_t y_plane_stride)
{
    extern __shared__ float sm[];
    const int r = ksize / 2, span = kBlurTile + 2 * r;
    float *in = sm;                      // [span][span + 1]
    float *tmp = sm + span * (span + 1); // [span][kBlurTile + 1]  horizontal pass
    const int64_t plane = blockIdx.z;
    const int x0 = blockIdx.x * kBlurTile, y0 = blockIdx.y * kBlurTile;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8 threads, no index divisions in the loops
    const float *xp = x + plane * (int64_t)H * W;
    for (int ly = ty; ly < span; ly += 8) {
        const float *row = xp + (int64_t)min(max(y0 + ly - r, 0), H - 1) * W;
        for (int lx = tx; lx < span; lx += 32) in[ly * (span + 1) + lx] = __ldg(row + min(max(x0 + lx - r, 0), W - 1));
    }
    __syncthreads();
    for (int ly = ty; ly < span; ly += 8) {
        float acc = 0.0f;
        for (int t = 0; t < ksize; ++t) acc += taps.w[t] * in[ly * (span + 1) + tx + t];
        tmp[ly * (kBlurTile + 1) + tx] = acc;
    }
    __syncthreads();
    float *yp = y + plane * y_plane_stride;
    const int gx = x0 + tx;
    if (gx < W) {
        for (int ly = ty; ly < kBlurTile; ly += 8) {
            const int gy = y0 + ly;
            if (gy >= H) break;
            float acc = 0.0f;
            for (int t = 0; t < ksize; ++t) acc += taps.w[t] * tmp[(ly + t) * (kBlurTile + 1) + tx];
            yp[(int64_t)gy * W + gx] = acc;
        }
    }
}

// F.avg_pool2d(x, 2, 2) on [planes][H][W] (H, W even) -> [planes][H/2][W/2].  grid = (ceil(Wo / 256), Ho, planes): no index divisions.
__global__ void __launch_bounds__(256) avg_pool2_kernel(const float *__restrict__ x, int64_t x_plane_stride, int H, int W, float *__restrict__ y)
{
    const int Ho = H / 2, Wo = W / 2;
    const int ox = blockIdx.x * 256 + threadIdx.x, oy = blockIdx.y;
    if (ox >= Wo) return;
    const int64_t p = blockIdx.z;
    const float *xp = x + p * x_plane_stride + (int64_t)(2 * oy) * W + 2 * ox;
    const float2 a = *reinterpret_cast<const float2 *>(xp);
    const float2 b = *reinterpret_cast<const float2 *>(xp + W);
    y[(p * Ho + oy) * (int64_t)Wo + ox] = (a.x + a.y + b.x + b.y) * 0.25f;
}

// F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False): src = (dst + 0.5) / 2 - 0.5 clamped at 0.
// grid = (ceil(W / 128), 2H, planes); each thread produces the two outputs 2*ix, 2*ix + 1 of one output row (one 8-byte store).
__global__ void __launch_bounds__(128) upsample2x_kernel(const float *__restrict__ x, int H, int W, float *__restrict__ y, int64_t y_plane_stride)
{
    const int ix = blockIdx.x * 128 + threadIdx.x, oy = blockIdx.y;
    if (ix >= W) return;
    const int64_t p = blockIdx.z;
    const float sy = fmaxf(0.5f * (oy + 0.5f) - 0.5f, 0.0f);
    const int y0 = (int)sy, y1 = min(y0 + 1, H - 1);
    const float ly = sy - y0, hy = 1.0f - ly;
    const float *r0 = x + (p * H + y0) * (int64_t)W, *r1 = x + (p * H + y1) * (int64_t)W;
    const int xm = max(ix - 1, 0), xp = min(ix + 1, W - 1);
    const float a0 = __ldg(r0 + xm), b0 = __ldg(r0 + ix), c0 = __ldg(r0 + xp);
    const float a1 = __ldg(r1 + xm), b1 = __ldg(r1 + ix), c1 = __ldg(r1 + xp);
    // output 2*ix: src = ix - 0.25 -> x0 = ix - 1 (lambda 0.75) except at the left border (src clamped to 0: x0 = 0, lambda 0);
    // output 2*ix + 1: src = ix + 0.25 -> x0 = ix (lambda 0.25), x1 = min(ix + 1, W - 1)
    float e, o;
    if (ix == 0) e = hy * (1.0f * b0 + 0.0f * c0) + ly * (1.0f * b1 + 0.0f * c1);
    else         e = hy * (0.25f * a0 + 0.75f * b0) + ly * (0.25f * a1 + 0.75f * b1);
    o = hy * (0.75f * b0 + 0.25f * c0) + ly * (0.75f * b1 + 0.25f * c1);
    *reinterpret_cast<float2 *>(y + p * y_plane_stride + (int64_t)oy * (2 * W) + 2 * ix) = make_float2(e, o);
}

// plane copy into the volume; grid = (ceil(hw / 1024), planes), float4 per thread (hw % 4 == 0)
__global__ void __launch_bounds__(256) copy_planes_kernel(const float4 *__restrict__ x, int64_t hw4, float4 *__restrict__ y, int64_t y_plane_stride4)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= hw4) return;
    const int64_t p = blockIdx.y;
    y[p * y_plane_stride4 + i] = ldg_stream(x + p * hw4 + i);
}

// grid_sampler_unnormalize + clip_coordinates (border padding, align_corners=False), no FMA contraction so that the
// sample position is computed with the reference's roundings
__device__ __forceinline__ float source_index(float coord, int size)
{
    float v = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(coord, 1.0f), (float)size), 1.0f), 2.0f);
    return fminf((float)(size - 1), fmaxf(v, 0.0f));
}

// One thread per output pixel (grid = (ceil(W / 128), H, N)); the 8 corner weights are shared by the C channels.
__global__ void __launch_bounds__(128) scale_space_warp_kernel(const float *__restrict__ volume, const float *__restrict__ motion,
                                                              const float *__restrict__ base_x, const float *__restrict__ base_y,
                                                              int C, int D, int H, int W, const float *__restrict__ x_cur,
                                                              float *__restrict__ x_pred, float *__restrict__ x_res)
{
    const int w = blockIdx.x * 128 + threadIdx.x, h = blockIdx.y;
    if (w >= W) return;
    const int64_t n = blockIdx.z, hw = (int64_t)H * W, p = (int64_t)h * W + w;
    const float *m = motion + n * 3 * hw + p;
    const float gx = __fadd_rn(__ldg(base_x + w), __ldg(m)), gy = __fadd_rn(__ldg(base_y + h), __ldg(m + hw)), gz = __ldg(m + 2 * hw);
    const float ix = source_index(gx, W), iy = source_index(gy, H), iz = source_index(gz, D);
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
    const float wx1 = ix - fx, wy1 = iy - fy, wz1 = iz - fz;                 // weights of the +1 corners
    const float wx0 = (fx + 1.0f) - ix, wy0 = (fy + 1.0f) - iy, wz0 = (fz + 1.0f) - iz;
    const bool x1ok = x0 + 1 < W, y1ok = y0 + 1 < H, z1ok = z0 + 1 < D;      // x0, y0, z0 are in range after clipping
    const float *vb = volume + n * C * D * hw;
    for (int c = 0; c < C; ++c) {
        const float *v0 = vb + ((int64_t)c * D + z0) * hw + (int64_t)y0 * W + x0;
        const float *v1 = v0 + hw;
        float acc = 0.0f;
        acc += __ldg(v0) * (wx0 * wy0 * wz0);
        if (x1ok) acc += __ldg(v0 + 1) * (wx1 * wy0 * wz0);
        if (y1ok) acc += __ldg(v0 + W) * (wx0 * wy1 * wz0);
        if (x1ok && y1ok) acc += __ldg(v0 + W + 1) * (wx1 * wy1 * wz0);
        if (z1ok) {
            acc += __ldg(v1) * (wx0 * wy0 * wz1);
            if (x1ok) acc += __ldg(v1 + 1) * (wx1 * wy0 * wz1);
            if (y1ok) acc += __ldg(v1 + W) * (wx0 * wy1 * wz1);
            if (x1ok && y1ok) acc += __ldg(v1 + W + 1) * (wx1 * wy1 * wz1);
        }
        const int64_t o = (n * C + c) * hw + p;
        x_pred[o] = acc;
        if (x_res) x_res[o] = __ldg(x_cur + o) - acc;
    }
}

// rgb2ycbcr / ycbcr2rgb, ITU-R BT.709 (compressai/transforms/functional.py:26-66), planar [N][3][HW] fp32; the operation order
// (and the absence of FMA contraction) follows the reference's expression so that results agree to the last bit or two
__global__ void __launch_bounds__(256) color_kernel(const float *__restrict__ x, int64_t hw, int to_ycbcr, float *__restrict__ y)
{
    const float Kr = 0.2126f, Kg = 0.7152f, Kb = 0.0722f;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= hw) return;
    const int64_t base = (int64_t)blockIdx.y * 3 * hw + i;
    const float a = x[base], b = x[base + hw], c = x[base + 2 * hw];
    float o0, o1, o2;
    if (to_ycbcr) {
        const float yy = __fadd_rn(__fadd_rn(__fmul_rn(Kr, a), __fmul_rn(Kg, b)), __fmul_rn(Kb, c));
        o0 = yy;
        o1 = __fadd_rn(__fdiv_rn(__fmul_rn(0.5f, __fsub_rn(c, yy)), (float)(1.0 - 0.0722)), 0.5f);
        o2 = __fadd_rn(__fdiv_rn(__fmul_rn(0.5f, __fsub_rn(a, yy)), (float)(1.0 - 0.2126)), 0.5f);
    } else {
        const float r = __fadd_rn(a, __fmul_rn((float)(2.0 - 2.0 * 0.2126), __fsub_rn(c, 0.5f)));
        const float bb = __fadd_rn(a, __fmul_rn((float)(2.0 - 2.0 * 0.0722), __fsub_rn(b, 0.5f)));
        const float g = __fdiv_rn(__fsub_rn(__fsub_rn(a, __fmul_rn(Kr, r)), __fmul_rn(Kb, bb)), Kg);
        o0 = r; o1 = g; o2 = bb;
    }
    y[base] = o0; y[base + hw] = o1; y[base + 2 * hw] = o2;
}

__global__ void __launch_bounds__(256) add_kernel(const float4 *__restrict__ a, const float4 *__restrict__ b, int64_t n4, float4 *__restrict__ out,
                                                 const float *at, const float *bt, float *ot, int tail)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 u = ldg_stream(a + i), v = ldg_stream(b + i);
        out[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) ot[threadIdx.x] = at[threadIdx.x] + bt[threadIdx.x];
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_gaussian_volume_workspace(int64_t planes, int H, int W, size_t *bytes)
{
    MMC_CHECK_ARG(bytes && planes >= 0 && H >= 1 && W >= 1, "mmc_gaussian_volume_workspace: bad argument");
    // pooled image + its blur at half resolution, plus two ping-pong upsampling buffers (largest: half resolution)
    *bytes = (size_t)planes * (size_t)(H / 2) * (size_t)(W / 2) * sizeof(float) * 4;
    return MMC_OK;
}

int mmc_gaussian_volume(const float *x, int64_t planes, int H, int W, const float *kernel1d_host, int ksize, int num_levels,
                        void *workspace, size_t ws_bytes, float *volume, void *stream)
{
    const char *name = "mmc_gaussian_volume";
    MMC_CHECK_ARG(planes >= 0 && H >= 1 && W >= 1 && num_levels >= 1, "%s: bad shape", name);
    MMC_CHECK_ARG(kernel1d_host && ksize >= 1 && (ksize & 1) && ksize <= kMaxBlurTaps, "%s: kernel size %d not odd or > %d", name, ksize, kMaxBlurTaps);
    const int div = 1 << (num_levels - 1);
    MMC_CHECK_ARG(H % div == 0 && W % div == 0, "%s: H and W must be multiples of %d (avg_pool2d / interpolate chain)", name, div);
    if (planes == 0) return MMC_OK;
    size_t need;
    mmc_gaussian_volume_workspace(planes, H, W, &need);
    MMC_CHECK_ARG(x && volume && (num_levels == 1 || (workspace && ws_bytes >= need)), "%s: NULL buffer or workspace too small", name);
    MMC_CHECK_ARG(planes <= 65535, "%s: too many planes", name);
    cudaStream_t st = (cudaStream_t)stream;
    BlurTaps taps;
    for (int i = 0; i < ksize; ++i) taps.w[i] = kernel1d_host[i];
    const int D = num_levels + 1;
    const int64_t hw = (int64_t)H * W, vstride = (int64_t)D * hw;
    const int span = kBlurTile + 2 * (ksize / 2);
    const size_t blur_smem = ((size_t)span * (span + 1) + (size_t)span * (kBlurTile + 1)) * sizeof(float);
    static PerDevice<int> attr_dev;
    if (!attr_dev.cur().load(std::memory_order_relaxed)) {
        MMC_CHECK_CUDA(cudaFuncSetAttribute(gaussian_blur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_dev.cur().store(1, std::memory_order_relaxed);
    }
    auto blur = [&](const float *src, int h, int w, float *dst, int64_t dst_stride) -> int {
        dim3 grid((unsigned)((w + kBlurTile - 1) / kBlurTile), (unsigned)((h + kBlurTile - 1) / kBlurTile), (unsigned)planes);
        gaussian_blur_kernel<<<grid, 256, blur_smem, st>>>(src, h, w, taps, ksize, dst, dst_stride);
        MMC_CHECK_LAUNCH(name);
        return MMC_OK;
    };
    // level 0: the frame itself; level 1: its blur
    MMC_CHECK_ARG(hw % 4 == 0 && aligned16(x) && aligned16(volume), "%s: H * W must be a multiple of 4 and the buffers 16-byte aligned", name);
    copy_planes_kernel<<<dim3((unsigned)((hw / 4 + 255) / 256), (unsigned)planes), 256, 0, st>>>((const float4 *)x, hw / 4, (float4 *)volume, vstride / 4);
    MMC_CHECK_LAUNCH(name);
    int rc = blur(x, H, W, volume + hw, vstride);
    if (rc) return rc;
    // levels 2..: pool the previous blurred image, blur it, upsample (level - 1) times by 2 into the volume slot
    float *ws = (float *)workspace;
    const size_t q = (size_t)planes * (H / 2) * (W / 2);
    float *pooled = ws, *blurred = ws + q, *up[2] = {ws + 2 * q, ws + 3 * q};
    const float *prev = volume + hw;   // blurred image at the previous resolution
    int64_t prev_stride = vstride;
    int h = H, w = W;
    for (int lvl = 2; lvl <= num_levels; ++lvl) {
        avg_pool2_kernel<<<dim3((unsigned)((w / 2 + 255) / 256), (unsigned)(h / 2), (unsigned)planes), 256, 0, st>>>(prev, prev_stride, h, w, pooled);
        MMC_CHECK_LAUNCH(name);
        h /= 2; w /= 2;
        rc = blur(pooled, h, w, blurred, (int64_t)h * w);
        if (rc) return rc;
        prev = blurred; prev_stride = (int64_t)h * w;
        const float *src = blurred;
        int uh = h, uw = w;
        for (int u = 0; u < lvl - 1; ++u) {
            const bool last = (u == lvl - 2);
            float *dst = last ? volume + (int64_t)lvl * hw : up[u & 1];
            upsample2x_kernel<<<dim3((unsigned)((uw + 127) / 128), (unsigned)(2 * uh), (unsigned)planes), 128, 0, st>>>(src, uh, uw, dst,
                                                                                                                  last ? vstride : (int64_t)uh * uw * 4);
            MMC_CHECK_LAUNCH(name);
            src = dst; uh *= 2; uw *= 2;
        }
    }
    return MMC_OK;
}

int mmc_scale_space_warp(const float *volume, const float *motion_info, const float *base_x, const float *base_y, int64_t N, int C, int D,
                         int H, int W, const float *x_cur, float *x_pred, float *x_res, void *stream)
{
    const char *name = "mmc_scale_space_warp";
    MMC_CHECK_ARG(N >= 0 && C >= 1 && D >= 1 && H >= 1 && W >= 1, "%s: bad shape", name);
    if (N == 0) return MMC_OK;
    MMC_CHECK_ARG(volume && motion_info && base_x && base_y && x_pred, "%s: NULL buffer", name);
    MMC_CHECK_ARG(!x_res || x_cur, "%s: x_res requested without x_cur", name);
    MMC_CHECK_ARG(N <= 65535 && H <= 65535, "%s: N and H must be <= 65535", name);
    scale_space_warp_kernel<<<dim3((unsigned)((W + 127) / 128), (unsigned)H, (unsigned)N), 128, 0, (cudaStream_t)stream>>>(volume, motion_info, base_x, base_y,
                                                                                                                      C, D, H, W, x_cur, x_pred, x_res);
    MMC_CHECK_LAUNCH(name);
    return MMC_OK;
}

int mmc_color_convert(const float *x, int64_t N, int64_t HW, int to_ycbcr, float *y, void *stream)
{
    MMC_CHECK_ARG(N >= 0 && HW >= 0 && N <= 65535, "mmc_color_convert: bad shape");
    if (N * HW == 0) return MMC_OK;
    MMC_CHECK_ARG(x && y, "mmc_color_convert: NULL buffer");
    color_kernel<<<dim3((unsigned)((HW + 255) / 256), (unsigned)N), 256, 0, (cudaStream_t)stream>>>(x, HW, to_ycbcr, y);
    MMC_CHECK_LAUNCH("mmc_color_convert");
    return MMC_OK;
}

int mmc_avg_pool2(const float *x, int64_t x_plane_stride, int64_t planes, int H, int W, float *y, void *stream)
{
    MMC_CHECK_ARG(planes >= 0 && planes <= 65535 && H >= 2 && W >= 2 && H / 2 <= 65535 && x_plane_stride >= (int64_t)H * W && W % 2 == 0,
                  "mmc_avg_pool2: bad shape (W must be even)");
    if (planes == 0) return MMC_OK;
    MMC_CHECK_ARG(x && y && (reinterpret_cast<uintptr_t>(x) & 7) == 0 && x_plane_stride % 2 == 0, "mmc_avg_pool2: NULL or unaligned buffer");
    avg_pool2_kernel<<<dim3((unsigned)((W / 2 + 255) / 256), (unsigned)(H / 2), (unsigned)planes), 256, 0, (cudaStream_t)stream>>>(x, x_plane_stride, H, W, y);
    MMC_CHECK_LAUNCH("mmc_avg_pool2");
    return MMC_OK;
}

int mmc_upsample2x_bilinear(const float *x, int64_t planes, int H, int W, float *y, int64_t y_plane_stride, void *stream)
{
    MMC_CHECK_ARG(planes >= 0 && planes <= 65535 && H >= 1 && W >= 1 && 2 * H <= 65535 && y_plane_stride >= 4ll * H * W && y_plane_stride % 2 == 0,
                  "mmc_upsample2x_bilinear: bad shape");
    if (planes == 0) return MMC_OK;
    MMC_CHECK_ARG(x && y && (reinterpret_cast<uintptr_t>(y) & 7) == 0, "mmc_upsample2x_bilinear: NULL or unaligned buffer");
    upsample2x_kernel<<<dim3((unsigned)((W + 127) / 128), (unsigned)(2 * H), (unsigned)planes), 128, 0, (cudaStream_t)stream>>>(x, H, W, y, y_plane_stride);
    MMC_CHECK_LAUNCH("mmc_upsample2x_bilinear");
    return MMC_OK;
}

int mmc_add(const float *a, const float *b, int64_t n, float *out, void *stream)
{
    MMC_CHECK_ARG(n >= 0, "mmc_add: n < 0");
    if (n == 0) return MMC_OK;
    MMC_CHECK_ARG(a && b && out && aligned16(a) && aligned16(b) && aligned16(out), "mmc_add: NULL or unaligned buffer");
    const int64_t n4 = n / 4;
    add_kernel<<<elementwise_grid(n4 > 0 ? n4 : 1, 256), 256, 0, (cudaStream_t)stream>>>((const float4 *)a, (const float4 *)b, n4, (float4 *)out,
                                                                                      a + 4 * n4, b + 4 * n4, out + 4 * n4, (int)(n - 4 * n4));
    MMC_CHECK_LAUNCH("mmc_add");
    return MMC_OK;
}

}  // extern "C"
