// common.cuh -- shared helpers for libmmcodec (sm_100a only).
#pragma once

#include <atomic>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "mmcodec.h"

namespace mmc {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define MMC_CHECK_ARG(cond, ...)             \
    do {                                     \
        if (!(cond)) {                       \
            mmc::set_error(__VA_ARGS__);     \
            return MMC_EINVAL;               \
        }                                    \
    } while (0)

#define MMC_UNSUPPORTED(cond, ...)           \
    do {                                     \
        if (cond) {                          \
            mmc::set_error(__VA_ARGS__);     \
            return MMC_EUNSUPPORTED;         \
        }                                    \
    } while (0)

// Called right after a kernel launch: catches launch-configuration errors without synchronising.
#define MMC_CHECK_LAUNCH(name)                                                          \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            mmc::set_error("%s: CUDA error %s", name, cudaGetErrorString(e__));         \
            return MMC_ECUDA;                                                           \
        }                                                                               \
        mmc::count_launch();                                                            \
    } while (0)

#define MMC_CHECK_CUDA(expr)                                                            \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            mmc::set_error("%s: CUDA error %s", #expr, cudaGetErrorString(e__));        \
            return MMC_ECUDA;                                                           \
        }                                                                               \
    } while (0)

// Per-device launch state (cudaFuncSetAttribute and occupancy queries apply to the CURRENT device only, so anything cached
// about them is keyed by cudaGetDevice(); plain process-wide statics would leave every device but the first unconfigured).
// Values are idempotent (every thread computes the same number), so relaxed atomics are all the synchronisation needed.
constexpr int kMaxDevices = 64;
template <typename T>
struct PerDevice {
    std::atomic<T> v[kMaxDevices];
    std::atomic<T> &cur()
    {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
        return v[d];
    }
};

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Grid for a grid-stride elementwise kernel: enough CTAs to fill the machine a few times over,
// always a multiple of the SM count so that the last wave is full.
inline int elementwise_grid(int64_t work_items, int block, int ctas_per_sm = 8)
{
    int64_t need = (work_items + block - 1) / block;
    int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
    if (need >= cap) return (int)cap;
    if (need <= 0) return 1;
    // round up to a multiple of the SM count when that does not more than double the grid
    int64_t r = ((need + kNumSMs - 1) / kNumSMs) * kNumSMs;
    return (int)((r <= 2 * need) ? r : need);
}

// torch.max(x, bound): NaN propagates (compressai/ops/bound_ops.py:36-37)
__device__ __forceinline__ float lower_bound_f(float x, float b) { return (x != x) ? x : fmaxf(x, b); }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum -> one atomicAdd per CTA.  `red` is a __shared__ float[32].
__device__ __forceinline__ void block_atomic_add(float v, float *red, float *dst)
{
    v = warp_sum(v);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int nw = (blockDim.x + 31) >> 5;
        float s = (lane < nw) ? red[lane] : 0.0f;
        s = warp_sum(s);
        if (lane == 0) atomicAdd(dst, s);
    }
}

__device__ __forceinline__ float4 ldg_stream(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace mmc
