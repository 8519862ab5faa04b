// probe_mma_rate.cu -- how many SM cycles does one tcgen05.mma (M = 128, K = 16, bf16, cta_group::1) cost as a function of
// N, of the shared-memory operand layout (SWIZZLE_128B / 64B / 32B / none, all K-major) and of where A lives (smem or TMEM)?
// Round-1 finding to explain: the conv main loop with N = 128 plateaus at ~52 % of the tensor pipe (~122 cycles per MMA
// instead of 64).  Operand CONTENTS do not matter for timing, so shared memory is just zero-filled; only the descriptors
// differ.  One CTA per SM, one thread issues `iters` x 4 K-steps cycling over `nslots` operand slots, then commits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/bin/probe_mma_rate profiles/probe_mma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// layout: 0 none, 2 = 128B swizzle, 4 = 64B, 6 = 32B (sm_100 descriptor encoding, bits 61-63)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// smem -> TMEM copy of 128 rows x 256 bits (one K = 16 slice of a bf16 A tile) through the tensor-core pipe
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc)
{
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

struct Cfg {
    int n, layout, a_tmem, iters, nslots;   // a_tmem: 0 = SS, 1 = TS (static A), 2 = tcgen05.cp of the A slice then TS, 3 = cp only
    int a_off[4], b_off[4];   // descriptor start-address offset of each K = 16 step
    int lbo, sbo;
    int a_bytes, b_bytes;
};

__global__ void __launch_bounds__(128, 1) probe(Cfg c, long long *cycles)
{
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int slot_bytes = c.a_bytes + c.b_bytes;
    for (int i = threadIdx.x; i < c.nslots * slot_bytes / 16; i += blockDim.x) ((uint4 *)smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (threadIdx.x < 32) {
        // the whole warp walks the loop (uniform control flow, uniform registers); the lane chosen by elect.sync issues -- the issue
        // scheme of conv_tc.cu.  (A first version issued from `if (threadIdx.x == 0)`: every tcgen05.mma then paid ~125 cycles of
        // R2UR / predication overhead and the probe measured its own issue loop.)
        const uint32_t idesc = make_idesc(c.n);
        const uint32_t base = smem_u32(smem);
        long long t0 = clock64();
        int slot = 0;
        for (int it = 0; it < c.iters; ++it) {
            const uint32_t a_addr = base + slot * slot_bytes, b_addr = a_addr + c.a_bytes;
            const uint64_t ad = make_desc(a_addr, c.lbo, c.sbo, c.layout), bd = make_desc(b_addr, c.lbo, c.sbo, c.layout);
            const uint32_t d = tmem + (uint32_t)(c.a_tmem ? (c.n <= 192 ? (it & 1) * c.n : 0) : (it & 1) * 256);
            const uint32_t a_t = tmem + (uint32_t)(448 + (it & 1) * 32);     // two A buffers of 32 columns each
            const uint32_t acc = it > 1;
            if (elect_one()) {
                if (c.layout == 2) {
                    // sw128: K steps are +32 B inside the swizzle atom for both operands (immediate offsets)
                    if (c.a_tmem == 0) {
                        mma_ss(d, ad, bd, idesc, acc); mma_ss(d, ad + 2, bd + 2, idesc, 1); mma_ss(d, ad + 4, bd + 4, idesc, 1); mma_ss(d, ad + 6, bd + 6, idesc, 1);
                    } else if (c.a_tmem == 1) {
                        mma_ts(d, a_t, bd, idesc, acc); mma_ts(d, a_t + 8, bd + 2, idesc, 1); mma_ts(d, a_t + 16, bd + 4, idesc, 1); mma_ts(d, a_t + 24, bd + 6, idesc, 1);
                    } else if (c.a_tmem == 2) {
                        cp_128x256b(a_t, ad); cp_128x256b(a_t + 8, ad + 2); cp_128x256b(a_t + 16, ad + 4); cp_128x256b(a_t + 24, ad + 6);
                        mma_ts(d, a_t, bd, idesc, acc); mma_ts(d, a_t + 8, bd + 2, idesc, 1); mma_ts(d, a_t + 16, bd + 4, idesc, 1); mma_ts(d, a_t + 24, bd + 6, idesc, 1);
                    } else {
                        cp_128x256b(a_t, ad); cp_128x256b(a_t + 8, ad + 2); cp_128x256b(a_t + 16, ad + 4); cp_128x256b(a_t + 24, ad + 6);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (c.a_tmem >= 2) cp_128x256b(a_t + k * 8, ad + (uint64_t)(c.a_off[k] >> 4));
                        if (c.a_tmem == 1 || c.a_tmem == 2) mma_ts(d, a_t + k * 8, bd + (uint64_t)(c.b_off[k] >> 4), idesc, acc | k);
                        else if (c.a_tmem == 0) mma_ss(d, ad + (uint64_t)(c.a_off[k] >> 4), bd + (uint64_t)(c.b_off[k] >> 4), idesc, acc | k);
                    }
                }
            }
            __syncwarp();
            if (++slot == c.nslots) slot = 0;
        }
        if (elect_one()) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// Issue-loop variants for the SS sw128 case (is ~94 cycles per N <= 128 MMA the tensor core or the issuing warp?):
//   kVariant 0: one elect.sync per 4 K steps, descriptors rebuilt per slot (the loop of `probe`, i.e. conv_tc.cu's shape)
//   kVariant 1: one elect.sync per kUnroll slots (4 * kUnroll MMAs), descriptors precomputed before the loop
//   kVariant 2: a single elected thread runs the whole loop (elect.sync once), descriptors precomputed
template <int kVariant, int kUnroll>
__global__ void __launch_bounds__(128, 1) probe_issue(int n, int iters, long long *cycles)
{
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int slot_bytes = 128 * 128 + 256 * 128;      // A tile + room for a 256-row B tile
    for (int i = threadIdx.x; i < kUnroll * slot_bytes / 16; i += blockDim.x) ((uint4 *)smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc(n);
        const uint32_t base = smem_u32(smem);
        uint64_t ad[kUnroll], bd[kUnroll];
#pragma unroll
        for (int s = 0; s < kUnroll; ++s) {
            ad[s] = make_desc(base + s * slot_bytes, 16, 1024, 2);
            bd[s] = make_desc(base + s * slot_bytes + 128 * 128, 16, 1024, 2);
        }
        long long t0 = clock64();
        if (kVariant == 2) {
            if (elect_one()) {
                for (int it = 0; it < iters; it += kUnroll) {
#pragma unroll
                    for (int s = 0; s < kUnroll; ++s) {
                        const uint32_t d = tmem + (uint32_t)((s & 1) * 256);
                        mma_ss(d, ad[s], bd[s], idesc, it > 1); mma_ss(d, ad[s] + 2, bd[s] + 2, idesc, 1);
                        mma_ss(d, ad[s] + 4, bd[s] + 4, idesc, 1); mma_ss(d, ad[s] + 6, bd[s] + 6, idesc, 1);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
        } else {
            for (int it = 0; it < iters; it += kUnroll) {
                if (kVariant == 1) {
                    if (elect_one()) {
#pragma unroll
                        for (int s = 0; s < kUnroll; ++s) {
                            const uint32_t d = tmem + (uint32_t)((s & 1) * 256);
                            mma_ss(d, ad[s], bd[s], idesc, it > 1); mma_ss(d, ad[s] + 2, bd[s] + 2, idesc, 1);
                            mma_ss(d, ad[s] + 4, bd[s] + 4, idesc, 1); mma_ss(d, ad[s] + 6, bd[s] + 6, idesc, 1);
                        }
                    }
                    __syncwarp();
                } else {
#pragma unroll
                    for (int s = 0; s < kUnroll; ++s) {
                        const uint32_t d = tmem + (uint32_t)((s & 1) * 256);
                        if (elect_one()) {
                            mma_ss(d, ad[s], bd[s], idesc, it > 1); mma_ss(d, ad[s] + 2, bd[s] + 2, idesc, 1);
                            mma_ss(d, ad[s] + 4, bd[s] + 4, idesc, 1); mma_ss(d, ad[s] + 6, bd[s] + 6, idesc, 1);
                        }
                        __syncwarp();
                    }
                }
            }
            if (elect_one()) {
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
        }
        asm volatile("{\n\t.reg .pred P1;\n\tW3:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D3;\n\tbra W3;\n\tD3:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int kVariant, int kUnroll>
static void run_issue(int sms, long long *d_cycles, long long *h)
{
    CHECK(cudaFuncSetAttribute(probe_issue<kVariant, kUnroll>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int n : {64, 128, 256}) {
        const int iters = 1998 / kUnroll * kUnroll;
        const size_t smem = (size_t)kUnroll * (128 * 128 + 256 * 128) + 1024;
        probe_issue<kVariant, kUnroll><<<sms, 128, smem>>>(n, iters, d_cycles);
        CHECK(cudaDeviceSynchronize());
        probe_issue<kVariant, kUnroll><<<sms, 128, smem>>>(n, iters, d_cycles);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(h, d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < sms; ++i) avg += (double)h[i];
        avg /= sms;
        printf("issue variant %d unroll %d  N %3d: %6.1f cycles per MMA (ideal %d)\n", kVariant, kUnroll, n, avg / (iters * 4.0), n / 2);
    }
}

// Does tcgen05.cp.128x256b of a SWIZZLE_128B K-major tile land in the layout the A-from-TMEM MMA expects (row = lane, K element k in
// 32-bit column k / 2)?  A tile holds the 16-bit pattern (row << 8 | k); 4 copies (K = 64), then every lane reads its 32 columns back.
__global__ void __launch_bounds__(128, 1) check_cp(int *mismatches, uint32_t *dump)
{
    extern __shared__ uint8_t raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {       // 16-byte chunks: row = i / 8, chunk j = i % 8 holds k = 8j .. 8j+7
        const int row = i >> 3, j = i & 7;
        uint16_t v[8];
        for (int e = 0; e < 8; ++e) v[e] = (uint16_t)((row << 8) | (j * 8 + e));
        *(uint4 *)(smem + row * 128 + ((j ^ (row & 7)) << 4)) = *(uint4 *)v;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (threadIdx.x == 0) {
        const uint64_t ad = make_desc(smem_u32(smem), 16, 1024, 2);
        for (int k = 0; k < 4; ++k) cp_128x256b(tmem + (uint32_t)(k * 8), ad + (uint64_t)(k * 2));
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred P1;\n\tW2:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D2;\n\tbra W2;\n\tD2:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = threadIdx.x;
    int bad = 0;
    for (int c0 = 0; c0 < 32; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int e = 0; e < 8; ++e) {
            const int k = (c0 + e) * 2;
            const uint32_t want = (uint32_t)((row << 8) | k) | ((uint32_t)((row << 8) | (k + 1)) << 16);
            if (r[e] != want) ++bad;
            if (row == 9) dump[c0 + e] = r[e];
        }
    }
    atomicAdd(mismatches, bad);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main()
{
    {
        int *d_bad; uint32_t *d_dump; int h_bad = -1; uint32_t h_dump[32];
        CHECK(cudaMalloc(&d_bad, 4)); CHECK(cudaMalloc(&d_dump, 128)); CHECK(cudaMemset(d_bad, 0, 4)); CHECK(cudaMemset(d_dump, 0, 128));
        CHECK(cudaFuncSetAttribute(check_cp, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        check_cp<<<1, 128, 32 * 1024>>>(d_bad, d_dump);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost)); CHECK(cudaMemcpy(h_dump, d_dump, 128, cudaMemcpyDeviceToHost));
        printf("check_cp (sw128 K-major tile -> TMEM, 128x256b x4): %d mismatching columns of 4096; row 9:", h_bad);
        for (int i = 0; i < 8; ++i) printf(" %08x", h_dump[i]);
        printf("\n");
    }
    int dev = 0, sms = 0;
    CHECK(cudaSetDevice(dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CHECK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long *d_cycles;
    CHECK(cudaMalloc(&d_cycles, sms * sizeof(long long)));
    long long *h = (long long *)malloc(sms * sizeof(long long));
    const char *lname[] = {"none", "", "sw128", "", "sw64", "", "sw32"};
    run_issue<0, 3>(sms, d_cycles, h);
    run_issue<1, 3>(sms, d_cycles, h);
    run_issue<2, 3>(sms, d_cycles, h);
    run_issue<1, 1>(sms, d_cycles, h);
    run_issue<2, 1>(sms, d_cycles, h);
    printf("layout a_src N grid cycles_per_mma ideal(N/2) us_total\n");
    for (int grid : {sms}) {
        for (int a_tmem = 0; a_tmem < 4; ++a_tmem) {
            for (int layout : {2}) {
                for (int n : {64, 128, 192, 256}) {
                    Cfg c;
                    c.n = n; c.layout = layout; c.a_tmem = a_tmem; c.iters = 2000; c.nslots = 3;
                    // one slot = a 128-row A tile and an n-row B tile, 64 K elements (128 B) per row, K-major
                    c.a_bytes = 128 * 128; c.b_bytes = n * 128;
                    for (int k = 0; k < 4; ++k) {
                        if (layout == 2) { c.a_off[k] = c.b_off[k] = 32 * k; }                                   // rows of 128 B, 8-row atoms of 1024 B
                        else if (layout == 4) { c.a_off[k] = (k >> 1) * 128 * 64 + (k & 1) * 32; c.b_off[k] = (k >> 1) * n * 64 + (k & 1) * 32; }   // [K half][rows][64 B]
                        else { c.a_off[k] = k * 128 * 32; c.b_off[k] = k * n * 32; }                             // [K step][rows][32 B]
                    }
                    if (layout == 2) { c.lbo = 16; c.sbo = 1024; }
                    else if (layout == 4) { c.lbo = 16; c.sbo = 512; }
                    else if (layout == 6) { c.lbo = 16; c.sbo = 256; }
                    else { c.lbo = 128; c.sbo = 256; }      // no swizzle: 8 x 16 B core matrices, the two K halves 128 B apart, 8-row groups 256 B apart
                    cudaEvent_t e0, e1;
                    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
                    probe<<<grid, 128, 3 * (c.a_bytes + c.b_bytes) + 1024>>>(c, d_cycles);   // warm-up
                    CHECK(cudaDeviceSynchronize());
                    CHECK(cudaEventRecord(e0));
                    probe<<<grid, 128, 3 * (c.a_bytes + c.b_bytes) + 1024>>>(c, d_cycles);
                    CHECK(cudaEventRecord(e1));
                    CHECK(cudaDeviceSynchronize());
                    float ms = 0;
                    CHECK(cudaEventElapsedTime(&ms, e0, e1));
                    CHECK(cudaMemcpy(h, d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                    double avg = 0;
                    for (int i = 0; i < grid; ++i) avg += (double)h[i];
                    avg /= grid;
                    printf("%-6s %-4s %3d %3d %8.1f %5d %9.1f\n", lname[layout], (a_tmem == 0 ? "smem" : a_tmem == 1 ? "tmem" : a_tmem == 2 ? "cp+ts" : "cp"), n, grid, avg / (c.iters * 4.0), n / 2, ms * 1000.0);
                }
            }
        }
    }
    return 0;
}
