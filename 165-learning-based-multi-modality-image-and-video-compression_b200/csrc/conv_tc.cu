// conv_tc.cu -- tensor-core implicit-GEMM convolution / transposed convolution for sm_100a.
//
// Replaces conv()/deconv() (compressai/models/utils.py:128-146: nn.Conv2d / nn.ConvTranspose2d,
// k in {1,3,5}, stride in {1,2}, padding k/2, output_padding stride-1) fused with what follows them in
// the transform stacks: bias, ReLU / LeakyReLU (compressai/models/google.py:254-269,363-377) and
// GDN / IGDN (compressai/layers/gdn.py:77-92), plus the |y| copy that feeds h_a (google.py:283).
//
// GEMM view: M = output pixels (128 per tile: a TH x TW patch of one image), N = Cout (<= 256 per
// pass), K = taps x Cin walked as (tap, 64-channel chunk) blocks.
//   * A operand: NHWC bf16 activations.  One 4-D TMA box {64 ch, TW, TH, 1} per K block lands in
//     shared memory as 128 rows x 128 B with the 128B swizzle, i.e. directly in the canonical
//     K-major UMMA layout.  Stride-2 convolutions use the tensor map's element strides, borders
//     (padding) are TMA out-of-bounds zero fill -- there is no im2col buffer and no halo logic.
//   * A transposed convolution is split into its stride^2 output phases; every phase is a stride-1
//     correlation over the taps of matching parity (3x3 / 3x2 / 2x3 / 2x2 for k=5), so no zero
//     insertion and no wasted MACs.
//   * B operand: weights pre-packed [tap][Cout][Cin] bf16, one 2-D TMA box {64, Ntile} per K block.
//   * tcgen05.mma (M=128, N=Ntile, K=16) issued by one elected thread, fp32 accumulators in TMEM in a ring of 2-4 stages so
//     that the epilogue of tile i overlaps the main loops of the next tiles.  CTA-pair variant (kPair): a 2-CTA cluster takes two
//     adjacent tiles, one cta_group::2 MMA (M=256) spans both SMs and each CTA supplies half of the weight rows.
//   * Epilogue (8 warps, two per TMEM lane quarter, one accumulator row = one pixel per thread): tcgen05.ld -> bias -> act ->
//     [GDN] -> bf16 / fp32 NHWC stores.  GDN is a second tensor-core contraction: the epilogue squares the activations into a
//     swizzled bf16 tile in shared memory, one thread issues D2[128 x C] = X2[128 x C] * gamma^T with gamma resident in
//     shared memory, and the result is combined as x * rsqrt(beta + D2) (sqrt for IGDN).
//   * Image-edge layers (Cin <= 8, MODE_PAD8) and the reconstruction layer (Cout <= 4, MODE_SCATTER: GEMM + col2im with two
//     independent epilogue teams) are described at their plan / epilogue code below.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2-9 = epilogue.
// Persistent grid: one CTA per SM (pairs: one cluster per SM pair), tiles strided across CTAs.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace mmc {

// warp 0 TMA, warp 1 MMA, then 4 * kParts epilogue warps: kParts warps per TMEM lane quarter split the accumulator columns
// kTeams = 2 (GDN epilogue, C <= 128): two such groups of epilogue warps work on alternate tiles, see epilogue_gdn
// Service warps: 0 = TMA producer (A patches in grouped mode), 1 = MMA issuer, 2 = weight-tile producer and 3 = second MMA issuer
// (both grouped mode only); the epilogue warps follow.
constexpr int kSvcWarps = 4;
constexpr int tc_threads(int parts, int teams = 1) { return 32 * kSvcWarps + 128 * parts * teams; }
// Pair kernel: the GDN norm contraction stays a per-CTA (cta_group::1) MMA on the CTA's own x^2 tile and a full copy of gamma, so the two
// epilogues of a pair never wait for each other (a pair-wide GDN MMA needs a cross-CTA hand-shake per tile and was slower).
constexpr int kMaxStages = 8;
constexpr int kMaxAccStages = 4;
constexpr int kMaxTaps = 32;
constexpr int kMaxCout = 1024;
constexpr int kABytes = 128 * 128;  // 128 pixels x 64 bf16

struct Tap {
    int16_t dy, dx;  // offset of the input patch for this tap (in A-grid units)
    int32_t brow;    // first row of this tap in the packed weight matrix
};

// How the layer is mapped onto the GEMM machinery
enum TcMode {
    MODE_STD = 0,     // NHWC bf16 input, K block = (tap, 64-channel chunk)
    MODE_PAD8 = 1,    // image-edge conv (Cin <= 8): zero-padded NHWC8 input, K block = one kernel row (8 px x 8 ch)
    MODE_SCATTER = 2  // narrow transposed conv as GEMM + col2im: N = k*k*Cout products per INPUT pixel, summed
                      // into output pixels through shared memory (no tap loop, every activation is read once)
};

// Tap group (MODE_STD, P.grouped): taps whose A tiles are row shifts of ONE patch.  The patch -- (TH + ntaps - 1) tile rows of TW = 8
// pixels, 64 channels, i.e. one 1024-byte swizzle atom per tile row -- is loaded once per 64-channel chunk; tap j reads it from row
// j on: a whole-atom offset of the UMMA descriptor, so the K-major SWIZZLE_128B layout stays intact.  Taps of a transposed
// convolution phase with the same dx, resp. taps of a strided convolution with the same kx and ky of the same parity (the rows of
// the patch are then every stride-th input row: TMA element strides), form a group.  Round 2: the per-SM L2 -> SM ingest
// (64 B / clk) bounded every large layer (32 KB of operands per 256 cycles of MMA); grouping cuts the A traffic 2-2.5x.
constexpr int kMaxGroups = 16;
struct Group {
    int16_t ox, oy;      // patch origin relative to the tile origin (A-grid units)
    uint8_t ntaps;       // taps reading this patch, at tile-row offsets 0 .. ntaps - 1
    uint8_t tap0;        // first entry of this group in gbrow[]
};

// n / d for 0 <= n < 2^31 without an integer division (the tile-index decomposition runs once per tile in every warp role; ncu source
// view of g_a.0: the four divisions of TileIter::init were ~10 % of the epilogue's issue slots).  Granlund-Montgomery round-up method.
struct FastDiv {
    uint32_t mul, shift, d;
    __host__ void set(int div)
    {
        d = (uint32_t)div;
        if (div <= 1) { mul = 0; shift = 0; return; }
        int l = 0;
        while ((1u << l) < d) ++l;                    // ceil(log2 d)
        const int p = 31 + l;
        mul = (uint32_t)((((uint64_t)1 << p) + d - 1) / d);
        shift = (uint32_t)(p - 32);
    }
    __device__ __forceinline__ int div(int n) const { return d <= 1 ? n : (int)(__umulhi((uint32_t)n, mul) >> shift); }
};

struct TcParams {
    CUtensorMap tmA, tmB, tmG, tmA2;   // tmA2: second activation source (channels Cin1 .. Cin-1), see mmc_conv_forward_tc2
    Tap taps[kMaxTaps];
    int phase_begin[5];  // taps of phase p are [phase_begin[p], phase_begin[p+1])
    Group groups[kMaxGroups];
    int32_t gbrow[kMaxTaps];   // first weight row of every grouped tap, group-major
    int group_begin[5];  // groups of phase p are [group_begin[p], group_begin[p+1])
    int grouped;         // 1: A-patch ring + B ring (producer / issuer loops below); 0: one (A tile, B tile) stage per K block
    int a_stage_bytes;   // grouped: bytes of one A patch (patch rows * 1024)
    int nA, nB;          // grouped: ring depths (per issuer)
    int issuers;         // grouped: 1 or 2 MMA-issuing threads (warps 1 and 3), each with its own A / B rings
    int n_phases;
    int mode;
    int a_sx, a_sy;  // A-box start = tile origin * (a_sx, a_sy) + tap offset
    int out_stride;  // output pixels per grid cell (deconv: stride; conv: 1)
    int B, Gh, Gw;   // per-phase pixel grid
    int Ho, Wo;
    int TH, TW, tiles_y, tiles_x;
    int step_y, step_x, off_y, off_x;   // tile origin = tile index * step - off (MODE_SCATTER tiles overlap by the halo)
    int k, pad, halo_lo, halo_hi;       // MODE_SCATTER geometry
    int spitch;                          // MODE_SCATTER: fp32 staging row pitch (floats)
    int Cout, Ntile, n_blocks;
    int kchunks;     // K boxes per tap (ceil(Cin / 64); 1 in MODE_PAD8)
    int kchunks1;    // K boxes that come from the first activation source (== kchunks with a single source)
    int ksteps;      // K=16 MMA steps per K box (4; 3 for a 5-wide kernel row in MODE_PAD8)
    int num_stages, acc_stages;
    int a_tmem;       // GDN: the x^2 operand of the norm contraction lives in TMEM (A-from-TMEM MMA), not in shared memory
    int bias_mma;     // the bias enters the accumulator through a constant-operand MMA (one N block per tile), not in the epilogue
    int late_release; // GDN epilogue: the accumulator stage is handed back by the thread that issues the norm MMAs, after issuing them
    int direct_store; // GDN epilogue: 32-byte vector stores straight from registers instead of the shared-memory staged copy-out
    int debug;       // profiling aid (env MMC_TC_DEBUG): 1 = no TMA traffic after priming, 2 = no main-loop MMAs, 3 = no GDN MMAs
    int pair;        // 1: CTA-pair kernel (cta_group::2): two adjacent tiles per MMA, each CTA holds half of the B rows
    int b_resident;  // whole packed weight matrix stays in shared memory (small layers); K blocks stream A only
    int act, gdn, out_f32, out2;
    int gdn_chunk;
    int tiles_per_phase, total_tiles;
    int epi_pipe;        // GDN epilogue software-pipelined over two tiles (epilogue_gdn_pipe)
    FastDiv fd_tpp, fd_nb, fd_tx, fd_ty, fd_grid, fd_nph;   // divisions of the tile-index decomposition
    int phase_inner;     // transposed convolutions: the stride^2 output phases of a wave of spatial tiles run back to back (see map_tile)
    int n_virtual;       // bound of the persistent loops' running index (== total_tiles unless phase_inner)
    int st_nb, st_tx, st_ty, st_b, st_ph;   // gridDim.x decomposed in the radices (n_blocks, tiles_x, tiles_y, B, phase)
    const float *bias;
    const float *beta;
    void *y;
    __nv_bfloat16 *y2;
};

__device__ __forceinline__ float act_tc(float v, int act)
{
    if (act == MMC_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == MMC_ACT_LEAKY_RELU) return v > 0.0f ? v : 0.01f * v;
    if (act == MMC_ACT_QRELU8) return fminf(fmaxf(v, 0.0f), 255.0f);
    return v;
}
// MUFU.RSQ without the denormal fix-up sequence of rsqrtf(): the GDN norm is beta + sum(gamma x^2) >= beta_min > 0
__device__ __forceinline__ float rsqrt_fast(float v)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
// D[tmem] (+)= A[tmem] * B[smem]^T: A rows = TMEM lanes, K elements packed two per 32-bit column
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ float sqrt_fast(float v)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}

struct TileCoord {
    int phase, b, y0, x0, n0;
};
// Tile index -> (phase, image, tile row, tile column, N block).  The persistent loop advances by gridDim.x tiles per
// iteration; instead of four integer divisions per tile the coordinates are kept as a mixed-radix counter and the
// step (decomposed on the host into the same radices) is added with carries.
__device__ __forceinline__ TileCoord make_coord(const TcParams &P, int phase, int b, int ty, int tx, int nb)
{
    TileCoord t;
    t.phase = phase; t.b = b;
    t.y0 = ty * P.step_y - P.off_y;
    t.x0 = tx * P.step_x - P.off_x;
    t.n0 = nb * P.Ntile;
    return t;
}
struct TileIter {
    int nb, tx, ty, b, phase;
    __device__ __forceinline__ void init(const TcParams &P, int tile)
    {
        phase = P.fd_tpp.div(tile);
        int r = tile - phase * P.tiles_per_phase, q;
        q = P.fd_nb.div(r); nb = r - q * P.n_blocks; r = q;
        q = P.fd_tx.div(r); tx = r - q * P.tiles_x;  r = q;
        q = P.fd_ty.div(r); ty = r - q * P.tiles_y;
        b = q;
    }
    __device__ __forceinline__ void advance(const TcParams &P)
    {
        nb += P.st_nb;
        int c = nb >= P.n_blocks; nb -= c ? P.n_blocks : 0;
        tx += P.st_tx + c;
        c = tx >= P.tiles_x; tx -= c ? P.tiles_x : 0;
        ty += P.st_ty + c;
        c = ty >= P.tiles_y; ty -= c ? P.tiles_y : 0;
        b += P.st_b + c;
        c = b >= P.B; b -= c ? P.B : 0;
        phase += P.st_ph + c;
    }
    __device__ __forceinline__ TileCoord coord(const TcParams &P) const { return make_coord(P, phase, b, ty, tx, nb); }
};

// Persistent-loop index -> tile.  Default: tile = index, i.e. phase-major -- every output phase of a transposed convolution sweeps the
// whole input, which is therefore read from HBM stride^2 times (round-2 ncu: g_s.4 reads 1.61 GB for a 0.40 GB input).  With
// P.phase_inner iteration k of CTA c is (wave w = k / n_phases, phase p = k % n_phases) -> spatial tile w * gridDim + c of phase p:
// all CTAs run the same phase at the same time (same work per CTA, the tap counts differ between phases), a CTA pair keeps two
// adjacent spatial tiles of ONE phase, and the 148 input patches of a wave are re-read from L2 by the three following phases.
// Returns false once the CTA has run out of spatial tiles (only ever at the end of its loop: every role breaks at the same index).
__device__ __forceinline__ bool map_tile(const TcParams &P, int v, int &tile)
{
    tile = v;
    if (!P.phase_inner) return true;
    const int k = P.fd_grid.div(v);
    const int w = P.fd_nph.div(k), p = k - w * P.n_phases;
    const int sp = w * (int)gridDim.x + (int)blockIdx.x;
    tile = p * P.tiles_per_phase + sp;
    return sp < P.tiles_per_phase;
}

__device__ __forceinline__ void load16f(const float *sm, float *o)
{
    const float4 *p = reinterpret_cast<const float4 *>(sm);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float4 t = p[i];
        o[4 * i] = t.x; o[4 * i + 1] = t.y; o[4 * i + 2] = t.z; o[4 * i + 3] = t.w;
    }
}

// 16 consecutive output channels of one pixel -> global memory (bf16 or fp32 NHWC) [+ bf16 secondary]
__device__ __forceinline__ void store16(const TcParams &P, int64_t off, const float *v)
{
    if (P.out_f32) {
        float4 *dst = reinterpret_cast<float4 *>((float *)P.y + off);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
        uint4 *dst = reinterpret_cast<uint4 *>((__nv_bfloat16 *)P.y + off);
        dst[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        dst[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
    }
    if (P.out2 == 1 || P.out2 == 2) {
        float w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = (P.out2 == 1) ? fabsf(v[i]) : v[i];
        uint4 *dst = reinterpret_cast<uint4 *>(P.y2 + off);
        dst[0] = make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
        dst[1] = make_uint4(pack_bf16(w[8], w[9]), pack_bf16(w[10], w[11]), pack_bf16(w[12], w[13]), pack_bf16(w[14], w[15]));
    }
}

// Constant operands that put the bias (and the GDN beta') into the accumulators through the tensor core instead of per-element
// shared-memory loads and adds in the epilogue (measured: the 32 LDS.128 + 128 FADD per thread and tile cost 13 % of g_s.4):
//   A "ones" tile [128 rows][K = 16]: columns 0 and 1 are 1.0;  B tile [rows][K = 16]: column 0 = bf16(v), column 1 = bf16(v - hi),
// so that one extra K = 16 MMA adds v (to ~2^-17 relative) to every accumulator row.  K-major, NO swizzle: 8-row x 16-byte core
// matrices, the two K halves 128 B apart (LBO), 8-row groups 256 B apart (SBO).
constexpr int kConstRowsMax = 192;
__device__ __forceinline__ uint64_t make_desc_ns(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(128 >> 4) << 16;    // leading byte offset: second K half
    d |= (uint64_t)(256 >> 4) << 32;    // stride byte offset: next 8 rows
    d |= (uint64_t)1 << 46;
    return d;                           // layout type 0: no swizzle
}
__device__ __forceinline__ void fill_const_tile(uint8_t *tile, int rows, const float *v, bool ones)
{
    // one thread per row: zero the row's 2 x 16 bytes, then set K columns 0 / 1
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        uint8_t *row = tile + (r >> 3) * 256 + (r & 7) * 16;
        float hi_f = 1.0f, lo_f = 1.0f;
        if (!ones) {
            const float val = v ? v[r] : 0.0f;
            const __nv_bfloat16 hi = __float2bfloat16_rn(val);
            hi_f = __bfloat162float(hi);
            lo_f = val - hi_f;
        }
        *reinterpret_cast<uint4 *>(row) = make_uint4(pack_bf16(hi_f, lo_f), 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(row + 128) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// ---------------------------------------------------------------------------------------------
// Fused GDN / IGDN epilogue (compressai/layers/gdn.py:77-92) for one 128-pixel x C-channel tile.
// Each of the 256 epilogue threads owns one pixel row and NCH 16-column chunks (C/2 channels); the
// biased activations stay in registers across the tensor-core norm contraction.
//   pass 1: x = acc + bias (all TMEM loads issued before one wait); x^2 -> bf16 swizzled K-major smem tile
//   MMA   : norm[128 x chunk] = X2[128 x C] * gamma[chunk rows]^T      (one thread issues, tcgen05.commit -> mbarrier)
//   pass 2: y = x * rsqrt(beta + norm)   (IGDN: x * sqrt(n) = x * n * rsqrt(n)); bf16 / fp32 NHWC stores
// ---------------------------------------------------------------------------------------------
struct GdnCtx {
    const TcParams &P;
    uint8_t *sA2, *sG;
    uint64_t *gdn_bar, *gload_bar;
    uint32_t tmem_base, acc_addr, lane_addr, norm_col;
    int row, half;   // half = which of the kParts column parts this thread owns
    bool valid;
    int64_t pix_off;
    int it;
    const int64_t *pix_off_s;   // per-pixel output offsets of this tile (-1: outside the image)
    uint32_t rank;              // pair mode: CTA rank in the cluster
    uint32_t ones, beta_tile;   // shared-memory addresses of the constant operands (see fill_const_tiles)
    uint32_t a_col;             // TMEM column of the x^2 operand (a_tmem mode)
    uint64_t *empty_bar;        // this tile's accumulator-release barrier (pair mode: the leader's, as a cluster address)
    uint32_t empty_leader;
    uint32_t bar_id;            // named barrier of this epilogue team (1 + team)
    bool first_warp;            // the team's first warp issues the norm MMAs
};

template <int NCH, int G, bool kPair, int kParts, int kTeams>
__device__ __forceinline__ void epilogue_gdn(const GdnCtx &g, const float *bias_s, const float *beta_s, uint32_t &gdn_phase)
{
    const TcParams &P = g.P;
    const bool inverse = P.gdn == MMC_GDN_INVERSE;
    constexpr int per = NCH / G;               // chunks of each norm group owned by this thread (its half of the group)
    constexpr int gch = kParts * per;          // 16-column chunks per norm group (gdn_chunk / 16)
    constexpr int kEpiThreads = 128 * kParts;
    float x[NCH][16];
    // chunk id of my j-th chunk: group (j / per), position (j % per) within my half of the group
    auto chunk_of = [&](int j) { return (j / per) * gch + g.half * per + (j % per); };
#pragma unroll
    for (int j = 0; j < NCH; ++j) tmem_ld16(g.acc_addr + (uint32_t)(chunk_of(j) << 4), x[j]);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int c0 = chunk_of(j) << 4;
        uint32_t pk[8];      // (the bias is already in the accumulator: see the constant-operand MMAs below)
        if (P.out2 == 3 && g.valid) {
            // training: keep the pre-GDN activations (bf16) for the backward pass
            uint4 *dst = reinterpret_cast<uint4 *>(P.y2 + g.pix_off + c0);
            dst[0] = make_uint4(pack_bf16(x[j][0], x[j][1]), pack_bf16(x[j][2], x[j][3]), pack_bf16(x[j][4], x[j][5]), pack_bf16(x[j][6], x[j][7]));
            dst[1] = make_uint4(pack_bf16(x[j][8], x[j][9]), pack_bf16(x[j][10], x[j][11]), pack_bf16(x[j][12], x[j][13]), pack_bf16(x[j][14], x[j][15]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(x[j][2 * i] * x[j][2 * i], x[j][2 * i + 1] * x[j][2 * i + 1]);
        if (P.a_tmem) {
            // A operand in TMEM: row = lane, K element k in 32-bit column k / 2 (two bf16 per column)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(g.tmem_base + g.lane_addr + g.a_col + (uint32_t)(c0 >> 1)),
                         "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
        } else {
            uint8_t *tile_base = g.sA2 + (size_t)(c0 >> 6) * kABytes + (size_t)g.row * 128;
            const int j0 = (c0 & 63) >> 3;   // 16-byte chunk index inside the 128-byte row
            *reinterpret_cast<uint4 *>(tile_base + (((j0) ^ (g.row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4 *>(tile_base + (((j0 + 1) ^ (g.row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
    }
    if (P.a_tmem) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
    // x is in registers (and x^2 staged): the conv accumulator can go back to the MMA warp -- but only AFTER the norm MMAs below
    // have been issued (P.late_release).  The tensor pipe executes in issue order: handing the accumulator back first (one arrival
    // per warp, right here) lets the MMA warp queue the whole main loop of the tile after next in front of this tile's norm
    // contraction, which then waits for all of it -- the epilogue and the main loop ran back to back instead of overlapped
    // (round-2 finding: g_s.4 8.3k cycles per tile = 6.1k main loop + 2.3k epilogue).
    tc_fence_before();
    __syncwarp();
    if (kTeams == 1 && !P.late_release && (threadIdx.x & 31) == 0) {
        if (kPair) mbar_arrive_cluster(g.empty_leader);
        else mbar_arrive(g.empty_bar);
    }
    asm volatile("bar.sync %0, %1;" ::"r"(g.bar_id), "n"(kEpiThreads) : "memory");
#pragma unroll
    for (int grp = 0; grp < G; ++grp) {
        const int g0 = grp * gch * 16;
        if (g.first_warp) {
            // first warp of the team: uniform control flow, one elected lane issues the (compile-time unrolled) MMAs
            if (g.it == 0 && grp == 0) mbar_wait(g.gload_bar, 0);
            tc_fence_after();
            const uint32_t idesc = make_idesc(gch * 16);
            const uint32_t a2 = smem_u32(g.sA2), gm = smem_u32(g.sG) + (uint32_t)(g0 * 128);
            if (elect_one()) {
#pragma unroll
                for (int kc = 0; kc < (P.debug == 3 ? 0 : NCH * kParts / 4); ++kc) {   // debug 3: profiling, no norm MMAs
                    const uint64_t adesc = make_desc(a2 + (uint32_t)(kc * kABytes));
                    const uint64_t bdesc = make_desc(gm + (uint32_t)(kc * P.Cout * 128));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (P.a_tmem) tc_mma_ts(g.tmem_base + g.norm_col, g.tmem_base + g.a_col + (uint32_t)((kc * 4 + k) * 8), bdesc + (uint64_t)(k * 2), idesc, (kc | k) != 0);
                        else tc_mma(g.tmem_base + g.norm_col, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kc | k) != 0);
                    }
                }
                // + 1 * beta': one K = 16 step against the constant operands (ones x [beta_hi, beta_lo, 0...])
                tc_mma(g.tmem_base + g.norm_col, make_desc_ns(g.ones), make_desc_ns(g.beta_tile + (uint32_t)((g0 >> 3) * 256)), idesc, 1);
                tc_commit(g.gdn_bar);
                if (kTeams == 1 && P.late_release && grp == 0) {
                    // every epilogue thread finished reading the accumulator before the bar.sync above; the norm MMAs are queued:
                    // now the main loop of the tile after next may follow them into the pipe (barrier count: see the kernel prologue)
                    if (kPair) mbar_arrive_cluster(g.empty_leader);
                    else mbar_arrive(g.empty_bar);
                }
            }
            __syncwarp();
        }
        mbar_wait(g.gdn_bar, gdn_phase);
        gdn_phase ^= 1;
        tc_fence_after();
        // my chunks of this group: j in [grp*per, (grp+1)*per), two norm chunks in flight at a time (register budget)
#pragma unroll
        for (int jb = 0; jb < per; jb += 2) {
            float nrm[2][16];
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (jb + u < per)
                    tmem_ld16(g.tmem_base + g.lane_addr + g.norm_col + (uint32_t)((chunk_of(grp * per + jb + u) << 4) - g0), nrm[u]);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (jb + u >= per) continue;
                const int j = grp * per + jb + u;
                const int c0 = chunk_of(j) << 4;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float n = nrm[u][i];          // beta' + gamma' x^2: beta' entered through the constant-operand MMA
                    x[j][i] *= inverse ? sqrt_fast(n) : rsqrt_fast(n);   // one MUFU either way
                }
                const float *y = x[j];
                if (G == 1 && !P.out_f32 && !P.out2 && P.direct_store) {
                    // one 256-bit store per 16-channel chunk: a full 32-byte sector, no staging round trip and no extra barriers
                    if (g.valid) {
                        __nv_bfloat16 *dst = (__nv_bfloat16 *)P.y + g.pix_off + c0;
                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst),
                                     "r"(pack_bf16(y[0], y[1])), "r"(pack_bf16(y[2], y[3])), "r"(pack_bf16(y[4], y[5])), "r"(pack_bf16(y[6], y[7])),
                                     "r"(pack_bf16(y[8], y[9])), "r"(pack_bf16(y[10], y[11])), "r"(pack_bf16(y[12], y[13])), "r"(pack_bf16(y[14], y[15]))
                                     : "memory");
                    }
                } else if (G == 1 && !P.out_f32 && !P.out2) {
                    // stage the bf16 result in the (now idle) x^2 tile, same swizzled [pixel][channel] layout
                    uint8_t *tile_base = g.sA2 + (size_t)(c0 >> 6) * kABytes + (size_t)g.row * 128;
                    const int j0 = (c0 & 63) >> 3;
                    *reinterpret_cast<uint4 *>(tile_base + (((j0) ^ (g.row & 7)) << 4)) =
                        make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
                    *reinterpret_cast<uint4 *>(tile_base + (((j0 + 1) ^ (g.row & 7)) << 4)) =
                        make_uint4(pack_bf16(y[8], y[9]), pack_bf16(y[10], y[11]), pack_bf16(y[12], y[13]), pack_bf16(y[14], y[15]));
                } else if (g.valid) {
                    store16(P, g.pix_off + c0, y);
                }
            }
        }
        // every epilogue thread is done with the norm columns (and, after the last group, with sA2)
        tc_fence_before();
        asm volatile("bar.sync %0, %1;" ::"r"(g.bar_id), "n"(kEpiThreads) : "memory");
    }
    if (G == 1 && !P.out_f32 && !P.out2 && !P.direct_store) {
        // Coalesced copy-out: consecutive lanes move consecutive 16-byte chunks of one pixel, so every warp store
        // covers whole 128-byte lines (the per-row direct stores touch 32 lines per instruction).
        constexpr int cpp = NCH * kParts * 2;      // 16-byte chunks per pixel (C / 8): 8 or 16
        constexpr int ppi = kEpiThreads / cpp;     // pixels covered per iteration (32 or 16, a multiple of 8)
        const int et = threadIdx.x - 32 * kSvcWarps;
        const int j = et % cpp, p0 = et / cpp;     // this thread always moves chunk j; pixel p0 + it * ppi
        const uint8_t *src = g.sA2 + (size_t)(j >> 3) * kABytes + (size_t)p0 * 128 + (((j & 7) ^ (p0 & 7)) << 4);
        __nv_bfloat16 *yo = (__nv_bfloat16 *)P.y + j * 8;
#pragma unroll
        for (int i = 0; i < 128 / ppi; ++i) {
            const int64_t off = g.pix_off_s[p0 + i * ppi];
            const uint4 v = *reinterpret_cast<const uint4 *>(src + (size_t)i * ppi * 128);
            if (off >= 0) *reinterpret_cast<uint4 *>(yo + off) = v;
        }
        asm volatile("bar.sync %0, %1;" ::"r"(g.bar_id), "n"(kEpiThreads) : "memory");   // staging tile free for the next tile's x^2
    }
}

// ---------------------------------------------------------------------------------------------
// Two-team GDN / IGDN epilogue (single-CTA kernel, C in {64, 128}).  With one team the epilogue is a serial chain per tile -- TMEM
// load of x (64 B / clk: 1024 cycles for 128 x 128 fp32), square, stage, barrier, norm-MMA round trip through the busy tensor pipe,
// TMEM load of the norm (another 1024), MUFU, stores: 4.5-6.3k cycles -- and it, not the MMAs, bounded g_a.0 and g_s.* (ncu r02:
// 35 % / 61 % tensor-pipe active).  Here two teams of FOUR warps take alternate tiles so that one team's TMEM reads overlap the
// other's MMA wait / MUFU / stores.  One thread owns a whole pixel row: x (C fp32 values) stays in registers between the passes
// (a second TMEM read of x was measured slower than the single team -- TMEM read bandwidth is the scarce resource; 8-warp teams
// would need 18 warps = 96 registers per thread, not enough for 64 + working set).  TMEM: the x^2 operand goes through TMEM as
// before (C / 2 columns per team) and the norm is written IN PLACE over the tile's accumulator stage, which therefore returns to
// the MMA warp only at the end of the epilogue; three stages (one being filled, one per team): 3 C + 2 C / 2 = 512 for C = 128.
// Same arithmetic as epilogue_gdn (bf16 x^2 operand, fp32 accumulate, beta through the constant-operand MMA): bit-identical.
// ---------------------------------------------------------------------------------------------
template <int NCH>     // 16-column chunks per thread = C / 16
__device__ __forceinline__ void epilogue_gdn_rows(const GdnCtx &g, uint32_t &gdn_phase)
{
    const TcParams &P = g.P;
    const bool inverse = P.gdn == MMC_GDN_INVERSE;
    constexpr int kTeamThreads = 128;
    float x[NCH][16];
    // ---- pass 1: x -> registers, x^2 -> this team's TMEM operand block ----
#pragma unroll
    for (int j = 0; j < NCH; j += 2) {
        tmem_ld16(g.acc_addr + (uint32_t)(j << 4), x[j]);
        tmem_ld16(g.acc_addr + (uint32_t)((j + 1) << 4), x[j + 1]);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int c0 = (j + u) << 4;
            const float *xv = x[j + u];
            if (P.out2 == 3 && g.valid) {
                // training: keep the pre-GDN activations (bf16) for the backward pass
                uint4 *dst = reinterpret_cast<uint4 *>(P.y2 + g.pix_off + c0);
                dst[0] = make_uint4(pack_bf16(xv[0], xv[1]), pack_bf16(xv[2], xv[3]), pack_bf16(xv[4], xv[5]), pack_bf16(xv[6], xv[7]));
                dst[1] = make_uint4(pack_bf16(xv[8], xv[9]), pack_bf16(xv[10], xv[11]), pack_bf16(xv[12], xv[13]), pack_bf16(xv[14], xv[15]));
            }
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(xv[2 * i] * xv[2 * i], xv[2 * i + 1] * xv[2 * i + 1]);
            // A operand in TMEM: row = lane, K element k in 32-bit column k / 2 (two bf16 per column)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(g.tmem_base + g.lane_addr + g.a_col + (uint32_t)(c0 >> 1)),
                         "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    asm volatile("bar.sync %0, %1;" ::"r"(g.bar_id), "n"(kTeamThreads) : "memory");
    if (g.first_warp) {
        // the whole team has read the accumulator: the norm may overwrite it
        if (g.it == 0) mbar_wait(g.gload_bar, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc(NCH * 16);
        const uint32_t gm = smem_u32(g.sG);
        if (elect_one()) {
#pragma unroll
            for (int kc = 0; kc < (P.debug == 3 ? 0 : NCH / 4); ++kc) {   // debug 3: profiling, no norm MMAs
                const uint64_t bdesc = make_desc(gm + (uint32_t)(kc * P.Cout * 128));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_ts(g.tmem_base + g.norm_col, g.tmem_base + g.a_col + (uint32_t)((kc * 4 + k) * 8), bdesc + (uint64_t)(k * 2), idesc, (kc | k) != 0);
            }
            // + 1 * beta': one K = 16 step against the constant operands (ones x [beta_hi, beta_lo, 0...])
            tc_mma(g.tmem_base + g.norm_col, make_desc_ns(g.ones), make_desc_ns(g.beta_tile), idesc, 1);
            tc_commit(g.gdn_bar);
        }
        __syncwarp();
    }
    mbar_wait(g.gdn_bar, gdn_phase);
    gdn_phase ^= 1;
    tc_fence_after();
    // ---- pass 2: y = x * rsqrt(beta' + gamma' x^2)  (IGDN: * sqrt), one MUFU per value, 256-bit stores ----
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int c0 = j << 4;
        float nrm[16];
        tmem_ld16(g.tmem_base + g.lane_addr + g.norm_col + (uint32_t)c0, nrm);
        tmem_ld_wait();
        float *y = x[j];
#pragma unroll
        for (int i = 0; i < 16; ++i) y[i] *= inverse ? sqrt_fast(nrm[i]) : rsqrt_fast(nrm[i]);
        if (g.valid) {
            if (!P.out_f32 && !(P.out2 == 1 || P.out2 == 2)) {
                __nv_bfloat16 *dst = (__nv_bfloat16 *)P.y + g.pix_off + c0;
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst),
                             "r"(pack_bf16(y[0], y[1])), "r"(pack_bf16(y[2], y[3])), "r"(pack_bf16(y[4], y[5])), "r"(pack_bf16(y[6], y[7])),
                             "r"(pack_bf16(y[8], y[9])), "r"(pack_bf16(y[10], y[11])), "r"(pack_bf16(y[12], y[13])), "r"(pack_bf16(y[14], y[15]))
                             : "memory");
            } else {
                store16(P, g.pix_off + c0, y);
            }
        }
    }
    // the norm (and with it the accumulator stage) has been consumed by every thread of the team
    tc_fence_before();
    asm volatile("bar.sync %0, %1;" ::"r"(g.bar_id), "n"(kTeamThreads) : "memory");
    if (g.first_warp && (threadIdx.x & 31) == 0) mbar_arrive(g.empty_bar);
}


// ---------------------------------------------------------------------------------------------
// Software-pipelined GDN / IGDN epilogue (single team, C <= 128, bf16 NHWC output without secondary outputs).
// epilogue_gdn is a serial chain per tile: read x, square, norm MMAs, WAIT for them (behind whatever the main loop has queued in the
// in-order tensor pipe: ~20 % of the chain, ncu source view of g_a.0), read the norm, scale, store; the chain -- 5.1k cycles per tile
// in g_a.0, 6.4k in g_s.4 -- bounds every fused layer whose main loop is shorter.  Here the wait of tile i is filled with pass 1 of
// tile i + 1:
//     pass 1 (i + 1): x -> bf16 pairs in registers (32 instead of 64 registers, which is what lets two tiles be in flight),
//                     x^2 -> TMEM operand block (i + 1) & 1
//     wait norm(i);  pass 2 (i): y = x * rsqrt(norm) (IGDN: * sqrt), 256-bit stores
//     hand-over (bar.arrive);  service warp 3 (norm_issuer_pipe) issues the norm MMAs (i + 1) and hands accumulator stage (i + 1) back
// TMEM: 2 accumulator stages + norm + two x^2 operand blocks = 2 C + C + 2 C / 2 = 512 columns at C = 128.  One non-blocking
// hand-over per tile instead of three team barriers.  x is rounded to bf16 BEFORE the scaling (the output is bf16 either way): |y - y_ref| <= 2^-8 |y| instead of
// 2^-9, inside the 1e-2 bf16 tolerance of BASELINE.json; the norm itself is computed from the fp32 x as before.
// Tried: the second warp of every TMEM lane quarter running the two passes in the opposite order (to keep the quarter's tcgen05.ld
// port busy while the other warp is in its MUFU / store phase) -- g_a.0 0.78 -> 1.2 ms: that warp starts with the norm wait.
// ---------------------------------------------------------------------------------------------
struct PipeTile {
    int64_t pix_off;
    int as;
    uint32_t aphase;
    bool valid;
};

template <int NCH, bool kPair, int kParts>
__device__ __forceinline__ void epilogue_gdn_pipe(const TcParams &P, uint32_t tmem_base, uint64_t *tmem_full_bar, uint64_t *gdn_bar)
{
    constexpr int kEpiThreads = 128 * kParts;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = (warp - kSvcWarps) >> 2;
    const int row = q * 32 + lane, th = row / P.TW, tw = row - th * P.TW;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t norm_col = (uint32_t)(P.acc_stages * P.Ntile), a_col0 = norm_col + (uint32_t)P.gdn_chunk;
    const bool inverse = P.gdn == MMC_GDN_INVERSE;
    uint32_t gdn_phase = 0;
    auto chunk_of = [&](int j) { return half * NCH + j; };

    // tile -> coordinates, accumulator stage (ring position of the it-th tile of this CTA)
    int acc_i = 0;
    uint32_t acc_ph = 0;
    TileIter ti;
    bool ti_started = false;
    auto setup = [&](int tile, PipeTile &t) {
        // consecutive tiles of a CTA are gridDim apart unless the phases are interleaved; a pair's padded last index must decode to
        // b == B (masked), which the carry chain of advance() does not do
        if (kPair || P.phase_inner || !ti_started) ti.init(P, tile);
        else ti.advance(P);
        ti_started = true;
        const TileCoord c = ti.coord(P);
        const int gy = c.y0 + th, gx = c.x0 + tw;
        const int py = P.out_stride == 2 ? (c.phase >> 1) : 0, px = P.out_stride == 2 ? (c.phase & 1) : 0;   // stride in {1, 2}
        t.valid = gy < P.Gh && gx < P.Gw && c.b < P.B;
        t.pix_off = (((int64_t)c.b * P.Ho + gy * P.out_stride + py) * P.Wo + gx * P.out_stride + px) * P.Cout + c.n0;
        t.as = acc_i; t.aphase = acc_ph;
        if (++acc_i == P.acc_stages) { acc_i = 0; acc_ph ^= 1; }
    };
    // pass 1: x -> packed bf16 registers, x^2 -> operand block r
    auto pass1 = [&](const PipeTile &t, uint32_t r, uint32_t (&xp)[NCH * 8]) {
        mbar_wait(&tmem_full_bar[t.as], t.aphase);
        tc_fence_after();
        const uint32_t acc = lane_base + (uint32_t)(t.as * P.Ntile), a_col = a_col0 + r * (uint32_t)(P.Cout / 2);
        // the loads of chunk pair j + 2 are in flight while pair j is squared and packed (tcgen05.wait::ld waits for ALL outstanding
        // loads, so the pipeline is two deep: load, wait, load next, compute, wait, compute)
        float x[2][2][16];
        tmem_ld16(acc + (uint32_t)(chunk_of(0) << 4), x[0][0]);
        tmem_ld16(acc + (uint32_t)(chunk_of(1) << 4), x[0][1]);
#pragma unroll
        for (int j = 0; j < NCH; j += 2) {
            const int b = (j >> 1) & 1;
            tmem_ld_wait();
            if (j + 2 < NCH) {
                tmem_ld16(acc + (uint32_t)(chunk_of(j + 2) << 4), x[b ^ 1][0]);
                tmem_ld16(acc + (uint32_t)(chunk_of(j + 3) << 4), x[b ^ 1][1]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int c0 = chunk_of(j + u) << 4;
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    pk[i] = pack_bf16(x[b][u][2 * i] * x[b][u][2 * i], x[b][u][2 * i + 1] * x[b][u][2 * i + 1]);
                    xp[(j + u) * 8 + i] = pack_bf16(x[b][u][2 * i], x[b][u][2 * i + 1]);
                }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + a_col + (uint32_t)(c0 >> 1)),
                             "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
    };
    // pass 2: y = x * rsqrt(norm) (IGDN: * sqrt), straight from registers to global memory
    auto pass2 = [&](const PipeTile &t, const uint32_t (&xp)[NCH * 8]) {
        mbar_wait(gdn_bar, gdn_phase);
        gdn_phase ^= 1;
        tc_fence_after();
        float nrm[2][2][16];
        tmem_ld16(lane_base + norm_col + (uint32_t)(chunk_of(0) << 4), nrm[0][0]);
        tmem_ld16(lane_base + norm_col + (uint32_t)(chunk_of(1) << 4), nrm[0][1]);
#pragma unroll
        for (int jb = 0; jb < NCH; jb += 2) {
            const int b = (jb >> 1) & 1;
            tmem_ld_wait();
            if (jb + 2 < NCH) {
                tmem_ld16(lane_base + norm_col + (uint32_t)(chunk_of(jb + 2) << 4), nrm[b ^ 1][0]);
                tmem_ld16(lane_base + norm_col + (uint32_t)(chunk_of(jb + 3) << 4), nrm[b ^ 1][1]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t w = xp[(jb + u) * 8 + i];
                    const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xffff0000u);
                    const float n0 = nrm[b][u][2 * i], n1 = nrm[b][u][2 * i + 1];
                    o[i] = pack_bf16(x0 * (inverse ? sqrt_fast(n0) : rsqrt_fast(n0)), x1 * (inverse ? sqrt_fast(n1) : rsqrt_fast(n1)));
                }
                if (t.valid) {
                    __nv_bfloat16 *dst = (__nv_bfloat16 *)P.y + t.pix_off + (chunk_of(jb + u) << 4);
                    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                                 "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
                }
            }
        }
        tc_fence_before();
    };
    // pass 1 of tile n chained with pass 2 of tile c: the first norm loads of c are issued before the last chunk pair of n is squared,
    // so the TMEM read port has work during that arithmetic as well
    auto square_pack = [&](const float (&xv)[2][16], int j, uint32_t a_col, uint32_t (&xp)[NCH * 8]) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int c0 = chunk_of(j + u) << 4;
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                pk[i] = pack_bf16(xv[u][2 * i] * xv[u][2 * i], xv[u][2 * i + 1] * xv[u][2 * i + 1]);
                xp[(j + u) * 8 + i] = pack_bf16(xv[u][2 * i], xv[u][2 * i + 1]);
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + a_col + (uint32_t)(c0 >> 1)),
                         "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
        }
    };
    auto scale_store = [&](const PipeTile &t, const float (&nv)[2][16], int jb, const uint32_t (&xp)[NCH * 8]) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t w = xp[(jb + u) * 8 + i];
                const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xffff0000u);
                const float n0 = nv[u][2 * i], n1 = nv[u][2 * i + 1];
                o[i] = pack_bf16(x0 * (inverse ? sqrt_fast(n0) : rsqrt_fast(n0)), x1 * (inverse ? sqrt_fast(n1) : rsqrt_fast(n1)));
            }
            if (t.valid) {
                __nv_bfloat16 *dst = (__nv_bfloat16 *)P.y + t.pix_off + (chunk_of(jb + u) << 4);
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                             "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
            }
        }
    };
    auto fused = [&](const PipeTile &c, const uint32_t (&xc)[NCH * 8], const PipeTile &n, uint32_t rn, uint32_t (&xn)[NCH * 8]) {
        mbar_wait(&tmem_full_bar[n.as], n.aphase);
        tc_fence_after();
        const uint32_t acc = lane_base + (uint32_t)(n.as * P.Ntile), a_col = a_col0 + rn * (uint32_t)(P.Cout / 2);
        const uint32_t nb = lane_base + norm_col;
        float buf[2][2][16];
        tmem_ld16(acc + (uint32_t)(chunk_of(0) << 4), buf[0][0]);
        tmem_ld16(acc + (uint32_t)(chunk_of(1) << 4), buf[0][1]);
        // ---- pass 1 (n): all chunk pairs but the last ----
#pragma unroll
        for (int j = 0; j + 2 < NCH; j += 2) {
            const int b = (j >> 1) & 1;
            tmem_ld_wait();
            tmem_ld16(acc + (uint32_t)(chunk_of(j + 2) << 4), buf[b ^ 1][0]);
            tmem_ld16(acc + (uint32_t)(chunk_of(j + 3) << 4), buf[b ^ 1][1]);
            square_pack(buf[b], j, a_col, xn);
        }
        constexpr int bl = ((NCH - 2) >> 1) & 1;       // buffer of the last x chunk pair
        tmem_ld_wait();
        // ---- norm(c) is needed from here on; its first chunk pair loads while the last x pair is squared ----
        mbar_wait(gdn_bar, gdn_phase);
        gdn_phase ^= 1;
        tc_fence_after();
        tmem_ld16(nb + (uint32_t)(chunk_of(0) << 4), buf[bl ^ 1][0]);
        tmem_ld16(nb + (uint32_t)(chunk_of(1) << 4), buf[bl ^ 1][1]);
        square_pack(buf[bl], NCH - 2, a_col, xn);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        // ---- pass 2 (c) ----
#pragma unroll
        for (int jb = 0; jb < NCH; jb += 2) {
            const int b = (bl ^ 1) ^ ((jb >> 1) & 1);
            tmem_ld_wait();
            if (jb + 2 < NCH) {
                tmem_ld16(nb + (uint32_t)(chunk_of(jb + 2) << 4), buf[b ^ 1][0]);
                tmem_ld16(nb + (uint32_t)(chunk_of(jb + 3) << 4), buf[b ^ 1][1]);
            }
            scale_store(c, buf[b], jb, xc);
        }
        tc_fence_before();
    };
    // The epilogue threads only ARRIVE: what the hand-over protects (x^2 of the next tile complete, its accumulator and the current
    // norm read by everybody) is needed by the thread that issues the next norm MMAs -- service warp 3 (norm_issuer_pipe), which
    // waits on the same named barrier.  No epilogue thread can arrive twice before the issuer has passed: its next arrival comes
    // after a wait for the norm that issuer produces.  (ncu: 11 % of the epilogue warps' time was the team barrier.)
    auto team_arrive = [&]() { asm volatile("bar.arrive 1, %0;" ::"n"(kEpiThreads + 32) : "memory"); };

    int v = blockIdx.x, tile;
    if (v >= P.n_virtual || !map_tile(P, v, tile)) return;
    PipeTile cur, nxt;
    uint32_t xa[NCH * 8], xb[NCH * 8];
    uint32_t r = 0;
    setup(tile, cur);
    pass1(cur, r, xa);
    team_arrive();
    for (;;) {
        // invariant: xa holds tile `cur`, whose norm MMAs have been issued from operand block r
        v += gridDim.x;
        const bool more = v < P.n_virtual && map_tile(P, v, tile);
        if (more) {
            setup(tile, nxt);
            fused(cur, xa, nxt, r ^ 1, xb);
        } else {
            pass2(cur, xa);
            break;
        }
        team_arrive();                        // norm(cur) consumed by this thread, its part of x^2(nxt) complete
        // roles swap: the compiler keeps xa / xb in registers because the loop is unrolled by two below
        v += gridDim.x;
        const bool more2 = v < P.n_virtual && map_tile(P, v, tile);
        if (more2) {
            setup(tile, cur);
            fused(nxt, xb, cur, r, xa);
        } else {
            pass2(nxt, xb);
            break;
        }
        team_arrive();
    }
}

// Service warp 3 of the pipelined GDN epilogue: per tile, wait for the epilogue threads' hand-over, issue the tile's norm contraction
// (norm = x^2 gamma'^T + beta' from the tile's x^2 operand block) and hand the accumulator stage back to the main-loop issuer.
template <int NCH, bool kPair, int kParts>
__device__ __forceinline__ void norm_issuer_pipe(const TcParams &P, uint32_t tmem_base, uint64_t *tmem_empty_bar, uint64_t *gdn_bar, uint64_t *gload_bar,
                                                 const uint8_t *sG, uint32_t s_ones, uint32_t s_beta)
{
    constexpr int kEpiThreads = 128 * kParts;
    const uint32_t norm_col = (uint32_t)(P.acc_stages * P.Ntile), a_col0 = norm_col + (uint32_t)P.gdn_chunk;
    const uint32_t empty_leader = kPair ? mapa_u32(smem_u32(&tmem_empty_bar[0]), 0) : 0u;
    const uint32_t idesc = make_idesc(P.Ntile), gm = smem_u32(sG);
    uint32_t r = 0;
    int as = 0;
    bool first = true;
    for (int v = blockIdx.x; v < P.n_virtual; v += gridDim.x) {
        int tile;
        if (!map_tile(P, v, tile)) break;
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads + 32) : "memory");
        if (first) { mbar_wait(gload_bar, 0); first = false; }
        tc_fence_after();
        if (elect_one()) {
            const uint32_t a_col = a_col0 + r * (uint32_t)(P.Cout / 2);
#pragma unroll
            for (int kc = 0; kc < NCH * kParts / 4; ++kc) {
                const uint64_t bdesc = make_desc(gm + (uint32_t)(kc * P.Cout * 128));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_ts(tmem_base + norm_col, tmem_base + a_col + (uint32_t)((kc * 4 + k) * 8), bdesc + (uint64_t)(k * 2), idesc, (kc | k) != 0);
            }
            tc_mma(tmem_base + norm_col, make_desc_ns(s_ones), make_desc_ns(s_beta), idesc, 1);
            tc_commit(gdn_bar);
            if (kPair) mbar_arrive_cluster(empty_leader + (uint32_t)(as * sizeof(uint64_t)));
            else mbar_arrive(&tmem_empty_bar[as]);
        }
        __syncwarp();
        r ^= 1;
        if (++as == P.acc_stages) as = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// col2im gather of the reconstruction layer (MODE_SCATTER, k = 5, stride 2), on raw shared-memory addresses.
// Round 2 (ncu source view of g_s.6, profiles/r02_ncu_scatter_source.txt): the first version walked generic pointers -- every one
// of the 25 product loads of a work item carried its own 64-bit address arithmetic (LEA / IMAD.WIDE / LD), and the item -> (pixel,
// channel) decomposition (three integer divisions) was redone for every tile: ~840 SASS instructions per item, the epilogue was
// ISSUE-bound (2.5k cycles per tile against 0.5k for the tile's TMA stream).  Here the 25 loads are ld.shared with immediate
// offsets off 9 (dy, dx) row bases, and the work items of a thread are decomposed once, before the tile loop.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float lds_f32(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

// base: byte address of S[centre pixel][c]; sp4: bytes per product row; rowp: bytes per tile row of product rows (TW * sp4).
// Output pixel (2a + py, 2b + px) sums the taps with ky = py, kx = px (mod 2); tap (ky, kx) comes from input pixel
// (a + (py + 2 - ky) / 2, b + (px + 2 - kx) / 2)  (oy = 2 iy - 2 + ky).  COUT = 0: channel count at run time.
template <int COUT>
__device__ __forceinline__ void gather_k5s2(uint32_t base, uint32_t sp4, uint32_t rowp, int cout_rt, float &o00, float &o01, float &o10, float &o11)
{
    const uint32_t cs = (uint32_t)(COUT > 0 ? COUT : cout_rt) * 4u;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
        const int py = ky & 1, dy = (py + 2 - ky) / 2;
        const uint32_t rb = base + (uint32_t)(dy * (int)rowp);
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
            const int px = kx & 1, dx = (px + 2 - kx) / 2;
            const float v = lds_f32(rb + (uint32_t)(dx * (int)sp4) + (uint32_t)(ky * 5 + kx) * cs);
            if (py == 0 && px == 0) o00 += v;
            else if (py == 0) o01 += v;
            else if (px == 0) o10 += v;
            else o11 += v;
        }
    }
}
constexpr int kScatterItems = 3;    // work items per gather thread: ceil(6 * 14 * 4 / 128)

enum { EPI_PLAIN = 0, EPI_GDN = 1, EPI_SCATTER = 2 };

template <int kEpi, int kNCH, bool kPair, int kParts, int kTeams>
__global__ void __launch_bounds__(tc_threads(kParts, kTeams), 1) conv_tc_kernel(const __grid_constant__ TcParams P)
{
    constexpr int kEpiThreads = 128 * kParts;     // threads of ONE epilogue team
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
    __shared__ uint64_t bfull_bar[kMaxStages], bempty_bar[kMaxStages];   // grouped mode: the weight-tile ring (full_bar / empty_bar: A patches)
    __shared__ uint64_t tmem_full_bar[kMaxAccStages], tmem_empty_bar[kMaxAccStages], gdn_bar[2], gload_bar, bres_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[kMaxCout];
    __shared__ __align__(16) float beta_s[256];
    __shared__ int64_t pix_off_s[128];   // GDN epilogue: global element offset of each tile pixel (-1 = masked)
    // ones | bias (<= 256 rows) | beta (GDN, <= 192 rows) tiles of the constant-operand MMAs
    __shared__ __align__(128) uint8_t s_const[kEpi == EPI_GDN ? 4096 + 2 * (kConstRowsMax / 8) * 256 : 16];
    uint8_t *s_ones = s_const, *s_biasB = s_const + (kEpi == EPI_GDN ? 4096 : 0), *s_betaB = s_const + (kEpi == EPI_GDN ? 4096 + (kConstRowsMax / 8) * 256 : 0);

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_tile_bytes = (kPair ? P.Ntile / 2 : P.Ntile) * 128;   // pair mode: this CTA's half of the weight rows
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const int stage_bytes = kABytes + (P.b_resident ? 0 : b_tile_bytes);
    // grouped: [nA A patches][nB weight tiles] instead of num_stages (A tile, B tile) pairs
    const size_t ring_bytes = P.grouped ? (size_t)P.issuers * ((size_t)P.nA * P.a_stage_bytes + (size_t)P.nB * b_tile_bytes) : (size_t)P.num_stages * stage_bytes;
    uint8_t *sG = smem + ring_bytes;       // gamma: (Cout/64) tiles of [Cout][64] bf16
    uint8_t *sA2 = sG + (size_t)P.Cout * P.Cout * 2;   // x^2:   (Cout/64) tiles of [128][64] bf16
    float *sStage = reinterpret_cast<float *>(sG);                  // MODE_SCATTER: [128][spitch] fp32 products
    // resident weights (b_resident): one [Ntile][64] bf16 tile per K block, after the GDN / staging region
    const size_t epi_bytes = (kEpi == EPI_GDN) ? (size_t)P.Cout * P.Cout * 2 + (P.a_tmem ? 0 : (size_t)(P.Cout / 64) * kABytes)
                           : (kEpi == EPI_SCATTER) ? kParts * (((size_t)128 * P.spitch * sizeof(float) + 1023) & ~(size_t)1023) : 0;
    uint8_t *sBres = sG + epi_bytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&bfull_bar[s], 1); mbar_init(&bempty_bar[s], 1); }
        for (int s = 0; s < kMaxAccStages; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], kEpi == EPI_SCATTER ? 4 : (kPair ? 2 : 1) * ((kEpi == EPI_GDN && (P.late_release || kTeams == 2)) ? 1 : kEpiThreads / 32)); }   // one arrival per epilogue warp (col2im: per team of 4)
        mbar_init(&gdn_bar[0], 1);
        mbar_init(&gdn_bar[1], 1);
        mbar_init(&gload_bar, 1);
        mbar_init(&bres_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&P.tmA);
        prefetch_tmap(&P.tmB);
        if (kEpi == EPI_GDN) prefetch_tmap(&P.tmG);
    }
    if (warp == 1) {
        if (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    // Programmatic dependent launch (launch_tc sets cudaLaunchAttributeProgrammaticStreamSerialization): everything above -- barrier
    // init, tensor-map prefetch, the TMEM allocation -- touches only this CTA's own state and may run while the previous kernel of the
    // stream is still draining; from here on global memory written by it (activations; on a first forward also packed weights and
    // the GDN parameters) is read, so every thread waits for its completion and memory flush first.  The early launch_dependents lets
    // the NEXT kernel's CTAs take over SMs as this (persistent, one CTA per SM) grid's CTAs exit one by one: its prologue and first
    // operand loads overlap this kernel's tail instead of following a full grid drain + launch.  Both are no-ops without the attribute.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (kEpi == EPI_GDN) {
        // pair mode: this CTA supplies weight (and bias) rows [rank * Ntile / 2, (rank + 1) * Ntile / 2) of the main-loop B operand
        const int brows = kPair ? P.Ntile / 2 : P.Ntile;
        fill_const_tile(s_ones, 128, nullptr, true);
        fill_const_tile(s_biasB, brows, P.bias ? P.bias + (kPair ? (int)rank * brows : 0) : nullptr, false);
        if (kEpi == EPI_GDN) fill_const_tile(s_betaB, P.Cout, P.beta, false);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kMaxCout; i += tc_threads(kParts, kTeams)) {
        bias_s[i] = (P.bias && i < P.Cout) ? P.bias[i] : 0.0f;
        if (i < 256) beta_s[i] = (kEpi == EPI_GDN && i < P.Cout) ? P.beta[i] : 1.0f;
    }
    tc_fence_before();
    __syncthreads();
    if (kPair) cluster_sync_all();     // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    // pair mode: tiles 2j and 2j+1 go to the two CTAs of a cluster (blockIdx.x = 2 * pair + rank, gridDim.x even, the host
    // pads tiles_per_phase to an even count; an index past the last image decodes to b == B and is masked everywhere)

    if (warp == 0) {
        // ===================== TMA producer =====================
        // ONE elected thread runs the whole loop (round 2: with the entire warp walking the K-block loop and an elect.sync per
        // block, the per-block bookkeeping -- pointer re-derivation, R2UR chains, divergence barriers -- cost more than the four
        // MMAs it feeds; see profiles/r02_probe_mma_rate_v4.txt).  Stage and barrier addresses are running 32-bit values.
        if (elect_one()) {
            if (kEpi == EPI_GDN) {
                const int grows = P.Cout;
                mbar_expect_tx(&gload_bar, (uint32_t)(grows * P.Cout * 2));
                for (int kc = 0; kc < P.Cout / 64; ++kc)
                    tma_load_2d(&P.tmG, &gload_bar, sG + (size_t)kc * grows * 128, kc * 64, 0);
            }
            if (P.b_resident) {
                const int nkb = P.phase_begin[1] * P.kchunks;
                mbar_expect_tx(&bres_bar, (uint32_t)(nkb * b_tile_bytes));
                for (int tp = 0; tp < P.phase_begin[1]; ++tp)
                    for (int kc = 0; kc < P.kchunks; ++kc)
                        tma_load_2d(&P.tmB, &bres_bar, sBres + (size_t)(tp * P.kchunks + kc) * b_tile_bytes, kc * 64, P.taps[tp].brow);
            }
            const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
            const uint32_t full0_leader = kPair ? mapa_u32(full0, 0) : 0u;   // pair: the leader's barrier collects both CTAs' bytes
            const uint32_t sbytes = (uint32_t)stage_bytes, nst = (uint32_t)P.num_stages;
            const int kchunks = P.kchunks, kchunks1 = P.kchunks1, half_n = P.Ntile / 2;
            const bool resident = P.b_resident != 0, dbg_no_tma = P.debug == 1;
            uint32_t stage = 0, phase = 0, a_s = smem0;
            TileIter ti;
            ti.init(P, blockIdx.x);
            if (P.grouped) {
                // ---- grouped: this warp streams the A patches, one per (group, chunk); warp 2 streams the weight tiles.  Even and odd
                //      tiles (the two MMA issuers) have their OWN rings of nA patches: a barrier has one phase bit, so every ring needs
                //      a single consumer that observes each of its phases ----
                const uint32_t abytes = (uint32_t)P.a_stage_bytes, nA = (uint32_t)P.nA;
                uint32_t st[2] = {0, 0}, ph[2] = {0, 0};
                uint32_t r = 0;
                const uint32_t rtoggle = P.issuers == 2 ? 1u : 0u;
                for (int v = blockIdx.x; v < P.n_virtual; v += gridDim.x, ti.advance(P), r ^= rtoggle) {
                    int tile;
                    if (!map_tile(P, v, tile)) break;
                    if (kPair || P.phase_inner) ti.init(P, tile);
                    const TileCoord t = ti.coord(P);
                    const int cx = t.x0 * P.a_sx, cy = t.y0 * P.a_sy;
                    const int g_end = P.group_begin[t.phase + 1];
                    uint32_t stage_r = st[r], phase_r = ph[r];
                    for (int gi = P.group_begin[t.phase]; gi < g_end; ++gi) {
                        const Group G = P.groups[gi];
                        const int ax = cx + G.ox, ay = cy + G.oy;
                        for (int kc = 0; kc < kchunks; ++kc) {
                            const bool live = !(dbg_no_tma && (phase_r != 0 || v >= (int)(blockIdx.x + 2 * gridDim.x)));
                            const uint32_t slot = r * nA + stage_r;
                            const uint32_t dst = smem0 + slot * abytes;
                            mbar_wait_a(empty0 + slot * 8, phase_r ^ 1);
                            if (!live) {
                                if (!kPair || rank == 0) mbar_arrive_a(full0 + slot * 8);
                            } else if (kPair) {
                                const uint32_t fb = full0_leader + slot * 8;
                                if (rank == 0) mbar_expect_tx_a(full0 + slot * 8, 2 * abytes);
                                if (kc < kchunks1) tma_load_4d_pair_a(&P.tmA, fb, dst, kc * 64, ax, ay, t.b);
                                else tma_load_4d_pair_a(&P.tmA2, fb, dst, (kc - kchunks1) * 64, ax, ay, t.b);
                            } else {
                                const uint32_t fb = full0 + slot * 8;
                                mbar_expect_tx_a(fb, abytes);
                                if (kc < kchunks1) tma_load_4d_a(&P.tmA, fb, dst, kc * 64, ax, ay, t.b);
                                else tma_load_4d_a(&P.tmA2, fb, dst, (kc - kchunks1) * 64, ax, ay, t.b);
                            }
                            if (++stage_r == nA) { stage_r = 0; phase_r ^= 1; }
                        }
                    }
                    st[r] = stage_r; ph[r] = phase_r;
                }
            } else
            for (int v = blockIdx.x; v < P.n_virtual; v += gridDim.x, ti.advance(P)) {
                int tile;
                if (!map_tile(P, v, tile)) break;
                if (kPair || P.phase_inner) ti.init(P, tile);
                const TileCoord t = ti.coord(P);
                const int cx = t.x0 * P.a_sx, cy = t.y0 * P.a_sy;
                const int tp_end = P.phase_begin[t.phase + 1];
                for (int tp = P.phase_begin[t.phase]; tp < tp_end; ++tp) {
                    const Tap tap = P.taps[tp];
                    const int ax = cx + tap.dx, ay = cy + tap.dy, brow = tap.brow + t.n0 + (kPair ? (int)rank * half_n : 0);
                    for (int kc = 0; kc < kchunks; ++kc) {
                        mbar_wait_a(empty0 + stage * 8, phase ^ 1);
                        if (kPair) {
                            const uint32_t fb = full0_leader + stage * 8;
                            if (rank == 0) mbar_expect_tx_a(full0 + stage * 8, 2 * sbytes);   // only the leader arrives on it
                            if (kc < kchunks1) tma_load_4d_pair_a(&P.tmA, fb, a_s, kc * 64, ax, ay, t.b);
                            else tma_load_4d_pair_a(&P.tmA2, fb, a_s, (kc - kchunks1) * 64, ax, ay, t.b);
                            tma_load_2d_pair_a(&P.tmB, fb, a_s + kABytes, kc * 64, brow);
                        } else if (dbg_no_tma && (phase != 0 || v != (int)blockIdx.x)) {   // profiling: MMA-only rate
                            mbar_arrive_a(full0 + stage * 8);
                        } else if (P.debug >= 4 && (phase != 0 || v != (int)blockIdx.x)) {
                            // profiling: 4 = no A loads after priming, 5 = no B loads (which operand stream bounds the layer?)
                            const uint32_t fb = full0 + stage * 8;
                            mbar_expect_tx_a(fb, P.debug == 4 ? sbytes - kABytes : (uint32_t)kABytes);
                            if (P.debug == 5) tma_load_4d_a(&P.tmA, fb, a_s, kc * 64, ax, ay, t.b);
                            else tma_load_2d_a(&P.tmB, fb, a_s + kABytes, kc * 64, brow);
                        } else {
                            const uint32_t fb = full0 + stage * 8;
                            mbar_expect_tx_a(fb, sbytes);
                            if (kc < kchunks1) tma_load_4d_a(&P.tmA, fb, a_s, kc * 64, ax, ay, t.b);
                            else tma_load_4d_a(&P.tmA2, fb, a_s, (kc - kchunks1) * 64, ax, ay, t.b);
                            if (!resident) tma_load_2d_a(&P.tmB, fb, a_s + kABytes, kc * 64, brow);
                        }
                        ++stage; a_s += sbytes;
                        if (stage == nst) { stage = 0; a_s = smem0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ===================== weight-tile producer (grouped mode) =====================
        // Its own thread: with one producer thread for both operand streams and one MMA-issuing thread, each spent ~500 cycles
        // per tap on ring bookkeeping (ncu r02 source view: the epilogue warps starved on tmem_full while those two threads were
        // busy executing, not waiting) -- twice the 256 cycles the four MMAs of a tap take.
        if (P.grouped && elect_one()) {
            const uint32_t smem0 = smem_u32(smem);
            const uint32_t bfull0 = smem_u32(&bfull_bar[0]), bempty0 = smem_u32(&bempty_bar[0]);
            const uint32_t bfull0_leader = kPair ? mapa_u32(bfull0, 0) : 0u;
            const uint32_t bbytes = (uint32_t)b_tile_bytes, nB = (uint32_t)P.nB, b_s0 = smem0 + (uint32_t)P.issuers * (uint32_t)P.nA * (uint32_t)P.a_stage_bytes;
            const int kchunks = P.kchunks, half_n = P.Ntile / 2;
            const bool dbg_no_tma = P.debug == 1;
            uint32_t st[2] = {0, 0}, ph[2] = {0, 0};
            uint32_t r = 0;
            const uint32_t rtoggle = P.issuers == 2 ? 1u : 0u;
            for (int v = blockIdx.x; v < P.n_virtual; v += gridDim.x, r ^= rtoggle) {
                int tile;
                if (!map_tile(P, v, tile)) break;
                const int tphase = P.fd_tpp.div(tile);
                int n0 = 0;
                if (P.n_blocks > 1) n0 = ((tile - tphase * P.tiles_per_phase) % P.n_blocks) * P.Ntile;
                const int brow_off = n0 + (kPair ? (int)rank * half_n : 0);
                const int g_end = P.group_begin[tphase + 1];
                uint32_t bi = st[r], bphase = ph[r];
                for (int gi = P.group_begin[tphase]; gi < g_end; ++gi) {
                    const int ntaps = P.groups[gi].ntaps, tap0 = P.groups[gi].tap0;
                    for (int kc = 0; kc < kchunks; ++kc) {
                        for (int j = 0; j < ntaps; ++j) {
                            const int brow = P.gbrow[tap0 + j] + brow_off;
                            const bool live = !(dbg_no_tma && (bphase != 0 || v >= (int)(blockIdx.x + 2 * gridDim.x)));
                            const uint32_t slot = r * nB + bi;
                            const uint32_t dst = b_s0 + slot * bbytes;
                            mbar_wait_a(bempty0 + slot * 8, bphase ^ 1);
                            if (!live) {
                                if (!kPair || rank == 0) mbar_arrive_a(bfull0 + slot * 8);
                            } else if (kPair) {
                                if (rank == 0) mbar_expect_tx_a(bfull0 + slot * 8, 2 * bbytes);
                                tma_load_2d_pair_a(&P.tmB, bfull0_leader + slot * 8, dst, kc * 64, brow);
                            } else {
                                mbar_expect_tx_a(bfull0 + slot * 8, bbytes);
                                tma_load_2d_a(&P.tmB, bfull0 + slot * 8, dst, kc * 64, brow);
                            }
                            if (++bi == nB) { bi = 0; bphase ^= 1; }
                        }
                    }
                }
                st[r] = bi; ph[r] = bphase;
            }
        }
        __syncwarp();
    } else if (warp == 3 && kEpi == EPI_GDN && kTeams == 1 && (kNCH == 4 || kNCH == 2) && P.epi_pipe) {
        // ===================== norm-contraction issuer of the pipelined GDN epilogue =====================
        if constexpr (kEpi == EPI_GDN && kTeams == 1 && (kNCH == 4 || kNCH == 2))
            norm_issuer_pipe<kNCH, kPair, kParts>(P, tmem_base, tmem_empty_bar, &gdn_bar[0], &gload_bar, sG, smem_u32(s_ones), smem_u32(s_betaB));
    } else if ((warp == 1 || (warp == 3 && P.grouped && P.issuers == 2)) && (!kPair || rank == 0)) {
        // ===================== MMA issuer (pair mode: the leader CTA issues for both) =====================
        // grouped mode: TWO issuing threads (warps 1 and 3) take alternate tiles -- different accumulator stages, so no ordering
        // between them is needed; each skips the ring slots of the other's tiles.
        // ONE elected thread: wait for the stage, four tcgen05.mma whose descriptors are the stage-0 descriptors plus a running
        // 16-byte-unit offset, commit.  Measured (profiles/r02_probe_mma_rate_v4.txt): a lean loop sustains the tensor core's
        // 64 cycles per N = 128 MMA; the round-1 loop took ~122 cycles per MMA whatever N <= 128 -- it was issue-bound.
        if (elect_one()) {
            const uint32_t idesc = kPair ? make_idesc_m256(P.Ntile) : make_idesc(P.Ntile);
            const uint64_t adesc0 = make_desc(smem_u32(smem));
            const uint64_t bres0 = make_desc(smem_u32(sBres));
            const uint64_t ones_desc = make_desc_ns(smem_u32(s_ones)), biasb_desc = make_desc_ns(smem_u32(s_biasB));
            const uint32_t step16 = (uint32_t)(stage_bytes >> 4), bres_step16 = (uint32_t)(b_tile_bytes >> 4), nst = (uint32_t)P.num_stages;
            const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
            const uint32_t tfull0 = smem_u32(&tmem_full_bar[0]), tempty0 = smem_u32(&tmem_empty_bar[0]);
            const bool resident = P.b_resident != 0, four = P.ksteps == 4, no_mma = P.debug == 2;
            const int kchunks = P.kchunks, ntile = P.Ntile, acc_stages = P.acc_stages;
            uint32_t stage = 0, phase = 0, off16 = 0;
            // grouped mode: A patches in full_bar / empty_bar's ring, weight tiles in their own ring behind it
            const uint32_t g_nA = (uint32_t)P.nA, g_nB = (uint32_t)P.nB, g_astep16 = (uint32_t)(P.a_stage_bytes >> 4), g_bstep16 = (uint32_t)(b_tile_bytes >> 4);
            const uint32_t g_bfull0 = smem_u32(&bfull_bar[0]) + ((P.grouped && warp == 3) ? (uint32_t)P.nB * 8 : 0u);
            const uint32_t g_bempty0 = smem_u32(&bempty_bar[0]) + ((P.grouped && warp == 3) ? (uint32_t)P.nB * 8 : 0u);
            // this issuer's rings: A patches [ring][nA], then weight tiles [ring][nB]; barrier slots likewise
            const uint32_t g_ring = (P.grouped && warp == 3) ? 1u : 0u;
            const uint64_t g_adesc0 = make_desc(smem_u32(smem) + g_ring * g_nA * (uint32_t)P.a_stage_bytes);
            const uint64_t g_bdesc0 = make_desc(smem_u32(smem) + (uint32_t)P.issuers * g_nA * (uint32_t)P.a_stage_bytes + g_ring * g_nB * (uint32_t)b_tile_bytes);
            const uint32_t g_full0 = smem_u32(&full_bar[0]) + g_ring * g_nA * 8, g_empty0 = smem_u32(&empty_bar[0]) + g_ring * g_nA * 8;
            uint32_t g_bi = 0, g_bphase = 0, g_boff16 = 0;
            if (resident) mbar_wait(&bres_bar, 0);
            uint32_t acc_i = 0, acc_ph = 0;
            const uint32_t my_parity = warp == 3 ? 1u : 0u;
            uint32_t it_par = 0;
            for (int v = blockIdx.x; v < P.n_virtual; v += gridDim.x, it_par ^= 1) {
                int tile;
                if (!map_tile(P, v, tile)) break;
                const int tphase = P.fd_tpp.div(tile);
                const int nkb = (P.phase_begin[tphase + 1] - P.phase_begin[tphase]) * kchunks;
                if (P.grouped && P.issuers == 2 && it_par != my_parity) {
                    // the other issuer's tile (it has its own operand rings): only the accumulator ring is shared
                    if (++acc_i == (uint32_t)acc_stages) { acc_i = 0; acc_ph ^= 1; }
                    continue;
                }
                mbar_wait_a(tempty0 + acc_i * 8, acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc_i * (uint32_t)ntile;
                uint64_t bres = bres0;
                if (P.grouped) {
                    // ---- grouped: tap j of a group reads the group's A patch from tile row j on (one 1024-byte swizzle atom per row) ----
                    uint32_t first = 0;
                    const int g_end = P.group_begin[tphase + 1];
                    for (int gi = P.group_begin[tphase]; gi < g_end; ++gi) {
                        const int ntaps = P.groups[gi].ntaps;
                        for (int kc = 0; kc < kchunks; ++kc) {
                            mbar_wait_a(g_full0 + stage * 8, phase);
                            tc_fence_after();
                            uint64_t adesc = g_adesc0 + off16;
                            for (int j = 0; j < ntaps; ++j) {
                                mbar_wait_a(g_bfull0 + g_bi * 8, g_bphase);
                                tc_fence_after();
                                const uint64_t bdesc = g_bdesc0 + g_boff16;
                                if (kPair) {
                                    tc_mma2(d_tmem, adesc, bdesc, idesc, first);
                                    tc_mma2(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                                    tc_mma2(d_tmem, adesc + 4, bdesc + 4, idesc, 1);
                                    tc_mma2(d_tmem, adesc + 6, bdesc + 6, idesc, 1);
                                    tc_commit2_a(g_bempty0 + g_bi * 8);
                                } else {
                                    if (!no_mma) {
                                        tc_mma(d_tmem, adesc, bdesc, idesc, first);
                                        tc_mma(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                                        tc_mma(d_tmem, adesc + 4, bdesc + 4, idesc, 1);
                                        tc_mma(d_tmem, adesc + 6, bdesc + 6, idesc, 1);
                                    }
                                    tc_commit_a(g_bempty0 + g_bi * 8);
                                }
                                first = 1;
                                adesc += 64;                                    // next tile row of the patch: + 1024 B
                                ++g_bi; g_boff16 += g_bstep16;
                                if (g_bi == g_nB) { g_bi = 0; g_boff16 = 0; g_bphase ^= 1; }
                            }
                            if (kPair) tc_commit2_a(g_empty0 + stage * 8);        // the patch is free once its last tap's MMAs have read it
                            else tc_commit_a(g_empty0 + stage * 8);
                            ++stage; off16 += g_astep16;
                            if (stage == g_nA) { stage = 0; off16 = 0; phase ^= 1; }
                        }
                    }
                } else
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait_a(full0 + stage * 8, phase);
                    tc_fence_after();
                    const uint64_t adesc = adesc0 + off16;
                    const uint64_t bdesc = resident ? bres : adesc + (uint64_t)(kABytes >> 4);
                    if (kPair) {
                        tc_mma2(d_tmem, adesc, bdesc, idesc, kb != 0);
                        tc_mma2(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                        tc_mma2(d_tmem, adesc + 4, bdesc + 4, idesc, 1);
                        tc_mma2(d_tmem, adesc + 6, bdesc + 6, idesc, 1);
                        tc_commit2_a(empty0 + stage * 8);                          // both CTAs' producers
                    } else {
                        if (!no_mma) {
                            // K=16 per step: +32 B (= +2 in descriptor units) inside the 128-byte swizzle atom
                            tc_mma(d_tmem, adesc, bdesc, idesc, kb != 0);
                            tc_mma(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                            tc_mma(d_tmem, adesc + 4, bdesc + 4, idesc, 1);
                            if (four) tc_mma(d_tmem, adesc + 6, bdesc + 6, idesc, 1);
                        }
                        tc_commit_a(empty0 + stage * 8);   // frees the smem slot once these MMAs have read it
                    }
                    bres += bres_step16;
                    ++stage; off16 += step16;
                    if (stage == nst) { stage = 0; off16 = 0; phase ^= 1; }
                }
                if (kPair) {
                    if (kEpi == EPI_GDN) tc_mma2(d_tmem, ones_desc, biasb_desc, idesc, 1);   // + bias
                    tc_commit2_a(tfull0 + acc_i * 8);      // both CTAs' epilogues
                } else {
                    if (kEpi == EPI_GDN) tc_mma(d_tmem, ones_desc, biasb_desc, idesc, 1);    // + bias
                    tc_commit_a(tfull0 + acc_i * 8);       // accumulator complete -> epilogue
                }
                if (++acc_i == (uint32_t)acc_stages) { acc_i = 0; acc_ph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= kSvcWarps && kEpi == EPI_GDN && kTeams == 1 && (kNCH == 4 || kNCH == 2) && P.epi_pipe) {
        // ===================== epilogue, software-pipelined over two tiles (GDN, C <= 128) ============
        if constexpr (kEpi == EPI_GDN && kTeams == 1 && (kNCH == 4 || kNCH == 2))
            epilogue_gdn_pipe<kNCH, kPair, kParts>(P, tmem_base, tmem_full_bar, &gdn_bar[0]);
    } else if (warp >= kSvcWarps) {
        // ===================== epilogue: 8 warps, 2 per TMEM lane quarter, each pair splits the columns ============
        const int q = warp & 3;                 // TMEM lane quarter this warp can access
        const int team = (warp - kSvcWarps) / (4 * kParts);                 // kTeams == 2: which of the two epilogue teams
        const int half = ((warp - kSvcWarps) % (4 * kParts)) >> 2;          // column part of this warp: 0 .. kParts - 1
        const int row = q * 32 + lane;          // accumulator row == pixel of the tile
        const int th = row / P.TW, tw = row - th * P.TW;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t norm_col = (uint32_t)(P.acc_stages * P.Ntile);
        uint32_t gdn_phase = 0;
            int it = 0;
        TileIter ti;
        ti.init(P, blockIdx.x);
        int acc_i = 0;
        uint32_t acc_ph = 0;
        const uint32_t empty_leader = kPair ? mapa_u32(smem_u32(&tmem_empty_bar[0]), 0) : 0u;
        // col2im gather: this thread's work items e = et + 128 i -> (channel c, interior input pixel (ay, ax)), tile independent
        uint32_t sc_soff[kScatterItems];    // byte offset of S[centre pixel][c] in the team's staging buffer
        int sc_yx[kScatterItems];           // (c << 24) | (ay << 16) | ax, -1: no item
        int64_t sc_out[kScatterItems];      // output offset relative to the tile's first interior output pixel of channel 0
        int sc_items = 0;                   // work items per tile (all threads)
        if constexpr (kEpi == EPI_SCATTER) {
            const int et = (threadIdx.x - 32 * kSvcWarps) & 127;
            const int ih = P.TH - P.halo_lo - P.halo_hi, iw = P.TW - P.halo_lo - P.halo_hi;
            const int items = ih * iw * P.Cout;
            sc_items = items;
#pragma unroll
            for (int i = 0; i < kScatterItems; ++i) {
                const int e = et + 128 * i;
                const int ax = e % iw, r2 = e / iw;
                const int ay = r2 % ih, c = r2 / ih;
                sc_yx[i] = e < items ? ((c << 24) | (ay << 16) | ax) : -1;
                sc_soff[i] = e < items ? (uint32_t)((((P.halo_lo + ay) * P.TW + P.halo_lo + ax) * P.spitch + c) * 4) : (uint32_t)(((P.halo_lo * P.TW + P.halo_lo) * P.spitch) * 4);    // no item: any valid product row
                sc_out[i] = ((int64_t)c * P.Ho + ay * P.out_stride) * P.Wo + ax * P.out_stride;
            }
        }
        for (int v = blockIdx.x; v < P.n_virtual; v += gridDim.x, ++it, ti.advance(P), acc_ph ^= (acc_i + 1 == P.acc_stages), acc_i = (acc_i + 1 == P.acc_stages) ? 0 : acc_i + 1) {
            int tile;
            if (!map_tile(P, v, tile)) break;
            if (kPair || P.phase_inner) ti.init(P, tile);
            // col2im epilogue: kParts (2-4) independent teams of 4 warps (one warp per TMEM lane quarter) take tiles round robin, each
            // with its own staging buffer and named barrier, so that one team's TMEM / shared-memory latencies overlap the others' work
            // (a team is a serial chain of dependent latencies: with two teams the layer ran at ~0.3 instructions per scheduler cycle).
            // acc_stages is a multiple of kParts there, hence every accumulator stage (and its barriers) always belongs to the same team.
            if (kEpi == EPI_SCATTER && (it % kParts) != half) continue;
            if (kTeams == 2 && (it & 1) != team) continue;             // GDN teams: alternate tiles
            const TileCoord t = ti.coord(P);
            const int as = acc_i;                       // accumulator ring position of this tile
            const uint32_t aphase = acc_ph;
            const int gy = t.y0 + th, gx = t.x0 + tw;
            const bool valid = gy < P.Gh && gx < P.Gw && t.b < P.B;
            const uint32_t acc_addr = tmem_base + lane_addr + (uint32_t)(as * P.Ntile);

            mbar_wait(&tmem_full_bar[as], aphase);
            tc_fence_after();

            if (kEpi == EPI_SCATTER) {
                // ---- GEMM + col2im: column n = (ky*k + kx)*Cout + c holds x[q] . w[:, c, ky, kx] for INPUT pixel q ----
                float *stage = sStage + (size_t)half * ((((size_t)128 * P.spitch * sizeof(float) + 1023) & ~(size_t)1023) / sizeof(float));
                const uint32_t stage_a = smem_u32(stage);
                {
                    const int nch = P.Ntile >> 4;
                    const uint32_t srow = stage_a + (uint32_t)(row * P.spitch) * 4u;
                    // scalar stores with an ODD row pitch: conflict-free here (lanes = rows) and in the gather below
                    // (lanes = neighbouring pixels = neighbouring rows); 16-byte stores would need a pitch that is a
                    // multiple of 4 and make the gather 4-way bank conflicted.  Four TMEM loads in flight per wait.
                    for (int ch = 0; ch < nch; ch += 4) {
                        float v[4][16];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (ch + u < nch) tmem_ld16(acc_addr + ((ch + u) << 4), v[u]);
                        tmem_ld_wait();
                        const uint32_t dst = srow + (uint32_t)(ch << 6);
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (ch + u < nch) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) sts_f32(dst + 64 * u + 4 * i, v[u][i]);
                            }
                    }
                }
                // the accumulator is drained: hand the TMEM stage back before the gather pass
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
                asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
                if (P.k == 5 && P.out_stride == 2) {
                    // Gather (k = 5, stride 2): one work item per (interior input-resolution pixel a, channel c) produces the 2 x 2
                    // output block at 2 a; every product S[q][(ky,kx,c)] is consumed exactly once.
                    const int oyb = (t.y0 + P.halo_lo) * 2, oxb = (t.x0 + P.halo_lo) * 2;
                    float *tile_out = (float *)P.y + ((int64_t)t.b * P.Cout * P.Ho + oyb) * P.Wo + oxb;
                    const uint32_t sp4 = (uint32_t)P.spitch * 4u, rowp = sp4 * (uint32_t)P.TW;
                    const bool even_w = (P.Wo & 1) == 0;
                    // all product loads of this thread's items are issued before the first use (a missing item reads the first interior
                    // pixel's products and is dropped at the store): the gather is a latency chain of 25 loads + adds per item otherwise
                    float o[kScatterItems][4];
#pragma unroll
                    for (int i = 0; i < kScatterItems; ++i) {
                        if (i * 128 >= sc_items) break;                       // uniform: item slots in use = ceil(items / 128)
                        const uint32_t base = stage_a + sc_soff[i];
                        const float b0 = bias_s[(sc_yx[i] >> 24) & 3];
                        o[i][0] = o[i][1] = o[i][2] = o[i][3] = b0;           // same summation order as the generic loop: bias, then the taps
                        if (P.Cout == 3) gather_k5s2<3>(base, sp4, rowp, 3, o[i][0], o[i][1], o[i][2], o[i][3]);
                        else if (P.Cout == 1) gather_k5s2<1>(base, sp4, rowp, 1, o[i][0], o[i][1], o[i][2], o[i][3]);
                        else gather_k5s2<0>(base, sp4, rowp, P.Cout, o[i][0], o[i][1], o[i][2], o[i][3]);
                    }
#pragma unroll
                    for (int i = 0; i < kScatterItems; ++i) {
                        if (i * 128 >= sc_items) break;
                        if (sc_yx[i] < 0) continue;
                        const int oy0 = oyb + 2 * ((sc_yx[i] >> 16) & 0xff), ox0 = oxb + 2 * (sc_yx[i] & 0xffff);
                        if (oy0 >= P.Ho || ox0 >= P.Wo) continue;
                        const float o00 = act_tc(o[i][0], P.act), o01 = act_tc(o[i][1], P.act);
                        const float o10 = act_tc(o[i][2], P.act), o11 = act_tc(o[i][3], P.act);
                        float *row0 = tile_out + sc_out[i];
                        const bool two_x = ox0 + 1 < P.Wo, two_y = oy0 + 1 < P.Ho;
                        if (two_x && even_w) {
                            *reinterpret_cast<float2 *>(row0) = make_float2(o00, o01);
                            if (two_y) *reinterpret_cast<float2 *>(row0 + P.Wo) = make_float2(o10, o11);
                        } else {
                            row0[0] = o00;
                            if (two_x) row0[1] = o01;
                            if (two_y) { row0[P.Wo] = o10; if (two_x) row0[P.Wo + 1] = o11; }
                        }
                    }
                } else {
                    // Generic gather (other kernel sizes / strides): one work item per (interior input-resolution pixel a, channel c)
                    // produces the s x s output block (s*a + p); every product S[q][(ky,kx,c)] is consumed exactly once.
                    const int s = P.out_stride, k = P.k;
                    const int ih = P.TH - P.halo_lo - P.halo_hi, iw = P.TW - P.halo_lo - P.halo_hi;   // interior (input res)
                    const int items = ih * iw * P.Cout;
                    float *yo = (float *)P.y;
                    for (int e = (threadIdx.x - 32 * kSvcWarps) & 127; e < items; e += 128) {
                        const int ax = e % iw;
                        const int r2 = e / iw;
                        const int ay = r2 % ih;
                        const int c = r2 / ih;
                        const int oy0 = (t.y0 + P.halo_lo + ay) * s, ox0 = (t.x0 + P.halo_lo + ax) * s;
                        if (oy0 >= P.Ho || ox0 >= P.Wo) continue;
                        const float b0 = bias_s[c];
                        {
                            for (int py = 0; py < s; ++py)
                                for (int px = 0; px < s; ++px) {
                                    const int oy = oy0 + py, ox = ox0 + px;
                                    if (oy >= P.Ho || ox >= P.Wo) continue;
                                    float acc = b0;
                                    for (int ky = (oy + P.pad) % s; ky < k; ky += s) {   // ky = oy + pad - s*iy
                                        const int iyl = (oy + P.pad - ky) / s - t.y0;
                                        if (iyl < 0 || iyl >= P.TH) continue;
                                        for (int kx = (ox + P.pad) % s; kx < k; kx += s) {
                                            const int ixl = (ox + P.pad - kx) / s - t.x0;
                                            if (ixl < 0 || ixl >= P.TW) continue;
                                            acc += stage[(size_t)(iyl * P.TW + ixl) * P.spitch + (ky * k + kx) * P.Cout + c];
                                        }
                                    }
                                    yo[(((int64_t)t.b * P.Cout + c) * P.Ho + oy) * P.Wo + ox] = act_tc(acc, P.act);
                                }
                        }
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");   // this team's staging buffer is free for its next tile
                continue;
            } else {
                const int py = t.phase / P.out_stride, px = t.phase - py * P.out_stride;
                const int oy = gy * P.out_stride + py, ox = gx * P.out_stride + px;
                const int64_t pix_off = (((int64_t)t.b * P.Ho + oy) * P.Wo + ox) * P.Cout + t.n0;
                const int nch = P.Ntile >> 4;                       // 16-column chunks in this tile
                const int ch_lo = (nch * half + kParts - 1) / kParts;
                const int ch_hi = (nch * (half + 1) + kParts - 1) / kParts;
                if (kEpi == EPI_PLAIN) {
                    // two TMEM loads in flight per wait
                    for (int ch = ch_lo; ch < ch_hi; ch += 2) {
                        const int c0 = ch << 4;
                        const bool two = ch + 1 < ch_hi;
                        float v[16], w[16], bs[16];
                        tmem_ld16(acc_addr + c0, v);
                        if (two) tmem_ld16(acc_addr + c0 + 16, w);
                        if (P.bias_mma) {
                            // the bias is already in the accumulator (constant-operand MMA)
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = act_tc(v[i], P.act);
                            if (valid) store16(P, pix_off + c0, v);
                            if (two) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) w[i] = act_tc(w[i], P.act);
                                if (valid) store16(P, pix_off + c0 + 16, w);
                            }
                            continue;
                        }
                        load16f(bias_s + t.n0 + c0, bs);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = act_tc(v[i] + bs[i], P.act);
                        if (valid) store16(P, pix_off + c0, v);
                        if (two) {
                            load16f(bias_s + t.n0 + c0 + 16, bs);
#pragma unroll
                            for (int i = 0; i < 16; ++i) w[i] = act_tc(w[i] + bs[i], P.act);
                            if (valid) store16(P, pix_off + c0 + 16, w);
                        }
                    }
                } else {
                    if (kTeams == 1 && half == 0) pix_off_s[row] = valid ? pix_off : -1;   // published by the bar.sync inside epilogue_gdn
                    GdnCtx g{P, sA2, sG, &gdn_bar[team], &gload_bar, tmem_base, acc_addr, lane_addr, norm_col, row, half, valid, pix_off,
                             kTeams == 2 ? (it >> 1) : it, pix_off_s,
                             rank, smem_u32(s_ones), smem_u32(s_betaB), norm_col + (uint32_t)P.gdn_chunk, &tmem_empty_bar[as],
                             kPair ? empty_leader + (uint32_t)(as * sizeof(uint64_t)) : 0u, (uint32_t)(1 + team), ((warp - kSvcWarps) % (4 * kParts)) == 0};
                    if constexpr (kTeams == 2) {
                        // norm in place over this tile's accumulator stage; the x^2 operand block of the team sits behind the stages
                        g.norm_col = (uint32_t)(as * P.Ntile);
                        g.a_col = norm_col + (uint32_t)(team * (P.Cout / 2));     // norm_col here = acc_stages * Ntile
                        epilogue_gdn_rows<(kNCH > 0 ? kNCH : 4)>(g, gdn_phase);
                    } else
                    epilogue_gdn<(kNCH > 0 ? kNCH : 2), (kNCH == 6 ? 2 : 1), kPair, kParts, kTeams>(g, bias_s, beta_s, gdn_phase);
                }
            }
            tc_fence_before();
            __syncwarp();          // every lane's TMEM reads of this accumulator are complete: one arrival per warp
            if (lane == 0 && kEpi != EPI_GDN) {      // (the GDN epilogue releases the accumulator right after its first pass)
                if (kPair) mbar_arrive_cluster(empty_leader + (uint32_t)(as * sizeof(uint64_t)));   // the leader's MMA issuer waits for both epilogues
                else mbar_arrive(&tmem_empty_bar[as]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (kPair) cluster_sync_all();     // nobody leaves while the peer may still signal its barriers / read its shared memory
    if (warp == 1) {
        tc_fence_after();
        if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---------------------------------------------------------------------------------------------
// host-side plan shared by the weight packer and the launcher
// ---------------------------------------------------------------------------------------------
struct Plan {
    int mode;
    int ntaps;
    Tap taps[kMaxTaps];
    int tap_ky[kMaxTaps], tap_kx[kMaxTaps];   // MODE_STD: kernel position; MODE_PAD8: ky only
    int phase_begin[5], n_phases;
    int a_sx, a_sy, out_stride, Gh, Gw;
    int kchunks, ksteps;
    int Ntile, n_blocks;
    int wrows, wcols;                          // packed weight matrix: [ntaps * wrows][wcols] bf16
    int Ho, Wo, Hp, Wp;
    int halo_lo, halo_hi;                      // MODE_SCATTER: input rows/cols around an output pair that contribute to it
};

static void pad8_extent(const mmc_conv_desc *d, int Ho, int Wo, int *Hp, int *Wp)
{
    const int pad = d->k / 2;
    int hp = d->H + 2 * pad, wp = d->W + 2 * pad;
    const int need_w = d->stride * (Wo - 1) + 8;   // every output column reads an 8-pixel window
    const int need_h = d->stride * (Ho - 1) + d->k;
    if (wp < need_w) wp = need_w;
    if (hp < need_h) hp = need_h;
    *Hp = hp;
    *Wp = (wp + 1) & ~1;
}

static int make_plan(const mmc_conv_desc *d, Plan &pl, const char *name)
{
    MMC_CHECK_ARG(d != nullptr, "%s: descriptor is NULL", name);
    MMC_CHECK_ARG(d->B >= 0 && d->H >= 1 && d->W >= 1 && d->Cin >= 1 && d->Cout >= 1, "%s: bad shape", name);
    MMC_CHECK_ARG(d->k == 1 || d->k == 3 || d->k == 5, "%s: kernel size %d not in {1,3,5}", name, d->k);
    MMC_CHECK_ARG(d->stride == 1 || d->stride == 2, "%s: stride %d not in {1,2}", name, d->stride);
    MMC_UNSUPPORTED(d->Cout > kMaxCout, "%s: Cout=%d exceeds %d", name, d->Cout, kMaxCout);
    memset(&pl, 0, sizeof(pl));
    const int k = d->k, s = d->stride, pad = k / 2;
    mmc_conv_out_size(d, &pl.Ho, &pl.Wo);
    if (d->in_layout == MMC_NHWC_PAD8) {
        // ---- image-edge convolution: one K box per kernel row = 8 pixels x 8 channels ----
        MMC_UNSUPPORTED(d->transposed || d->Cin > 8 || d->Cout % 16 != 0, "%s: NHWC_PAD8 input needs a forward conv with Cin <= 8, Cout %% 16 == 0", name);
        pl.mode = MODE_PAD8;
        pl.n_phases = 1; pl.a_sx = 1; pl.a_sy = s; pl.out_stride = 1; pl.Gh = pl.Ho; pl.Gw = pl.Wo;
        pl.kchunks = 1; pl.ksteps = (k * 8 + 15) / 16;
        for (int ky = 0; ky < k; ++ky) {
            pl.taps[pl.ntaps] = Tap{(int16_t)ky, 0, ky * d->Cout};
            pl.tap_ky[pl.ntaps++] = ky;
        }
        pl.phase_begin[0] = 0; pl.phase_begin[1] = pl.ntaps;
        pl.wrows = d->Cout; pl.wcols = 64;
        pad8_extent(d, pl.Ho, pl.Wo, &pl.Hp, &pl.Wp);
    } else if (d->transposed && s == 2 && k >= 3 && k * k * d->Cout <= 256 && d->Cout <= 4) {
        // ---- narrow transposed conv as GEMM + col2im: P[q][(ky,kx,c)] = x[q] . w[:, c, ky, kx] ----
        MMC_UNSUPPORTED(d->Cin % 8 != 0 || d->Cin < 32, "%s: tensor-core path needs Cin %% 8 == 0 and Cin >= 32 (got %d)", name, d->Cin);
        pl.mode = MODE_SCATTER;
        pl.n_phases = 1; pl.a_sx = pl.a_sy = 1; pl.out_stride = s; pl.Gh = d->H; pl.Gw = d->W;
        pl.kchunks = (d->Cin + 63) / 64; pl.ksteps = 4;
        pl.taps[0] = Tap{0, 0, 0};
        pl.ntaps = 1;
        pl.phase_begin[0] = 0; pl.phase_begin[1] = 1;
        pl.wrows = (k * k * d->Cout + 15) & ~15; pl.wcols = d->Cin;
        // output rows s*a .. s*a+s-1 receive input rows a - halo_lo .. a + halo_hi   (ky = oy + pad - s*iy in [0, k))
        pl.halo_lo = -((pad - k + 1) / s);          // = ceil((k - 1 - pad) / s) for k - 1 >= pad
        if ((k - 1 - pad) % s) pl.halo_lo = (k - 1 - pad + s - 1) / s;
        pl.halo_hi = (s - 1 + pad) / s;
    } else {
        MMC_UNSUPPORTED(d->Cin % 8 != 0 || d->Cin < 32, "%s: tensor-core path needs Cin %% 8 == 0 and Cin >= 32 (got %d); use the direct kernel", name, d->Cin);
        MMC_UNSUPPORTED(d->Cout % 16 != 0, "%s: tensor-core path needs Cout %% 16 == 0 (got %d); use the direct kernel", name, d->Cout);
        pl.mode = MODE_STD;
        pl.kchunks = (d->Cin + 63) / 64; pl.ksteps = 4;
        pl.wrows = d->Cout; pl.wcols = d->Cin;
        if (!d->transposed) {
            pl.n_phases = 1; pl.a_sx = pl.a_sy = s; pl.out_stride = 1; pl.Gh = pl.Ho; pl.Gw = pl.Wo;
            for (int ky = 0; ky < k; ++ky)
                for (int kx = 0; kx < k; ++kx) {
                    pl.taps[pl.ntaps] = Tap{(int16_t)(ky - pad), (int16_t)(kx - pad), pl.ntaps * d->Cout};
                    pl.tap_ky[pl.ntaps] = ky; pl.tap_kx[pl.ntaps] = kx;
                    ++pl.ntaps;
                }
            pl.phase_begin[1] = pl.ntaps;
        } else {
            // oy = iy*s - pad + ky  =>  for output phase py: ky == (py + pad) mod s, iy = qy + (py + pad - ky)/s
            pl.n_phases = s * s; pl.a_sx = pl.a_sy = 1; pl.out_stride = s; pl.Gh = d->H; pl.Gw = d->W;
            for (int ph = 0; ph < s * s; ++ph) {
                const int py = ph / s, px = ph % s;
                pl.phase_begin[ph] = pl.ntaps;
                for (int ky = 0; ky < k; ++ky) {
                    if (((py + pad - ky) % s) != 0) continue;
                    for (int kx = 0; kx < k; ++kx) {
                        if (((px + pad - kx) % s) != 0) continue;
                        pl.taps[pl.ntaps] = Tap{(int16_t)((py + pad - ky) / s), (int16_t)((px + pad - kx) / s), pl.ntaps * d->Cout};
                        pl.tap_ky[pl.ntaps] = ky; pl.tap_kx[pl.ntaps] = kx;
                        ++pl.ntaps;
                    }
                }
            }
            pl.phase_begin[s * s] = pl.ntaps;
        }
    }
    return MMC_OK;
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 torch layout -> bf16 [tap][rows][cols] in the plan's K order
// ---------------------------------------------------------------------------------------------
struct PackParams {
    int mode, transposed, Cin, Cout, k, stride, ntaps, wrows, wcols;
    int16_t ta[kMaxTaps], tb[kMaxTaps];
};

__global__ void __launch_bounds__(256) pack_weights_kernel(const float *__restrict__ w, PackParams q, __nv_bfloat16 *__restrict__ out)
{
    const int kk = q.k * q.k;
    int64_t n = (int64_t)q.ntaps * q.wrows * q.wcols;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int col = (int)(i % q.wcols);
        int64_t r = i / q.wcols;
        int row = (int)(r % q.wrows);
        int tap = (int)(r / q.wrows);
        float v = 0.0f;
        if (q.mode == MODE_STD) {
            int kpos = q.ta[tap] * q.k + q.tb[tap];
            int64_t src = q.transposed ? (((int64_t)col * q.Cout + row) * kk + kpos) : (((int64_t)row * q.Cin + col) * kk + kpos);
            v = __ldg(w + src);
        } else if (q.mode == MODE_PAD8) {
            int kx = col >> 3, c = col & 7, ky = q.ta[tap];
            if (kx < q.k && c < q.Cin) v = __ldg(w + (((int64_t)row * q.Cin + c) * kk + ky * q.k + kx));
        } else if (q.mode == MODE_SCATTER) {
            int kpos = row / q.Cout, c = row - kpos * q.Cout;   // row n = (ky*k + kx)*Cout + c
            if (kpos < kk) v = __ldg(w + (((int64_t)col * q.Cout + c) * kk + kpos));
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

// NCHW fp32 image -> zero-padded NHWC bf16 with 8 channels per pixel ([B][Hp][Wp][8], 16 B per pixel).
// One thread converts 4 horizontally adjacent pixels: up to 8 independent plane loads in flight, 64 B stored.
__global__ void __launch_bounds__(128) pad8_kernel(const float *__restrict__ x, int C, int H, int W, int pad, int Hp, int Wp,
                                                  uint4 *__restrict__ out)
{
    // grid = (ceil(quads per row / 128), Hp, B): no index divisions
    const int qx = blockIdx.x * 128 + threadIdx.x;
    if (qx * 4 >= Wp) return;
    const int py = blockIdx.y;
    const int64_t b = blockIdx.z;
    const int iy = py - pad;
    float v[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 8; ++c) v[p][c] = 0.0f;
    if (iy >= 0 && iy < H) {
        const float *row = x + (b * C * H + iy) * (int64_t)W;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c >= C) break;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int ix = qx * 4 + p - pad;
                if (ix >= 0 && ix < W) v[p][c] = __ldg(row + (int64_t)c * H * W + ix);
            }
        }
    }
    uint4 *dst = out + (b * Hp + py) * (int64_t)Wp + qx * 4;
#pragma unroll
    for (int p = 0; p < 4; ++p)
        if (qx * 4 + p < Wp)
            dst[p] = make_uint4(pack_bf16(v[p][0], v[p][1]), pack_bf16(v[p][2], v[p][3]), pack_bf16(v[p][4], v[p][5]), pack_bf16(v[p][6], v[p][7]));
}

static int pick_ntile(int cout)
{
    for (int n = 256; n >= 16; n -= 16)
        if (cout % n == 0) return n;
    return 0;
}

static void pick_tile(int gh, int gw, int sx, int sy, int *TH, int *TW)
{
    const int cand[][2] = {{8, 16}, {16, 8}, {4, 32}, {32, 4}, {2, 64}, {64, 2}, {1, 128}, {128, 1}};
    int64_t best = -1;
    for (auto &c : cand) {
        if (c[0] * sy > 256 || c[1] * sx > 256) continue;   // TMA box limit
        int64_t tiles = (int64_t)((gh + c[0] - 1) / c[0]) * ((gw + c[1] - 1) / c[1]);
        if (best < 0 || tiles < best) { best = tiles; *TH = c[0]; *TW = c[1]; }
    }
}

template <int kEpi, int kNCH, bool kPair = false, int kParts = 2, int kTeams = 1>
static int launch_tc(const TcParams &P, size_t fixed, size_t stage_bytes, cudaStream_t st, const char *name)
{
    // dynamic shared memory available next to the kernel's static allocation (227 KB per CTA on sm_100)
    static PerDevice<size_t> budget_dev;   // one slot per device and per template instantiation
    size_t budget = budget_dev.cur().load(std::memory_order_relaxed);
    if (budget == 0) {
        cudaFuncAttributes fa;
        MMC_CHECK_CUDA(cudaFuncGetAttributes(&fa, conv_tc_kernel<kEpi, kNCH, kPair, kParts, kTeams>));
        size_t avail = 227 * 1024 - fa.sharedSizeBytes;
        MMC_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<kEpi, kNCH, kPair, kParts, kTeams>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)avail));
        budget = avail;
        budget_dev.cur().store(avail, std::memory_order_relaxed);
    }
    TcParams Q = P;
    size_t smem;
    if (P.grouped) {
        const size_t bbytes = stage_bytes - kABytes;
        smem = fixed + (size_t)P.issuers * ((size_t)P.nA * P.a_stage_bytes + (size_t)P.nB * bbytes);
        MMC_UNSUPPORTED(smem > budget, "%s: shared memory budget exceeded (grouped rings)", name);
        Q.num_stages = P.nA;
    } else {
        MMC_UNSUPPORTED(fixed + 2 * stage_bytes > budget, "%s: shared memory budget exceeded", name);
        int stages = (int)((budget - fixed) / stage_bytes);
        if (stages > kMaxStages) stages = kMaxStages;
        Q.num_stages = stages;
        smem = fixed + (size_t)stages * stage_bytes;
    }
    // phase_inner: a CTA's index selects its spatial tiles, every CTA walks all phases (map_tile)
    const int work = Q.phase_inner ? Q.tiles_per_phase : Q.total_tiles;
    int grid = work < kNumSMs ? work : kNumSMs;
    auto set_virtual = [&](int g) {
        Q.n_virtual = Q.phase_inner ? ((Q.tiles_per_phase + g - 1) / g) * Q.n_phases * g : Q.total_tiles;
        Q.fd_tpp.set(Q.tiles_per_phase); Q.fd_nb.set(Q.n_blocks); Q.fd_tx.set(Q.tiles_x); Q.fd_ty.set(Q.tiles_y);
        Q.fd_grid.set(g); Q.fd_nph.set(Q.n_phases);
    };
    if (const char *g = getenv("MMC_TC_DEBUG")) Q.debug = atoi(g);
    Q.direct_store = 1;   // measured: g_a.0 1.04 -> 0.99 ms, g_s.2 0.38 -> 0.37 ms vs the staged, coalesced copy-out (MMC_TC_GDN_DIRECT=0)
    if (const char *g = getenv("MMC_TC_GDN_DIRECT")) Q.direct_store = atoi(g);
    if (Q.a_tmem) Q.direct_store = 1;    // no shared-memory x^2 tile to stage the copy-out in
    Q.late_release = 1;
    if (const char *g = getenv("MMC_TC_LATE_RELEASE")) Q.late_release = atoi(g) != 0;   // 0: round-1 behaviour (measurement aid)
    if (Q.epi_pipe) Q.late_release = 1;  // the pipelined epilogue hands the accumulator back from the norm-issuing thread (one arrival)
    if (const char *g = getenv("MMC_TC_GRID")) {   // profiling aid: restrict the persistent grid (profiles/probe_grid.py)
        int v = atoi(g);
        if (v >= 1 && v < grid) grid = v;
    }
    {
        int r = grid;   // step of the persistent loop, in the tile counter's radices
        Q.st_nb = r % Q.n_blocks; r /= Q.n_blocks;
        Q.st_tx = r % Q.tiles_x;  r /= Q.tiles_x;
        Q.st_ty = r % Q.tiles_y;  r /= Q.tiles_y;
        Q.st_b = r % Q.B;
        Q.st_ph = r / Q.B;
    }
    // Programmatic dependent launch between consecutive layers (see the griddepcontrol pair in the kernel).  MMC_TC_PDL=0: plain
    // stream order (measurement aid).
    static const bool use_pdl = getenv("MMC_TC_PDL") ? atoi(getenv("MMC_TC_PDL")) != 0 : false;   // opt-in until measured
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attrs[2];
    int nattr = 0;
    if (use_pdl) {
        attrs[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[nattr].val.programmaticStreamSerializationAllowed = 1;
        ++nattr;
    }
    cfg.blockDim = dim3(tc_threads(kParts, kTeams)); cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = attrs;
    if (kPair) {
        // 2-CTA clusters: even grid, as many pairs as the device can keep resident
        cudaLaunchAttribute &attr = attrs[nattr++];
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.numAttrs = nattr;
        static PerDevice<int> max_pairs_dev;
        int max_pairs = max_pairs_dev.cur().load(std::memory_order_relaxed);
        if (max_pairs == 0) {
            cfg.gridDim = dim3(kNumSMs & ~1);
            int n = 0;
            MMC_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<kEpi, kNCH, kPair, kParts, kTeams>, &cfg));
            max_pairs = n > 0 ? n : 1;
            max_pairs_dev.cur().store(max_pairs, std::memory_order_relaxed);
        }
        grid &= ~1;
        if (grid > 2 * max_pairs) grid = 2 * max_pairs;
        if (grid < 2) grid = 2;
        cfg.gridDim = dim3(grid);
        set_virtual(grid);
        MMC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<kEpi, kNCH, kPair, kParts, kTeams>, Q));
        count_launch();
        return MMC_OK;
    }
    cfg.numAttrs = nattr;
    cfg.gridDim = dim3(grid);
    set_virtual(grid);
    MMC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<kEpi, kNCH, kPair, kParts, kTeams>, Q));
    count_launch();
    return MMC_OK;
}

}  // namespace mmc

using namespace mmc;

extern "C" {

int mmc_conv_pad8_size(const mmc_conv_desc *d, int *Hp, int *Wp)
{
    MMC_CHECK_ARG(d && Hp && Wp, "mmc_conv_pad8_size: NULL argument");
    int Ho, Wo;
    int rc = mmc_conv_out_size(d, &Ho, &Wo);
    if (rc) return rc;
    pad8_extent(d, Ho, Wo, Hp, Wp);
    return MMC_OK;
}

int mmc_pad_nchw_to_nhwc8(const float *x, int64_t B, int C, int H, int W, int pad, int Hp, int Wp, void *out, void *stream)
{
    MMC_CHECK_ARG(B >= 0 && C >= 1 && C <= 8 && H >= 1 && W >= 1 && pad >= 0 && Hp >= H + pad && Wp >= W + pad,
                  "mmc_pad_nchw_to_nhwc8: bad shape");
    if (B == 0) return MMC_OK;
    MMC_CHECK_ARG(x && out && aligned16(out), "mmc_pad_nchw_to_nhwc8: NULL or unaligned buffer");
    MMC_CHECK_ARG(B <= 65535 && Hp <= 65535, "mmc_pad_nchw_to_nhwc8: B and the padded height must be <= 65535");
    const int qpr = (Wp + 3) / 4;
    pad8_kernel<<<dim3((unsigned)((qpr + 127) / 128), (unsigned)Hp, (unsigned)B), 128, 0, (cudaStream_t)stream>>>(x, C, H, W, pad, Hp, Wp, (uint4 *)out);
    MMC_CHECK_LAUNCH("mmc_pad_nchw_to_nhwc8");
    return MMC_OK;
}

int mmc_conv_pack_weights(const mmc_conv_desc *d, const float *w, void *w_packed, size_t *bytes, void *stream)
{
    Plan pl;
    int rc = make_plan(d, pl, "mmc_conv_pack_weights");
    if (rc) return rc;
    int64_t n = (int64_t)pl.ntaps * pl.wrows * pl.wcols;
    if (bytes) *bytes = (size_t)n * sizeof(__nv_bfloat16);
    if (!w_packed) return MMC_OK;
    MMC_CHECK_ARG(w != nullptr, "mmc_conv_pack_weights: w is NULL");
    PackParams q;
    q.mode = pl.mode; q.transposed = d->transposed; q.Cin = d->Cin; q.Cout = d->Cout; q.k = d->k; q.stride = d->stride;
    q.ntaps = pl.ntaps; q.wrows = pl.wrows; q.wcols = pl.wcols;
    for (int t = 0; t < kMaxTaps; ++t) { q.ta[t] = (int16_t)pl.tap_ky[t]; q.tb[t] = (int16_t)pl.tap_kx[t]; }
    pack_weights_kernel<<<elementwise_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(w, q, (__nv_bfloat16 *)w_packed);
    MMC_CHECK_LAUNCH("mmc_conv_pack_weights");
    return MMC_OK;
}

static int conv_forward_tc_impl(const mmc_conv_desc *d, const void *x, const void *x2, int cin1, const void *w_packed, const float *bias,
                                const float *beta_eff, const void *gamma_eff_bf16, void *y, void *y2, void *stream, const char *name);

int mmc_conv_forward_tc(const mmc_conv_desc *d, const void *x, const void *w_packed, const float *bias,
                        const float *beta_eff, const void *gamma_eff_bf16, void *y, void *y2, void *stream)
{
    return conv_forward_tc_impl(d, x, nullptr, d ? d->Cin : 0, w_packed, bias, beta_eff, gamma_eff_bf16, y, y2, stream, "mmc_conv_forward_tc");
}

int mmc_conv_forward_tc2(const mmc_conv_desc *d, const void *x1, int cin1, const void *x2, const void *w_packed, const float *bias,
                         const float *beta_eff, const void *gamma_eff_bf16, void *y, void *y2, void *stream)
{
    MMC_CHECK_ARG(d && x2 && cin1 > 0 && cin1 < d->Cin, "mmc_conv_forward_tc2: needs two sources with 0 < cin1 < Cin");
    MMC_UNSUPPORTED(cin1 % 64 != 0 || (d->Cin - cin1) % 8 != 0 || d->in_layout != MMC_NHWC,
                    "mmc_conv_forward_tc2: the first source must have a multiple of 64 channels, the second a multiple of 8 (NHWC bf16)");
    return conv_forward_tc_impl(d, x1, x2, cin1, w_packed, bias, beta_eff, gamma_eff_bf16, y, y2, stream, "mmc_conv_forward_tc2");
}

static int conv_forward_tc_impl(const mmc_conv_desc *d, const void *x, const void *x2, int cin1, const void *w_packed, const float *bias,
                                const float *beta_eff, const void *gamma_eff_bf16, void *y, void *y2, void *stream, const char *name)
{
    Plan pl;
    int rc = make_plan(d, pl, name);
    if (rc) return rc;
    MMC_CHECK_ARG(d->in_dtype == MMC_BF16 && (d->in_layout == MMC_NHWC || d->in_layout == MMC_NHWC_PAD8), "%s: input must be NHWC bf16", name);
    MMC_CHECK_ARG((d->act >= 0 && d->act <= MMC_ACT_LEAKY_RELU) || d->act == MMC_ACT_QRELU8, "%s: bad act", name);
    MMC_CHECK_ARG(d->gdn >= 0 && d->gdn <= MMC_GDN_INVERSE, "%s: bad gdn mode", name);
    MMC_CHECK_ARG((d->out2_bf16 >= 0 && d->out2_bf16 <= 2) || (d->out2_bf16 == 3 && d->gdn != MMC_GDN_NONE), "%s: bad out2_bf16 (3 = pre-GDN activations, needs a fused GDN)", name);
    MMC_CHECK_ARG(d->gdn == MMC_GDN_NONE || (beta_eff && gamma_eff_bf16), "%s: GDN needs beta/gamma", name);
    MMC_CHECK_ARG(!d->out2_bf16 || y2, "%s: out2_bf16 set but y2 is NULL", name);
    if (pl.mode == MODE_SCATTER) {
        MMC_CHECK_ARG(d->out_layout == MMC_NCHW && d->out_dtype == MMC_F32 && d->gdn == MMC_GDN_NONE && !d->out2_bf16,
                      "%s: narrow transposed conv writes planar fp32 NCHW without GDN / secondary output", name);
    } else {
        MMC_CHECK_ARG(d->out_layout == MMC_NHWC, "%s: output must be NHWC", name);
        MMC_UNSUPPORTED(d->gdn != MMC_GDN_NONE && (d->Cout > 192 || d->Cout % 64 != 0),
                        "%s: fused GDN supports Cout in {64,128,192} (got %d)", name, d->Cout);
    }
    if (d->B == 0) return MMC_OK;
    MMC_CHECK_ARG(x && w_packed && y, "%s: NULL buffer", name);
    MMC_CHECK_ARG(!x2 || (aligned16(x2) && pl.mode != MODE_PAD8), "%s: second source must be 16-byte aligned NHWC", name);
    MMC_CHECK_ARG(aligned16(x) && aligned16(w_packed) && aligned16(y) && (!y2 || aligned16(y2)), "%s: buffers must be 16-byte aligned", name);

    TcParams P;
    memset(&P, 0, sizeof(P));
    P.mode = pl.mode;
    P.Ho = pl.Ho; P.Wo = pl.Wo; P.B = d->B; P.Cout = d->Cout;
    P.kchunks = pl.kchunks; P.ksteps = pl.ksteps;
    P.kchunks1 = x2 ? cin1 / 64 : pl.kchunks;
    P.act = d->act; P.gdn = d->gdn; P.out_f32 = (d->out_dtype == MMC_F32); P.out2 = d->out2_bf16;
    P.bias = bias; P.beta = beta_eff; P.y = y; P.y2 = (__nv_bfloat16 *)y2;
    P.n_phases = pl.n_phases; P.a_sx = pl.a_sx; P.a_sy = pl.a_sy; P.out_stride = pl.out_stride; P.Gh = pl.Gh; P.Gw = pl.Gw;
    for (int t = 0; t < pl.ntaps; ++t) P.taps[t] = pl.taps[t];
    for (int i = 0; i < 5; ++i) P.phase_begin[i] = pl.phase_begin[i];

    // ---- tiling ----
    P.k = d->k; P.pad = d->k / 2; P.halo_lo = pl.halo_lo; P.halo_hi = pl.halo_hi;
    if (pl.mode == MODE_SCATTER) P.Ntile = pl.wrows;
    else P.Ntile = (d->gdn != MMC_GDN_NONE) ? d->Cout : pick_ntile(d->Cout);
    MMC_UNSUPPORTED(P.Ntile == 0 || P.Ntile > 256, "%s: no valid N tile for Cout=%d", name, d->Cout);
    P.n_blocks = (pl.mode == MODE_SCATTER) ? 1 : d->Cout / P.Ntile;
    if (pl.mode == MODE_SCATTER) {
        // overlapping input patches: consecutive tiles advance by the patch size minus the halo
        P.TH = 8; P.TW = 16;
        P.step_y = P.TH - pl.halo_lo - pl.halo_hi; P.step_x = P.TW - pl.halo_lo - pl.halo_hi;
        P.off_y = P.off_x = pl.halo_lo;
        P.spitch = P.Ntile | 1;                          // odd pitch (floats): see the staging stores in the kernel
        MMC_UNSUPPORTED(P.step_y * P.step_x * d->Cout > 128 * kScatterItems, "%s: more than %d col2im work items per tile", name, 128 * kScatterItems);
    } else {
        pick_tile(P.Gh, P.Gw, pl.mode == MODE_PAD8 ? 1 : P.a_sx, P.a_sy, &P.TH, &P.TW);
        // Tap groups (see struct Group): taps with the same x offset whose y offsets differ by multiples of the A-grid stride
        // read row shifts of one patch.  Needs tiles of 16 rows x 8 pixels (one swizzle atom per tile row).
        // OPT-IN (MMC_TC_GROUPED=1): with tap groups + epilogue teams enabled an intermittent illegal-address fault was seen in ~5 of
        // ~110 fresh-process runs (and in 1 of 8 ranks of an 8-GPU run), none in ~70 runs without them; not localised yet
        // (DESIGN.md section 8), so the default is the main loop that has no such record.  Measured gain when on: step 5.3 -> 5.05 ms.
        bool want_grouped = false;
        if (const char *g = getenv("MMC_TC_GROUPED")) want_grouped = atoi(g) != 0;
        want_grouped = want_grouped && pl.mode == MODE_STD && pl.ntaps > 1;
        if (want_grouped) {
            int ng = 0, nt = 0, max_taps = 1;
            bool ok = true;
            for (int ph = 0; ph < pl.n_phases && ok; ++ph) {
                P.group_begin[ph] = ng;
                bool used[kMaxTaps] = {false};
                for (int t0 = pl.phase_begin[ph]; t0 < pl.phase_begin[ph + 1] && ok; ++t0) {
                    if (used[t0]) continue;
                    // collect the taps of this phase with t0's dx and dy == t0's dy (mod a_sy), in ascending dy
                    int members[kMaxTaps], nm = 0;
                    for (int t = pl.phase_begin[ph]; t < pl.phase_begin[ph + 1]; ++t) {
                        if (used[t] || pl.taps[t].dx != pl.taps[t0].dx) continue;
                        if (((pl.taps[t].dy - pl.taps[t0].dy) % P.a_sy) != 0) continue;
                        members[nm++] = t;
                    }
                    for (int a = 0; a < nm; ++a)
                        for (int b2 = a + 1; b2 < nm; ++b2)
                            if (pl.taps[members[b2]].dy < pl.taps[members[a]].dy) { int tmp = members[a]; members[a] = members[b2]; members[b2] = tmp; }
                    // split where consecutive members are not exactly one A-grid row apart
                    int a = 0;
                    while (a < nm && ok) {
                        int b2 = a + 1;
                        while (b2 < nm && pl.taps[members[b2]].dy - pl.taps[members[b2 - 1]].dy == P.a_sy) ++b2;
                        if (ng >= kMaxGroups) { ok = false; break; }
                        P.groups[ng].ox = pl.taps[members[a]].dx;
                        P.groups[ng].oy = pl.taps[members[a]].dy;
                        P.groups[ng].ntaps = (uint8_t)(b2 - a);
                        P.groups[ng].tap0 = (uint8_t)nt;
                        for (int m = a; m < b2; ++m) { P.gbrow[nt++] = pl.taps[members[m]].brow; used[members[m]] = true; }
                        if (b2 - a > max_taps) max_taps = b2 - a;
                        ++ng;
                        a = b2;
                    }
                }
            }
            P.group_begin[pl.n_phases] = ng;
            if (ok && nt == pl.ntaps) {
                P.grouped = 1;
                P.TH = 16; P.TW = 8;
                P.a_stage_bytes = (P.TH + max_taps - 1) * P.TW * 128;
                if ((P.TH + max_taps - 1) * P.a_sy > 256) P.grouped = 0;     // TMA box limit
            }
            if (!P.grouped) pick_tile(P.Gh, P.Gw, P.a_sx, P.a_sy, &P.TH, &P.TW);
        }
        P.step_y = P.TH; P.step_x = P.TW; P.off_y = P.off_x = 0;
    }
    P.tiles_y = (P.Gh + P.step_y - 1) / P.step_y;
    P.tiles_x = (P.Gw + P.step_x - 1) / P.step_x;
    int64_t tpp = (int64_t)d->B * P.tiles_y * P.tiles_x * P.n_blocks;
    // CTA-pair kernel (cta_group::2): the 128-channel GDN layers.  One MMA covers two adjacent tiles (M = 256) and each CTA
    // supplies half of the weight rows, which halves the weight traffic and takes the shared-memory operand reads per MMA from
    // 8 KB to 6 KB per SM -- the N = 128 single-CTA MMA is bound by exactly that read bandwidth (profiles/README.md).
    const bool pair_ok = pl.mode == MODE_STD && d->Cout == 128 && P.Ntile == 128 && P.n_blocks == 1 && !d->out2_bf16;
    // Only where the main loop is long enough to stay the bottleneck once it runs twice as fast: with fewer than ~24 K blocks per
    // tile (the 2x2 .. 3x3-tap phases of the transposed convolutions) the GDN epilogue (~5.8k cycles per tile) takes over and
    // the cross-CTA hand-shake of the pair kernel only adds to it (measured: g_s.4 1.54 -> 1.66 ms, g_a.2 1.27 -> 1.07 ms).
    int min_kb = 1 << 30;
    for (int ph = 0; ph < pl.n_phases; ++ph) {
        const int kb = (pl.phase_begin[ph + 1] - pl.phase_begin[ph]) * pl.kchunks;
        if (kb < min_kb) min_kb = kb;
    }
    // (the same threshold holds for the plain bias / activation epilogue: ssf2020's 128 -> 128 encoder layers 0.90 -> 0.73 ms per GOP
    //  with the pair kernel, its transposed-conv decoder layers 0.92 -> 0.98 ms)
    // Round 2, after the lean issue loops and the late accumulator hand-back: the fused-GDN transposed convolutions gain from the
    // pair kernel too (probe on cfg 2, batch 64: g_s.0 0.131 -> 0.120 ms, g_s.2 0.364 -> 0.342, g_s.4 1.400 -> 1.334), so layers
    // with a GDN epilogue take it from 8 K blocks per tile on; the plain epilogue keeps the round-1 threshold.
    int need_kb = (d->gdn != MMC_GDN_NONE) ? 8 : 24;
    if (const char *g = getenv("MMC_TC_PAIR_MINKB")) need_kb = atoi(g);
    P.pair = (pair_ok && min_kb >= need_kb && tpp * pl.n_phases >= 4 * kNumSMs) ? 1 : 0;
    if (const char *g = getenv("MMC_TC_PAIR")) {   // 0: never, 2: whenever the shape allows it (tests), else the default rule
        if (atoi(g) == 0) P.pair = 0;
        if (atoi(g) == 2) P.pair = pair_ok ? 1 : 0;
    }
    if (P.pair) tpp = (tpp + 1) & ~(int64_t)1;     // even tile count per phase: a pair never straddles two phases
    MMC_CHECK_ARG(tpp * P.n_phases < (1ll << 31), "%s: too many tiles", name);
    P.tiles_per_phase = (int)tpp;
    P.total_tiles = (int)(tpp * P.n_phases);
    // Transposed convolutions with enough spatial tiles for many waves: run the stride^2 phases of a wave back to back, so that the
    // input is read from HBM once instead of once per phase (map_tile).  Small layers keep the phase-major order: it spreads
    // phases x tiles over the SMs, the other order only tiles.  MMC_TC_PHASE_INNER=0 / 1: never / whenever there are phases.
    // The phases of the LAST wave run on a partly empty machine, so the order pays only when that wave is a small part of the layer:
    // cfg 2, batch 64: g_s.4 (12 288 spatial tiles) 1.397 -> 1.343 ms, DRAM reads 1.61 -> 0.40 GB; g_s.2 (3 072) 0.334 -> 0.332 ms,
    // 0.40 -> 0.10 GB; g_s.0 (768 tiles = 5.2 waves) 0.105 -> 0.122 ms, hence the 4 % rule.
    {
        const int64_t waves = (tpp + kNumSMs - 1) / kNumSMs;
        P.phase_inner = (pl.mode == MODE_STD && pl.n_phases > 1 && waves * kNumSMs * 100 <= tpp * 104) ? 1 : 0;
    }
    if (const char *g = getenv("MMC_TC_PHASE_INNER")) P.phase_inner = (pl.mode == MODE_STD && pl.n_phases > 1 && atoi(g) != 0) ? 1 : 0;
    MMC_CHECK_ARG((tpp + kNumSMs) * P.n_phases < (1ll << 31), "%s: too many tiles", name);
    // GDN kernels only: for the plain bias / activation epilogue the extra MMA and the shared memory of the constant tiles cost more
    // than the 16 adds per chunk they replace (measured on ssf2020: 12.8 -> 14.1 ms per GOP with it)
    P.bias_mma = (d->gdn != MMC_GDN_NONE) ? 1 : 0;
    P.gdn_chunk = 0;
    if (d->gdn != MMC_GDN_NONE) P.gdn_chunk = (3 * P.Ntile <= 512) ? P.Ntile : P.Ntile / 2;
    // GDN with C <= 128: the x^2 operand of the norm contraction goes through TMEM (C / 2 columns), which frees its shared-memory
    // tile for one more pipeline stage and removes its write + read from the shared-memory port
    P.a_tmem = (d->gdn != MMC_GDN_NONE && d->Cout <= 128) ? 1 : 0;
    if (const char *g = getenv("MMC_TC_ATMEM")) P.a_tmem = P.a_tmem && atoi(g) != 0;
    P.acc_stages = (512 - P.gdn_chunk - (P.a_tmem ? d->Cout / 2 : 0)) / P.Ntile;      // as many accumulator stages as TMEM holds
    if (P.acc_stages > kMaxAccStages) P.acc_stages = kMaxAccStages;
    if (P.acc_stages < 1) P.acc_stages = 1;
    // Two GDN epilogue teams (single-CTA kernel, C in {64, 128}; see epilogue_gdn_teams): x^2 through ONE shared-memory tile, the
    // norm in two halves into a team-private scratch block, three accumulator stages: 3 C + 2 C / 2 <= 512 TMEM columns.
    // OPT-IN (MMC_TC_TEAMS=2), see the note at MMC_TC_GROUPED.  Measured gain when on: g_a.0 0.92 -> 0.68 ms.
    bool teams = false;
    if (const char *g = getenv("MMC_TC_TEAMS")) teams = atoi(g) == 2;
    teams = teams && d->gdn != MMC_GDN_NONE && !P.pair && (d->Cout == 64 || d->Cout == 128) && P.Ntile == d->Cout;
    if (teams) { P.acc_stages = 3; P.a_tmem = 1; P.gdn_chunk = 0; }
    // Software-pipelined GDN epilogue (epilogue_gdn_pipe): C in {64, 128}, bf16 NHWC output, no secondary output.  Two x^2 operand
    // blocks in TMEM (2 x C / 2 columns) next to the norm block, which fit the 64 columns the single-team layout leaves free at C = 128 (2 accumulator
    // stages either way).  MMC_TC_EPI_PIPE=0 / 1 overrides.
    bool epi_pipe = true;     // measured (cfg 2, batch 64): g_a.0 0.859 -> 0.783 ms, g_s.2 0.330 -> 0.321, g_s.4 1.306 -> 1.289; step -2 %
    if (const char *g = getenv("MMC_TC_EPI_PIPE")) epi_pipe = atoi(g) != 0;
    epi_pipe = epi_pipe && d->gdn != MMC_GDN_NONE && !teams && P.a_tmem && (d->Cout == 64 || d->Cout == 128) && P.Ntile == d->Cout &&
               !P.out_f32 && !P.out2 && pl.mode != MODE_SCATTER &&
               !(getenv("MMC_TC_ISSUERS") && atoi(getenv("MMC_TC_ISSUERS")) == 2);    // warp 3 is the norm issuer here
    if (epi_pipe) {
        P.epi_pipe = 1;
        P.acc_stages = (512 - P.gdn_chunk - d->Cout) / P.Ntile;
        if (P.acc_stages > kMaxAccStages) P.acc_stages = kMaxAccStages;
    }
    // col2im epilogue teams (4 warps each).  Default 3: measured on cfg 2 (g_s.6, batch 64) -- see the table in DESIGN.md 4.1
    int sc_teams = 3;
    if (const char *g = getenv("MMC_TC_SCATTER_TEAMS")) { const int v = atoi(g); if (v >= 2 && v <= 4) sc_teams = v; }
    if (pl.mode == MODE_SCATTER) {
        // every team has its own fp32 staging tile: keep room for the resident weights and three activation stages
        const size_t staging = (((size_t)128 * P.spitch * sizeof(float)) + 1023) & ~(size_t)1023;
        const size_t b_res = (size_t)pl.ntaps * pl.kchunks * P.Ntile * 128;
        while (sc_teams > 2 && (P.acc_stages < sc_teams || sc_teams * staging + b_res + 3 * kABytes > (size_t)200 * 1024)) --sc_teams;
        P.acc_stages -= P.acc_stages % sc_teams;    // the col2im epilogue teams own the accumulator stages round robin
        MMC_UNSUPPORTED(P.acc_stages < 2, "%s: the reconstruction kernel needs two accumulator stages (N tile %d)", name, P.Ntile);
    }
    MMC_UNSUPPORTED((teams ? 3 * P.Ntile + d->Cout : P.acc_stages * P.Ntile + P.gdn_chunk + (P.a_tmem ? (P.epi_pipe ? d->Cout : d->Cout / 2) : 0)) > 512 || (P.gdn_chunk % 16) != 0 || (P.epi_pipe && P.acc_stages < 2),
                    "%s: TMEM budget exceeded", name);

    size_t fixed = 1024;  // alignment slack
    if (d->gdn != MMC_GDN_NONE) fixed += (size_t)d->Cout * d->Cout * 2 + (P.a_tmem ? 0 : (size_t)(d->Cout / 64) * kABytes);
    if (pl.mode == MODE_SCATTER) fixed += sc_teams * ((((size_t)128 * P.spitch * sizeof(float)) + 1023) & ~(size_t)1023);   // one staging buffer per epilogue team
    // Small layers (image-edge conv, reconstruction deconv): keep every weight tile resident in shared memory so that
    // the K blocks stream activations only (halves the L2 -> SM traffic of those layers).
    const size_t b_total = (size_t)pl.ntaps * pl.kchunks * P.Ntile * 128;
    P.b_resident = (!P.pair && P.n_blocks == 1 && P.n_phases == 1 && b_total <= 96 * 1024 && fixed + b_total + 3 * kABytes <= 204 * 1024) ? 1 : 0;
    if (P.b_resident) fixed += b_total;
    const size_t stage_bytes = kABytes + (P.b_resident ? 0 : (size_t)(P.pair ? P.Ntile / 2 : P.Ntile) * 128);
    if (P.grouped) {
        // ring depths from a conservative dynamic shared-memory budget (227 KB minus the largest static footprint of the kernels);
        // launch_tc re-checks against the real one
        const size_t avail = (size_t)206 * 1024 - fixed, bbytes = (size_t)(P.pair ? P.Ntile / 2 : P.Ntile) * 128;
        // rings per issuer: nA patches and nB weight tiles.  Default ONE issuer: two issuers (MMC_TC_ISSUERS=2) halve the ring depth
        // each one sees, and the operand stream is latency-bound (measured: g_a.2 1.03 -> 1.15 ms, g_s.4 1.36 -> 1.43 ms)
        int issuers = 1;
        if (const char *g = getenv("MMC_TC_ISSUERS")) issuers = atoi(g) == 2 ? 2 : 1;
        int nA = issuers == 2 ? 2 : 3;
        if (avail < (size_t)issuers * ((size_t)nA * P.a_stage_bytes + 3 * bbytes)) nA = 2;
        int nB = (P.b_resident || avail < (size_t)issuers * nA * P.a_stage_bytes) ? 0 : (int)((avail - (size_t)issuers * nA * P.a_stage_bytes) / (issuers * bbytes));
        if (nB > kMaxStages / issuers) nB = kMaxStages / issuers;
        P.issuers = issuers;
        if (P.b_resident || nB < 2) {
            // does not fit (or resident weights): fall back to one (A tile, B tile) stage per K block -- the tile shape stays valid
            P.grouped = 0;
        } else {
            P.nA = nA; P.nB = nB;
        }
    }

    // ---- tensor maps ----
    if (pl.mode == MODE_PAD8) {
        // Overlapping-window view of the padded NHWC8 image: dim0 = 64 contiguous elements (8 pixels x 8 ch),
        // dim1 = output column (stride * 16 B apart), dim2 = padded row, dim3 = image.
        uint64_t dims[4] = {64, (uint64_t)pl.Wo, (uint64_t)pl.Hp, (uint64_t)d->B};
        uint64_t str[3] = {(uint64_t)d->stride * 16, (uint64_t)pl.Wp * 16, (uint64_t)pl.Hp * pl.Wp * 16};
        uint32_t box[4] = {64, (uint32_t)P.TW, (uint32_t)(P.TH * P.a_sy), 1};
        uint32_t es[4] = {1, 1, (uint32_t)P.a_sy, 1};
        rc = encode_map(&P.tmA, x, 4, dims, str, box, es, "padded image");
    } else {
        // one tensor map per activation source: the K loop takes its 64-channel boxes from the first source, then from the second
        // (the concatenation of two feature maps along the channels is never materialised)
        const int c1 = x2 ? cin1 : d->Cin, c2 = d->Cin - c1;
        uint64_t dims[4] = {(uint64_t)c1, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
        uint64_t str[3] = {(uint64_t)c1 * 2, (uint64_t)d->W * c1 * 2, (uint64_t)d->H * d->W * c1 * 2};
        // grouped: the box is the whole patch (tile rows + the rows the other taps of a group reach)
        const int box_rows = P.grouped ? P.a_stage_bytes / (P.TW * 128) : P.TH;
        uint32_t box[4] = {64, (uint32_t)(P.TW * P.a_sx), (uint32_t)(box_rows * P.a_sy), 1};
        uint32_t es[4] = {1, (uint32_t)P.a_sx, (uint32_t)P.a_sy, 1};
        rc = encode_map(&P.tmA, x, 4, dims, str, box, es, "activations");
        if (!rc && x2) {
            uint64_t dims2[4] = {(uint64_t)c2, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
            uint64_t str2[3] = {(uint64_t)c2 * 2, (uint64_t)d->W * c2 * 2, (uint64_t)d->H * d->W * c2 * 2};
            rc = encode_map(&P.tmA2, x2, 4, dims2, str2, box, es, "activations (second source)");
        }
    }
    if (rc) return rc;
    {
        uint64_t dims[2] = {(uint64_t)pl.wcols, (uint64_t)pl.ntaps * pl.wrows};
        uint64_t str[1] = {(uint64_t)pl.wcols * 2};
        uint32_t box[2] = {64, (uint32_t)(P.pair ? P.Ntile / 2 : P.Ntile)};
        uint32_t es[2] = {1, 1};
        rc = encode_map(&P.tmB, w_packed, 2, dims, str, box, es, "weights");
        if (rc) return rc;
    }
    if (d->gdn != MMC_GDN_NONE) {
        MMC_CHECK_ARG(aligned16(gamma_eff_bf16), "%s: gamma must be 16-byte aligned", name);
        uint64_t dims[2] = {(uint64_t)d->Cout, (uint64_t)d->Cout};
        uint64_t str[1] = {(uint64_t)d->Cout * 2};
        uint32_t box[2] = {64, (uint32_t)d->Cout};
        uint32_t es[2] = {1, 1};
        rc = encode_map(&P.tmG, gamma_eff_bf16, 2, dims, str, box, es, "gamma");
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (pl.mode == MODE_SCATTER) {
        if (sc_teams == 4) return launch_tc<EPI_SCATTER, 0, false, 4>(P, fixed, stage_bytes, st, name);
        if (sc_teams == 3) return launch_tc<EPI_SCATTER, 0, false, 3>(P, fixed, stage_bytes, st, name);
        return launch_tc<EPI_SCATTER, 0>(P, fixed, stage_bytes, st, name);
    }
    if (d->gdn != MMC_GDN_NONE) {
        // one kernel per channel count (16-column chunks per epilogue thread = Cout / 32) so that each gets its own
        // register allocation: C=128 keeps 64 activations per thread in registers, C=192 keeps 96
        // 8 epilogue warps (2 column parts); C = 192 has two 96-column norm groups
        // (measured: 16 epilogue warps are SLOWER than 8 -- g_a.0 1.05 -> 1.39 ms -- so 8 stays the default; the 4-part
        //  instantiation is kept behind MMC_TC_PARTS=4 for experiments)
        const bool wide = getenv("MMC_TC_PARTS") ? atoi(getenv("MMC_TC_PARTS")) == 4 : false;
        if (d->Cout == 128 && P.pair) return wide ? launch_tc<EPI_GDN, 2, true, 4>(P, fixed, stage_bytes, st, name)
                                                  : launch_tc<EPI_GDN, 4, true, 2>(P, fixed, stage_bytes, st, name);
        if (d->Cout == 128 && teams) return launch_tc<EPI_GDN, 8, false, 1, 2>(P, fixed, stage_bytes, st, name);
        if (d->Cout == 128) return wide ? launch_tc<EPI_GDN, 2, false, 4>(P, fixed, stage_bytes, st, name)
                                        : launch_tc<EPI_GDN, 4, false, 2>(P, fixed, stage_bytes, st, name);
        if (d->Cout == 192) return launch_tc<EPI_GDN, 6>(P, fixed, stage_bytes, st, name);
        if (teams) return launch_tc<EPI_GDN, 4, false, 1, 2>(P, fixed, stage_bytes, st, name);
        return launch_tc<EPI_GDN, 2>(P, fixed, stage_bytes, st, name);
    }
    if (P.pair) return launch_tc<EPI_PLAIN, 0, true, 2>(P, fixed, stage_bytes, st, name);
    return launch_tc<EPI_PLAIN, 0>(P, fixed, stage_bytes, st, name);
}

}  // extern "C"
