// 3x3 (+ fused 1x1 skip) triplane convolution as an implicit GEMM on the sm_100a tensor cores.
//
//   reference: TriplaneConv.forward, src/diffusion/unet_triplane.py:31-60 (the per-plane nn.Conv2d 3x3, pad 1)
//              + TriplaneResBlock skip_connection 1x1 (unet_triplane.py:255, 308-311)
//
// GEMM view (per plane, per sample):  D[pixel, co] = sum_{tap, c} A[pixel + tap, c] * W[co, tap*C + c]
//   M tile  = 128 pixels = an 8 x 16 image patch; the A operand of tap (kh, kw) is the same patch shifted by
//             (kh-1, kw-1), fetched by ONE 5-D TMA box {64 ch, 16, 8, 1, 1} whose out-of-bounds elements
//             (negative / >= size coordinates) are zero-filled by the TMA unit == the conv's zero padding.
//   N tile  = 64 output channels, K chunk = 64 input channels of one tap (128 B rows, SWIZZLE_128B).
//   The rollout's broadcast channels never enter this GEMM (folded into the 1-D terms added in the epilogue).
//
// Precision: operands are fp16 (hi, lo) pairs, v = hi + lo/2048.  NSPLIT == 3 issues
//   D1 += Ah*Bh ; D2 += Al*Bh ; D2 += Ah*Bl      (tcgen05.mma kind::f16, fp32 accumulate in TMEM)
// and the epilogue forms D1 + D2/2048 (error ~2^-22, i.e. fp32-grade).  NSPLIT == 1 issues only Ah*Bh.
//
// Warp roles (192 threads, 1 CTA/SM): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> +bias +rollout 1-D terms +residual -> fp32 NHWC).
#pragma once
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace s3d {

constexpr int kTileH = 16, kTileW = 8;     // conv M tile = 16 rows x 8 cols of one plane (m = h*8 + w)
// Halo patch of one 64-channel block: 18 rows x kHaloW pixels, one 128-byte line per pixel.  The tensor core takes the swizzle phase
// from the absolute shared-memory address and the distance between the 8-pixel row groups from the descriptor's SBO, so the row
// pitch needs no power-of-two padding: 10 pixels (tile width + 2) instead of 16 shrinks the patch from 36 to 22.5 KiB, and the
// shared memory that frees buys a weight ring deep enough to cover the L2 latency of the weight tiles.
#ifndef S3D_HALO_W
#define S3D_HALO_W 10
#endif
constexpr int kHaloH = kTileH + 2, kHaloW = S3D_HALO_W;
static_assert(kHaloW >= kTileW + 2, "halo width");
constexpr int kBM = 128, kBN = 64, kBK = 64;
constexpr int kABytes = kBM * kBK * 2;             // 16 KiB: plain 128-row A tile (skip / rollout groups)
constexpr int kAHaloBytes = kHaloH * kHaloW * kBK * 2;   // bytes one halo TMA box delivers
constexpr int kAHaloPad = (kAHaloBytes + 1023) / 1024 * 1024;   // its footprint: the next operand starts 1024-byte aligned
constexpr int kBBytes = kBN * kBK * 2;             //  8 KiB
constexpr int kConvThreads = 352;                  // warps: 0 A-producer, 1 MMA, 2..9 epilogue, 10 B-producer
constexpr int kEpiWarps = 8, kEpiThreads = kEpiWarps * 32;
constexpr int kRollThreads = 192;

struct ConvTcMaps {
    CUtensorMap a[3];   // activations (C, cols, rows, B, 2) fp16, box {64, 16, 18}: halo patch
    CUtensorMap x[3];   // skip input  (Cs, cols, rows, B, 2) fp16, box {64, 8, 16} (unused when Cs == 0)
    CUtensorMap w[3];   // weights     (Ktot, Cout, 2) fp16, K = tap*C + c, then the Cs skip channels
};

struct ConvTcArgs {
    TriDims d;
    int C, Cout, Cs;
    int tiles_x[3];
    int tile_start[4];   // prefix sum of tiles per plane
    ConvEpi e;
    StatsSink sink;      // GroupNorm group sums of the output (sink.acc == nullptr: none); needs 64 % (Cout/32) == 0
    // Fused TriplaneDownsample2x (unet_triplane.py:127-145, avg_pool2d k2 s2, floor on odd sizes): when pool.p[0] != nullptr the
    // epilogue also writes the 2x2-averaged output [B][rows/2][cols/2][Cout] and its GroupNorm group sums (a tile is 16 x 8 pixels
    // at an even origin, so every pooled pixel lies inside one warp's rows).  Saves the k_avgpool2 launch and its re-read.
    TriF pool;
    StatsSink pool_sink;
    Trace tr;            // opt-in phase stamps (common.cuh)
    int bo_kw;           // bring-up switch (S3D_HALO_BO_KW=1): put kw into the descriptor base_offset (measured WRONG on B200:
                         // the tensor core derives the swizzle phase from the absolute shared-memory address)
};

struct RollTcMaps {
    CUtensorMap a[6];   // means of source s: (C, L, 1, B, 2) fp16, box {64, 128, 1}  (stand-alone k_roll_tc only)
    CUtensorMap w[6];   // class-summed 1-D weights: (3C, 4*Cout, 2) fp16, K = along*C + c
};
struct RollTcArgs {
    int L[6], ncls[6];
    int tm[6];          // positions per roll tile of source s (<= 96): L is cut into equal tiles so that no tile's A patch (64-bit axis
                        // sums read through ONE SM's load path: what bounds the roll chain every conv epilogue waits for) is a straggler
    float* T[6];        // [B][4][L][Cout]
    int tile_start[7];
    int C, Cout;
    int soff[6];        // position offset of source s inside a sample's block of axis sums
    float scale[6];     // 2^-24 / (length of the averaged axis): fixed-point sum -> mean
};
// Fused launch: the rollout 1-D GEMM tiles ("roll tiles") are the first tile indices of the persistent conv kernel; the
// conv tiles' epilogues wait on a device counter until every roll tile has been written (roll tiles never wait on
// anything and are the first tiles of the lowest-numbered CTAs, so the wait cannot deadlock).
// A roll tile's A operand is built in place: the (otherwise idle) epilogue warps read the 64-bit fixed-point axis sums that
// k_gn_silu accumulated, turn them into (hi, lo) fp16 means and write them into an A slot in the 128-byte-swizzled layout the
// tensor core expects — one 130-row patch per 64-channel block, which the three taps of the 1-D conv address with a start
// offset of 0 / 1 / 2 rows.  No conversion pass, no grid-wide wait, no TMA round trip in front of the roll tiles.
struct FusedRoll {
    RollTcArgs R;
    int n_roll;              // number of roll tiles (0: none; Trow/Tcol come from a separate launch or are absent)
    int ntn;                 // N tiles per roll M tile
    unsigned int* counters;  // [2]: roll tiles done, CTAs exited (both return to 0 when the kernel ends)
    const unsigned long long* sums;   // [B][total_len][C] (re-zeroed by the next k_gn_silu of the step)
    int total_len;
};

// Precision modes (operands are fp16 (hi, lo) pairs, v = hi + lo/2048; fp32 accumulation in TMEM):
//   1: Ah*Bh                           one MMA, fp16-grade products
//   2: Ah*Bh + Ah*Bl                   one N=128 MMA (lo weight tile behind the hi tile): exact weights, fp16-rounded activations
//   4: Ah*Bh + Al*Bh                   two MMAs: exact activations, fp16-rounded weights
//   3: Ah*Bh + Ah*Bl + Al*Bh           fp32-grade (error ~2^-22)
//   5: mode 2 for the conv tiles, mode 3 for the rollout 1-D GEMM tiles (their A operand is built in shared memory from the fp32
//      axis means, so its lo half costs no memory traffic: it sits behind the hi patch inside the same A slot)
template <int MODE>
struct ConvTcCfg {
    static constexpr bool kAlo = MODE == 3 || MODE == 4;      // the lo halves of the activations are loaded and multiplied
    static constexpr bool kBlo = MODE == 2 || MODE == 3 || MODE == 5;      // the lo halves of the weights
    static constexpr bool kTwo = MODE != 1;                   // a second accumulator D2 (scaled by 1/2048 in the epilogue)
    static constexpr bool kRollAlo = kAlo || MODE == 5;       // roll tiles multiply the lo halves of the means
    static constexpr int kRollLoOff = kAlo ? kAHaloPad : 18 * 1024;   // where a roll tile's lo patch sits inside its A slot
    static constexpr int kASlotBytes = kAlo ? 2 * kAHaloPad : (MODE == 5 ? 36 * 1024 : kAHaloPad);   // hi at +0, lo at +kAHaloPad
    static constexpr int kBSlotBytes = (kBlo ? 2 : 1) * kBBytes;       // hi at +0, lo at +kBBytes
    static constexpr int kASlots = kAlo ? 2 : 3;
    static constexpr int kEpiBytes = kEpiWarps * 32 * 32 * 4 + 64;   // epilogue staging: [32 rows][32 cols] fp32 per epilogue warp
    // weight ring: whatever shared memory is left (the barrier block holds 30 mbarriers at most)
    static constexpr int kBBudget = (227 * 1024 - 1024 - 256 - kEpiBytes - kASlots * kASlotBytes) / kBSlotBytes;
    static constexpr int kBCap = (30 - 5 - 2 * kASlots) / 2;
    static constexpr int kBSlots = kBBudget < kBCap ? kBBudget : kBCap;
    static_assert(kBSlots >= 3, "weight ring too shallow");
    static_assert(kABytes + (kAlo ? kAHaloPad : 0) <= kASlotBytes && kAHaloPad >= kABytes, "a plain skip tile must fit the A slot");
    static constexpr int kRingBytes = kASlots * kASlotBytes + kBSlots * kBSlotBytes;
    static constexpr int kAccCols = kTwo ? 128 : 64;            // TMEM columns of one accumulator stage
    static constexpr int kTmemCols = 2 * kAccCols;              // double-buffered accumulators
    static constexpr int kSmemBytes = kRingBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiBytes;
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
    static_assert(2 * kASlots + 2 * kBSlots + 5 <= 30, "barrier block");
    // stand-alone k_roll_tc (MODE 1 or 3 only) keeps the simple 4-stage {A,B} ring
    static constexpr int kStageBytes = (MODE == 3 ? 2 : 1) * (kABytes + kBBytes);
    static constexpr int kStages = MODE == 3 ? 4 : 8;
    static constexpr int kRollSmemBytes = kStages * kStageBytes + 1024 + 256;
};

struct ConvTile {
    int ip;    // tile index inside its plane
    int plane, h0, w0, n0, b;
};
__device__ __forceinline__ ConvTile conv_tile_decode(const ConvTcArgs& A, int t) {
    const int ts = A.tile_start[3];
    const int nt_n = A.Cout / kBN;
    ConvTile T;
    T.b = t / (ts * nt_n);
    int rem = t - T.b * ts * nt_n;
    const int nt = rem / ts;
    rem -= nt * ts;
    T.n0 = nt * kBN;
    T.plane = rem >= A.tile_start[2] ? 2 : (rem >= A.tile_start[1] ? 1 : 0);
    const int ip = rem - A.tile_start[T.plane];
    T.ip = ip;
    const int ty = ip / A.tiles_x[T.plane];
    T.h0 = ty * kTileH;
    T.w0 = (ip - ty * A.tiles_x[T.plane]) * kTileW;
    return T;
}

struct RollTile {
    int src, p0, n0, b;
};
__device__ __forceinline__ RollTile roll_tile_decode(const FusedRoll& F, int t) {
    const int mt_total = F.R.tile_start[6];
    RollTile T;
    T.b = t / (mt_total * F.ntn);
    int rem = t - T.b * mt_total * F.ntn;
    const int nt = rem / mt_total;
    rem -= nt * mt_total;
    T.n0 = nt * kBN;
    int src = 0;
#pragma unroll
    for (int k = 1; k < 6; ++k)
        if (rem >= F.R.tile_start[k]) src = k;
    T.src = src;
    T.p0 = (rem - F.R.tile_start[src]) * F.R.tm[src];
    return T;
}

constexpr int kRollTmMax = 96;           // most positions a roll tile produces (the MMA is M = 128 regardless: rows beyond tm are don't-care)
constexpr int kRollRows = kRollTmMax + 2;   // most rows of a roll tile's A patch: positions p0-1 .. p0+tm
static_assert((kRollRows + 32) * 128 <= 18 * 1024 && (kRollRows + 32) * 128 <= kAHaloPad,
              "a roll tile's hi patch (the M = 128 MMA reads 2 + 128 rows) must end before its lo patch and inside an A slot");

// Persistent implicit-GEMM convolution, one CTA per SM, tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...
//
// Operand traffic is what bounds the main loop (L2 -> SM), so the A operand is fetched ONCE per 64-channel block as an
// 18 x 16-pixel halo patch and all nine taps read it in place: tap (kh, kw) is the same shared-memory patch addressed from
// row (kh*16 + kw) with an 8-row-group stride of 2048 B (UMMA descriptor start / SBO), i.e. 1/6 of the activation bytes of a
// tap-by-tap im2col.  Only the 8 KiB weight tile changes per tap.
// Warps: 0 = A producer (TMA), 6 = B producer (TMA), 1 = MMA issuer, 2..5 = epilogue.  Accumulators are double-buffered in
// TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
// Epilogue: every epilogue warp is self-contained (no CTA-wide barrier): it owns the 32 accumulator rows of its TMEM lane
// quarter (4 image rows x 8 pixels), moves them 32 columns at a time TMEM -> registers -> its own 4 KiB staging block
// (16-byte XOR swizzle), and re-reads the block transposed so that 8 consecutive lanes cover 128 contiguous bytes of one pixel:
// + bias + rollout 1-D terms + residual, fp32 NHWC store, and the GroupNorm group sums of what it stored (fixed-point atomics).
template <int MODE>
__global__ void __launch_bounds__(kConvThreads, 1) k_conv_tc(const __grid_constant__ ConvTcMaps M,
                                                             const __grid_constant__ RollTcMaps RM, const ConvTcArgs A,
                                                             const FusedRoll F, const int total_tiles) {
    using Cfg = ConvTcCfg<MODE>;
    constexpr bool kAlo = Cfg::kAlo, kBlo = Cfg::kBlo, kTwo = Cfg::kTwo;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operands need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + Cfg::kASlots * Cfg::kASlotBytes;
    uint64_t* fullA = reinterpret_cast<uint64_t*>(smem + Cfg::kRingBytes);
    uint64_t* emptyA = fullA + Cfg::kASlots;
    uint64_t* fullB = emptyA + Cfg::kASlots;
    uint64_t* emptyB = fullB + Cfg::kBSlots;
    uint64_t* tmem_full_bar = emptyB + Cfg::kBSlots;         // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;            // [2]
    uint64_t* roll_filled_bar = tmem_empty_bar + 2;          // the CTA's last roll tile has written its A patches
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(roll_filled_bar + 1);
    float* stage = reinterpret_cast<float*>(smem + Cfg::kRingBytes + 256);      // [4 warps][32 rows][32 cols], 16-byte chunk k of row r at k ^ (r & 7)

    // warp index through a shuffle: provably warp-uniform, so that everything the single-issuer roles compute from it lives in
    // uniform registers (a role written under `if (lane == 0)` makes the compiler wrap every TMA / MMA issue in an
    // elect-and-broadcast loop: measured ~117 cycles per tcgen05.mma instead of the 32-64 the tensor core needs)
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // trace slots: 0 entry, 1 set-up done, 3 first roll patch filled, 22 exit; per local tile lt0 <= lt < lt0 + 3 (S3D_TRACE_LT0, default 0) at 4 + 6*(lt - lt0):
    // +0 first operands landed, +1 all MMAs issued, +2 epilogue addends of the first batch requested, +3 accumulator complete,
    // +4 epilogue done, +5 A loads issued
    if (threadIdx.x == 0) trace_mark(A.tr, 0);
    const int cblks = A.C / kBK;
    const int nskip = A.Cs / kBK;
    constexpr uint32_t kALo = kAHaloPad;
    constexpr uint32_t kAStdTx = (kAlo ? 2 : 1) * kABytes, kAHaloTx = (kAlo ? 2 : 1) * kAHaloBytes;

    if (warp == 0 && lane == 0) {
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            ptx::prefetch_tmap(&M.a[p]);
            ptx::prefetch_tmap(&M.w[p]);
            if (A.Cs) ptx::prefetch_tmap(&M.x[p]);
        }
        if (F.n_roll) {
#pragma unroll
            for (int p = 0; p < 6; ++p) ptx::prefetch_tmap(&RM.w[p]);
        }
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < Cfg::kASlots; ++s) {
                ptx::mbar_init(&fullA[s], 1);
                ptx::mbar_init(&emptyA[s], 1);
            }
            for (int s = 0; s < Cfg::kBSlots; ++s) {
                ptx::mbar_init(&fullB[s], 1);
                ptx::mbar_init(&emptyB[s], 1);
            }
            for (int s = 0; s < 2; ++s) {
                ptx::mbar_init(&tmem_full_bar[s], 1);
                ptx::mbar_init(&tmem_empty_bar[s], kEpiWarps);       // one arrive per epilogue warp
            }
            ptx::mbar_init(roll_filled_bar, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (threadIdx.x == 0) trace_mark(A.tr, 1);
    // PDL: everything above (barriers, TMEM, descriptor prefetch) and the weight-tile producer below overlap the tail of the
    // previous kernel; threads that touch activations / statistics / rollout terms wait for it first.
    pdl_trigger();

    if (warp == 0) {
        // ===================== TMA producer: A operand groups (conv tiles; roll tiles are filled by the epilogue warps) =====
        // The whole warp walks the loop converged; one elected lane issues.
        if (lane == 0) trace_mark(A.tr, 24);
        pdl_wait();
        int ga = 0, ltp = 0;
        bool ring_synced = false;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ltp) {
            if (t < F.n_roll) {
                ga += cblks;
                continue;
            }
            // The A ring has two producers: the epilogue warps fill it for the roll tiles, this warp for the conv tiles.  A
            // parity wait only tells neighbouring phases apart, so this warp may not enter the ring two or more phases ahead
            // of the consumer: when the CTA's roll tiles take 2 * kASlots groups or more (several roll tiles per CTA at large
            // batch), wait until the last of them has been written (then every earlier group but the last per slot has been
            // released, and the in-order chain below is exact again).
            if (!ring_synced) {
                ring_synced = true;
                if (ga >= 2 * Cfg::kASlots) ptx::mbar_wait(roll_filled_bar, 0);
            }
            const ConvTile T = conv_tile_decode(A, t - F.n_roll);
            if (ltp == 0 && lane == 0) trace_mark(A.tr, 25);
            for (int cb = 0; cb < cblks; ++cb, ++ga) {
                const int s = ga % Cfg::kASlots;
                ptx::mbar_wait(&emptyA[s], ((ga / Cfg::kASlots) & 1) ^ 1);
                if (ltp == 0 && cb == 0 && lane == 0) trace_mark(A.tr, 26);
                uint8_t* st = smem_a + s * Cfg::kASlotBytes;
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(&fullA[s], kAHaloTx);
                    ptx::tma_load_5d(st, &M.a[T.plane], &fullA[s], cb * kBK, T.w0 - 1, T.h0 - 1, T.b, 0);
                    if (kAlo) ptx::tma_load_5d(st + kALo, &M.a[T.plane], &fullA[s], cb * kBK, T.w0 - 1, T.h0 - 1, T.b, 1);
                }
                __syncwarp();
            }
            for (int j = 0; j < nskip; ++j, ++ga) {
                const int s = ga % Cfg::kASlots;
                ptx::mbar_wait(&emptyA[s], ((ga / Cfg::kASlots) & 1) ^ 1);
                uint8_t* st = smem_a + s * Cfg::kASlotBytes;
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(&fullA[s], kAStdTx);
                    ptx::tma_load_5d(st, &M.x[T.plane], &fullA[s], j * kBK, T.w0, T.h0, T.b, 0);
                    if (kAlo) ptx::tma_load_5d(st + kALo, &M.x[T.plane], &fullA[s], j * kBK, T.w0, T.h0, T.b, 1);
                }
                __syncwarp();
            }
            if (static_cast<unsigned>(ltp - A.tr.lt0) < 3u && lane == 0) trace_mark(A.tr, 4 + 6 * (ltp - A.tr.lt0) + 5);
        }
    } else if (warp == 2 + kEpiWarps) {
        // ===================== TMA producer: B (weight) tiles =====================
        if (lane == 0) trace_mark(A.tr, 27);
        int gb = 0;
        auto load_b = [&](const CUtensorMap* wm, int kchunk, int n0) {
            const int s = gb % Cfg::kBSlots;
            ptx::mbar_wait(&emptyB[s], ((gb / Cfg::kBSlots) & 1) ^ 1);
            uint8_t* st = smem_b + s * Cfg::kBSlotBytes;
            if (ptx::elect_one()) {
                // ONE box {64 K, 64 N, hi|lo}: the lo tile lands right behind the hi tile (a TMA op costs ~250 cycles of producer
                // time whatever its size, tools/tma_bw.cu)
                ptx::mbar_arrive_expect_tx(&fullB[s], Cfg::kBSlotBytes);
                ptx::tma_load_3d(st, wm, &fullB[s], kchunk * kBK, n0, 0);
            }
            __syncwarp();
            if (gb == 0 && lane == 0) trace_mark(A.tr, 28);
            if (gb == 8 && lane == 0) trace_mark(A.tr, 29);
            ++gb;
        };
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            if (t < F.n_roll) {
                const RollTile T = roll_tile_decode(F, t);
                for (int cb = 0; cb < cblks; ++cb)
                    for (int al = 0; al < 3; ++al) load_b(&RM.w[T.src], al * cblks + cb, T.n0);
                continue;
            }
            const ConvTile T = conv_tile_decode(A, t - F.n_roll);
            for (int cb = 0; cb < cblks; ++cb)
                for (int tap = 0; tap < 9; ++tap) load_b(&M.w[T.plane], tap * cblks + cb, T.n0);
            for (int j = 0; j < nskip; ++j) load_b(&M.w[T.plane], 9 * cblks + j, T.n0);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = ptx::make_idesc_f16(kBM, kBN);
        constexpr uint32_t idesc2 = ptx::make_idesc_f16(kBM, 2 * kBN);
        int ga = 0, gb = 0, lt = 0;
        if (lane == 0) trace_mark(A.tr, 31);
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
            const int as = lt & 1;
            ptx::mbar_wait(&tmem_empty_bar[as], ((lt >> 1) & 1) ^ 1);        // epilogue has drained this accumulator stage
            ptx::tc_fence_after();
            const uint32_t d1 = tmem_base + as * Cfg::kAccCols, d2 = d1 + kBN;
            const bool roll = t < F.n_roll;
            const int n_patch = cblks;                       // groups whose taps share one A patch (halo / roll patch)
            const int n_groups = roll ? cblks : cblks + nskip;
            uint32_t acc = 0u;                               // 0 for the very first MMA of the tile
            for (int g = 0; g < n_groups; ++g, ++ga) {
                const int sa = ga % Cfg::kASlots;
                ptx::mbar_wait(&fullA[sa], (ga / Cfg::kASlots) & 1);
                const uint32_t a_base = ptx::smem_u32(smem_a + sa * Cfg::kASlotBytes);
                const bool patch = g < n_patch;
                const int nb = patch ? (roll ? 3 : 9) : 1;
                // start offset of tap 0 and its step per tap / per row of taps inside the shared patch
                const uint32_t sbo = (patch && !roll) ? kHaloW * 128u : 1024u;
                for (int tap = 0; tap < nb; ++tap, ++gb) {
                    const int sb = gb % Cfg::kBSlots;
                    ptx::mbar_wait(&fullB[sb], (gb / Cfg::kBSlots) & 1);
                    ptx::tc_fence_after();
                    if (acc == 0u && static_cast<unsigned>(lt - A.tr.lt0) < 3u && lane == 0) trace_mark(A.tr, 4 + 6 * (lt - A.tr.lt0) + 0);
                    const uint32_t b_base = ptx::smem_u32(smem_b + sb * Cfg::kBSlotBytes);
                    uint32_t off = 0;
                    if (patch && !roll) {
                        const int kh = tap / 3, kw = tap - kh * 3;
                        off = static_cast<uint32_t>(kh * kHaloW + kw) * 128u;
                    } else if (patch) {
                        off = static_cast<uint32_t>(tap) * 128u;     // 1-D tap: rows tap .. tap+127 of the 130-row patch
                    }
                    const uint64_t a_hi = ptx::make_sw128_desc(a_base + off, sbo, 0);
                    const uint64_t a_lo = ptx::make_sw128_desc(a_base + (roll ? Cfg::kRollLoOff : kALo) + off, sbo, 0);
                    const bool use_alo = kAlo || (Cfg::kRollAlo && roll);
                    const uint64_t b_hi = ptx::make_sw128_desc(b_base, 1024u, 0);
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k) {
                            const uint64_t ko = static_cast<uint64_t>((k * 32) >> 4);   // +32 B along K inside the swizzle atom
                            if (kBlo) {
                                // [D1 | D2] (+)= Ah * [Bh | Bl]  (one N=128 MMA: the lo weight tile sits right behind the hi
                                // tile in shared memory and D2 right behind D1 in TMEM), then D2 += Al * Bh
                                ptx::umma_f16(d1, a_hi + ko, b_hi + ko, idesc2, k == 0 ? acc : 1u);
                                if (use_alo) ptx::umma_f16(d2, a_lo + ko, b_hi + ko, idesc, 1u);
                            } else {
                                ptx::umma_f16(d1, a_hi + ko, b_hi + ko, idesc, k == 0 ? acc : 1u);
                                if (kAlo) ptx::umma_f16(d2, a_lo + ko, b_hi + ko, idesc, k == 0 ? acc : 1u);     // D2 (+)= Al * Bh
                            }
                        }
                        ptx::umma_commit(&emptyB[sb]);    // weight slot free once the MMAs above retire
                    }
                    __syncwarp();
                    acc = 1u;
                }
                if (ptx::elect_one()) ptx::umma_commit(&emptyA[sa]);        // A patch free once every tap has read it
                __syncwarp();
            }
            if (ptx::elect_one()) ptx::umma_commit(&tmem_full_bar[as]);     // this tile's accumulators are complete
            __syncwarp();
            if (static_cast<unsigned>(lt - A.tr.lt0) < 3u && lane == 0) trace_mark(A.tr, 4 + 6 * (lt - A.tr.lt0) + 1);
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;             // which 32 of the tile's 64 output channels: the two warps of a quarter split them
        const int m = quarter * 32 + lane;            // accumulator row (pixel of the tile / position of the roll tile)
        const int et = threadIdx.x - 64;              // 0..255 among the epilogue threads
        const int k8 = lane & 7, rloc = lane >> 3;    // transposed pass: 16-byte chunk k8 of staged rows rloc, rloc + 4, ...
        float* wst = stage + (warp - 2) * 1024;       // this warp's staging block
        bool roll_ready = F.n_roll == 0;
        int lt = 0, ga = 0;
        if (et == 0) trace_mark(A.tr, 30);
        pdl_wait();
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
            const int as = lt & 1;
            const uint32_t aph = (lt >> 1) & 1;
            const uint32_t lane_addr = tmem_base + as * Cfg::kAccCols + (static_cast<uint32_t>(quarter * 32) << 16);
            if (t < F.n_roll) {
                // ---------- rollout 1-D GEMM tile ----------
                const RollTile T = roll_tile_decode(F, t);
                const int L = F.R.L[T.src];
                // (1) A operand: axis sums -> (hi, lo) fp16 means, written swizzled into the A slots (one per 64-channel block)
                {
                    const unsigned long long* sp = F.sums + (static_cast<size_t>(T.b) * F.total_len + F.R.soff[T.src]) * A.C;
                    const float scale = F.R.scale[T.src];
                    constexpr int kChunksMax = kRollRows * 8;               // 16-byte chunks (8 channels) of one patch
                    const int kChunks = (F.R.tm[T.src] + 2) * 8;
                    constexpr int kIters = (kChunksMax + kEpiThreads - 1) / kEpiThreads, kBatch = kIters;     // every load of a patch in flight at once
                    for (int cb = 0; cb < cblks; ++cb, ++ga) {
                        const int s = ga % Cfg::kASlots;
                        if (et == 0 && lt == 0 && cb == 0) trace_mark(A.tr, 2);
                        ptx::mbar_wait(&emptyA[s], ((ga / Cfg::kASlots) & 1) ^ 1);
                        if (et == 0 && lt == 0 && cb == 0) trace_mark(A.tr, 16);
                        uint8_t* hi = smem_a + s * Cfg::kASlotBytes;
                        uint8_t* lo = hi + Cfg::kRollLoOff;
#pragma unroll
                        for (int hb = 0; hb < 1; ++hb) {
                            ulonglong2 raw[kBatch][4];
#pragma unroll
                            for (int i = 0; i < kBatch; ++i) {
                                const int q = et + kEpiThreads * (hb * kBatch + i), j = q >> 3, k = q & 7, pos = T.p0 - 1 + j;
                                const bool ok = q < kChunks && pos >= 0 && pos < L;
                                const ulonglong2* src = reinterpret_cast<const ulonglong2*>(sp + static_cast<size_t>(ok ? pos : 0) * A.C + cb * kBK + k * 8);
#pragma unroll
                                for (int e = 0; e < 4; ++e) raw[i][e] = ok ? __ldcg(src + e) : make_ulonglong2(0ull, 0ull);
                            }
                            if (et == 0 && lt == 0 && cb == 0) trace_mark(A.tr, 17);
#pragma unroll
                            for (int i = 0; i < kBatch; ++i) {
                                const int q = et + kEpiThreads * (hb * kBatch + i), j = q >> 3, k = q & 7;
                                if (q < kChunks) {
                                    uint32_t h[4], l[4];
#pragma unroll
                                    for (int e = 0; e < 4; ++e)
                                        split_f16x2(__ll2float_rn(static_cast<long long>(raw[i][e].x)) * scale,
                                                    __ll2float_rn(static_cast<long long>(raw[i][e].y)) * scale, h[e], l[e]);
                                    const uint32_t off = static_cast<uint32_t>(j) * 128u + (static_cast<uint32_t>(k ^ (j & 7)) << 4);
                                    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                                    if (Cfg::kRollAlo) *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
                                }
                            }
                        }
                        if (et == 0 && lt == 0 && cb == 0) trace_mark(A.tr, 18);
                        ptx::fence_proxy_async();                           // generic-proxy writes -> visible to the tensor core
                        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
                        if (et == 0) ptx::mbar_arrive(&fullA[s]);
                    }
                    if (et == 0 && t + static_cast<int>(gridDim.x) >= F.n_roll) ptx::mbar_arrive(roll_filled_bar);
                    if (et == 0 && lt == 0) trace_mark(A.tr, 3);
                }
                // (2) T[b][cls][pos][co] = accumulator, through the warp's staging block so that the global stores are coalesced
                const int cls = T.n0 / A.Cout, co0 = T.n0 - cls * A.Cout;
                float* __restrict__ outp = F.R.T[T.src] + ((static_cast<size_t>(T.b) * 4 + cls) * L + T.p0 + quarter * 32) * A.Cout + co0;
                ptx::mbar_wait(&tmem_full_bar[as], aph);
                if (et == 0 && static_cast<unsigned>(lt - A.tr.lt0) < 3u) trace_mark(A.tr, 4 + 6 * (lt - A.tr.lt0) + 3);
                __syncwarp();
                ptx::tc_fence_after();
                {
                    uint32_t v1[32], v2[32];
                    ptx::tmem_ld_32x32b_x32(lane_addr + half * 32, v1);
                    if (kTwo) ptx::tmem_ld_32x32b_x32(lane_addr + kBN + half * 32, v2);
                    ptx::tmem_ld_wait();
                    {
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[as]);
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float o[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            o[q] = __uint_as_float(v1[j + q]);
                            if (kTwo) o[q] = fmaf(__uint_as_float(v2[j + q]), 1.f / kLoScale, o[q]);
                        }
                        *reinterpret_cast<float4*>(wst + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int ml = it * 4 + rloc;
                        if (quarter * 32 + ml < F.R.tm[T.src] && T.p0 + quarter * 32 + ml < L)
                            __stcg(reinterpret_cast<float4*>(outp + static_cast<size_t>(ml) * A.Cout + half * 32) + k8,
                                   *reinterpret_cast<const float4*>(wst + ml * 32 + ((k8 ^ (ml & 7)) << 2)));
                    }
                    __syncwarp();
                }
                if (et == 0 && lt == 0) trace_mark(A.tr, 19);
                if (et == 0 && lt == 0) trace_mark(A.tr, 20);
                // publish: CTA barrier, then ONE release-scoped reduction (cumulative over the stores the barrier ordered before
                // it) instead of a device-wide fence in each of the 128 threads
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
                if (et == 0) {
                    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(F.counters), "r"(1u) : "memory");
                    if (lt == 0) trace_mark(A.tr, 21);
                }
                if (et == 0 && static_cast<unsigned>(lt - A.tr.lt0) < 3u) trace_mark(A.tr, 4 + 6 * (lt - A.tr.lt0) + 4);
                continue;
            }
            // ---------- conv tile ----------
            ga += cblks + nskip;                       // A groups this tile consumes (keeps the ring index of the roll fills in step)
            const ConvTile T = conv_tile_decode(A, t - F.n_roll);
            const int plane = T.plane, n0 = T.n0, b = T.b;
            const int rows = A.d.rows[plane], cols = A.d.cols[plane];
            const size_t px0 = static_cast<size_t>(b) * rows * cols;
            const int Cout = A.Cout;
            const float* __restrict__ resid = A.e.resid.p[plane];
            const float* __restrict__ Trow = A.e.Trow.p[plane];
            const float* __restrict__ Tcol = A.e.Tcol.p[plane];
            float* __restrict__ outb = A.e.out.p[plane];
            if (Trow && !roll_ready) {
                // the rollout terms are produced by this very launch (roll tiles): wait until all of them are written
                if (lane == 0) {
                    long long spins = 0;
                    for (;;) {
                        unsigned int v;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(F.counters) : "memory");
                        if (v >= static_cast<unsigned int>(F.n_roll)) break;
                        if (++spins > (1LL << 31)) __trap();     // never hang the device
                    }
                }
                __syncwarp();
                roll_ready = true;
                if (et == 0) trace_mark(A.tr, 23);
            }
            // Geometry of the transposed pass: lane (rloc = lane / 8, k8 = lane % 8) reads 16-byte chunk k8 of staged row
            // ml = it*4 + rloc, it = 0..7, i.e. pixel (r0 + it/2, w0 + (it & 1)*4 + rloc): four image rows and two columns per
            // thread.  What is added to the accumulator of one 32-channel half — residual, row-indexed and column-indexed
            // rollout terms — is requested before the accumulator is awaited.
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const int r0 = T.h0 + quarter * 4;
            const int cx[2] = {T.w0 + rloc, T.w0 + 4 + rloc};
            // Requests are issued as a block (`fetch`), folded into one addend per pixel (`fold`) only when they are needed: the
            // first half's requests fly while the MMAs finish, the second half's while the first half is stored.
            struct Addends {
                float4 rs[8], trow[2][4], tcol[2][4], bias;
            };
            auto fetch = [&](int hf, Addends& D) {
                const int ch = n0 + hf * 32;
                // per-channel addends of this thread's quad: bias (+ additive timestep embedding)
                D.bias = __ldg(reinterpret_cast<const float4*>(A.e.bias.p[plane] + ch) + k8);
                if (A.e.embadd) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(
                        A.e.embadd + static_cast<size_t>(A.e.film_row ? A.e.film_row[b] : b) * A.e.film_dim + A.e.film_off + ch) + k8);
                    D.bias.x += v.x; D.bias.y += v.y; D.bias.z += v.z; D.bias.w += v.w;
                }
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = r0 + (it >> 1), c = cx[it & 1];
                    D.rs[it] = zero4;
                    if (resid && r < rows && c < cols)
                        D.rs[it] = __ldg(reinterpret_cast<const float4*>(resid + (px0 + static_cast<size_t>(r) * cols + c) * Cout + ch) + k8);
                }
#pragma unroll
                for (int x = 0; x < 2; ++x)
#pragma unroll
                    for (int j = 0; j < 4; ++j) D.trow[x][j] = D.tcol[x][j] = zero4;
                if (Trow) {
                    const size_t bo = static_cast<size_t>(b) * 4;
                    // written by this launch: L2-coherent loads.  Equal edge classes share one load.
                    const int cc0 = edge_class(min(cx[0], cols - 1), cols), cc1 = edge_class(min(cx[1], cols - 1), cols);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = min(r0 + j, rows - 1);
                        D.trow[0][j] = __ldcg(reinterpret_cast<const float4*>(Trow + ((bo + cc0) * rows + r) * Cout + ch) + k8);
                        D.trow[1][j] = cc1 == cc0 ? D.trow[0][j]
                                                  : __ldcg(reinterpret_cast<const float4*>(Trow + ((bo + cc1) * rows + r) * Cout + ch) + k8);
                    }
#pragma unroll
                    for (int x = 0; x < 2; ++x) {
                        const int c = min(cx[x], cols - 1);
                        int prev = -1;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int rc = edge_class(min(r0 + j, rows - 1), rows);
                            if (j > 0 && rc == prev) D.tcol[x][j] = D.tcol[x][j - 1];
                            else D.tcol[x][j] = __ldcg(reinterpret_cast<const float4*>(Tcol + ((bo + rc) * cols + c) * Cout + ch) + k8);
                            prev = rc;
                        }
                    }
                }
            };
            auto fold = [&](const Addends& D, float4 (&add)[8]) {
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const float4 tr = D.trow[it & 1][it >> 1], tc = D.tcol[it & 1][it >> 1], rs = D.rs[it];
                    add[it].x = D.bias.x + rs.x + (tr.x + tc.x);
                    add[it].y = D.bias.y + rs.y + (tr.y + tc.y);
                    add[it].z = D.bias.z + rs.z + (tr.z + tc.z);
                    add[it].w = D.bias.w + rs.w + (tr.w + tc.w);
                }
            };
            Addends D;
            float4 add[8];
            fetch(half, D);
            if (et == 0 && static_cast<unsigned>(lt - A.tr.lt0) < 3u) trace_mark(A.tr, 4 + 6 * (lt - A.tr.lt0) + 2);
            ptx::mbar_wait(&tmem_full_bar[as], aph);
            if (et == 0 && static_cast<unsigned>(lt - A.tr.lt0) < 3u) trace_mark(A.tr, 4 + 6 * (lt - A.tr.lt0) + 3);
            __syncwarp();
            ptx::tc_fence_after();
            fold(D, add);
            const bool do_stats = A.sink.acc != nullptr;
            float4 ssum = zero4, ssq = zero4;
            float* __restrict__ poolb = A.pool.p[plane];
            float4 pv[2][2] = {{zero4, zero4}, {zero4, zero4}};      // [row pair][column half]: vertical sums of this thread's pixels
            {
                uint32_t v1[32], v2[32];
                ptx::tmem_ld_32x32b_x32(lane_addr + half * 32, v1);
                if (kTwo) ptx::tmem_ld_32x32b_x32(lane_addr + kBN + half * 32, v2);
                ptx::tmem_ld_wait();
                {
                    // this warp's TMEM reads of the tile are done: hand the accumulator stage back to the MMA issuer (8 arrivals)
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[as]);
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float o[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        o[q] = __uint_as_float(v1[j + q]);
                        if (kTwo) o[q] = fmaf(__uint_as_float(v2[j + q]), 1.f / kLoScale, o[q]);
                    }
                    *reinterpret_cast<float4*>(wst + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) = make_float4(o[0], o[1], o[2], o[3]);
                }
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int ml = it * 4 + rloc, r = r0 + (it >> 1), c = cx[it & 1];
                    if (r < rows && c < cols) {
                        const float4 a = *reinterpret_cast<const float4*>(wst + ml * 32 + ((k8 ^ (ml & 7)) << 2));
                        float4 o;
                        o.x = a.x + add[it].x;
                        o.y = a.y + add[it].y;
                        o.z = a.z + add[it].z;
                        o.w = a.w + add[it].w;
                        *reinterpret_cast<float4*>(outb + (px0 + static_cast<size_t>(r) * cols + c) * Cout + n0 + half * 32 + 4 * k8) = o;
                        ssum.x += o.x; ssum.y += o.y; ssum.z += o.z; ssum.w += o.w;
                        ssq.x = fmaf(o.x, o.x, ssq.x); ssq.y = fmaf(o.y, o.y, ssq.y);
                        ssq.z = fmaf(o.z, o.z, ssq.z); ssq.w = fmaf(o.w, o.w, ssq.w);
                        if (poolb) {
                            float4& a2 = pv[it >> 2][it & 1];
                            a2.x += o.x; a2.y += o.y; a2.z += o.z; a2.w += o.w;
                        }
                    }
                }
                __syncwarp();                      // the staging block is rewritten by the next tile
            }
            // (sum, sum-sq) per channel over the warp's pixels: the four row groups of a lane column are added in a fixed order, then
            // every GroupNorm group goes out as one fixed-point atomic (exact, so the order of the warps and tiles does not matter)
            auto emit_stats = [&](const StatsSink& sink, const float4& su, const float4& sq) {
                constexpr unsigned kFull = 0xffffffffu;
                const int cpg = Cout / kGroups;              // 2, 4 or 8 (host-checked)
                float v[8] = {su.x, su.y, su.z, su.w, sq.x, sq.y, sq.z, sq.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    v[e] += __shfl_xor_sync(kFull, v[e], 8);
                    v[e] += __shfl_xor_sync(kFull, v[e], 16);
                }
                // lanes 0..7 now hold channels n0 + half*32 + 4*k8 .. +3
                double gs[2][2];                          // [group inside the quad][sum / sum-sq]
                if (cpg == 2) {
                    gs[0][0] = static_cast<double>(v[0]) + static_cast<double>(v[1]);
                    gs[1][0] = static_cast<double>(v[2]) + static_cast<double>(v[3]);
                    gs[0][1] = static_cast<double>(v[4]) + static_cast<double>(v[5]);
                    gs[1][1] = static_cast<double>(v[6]) + static_cast<double>(v[7]);
                } else {
                    gs[0][0] = (static_cast<double>(v[0]) + static_cast<double>(v[1])) + (static_cast<double>(v[2]) + static_cast<double>(v[3]));
                    gs[0][1] = (static_cast<double>(v[4]) + static_cast<double>(v[5])) + (static_cast<double>(v[6]) + static_cast<double>(v[7]));
                    gs[1][0] = gs[1][1] = 0.0;
                    if (cpg == 8) {                       // a group spans two neighbouring lanes
                        gs[0][0] += __shfl_xor_sync(kFull, gs[0][0], 1);
                        gs[0][1] += __shfl_xor_sync(kFull, gs[0][1], 1);
                    }
                }
                if (lane < 8) {
                    const int ch = n0 + half * 32 + 4 * k8;
                    unsigned long long* acc = gn_acc(sink.acc, b, plane, T.ip * 4 + quarter);
                    if (cpg == 2) {
                        gn_fix_add(acc + (ch / 2) * 2, gs[0][0]);
                        gn_fix_add(acc + (ch / 2) * 2 + 1, gs[0][1]);
                        gn_fix_add(acc + (ch / 2 + 1) * 2, gs[1][0]);
                        gn_fix_add(acc + (ch / 2 + 1) * 2 + 1, gs[1][1]);
                    } else if (cpg == 4 || (lane & 1) == 0) {
                        gn_fix_add(acc + (ch / cpg) * 2, gs[0][0]);
                        gn_fix_add(acc + (ch / cpg) * 2 + 1, gs[0][1]);
                    }
                }
            };
            if (do_stats) emit_stats(A.sink, ssum, ssq);
            if (poolb) {
                // 2x2 average: the vertical pair was added above, the horizontal neighbour (column +1) lives 8 lanes up
                const int prow_n = rows >> 1, pcol_n = cols >> 1;
                float4 psum = zero4, psq = zero4;
#pragma unroll
                for (int pr = 0; pr < 2; ++pr)
#pragma unroll
                    for (int hc = 0; hc < 2; ++hc) {
                        float4 q = pv[pr][hc];
                        q.x += __shfl_xor_sync(0xffffffffu, q.x, 8);
                        q.y += __shfl_xor_sync(0xffffffffu, q.y, 8);
                        q.z += __shfl_xor_sync(0xffffffffu, q.z, 8);
                        q.w += __shfl_xor_sync(0xffffffffu, q.w, 8);
                        const int prw = (r0 >> 1) + pr, pcl = (T.w0 >> 1) + hc * 2 + (rloc >> 1);
                        if ((rloc & 1) == 0 && prw < prow_n && pcl < pcol_n) {
                            q.x *= 0.25f; q.y *= 0.25f; q.z *= 0.25f; q.w *= 0.25f;
                            *reinterpret_cast<float4*>(poolb + ((static_cast<size_t>(b) * prow_n + prw) * pcol_n + pcl) * Cout + n0 + half * 32 + 4 * k8) = q;
                            psum.x += q.x; psum.y += q.y; psum.z += q.z; psum.w += q.w;
                            psq.x = fmaf(q.x, q.x, psq.x); psq.y = fmaf(q.y, q.y, psq.y);
                            psq.z = fmaf(q.z, q.z, psq.z); psq.w = fmaf(q.w, q.w, psq.w);
                        }
                    }
                if (A.pool_sink.acc) emit_stats(A.pool_sink, psum, psq);
            }
            if (et == 0 && static_cast<unsigned>(lt - A.tr.lt0) < 3u) trace_mark(A.tr, 4 + 6 * (lt - A.tr.lt0) + 4);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
    if (threadIdx.x == 0) trace_mark(A.tr, 22);
    if (F.n_roll && threadIdx.x == 0) {
        // the last CTA to leave re-arms the counters for the next launch
        const unsigned int prev = atomicAdd(&F.counters[1], 1u);
        if (prev == gridDim.x - 1) {
            F.counters[0] = 0u;
            F.counters[1] = 0u;
        }
    }
}

// =====================================================================================
// Rollout 1-D terms on the tensor cores, stand-alone launch (S3D_FUSE_ROLL=0; the default fuses these tiles into k_conv_tc).
//   T[b][cls][pos][co] = sum_{along, c} mean[pos+along-1][c] * wc[cls*Cout + co][along*C + c]
// Same pipeline as k_conv_tc with a 1 x 128 "patch": M tile = 128 consecutive positions of one source's mean
// vector (5-D TMA box {64 ch, 128, 1, 1, 1}; the +-1 tap shift and both ends are the TMA zero fill), 3 taps,
// N tile = 64 of the 4*Cout class-summed output columns.  fp16 (hi, lo) means come from k_gn_silu's tail.
// grid (sum over the 6 sources of ceil(L/128), 4*Cout/64, B)
// =====================================================================================
template <int NSPLIT>
__global__ void __launch_bounds__(kRollThreads, 1) k_roll_tc(const __grid_constant__ RollTcMaps M, const RollTcArgs A) {
    pdl_wait();
    pdl_trigger();
    using Cfg = ConvTcCfg<NSPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + Cfg::kStages;
    uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int src = 0;
#pragma unroll
    for (int k = 1; k < 6; ++k)
        if (static_cast<int>(blockIdx.x) >= A.tile_start[k]) src = k;
    const int p0 = (blockIdx.x - A.tile_start[src]) * A.tm[src];
    const int n0 = blockIdx.y * kBN;
    const int b = blockIdx.z;
    if (A.T[src] == nullptr || n0 >= A.ncls[src] * A.Cout) return;      // uniform for the whole CTA
    const int cblks = A.C / kBK;
    const int nk = 3 * cblks;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&M.a[src]);
        ptx::prefetch_tmap(&M.w[src]);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < Cfg::kStages; ++s) {
                ptx::mbar_init(&full_bar[s], 1);
                ptx::mbar_init(&empty_bar[s], 1);
            }
            ptx::mbar_init(tmem_full_bar, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc<Cfg::kAccCols>(tmem_ptr_smem);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nk; ++i) {
                const int s = i % Cfg::kStages;
                const uint32_t ph = (i / Cfg::kStages) & 1;
                ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = smem + s * Cfg::kStageBytes;
                ptx::mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
                const int al = i / cblks, cb = i - al * cblks;
                // stage layout: [A hi][A lo][B hi][B lo]  (B lo directly behind B hi: one N=128 MMA reads both)
                constexpr int kBOff = (NSPLIT == 3 ? 2 : 1) * kABytes;
                ptx::tma_load_5d(st, &M.a[src], &full_bar[s], cb * kBK, p0 + al - 1, 0, b, 0);
                ptx::tma_load_3d(st + kBOff, &M.w[src], &full_bar[s], i * kBK, n0, 0);     // box {64, 64, hi|lo}: both weight tiles
                if (NSPLIT == 3) ptx::tma_load_5d(st + kABytes, &M.a[src], &full_bar[s], cb * kBK, p0 + al - 1, 0, b, 1);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(kBM, kBN);
            constexpr uint32_t idesc2 = ptx::make_idesc_f16(kBM, 2 * kBN);
            const uint32_t d1 = tmem_base, d2 = tmem_base + kBN;
            for (int i = 0; i < nk; ++i) {
                const int s = i % Cfg::kStages;
                const uint32_t ph = (i / Cfg::kStages) & 1;
                ptx::mbar_wait(&full_bar[s], ph);
                ptx::tc_fence_after();
                const uint32_t st = ptx::smem_u32(smem + s * Cfg::kStageBytes);
                constexpr uint32_t kBOff = (NSPLIT == 3 ? 2 : 1) * kABytes;
                const uint64_t a_hi = ptx::make_sw128_desc1024(st);
                const uint64_t a_lo = ptx::make_sw128_desc1024(st + kABytes);
                const uint64_t b_hi = ptx::make_sw128_desc1024(st + kBOff);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                    const uint64_t ko = static_cast<uint64_t>((k * 32) >> 4);
                    const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
                    if (NSPLIT == 3) {
                        ptx::umma_f16(d1, a_hi + ko, b_hi + ko, idesc2, acc);     // [D1 | D2] (+)= Ah * [Bh | Bl]
                        ptx::umma_f16(d2, a_lo + ko, b_hi + ko, idesc, 1u);       // D2 += Al * Bh
                    } else {
                        ptx::umma_f16(d1, a_hi + ko, b_hi + ko, idesc, acc);
                    }
                }
                ptx::umma_commit(&empty_bar[s]);
            }
            ptx::umma_commit(tmem_full_bar);
        }
    } else {
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const int pos = p0 + m, L = A.L[src];
        const bool valid = pos < L && m < A.tm[src];
        const int cls = n0 / A.Cout, co0 = n0 - cls * A.Cout;
        float* __restrict__ outp = A.T[src] + ((static_cast<size_t>(b) * 4 + cls) * L + pos) * A.Cout + co0;
        ptx::mbar_wait(tmem_full_bar, 0);
        __syncwarp();
        ptx::tc_fence_after();
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t v1[32], v2[32];
            ptx::tmem_ld_32x32b_x32(lane_addr + half * 32, v1);
            if (NSPLIT == 3) ptx::tmem_ld_32x32b_x32(lane_addr + kBN + half * 32, v2);
            ptx::tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float o[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float acc = __uint_as_float(v1[j + q]);
                        if (NSPLIT == 3) acc = fmaf(__uint_as_float(v2[j + q]), 1.f / kLoScale, acc);
                        o[q] = acc;
                    }
                    *reinterpret_cast<float4*>(outp + half * 32 + j) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc<Cfg::kAccCols>(tmem_base);
    }
}

}  // namespace s3d
