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
constexpr int kHaloH = kTileH + 2, kHaloW = 16;   // halo patch rows / (padded) cols: row pitch 16 px = 2048 B keeps every
                                                   // 8-pixel row group on the same 128B-swizzle phase
constexpr int kBM = 128, kBN = 64, kBK = 64;
constexpr int kABytes = kBM * kBK * 2;             // 16 KiB: plain 128-row A tile (skip / rollout groups)
constexpr int kAHaloBytes = kHaloH * kHaloW * kBK * 2;   // 36 KiB: halo patch of one 64-channel block
constexpr int kBBytes = kBN * kBK * 2;             //  8 KiB
constexpr int kConvThreads = 224;                  // warps: 0 A-producer, 1 MMA, 2..5 epilogue, 6 B-producer
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
    Trace tr;            // opt-in phase stamps (common.cuh)
    int bo_kw;           // bring-up switch (S3D_HALO_BO_KW=1): put kw into the descriptor base_offset (measured WRONG on B200:
                         // the tensor core derives the swizzle phase from the absolute shared-memory address)
};

struct RollTcMaps {
    CUtensorMap a[6];   // means of source s: (C, L, 1, B, 2) fp16, box {64, 128, 1}
    CUtensorMap w[6];   // class-summed 1-D weights: (3C, 4*Cout, 2) fp16, K = along*C + c
};
struct RollTcArgs {
    int L[6], ncls[6];
    float* T[6];        // [B][4][L][Cout]
    int tile_start[7];
    int C, Cout;
};
// Fused launch: the rollout 1-D GEMM tiles ("roll tiles") are the first tile indices of the persistent conv kernel; the
// conv tiles' epilogues wait on a device counter until every roll tile has been written (roll tiles never wait on
// anything and are the first tiles of the lowest-numbered CTAs, so the wait cannot deadlock).
struct FusedRoll {
    RollTcArgs R;
    int n_roll;              // number of roll tiles (0: none; Trow/Tcol come from a separate launch or are absent)
    int ntn;                 // N tiles per roll M tile
    unsigned int* counters;  // [3]: roll tiles done, CTAs exited, CTAs whose share of the means is converted
                             //      (all return to 0 when the kernel ends)
    // Phase 0 (every CTA, epilogue warps): axis sums (64-bit fixed point, accumulated by k_gn_silu) -> fp16 (hi, lo) means,
    // the A operand the roll tiles then fetch by TMA; the accumulators are re-zeroed on the way.
    unsigned long long* sums;     // [B][total_len][C]   nullptr: means16 already finalised by k_gn_silu
    __half* means16;              // [2][B][total_len][C]
    int total_len, B;
    int seg_end[6];               // cumulative segment ends (positions) in a sample's block
    float seg_scale[6];           // 2^-24 / (length of the averaged axis)
};

template <int NSPLIT>
struct ConvTcCfg {
    static constexpr int kASlotBytes = (NSPLIT == 3 ? 2 : 1) * kAHaloBytes;   // hi at +0, lo at +kAHaloBytes
    static constexpr int kBSlotBytes = (NSPLIT == 3 ? 2 : 1) * kBBytes;       // hi at +0, lo at +kBBytes
    static constexpr int kASlots = 2;
    static constexpr int kBSlots = NSPLIT == 3 ? 3 : 6;
    static constexpr int kRingBytes = kASlots * kASlotBytes + kBSlots * kBSlotBytes;
    static constexpr int kAccCols = NSPLIT == 3 ? 128 : 64;     // TMEM columns of one accumulator stage
    static constexpr int kTmemCols = 2 * kAccCols;              // double-buffered accumulators
    static constexpr int kStatBytes = kBM * 33 * 4 + 4 * 2 * 32 * 4 + 2 * kBN * 4 + 64 * 8 * 8 + 16;   // epilogue statistics scratch
    static constexpr int kSmemBytes = kRingBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kStatBytes;
    // stand-alone k_roll_tc keeps the simple 4-stage {A,B} ring
    static constexpr int kStageBytes = (NSPLIT == 3 ? 2 : 1) * (kABytes + kBBytes);
    static constexpr int kStages = NSPLIT == 3 ? 4 : 8;
    static constexpr int kRollSmemBytes = kStages * kStageBytes + 1024 + 256;
};

struct ConvTile {
    int ip;    // tile index inside its plane (statistics slot)
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
    T.p0 = (rem - F.R.tile_start[src]) * kBM;
    return T;
}

// Persistent implicit-GEMM convolution, one CTA per SM, tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...
//
// Operand traffic is what bounds this kernel (64 B/clk per SM from L2), so the A operand is fetched ONCE per 64-channel
// block as an 18 x 16-pixel halo patch and all nine taps read it in place: tap (kh, kw) is the same shared-memory patch
// addressed from row (kh*16 + kw) with an 8-row-group stride of 2048 B (UMMA descriptor start / SBO / base_offset), i.e.
// 1/6 of the activation bytes of a tap-by-tap im2col.  Only the 8 KiB weight tile changes per tap.
// Two independent TMA producers (A patches, B tiles), one MMA issuer, four epilogue warps; accumulators double-buffered
// in TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
template <int NSPLIT>
__global__ void __launch_bounds__(kConvThreads, 1) k_conv_tc(const __grid_constant__ ConvTcMaps M,
                                                             const __grid_constant__ RollTcMaps RM, const ConvTcArgs A,
                                                             const FusedRoll F, const int total_tiles) {
    using Cfg = ConvTcCfg<NSPLIT>;
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
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    // epilogue statistics scratch (behind the 256-byte barrier block)
    double* stat_fin = reinterpret_cast<double*>(smem + Cfg::kRingBytes + 256);                     // [64*8] (spare)
    float* stat_stage = reinterpret_cast<float*>(stat_fin + 64 * 8);                               // [128][33]
    float* stat_colp = stat_stage + kBM * 33;                                                      // [4][2][32]
    float* stat_tot = stat_colp + 4 * 2 * 32;                                                      // [2][64]
    int* stat_flag = reinterpret_cast<int*>(stat_tot + 2 * kBN);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // trace slots: 0 entry, 1 set-up done, 2 phase 0 done, 3 means visible to the A producer, 22 exit; per local tile lt < 3 at
    // 4 + 6*lt: +0 first operands landed, +1 all MMAs issued, +2 epilogue addends gathered, +3 accumulator complete,
    // +4 epilogue done, +5 A loads issued
    if (threadIdx.x == 0) trace_mark(A.tr, 0);
    const int cblks = A.C / kBK;
    const int nskip = A.Cs / kBK;
    constexpr uint32_t kALo = kAHaloBytes, kBLo = kBBytes;
    constexpr uint32_t kAStdTx = (NSPLIT == 3 ? 2 : 1) * kABytes, kAHaloTx = (NSPLIT == 3 ? 2 : 1) * kAHaloBytes;

    if (warp == 0 && lane == 0) {
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            ptx::prefetch_tmap(&M.a[p]);
            ptx::prefetch_tmap(&M.w[p]);
            if (A.Cs) ptx::prefetch_tmap(&M.x[p]);
        }
        if (F.n_roll) {
#pragma unroll
            for (int p = 0; p < 6; ++p) {
                ptx::prefetch_tmap(&RM.a[p]);
                ptx::prefetch_tmap(&RM.w[p]);
            }
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
                ptx::mbar_init(&tmem_empty_bar[s], 4);       // one arrive per epilogue warp
            }
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
        // ===================== TMA producer: A operand groups =====================
        if (lane == 0) {
            pdl_wait();
            int ga = 0, ltp = 0;
            bool means_ready = false;
            auto slot_wait = [&](uint32_t tx) -> uint8_t* {
                const int s = ga % Cfg::kASlots;
                ptx::mbar_wait(&emptyA[s], ((ga / Cfg::kASlots) & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(&fullA[s], tx);
                return smem_a + s * Cfg::kASlotBytes;
            };
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ltp) {
                if (t < F.n_roll) {
                    if (F.sums && !means_ready) {
                        // every CTA converts a share of the means in its phase 0: wait for all of them, then order the
                        // generic-proxy observation before the async-proxy (TMA) reads
                        const volatile unsigned int* cnt = F.counters + 2;
                        long long spins = 0;
                        while (*cnt < gridDim.x) {
                            if (++spins > (1LL << 31)) __trap();     // never hang the device
                        }
                        __threadfence();
                        asm volatile("fence.proxy.async;" ::: "memory");
                        means_ready = true;
                        trace_mark(A.tr, 3);
                    }
                    const RollTile T = roll_tile_decode(F, t);
                    for (int i = 0; i < 3 * cblks; ++i, ++ga) {
                        uint8_t* st = slot_wait(kAStdTx);
                        const int s = ga % Cfg::kASlots;
                        const int al = i / cblks, cb = i - al * cblks;
                        ptx::tma_load_5d(st, &RM.a[T.src], &fullA[s], cb * kBK, T.p0 + al - 1, 0, T.b, 0);
                        if (NSPLIT == 3) ptx::tma_load_5d(st + kALo, &RM.a[T.src], &fullA[s], cb * kBK, T.p0 + al - 1, 0, T.b, 1);
                    }
                    if (ltp < 3) trace_mark(A.tr, 4 + 6 * ltp + 5);
                    continue;
                }
                const ConvTile T = conv_tile_decode(A, t - F.n_roll);
                for (int cb = 0; cb < cblks; ++cb, ++ga) {
                    uint8_t* st = slot_wait(kAHaloTx);
                    const int s = ga % Cfg::kASlots;
                    ptx::tma_load_5d(st, &M.a[T.plane], &fullA[s], cb * kBK, T.w0 - 1, T.h0 - 1, T.b, 0);
                    if (NSPLIT == 3) ptx::tma_load_5d(st + kALo, &M.a[T.plane], &fullA[s], cb * kBK, T.w0 - 1, T.h0 - 1, T.b, 1);
                }
                for (int j = 0; j < nskip; ++j, ++ga) {
                    uint8_t* st = slot_wait(kAStdTx);
                    const int s = ga % Cfg::kASlots;
                    ptx::tma_load_5d(st, &M.x[T.plane], &fullA[s], j * kBK, T.w0, T.h0, T.b, 0);
                    if (NSPLIT == 3) ptx::tma_load_5d(st + kALo, &M.x[T.plane], &fullA[s], j * kBK, T.w0, T.h0, T.b, 1);
                }
                if (ltp < 3) trace_mark(A.tr, 4 + 6 * ltp + 5);
            }
        }
    } else if (warp == 6) {
        // ===================== TMA producer: B (weight) tiles =====================
        if (lane == 0) {
            int gb = 0;
            auto load_b = [&](const CUtensorMap* wm, int kchunk, int n0) {
                const int s = gb % Cfg::kBSlots;
                ptx::mbar_wait(&emptyB[s], ((gb / Cfg::kBSlots) & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(&fullB[s], Cfg::kBSlotBytes);
                uint8_t* st = smem_b + s * Cfg::kBSlotBytes;
                ptx::tma_load_3d(st, wm, &fullB[s], kchunk * kBK, n0, 0);
                if (NSPLIT == 3) ptx::tma_load_3d(st + kBLo, wm, &fullB[s], kchunk * kBK, n0, 1);
                ++gb;
            };
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                if (t < F.n_roll) {
                    const RollTile T = roll_tile_decode(F, t);
                    for (int i = 0; i < 3 * cblks; ++i) load_b(&RM.w[T.src], i, T.n0);
                    continue;
                }
                const ConvTile T = conv_tile_decode(A, t - F.n_roll);
                for (int cb = 0; cb < cblks; ++cb)
                    for (int tap = 0; tap < 9; ++tap) load_b(&M.w[T.plane], tap * cblks + cb, T.n0);
                for (int j = 0; j < nskip; ++j) load_b(&M.w[T.plane], 9 * cblks + j, T.n0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(kBM, kBN);
            constexpr uint32_t idesc2 = ptx::make_idesc_f16(kBM, 2 * kBN);
            int ga = 0, gb = 0, lt = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
                const int as = lt & 1;
                ptx::mbar_wait(&tmem_empty_bar[as], ((lt >> 1) & 1) ^ 1);        // epilogue has drained this accumulator stage
                ptx::tc_fence_after();
                const uint32_t d1 = tmem_base + as * Cfg::kAccCols, d2 = d1 + kBN;
                const bool roll = t < F.n_roll;
                const int n_halo = roll ? 0 : cblks;
                const int n_groups = roll ? 3 * cblks : cblks + nskip;
                bool first = true;
                for (int g = 0; g < n_groups; ++g, ++ga) {
                    const int sa = ga % Cfg::kASlots;
                    ptx::mbar_wait(&fullA[sa], (ga / Cfg::kASlots) & 1);
                    const uint32_t a_base = ptx::smem_u32(smem_a + sa * Cfg::kASlotBytes);
                    const bool halo = g < n_halo;
                    const int nb = halo ? 9 : 1;
                    for (int tap = 0; tap < nb; ++tap, ++gb) {
                        const int sb = gb % Cfg::kBSlots;
                        ptx::mbar_wait(&fullB[sb], (gb / Cfg::kBSlots) & 1);
                        ptx::tc_fence_after();
                        if (first && lt < 3) trace_mark(A.tr, 4 + 6 * lt + 0);
                        const uint32_t b_base = ptx::smem_u32(smem_b + sb * Cfg::kBSlotBytes);
                        uint64_t a_hi, a_lo;
                        if (halo) {
                            const int kh = tap / 3, kw = tap - kh * 3;
                            const uint32_t off = static_cast<uint32_t>(kh * kHaloW + kw) * 128u;
                            const uint32_t bo = A.bo_kw ? static_cast<uint32_t>(kw) : 0u;
                            a_hi = ptx::make_sw128_desc(a_base + off, kHaloW * 128u, bo);
                            a_lo = ptx::make_sw128_desc(a_base + kALo + off, kHaloW * 128u, bo);
                        } else {
                            a_hi = ptx::make_sw128_desc(a_base, 1024u, 0);
                            a_lo = ptx::make_sw128_desc(a_base + kALo, 1024u, 0);
                        }
                        const uint64_t b_hi = ptx::make_sw128_desc(b_base, 1024u, 0);
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k) {
                            const uint64_t ko = static_cast<uint64_t>((k * 32) >> 4);   // +32 B along K inside the swizzle atom
                            const uint32_t acc = (!first || k > 0) ? 1u : 0u;
                            if (NSPLIT == 3) {
                                // [D1 | D2] (+)= Ah * [Bh | Bl]  (one N=128 MMA: the lo weight tile sits right behind the hi
                                // tile in shared memory and D2 right behind D1 in TMEM), then D2 += Al * Bh
                                ptx::umma_f16(d1, a_hi + ko, b_hi + ko, idesc2, acc);
                                ptx::umma_f16(d2, a_lo + ko, b_hi + ko, idesc, 1u);
                            } else {
                                ptx::umma_f16(d1, a_hi + ko, b_hi + ko, idesc, acc);
                            }
                        }
                        first = false;
                        ptx::umma_commit(&emptyB[sb]);    // weight slot free once the MMAs above retire
                    }
                    ptx::umma_commit(&emptyA[sa]);        // A patch free once every tap has read it
                }
                ptx::umma_commit(&tmem_full_bar[as]);     // this tile's accumulators are complete
                if (lt < 3) trace_mark(A.tr, 4 + 6 * lt + 1);
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int m = quarter * 32 + lane;
        const int et = threadIdx.x - 64;              // 0..127 among the epilogue threads
        bool roll_ready = F.n_roll == 0;
        int lt = 0;
        pdl_wait();
        if (F.n_roll && F.sums) {
            const int C = A.C;
            const long long per_sample = static_cast<long long>(F.total_len) * C;
            const long long total = per_sample * F.B;
            const size_t lo_off = static_cast<size_t>(total);
            for (long long i = static_cast<long long>(blockIdx.x) * 128 + et; i < total; i += static_cast<long long>(gridDim.x) * 128) {
                const int pos = static_cast<int>((i % per_sample) / C);
                int seg = 0;
#pragma unroll
                for (int k = 0; k < 5; ++k)
                    if (pos >= F.seg_end[k]) seg = k + 1;
                const long long sv = static_cast<long long>(__ldcg(F.sums + i));
                F.sums[i] = 0ull;
                const float mean = __ll2float_rn(sv) * F.seg_scale[seg];
                __half hi, lo;
                split_f16(mean, hi, lo);
                F.means16[i] = hi;
                F.means16[lo_off + i] = lo;
            }
            __threadfence();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) atomicAdd(&F.counters[2], 1u);
            if (et == 0) trace_mark(A.tr, 2);
        }
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
            const int as = lt & 1;
            const uint32_t aph = (lt >> 1) & 1;
            const uint32_t lane_addr = tmem_base + as * Cfg::kAccCols + (static_cast<uint32_t>(quarter * 32) << 16);
            if (t < F.n_roll) {
                // ---------- rollout 1-D GEMM tile: T[b][cls][pos][co] = accumulator ----------
                const RollTile T = roll_tile_decode(F, t);
                const int pos = T.p0 + m, L = F.R.L[T.src];
                const int cls = T.n0 / A.Cout, co0 = T.n0 - cls * A.Cout;
                float* __restrict__ outp = F.R.T[T.src] + ((static_cast<size_t>(T.b) * 4 + cls) * L + pos) * A.Cout + co0;
                ptx::mbar_wait(&tmem_full_bar[as], aph);
                if (et == 0 && lt < 3) trace_mark(A.tr, 4 + 6 * lt + 3);
                __syncwarp();
                ptx::tc_fence_after();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t v1[32], v2[32];
                    ptx::tmem_ld_32x32b_x32(lane_addr + half * 32, v1);
                    if (NSPLIT == 3) ptx::tmem_ld_32x32b_x32(lane_addr + kBN + half * 32, v2);
                    ptx::tmem_ld_wait();
                    if (half == 1) {
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[as]);
                    }
                    if (pos < L) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float o[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                o[q] = __uint_as_float(v1[j + q]);
                                if (NSPLIT == 3) o[q] = fmaf(__uint_as_float(v2[j + q]), 1.f / kLoScale, o[q]);
                            }
                            __stcg(reinterpret_cast<float4*>(outp + half * 32 + j), make_float4(o[0], o[1], o[2], o[3]));
                        }
                    }
                }
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (et == 0) atomicAdd(&F.counters[0], 1u);
                if (et == 0 && lt < 3) trace_mark(A.tr, 4 + 6 * lt + 4);
                continue;
            }
            const ConvTile T = conv_tile_decode(A, t - F.n_roll);
            const int plane = T.plane, n0 = T.n0, b = T.b;
            const int r = T.h0 + (m >> 3), c = T.w0 + (m & 7);
            const int rows = A.d.rows[plane], cols = A.d.cols[plane];
            const bool valid = r < rows && c < cols;
            const size_t px = static_cast<size_t>(b) * rows * cols + static_cast<size_t>(r) * cols + c;
            // Everything the epilogue adds to the accumulator (bias + rollout 1-D terms + additive embedding + identity
            // residual) is gathered into registers while the MMA pipeline of this tile is still running.
            float pre[kBN];
            {
                const float4* bias4 = reinterpret_cast<const float4*>(A.e.bias.p[plane] + n0);
#pragma unroll
                for (int j = 0; j < kBN / 4; ++j) {
                    const float4 v = __ldg(bias4 + j);
                    pre[4 * j] = v.x; pre[4 * j + 1] = v.y; pre[4 * j + 2] = v.z; pre[4 * j + 3] = v.w;
                }
                if (A.e.embadd) {
                    const float4* e4 = reinterpret_cast<const float4*>(
                        A.e.embadd + static_cast<size_t>(A.e.film_row ? A.e.film_row[b] : b) * A.e.film_dim + A.e.film_off + n0);
#pragma unroll
                    for (int j = 0; j < kBN / 4; ++j) {
                        const float4 v = __ldg(e4 + j);
                        pre[4 * j] += v.x; pre[4 * j + 1] += v.y; pre[4 * j + 2] += v.z; pre[4 * j + 3] += v.w;
                    }
                }
                if (valid && A.e.resid.p[plane]) {
                    const float4* rs = reinterpret_cast<const float4*>(A.e.resid.p[plane] + px * A.Cout + n0);
#pragma unroll
                    for (int j = 0; j < kBN / 4; ++j) {
                        const float4 v = __ldg(rs + j);
                        pre[4 * j] += v.x; pre[4 * j + 1] += v.y; pre[4 * j + 2] += v.z; pre[4 * j + 3] += v.w;
                    }
                }
                if (A.e.Trow.p[plane] && !roll_ready) {
                    // the rollout terms are produced by this very launch (roll tiles): wait until all of them are written
                    if (lane == 0) {
                        const volatile unsigned int* cnt = F.counters;
                        long long spins = 0;
                        while (*cnt < static_cast<unsigned int>(F.n_roll)) {
                            if (++spins > (1LL << 31)) __trap();     // never hang the device
                        }
                    }
                    __syncwarp();
                    __threadfence();
                    roll_ready = true;
                }
                if (valid && A.e.Trow.p[plane]) {
                    const size_t bo = static_cast<size_t>(b) * 4;
                    const float4* tr = reinterpret_cast<const float4*>(
                        A.e.Trow.p[plane] + ((bo + edge_class(c, cols)) * rows + r) * A.Cout + n0);
                    const float4* tc = reinterpret_cast<const float4*>(
                        A.e.Tcol.p[plane] + ((bo + edge_class(r, rows)) * cols + c) * A.Cout + n0);
#pragma unroll
                    for (int j = 0; j < kBN / 4; ++j) {
                        const float4 v = __ldcg(tr + j), u = __ldcg(tc + j);      // written by this launch: L2-coherent loads
                        pre[4 * j] += v.x + u.x; pre[4 * j + 1] += v.y + u.y; pre[4 * j + 2] += v.z + u.z;
                        pre[4 * j + 3] += v.w + u.w;
                    }
                }
            }
            float* __restrict__ outp = A.e.out.p[plane] + px * A.Cout + n0;
            if (et == 0 && lt < 3) trace_mark(A.tr, 4 + 6 * lt + 2);
            ptx::mbar_wait(&tmem_full_bar[as], aph);
            if (et == 0 && lt < 3) trace_mark(A.tr, 4 + 6 * lt + 3);
            __syncwarp();
            ptx::tc_fence_after();
            const bool do_stats = A.sink.acc != nullptr;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v1[32], v2[32];
                ptx::tmem_ld_32x32b_x32(lane_addr + half * 32, v1);
                if (NSPLIT == 3) ptx::tmem_ld_32x32b_x32(lane_addr + kBN + half * 32, v2);
                ptx::tmem_ld_wait();
                if (half == 1) {
                    // all TMEM reads of this tile are done: hand the accumulator stage back to the MMA issuer
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[as]);
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float o[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float acc = __uint_as_float(v1[j + q]);
                        if (NSPLIT == 3) acc = fmaf(__uint_as_float(v2[j + q]), 1.f / kLoScale, acc);
                        o[q] = acc + pre[half * 32 + j + q];
                    }
                    if (valid) *reinterpret_cast<float4*>(outp + half * 32 + j) = make_float4(o[0], o[1], o[2], o[3]);
                    if (do_stats) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) stat_stage[m * 33 + j + q] = valid ? o[q] : 0.f;
                    }
                }
                if (do_stats) {
                    // per-channel (sum, sum-sq) of this tile's 128 pixels, fixed summation order
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    {
                        const int cc = et & 31, qq = et >> 5;
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
                        for (int p = 0; p < 32; ++p) {
                            const float v = stat_stage[(qq * 32 + p) * 33 + cc];
                            s1 += v;
                            s2 = fmaf(v, v, s2);
                        }
                        stat_colp[(qq * 2 + 0) * 32 + cc] = s1;
                        stat_colp[(qq * 2 + 1) * 32 + cc] = s2;
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (et < 64) {
                        const int which = et >> 5, cc = et & 31;
                        stat_tot[which * kBN + half * 32 + cc] = (stat_colp[(0 * 2 + which) * 32 + cc] + stat_colp[(1 * 2 + which) * 32 + cc]) +
                                                                 (stat_colp[(2 * 2 + which) * 32 + cc] + stat_colp[(3 * 2 + which) * 32 + cc]);
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
            }
            if (do_stats) {
                const int cpg = A.Cout / kGroups, gpt = kBN / cpg, g0 = n0 / cpg;
                if (et < 2 * gpt) {
                    const int gl = et >> 1, which = et & 1;
                    double acc = 0.0;
                    for (int cc = gl * cpg; cc < (gl + 1) * cpg; ++cc) acc += static_cast<double>(stat_tot[which * kBN + cc]);
                    gn_fix_add(A.sink.acc + (static_cast<size_t>(b) * 3 + plane) * 64 + (g0 + gl) * 2 + which, acc);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");      // stat_tot is reused by the next tile
            }
            if (et == 0 && lt < 3) trace_mark(A.tr, 4 + 6 * lt + 4);
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
            F.counters[2] = 0u;
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
    const int p0 = (blockIdx.x - A.tile_start[src]) * kBM;
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
                ptx::tma_load_3d(st + kBOff, &M.w[src], &full_bar[s], i * kBK, n0, 0);
                if (NSPLIT == 3) {
                    ptx::tma_load_5d(st + kABytes, &M.a[src], &full_bar[s], cb * kBK, p0 + al - 1, 0, b, 1);
                    ptx::tma_load_3d(st + kBOff + kBBytes, &M.w[src], &full_bar[s], i * kBK, n0, 1);
                }
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
        const bool valid = pos < L;
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
