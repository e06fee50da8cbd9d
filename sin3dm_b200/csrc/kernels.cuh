// Bandwidth-bound / small kernels of the triplane UNet and the scheduler.
// Activations: per-plane NHWC fp32 residual stream [B][rows][cols][C]; conv operands are fp16
// (hi, lo) pairs laid out [2][B][rows][cols][C] so one 5-D TMA map serves both halves.
#pragma once
#include "common.cuh"

namespace s3d {

// =====================================================================================
// GroupNorm statistics -> per-channel affine coefficients.
// reference src/diffusion/nn.py:17-19 (GroupNorm32 computes in fp32), unet_triplane.py:63-84, FiLM :285-297
// Every kernel that writes a tensor which is normalised next also adds its share of the 32 groups' (sum, sum-sq) to 64
// accumulators per (sample, plane): 64-bit fixed point (value * 2^20) with integer atomics, so the totals are exact sums of
// the per-CTA partials — independent of CTA order, batch composition and GPU count.  The consumer turns them into
//     y = x * coefA[c] + coefB[c]   ==  GroupNorm(x) * gamma + beta   (optionally * (1 + scale) + shift)
// with one 64-value load per CTA (no statistics pass, no partial-slot reduction).  The accumulators of a tensor are re-zeroed
// by the NEXT consumer kernel in the step (its `zero` job), i.e. after their only reader has finished.
// =====================================================================================
constexpr int kGnRep = 8;                         // replicas of the accumulators: same-address atomics serialise in the L2
                                                  // (~25 ns each, measured), so producers spread over kGnRep copies
constexpr double kGnFix = 1048576.0;              // 2^20
constexpr double kGnFixInv = 1.0 / 1048576.0;
struct StatsSink {                 // producer side
    unsigned long long* acc;  // [B][3][kGnRep][64]  (sum, sum-sq) of the 32 groups, fixed point; nullptr: not wanted
    int C;
};
struct StatsSrc {                  // consumer side
    const unsigned long long* acc;   // [B][3][kGnRep][64]
    TriCF gamma, beta;        // consumer norm parameters [C]
    const float* film;        // [rows][film_dim] or nullptr
    const int* film_row;
    int film_dim, film_off;   // scale at film_off, shift at film_off + C
    unsigned long long* zero; // accumulators of the previous consumer's tensor (already read): re-zeroed by CTA 0
    int zero_n;
};
__device__ __forceinline__ void gn_fix_add(unsigned long long* p, double v) {
    atomicAdd(p, static_cast<unsigned long long>(__double2ll_rn(v * kGnFix)));
}
// accumulator block of (sample b, plane) for a producer that identifies itself by `who` (tile / CTA index)
__device__ __forceinline__ unsigned long long* gn_acc(unsigned long long* acc, int b, int plane, int who) {
    return acc + ((static_cast<size_t>(b) * 3 + plane) * kGnRep + (who & (kGnRep - 1))) * 64;
}

// Consumer prologue (all threads of the CTA, nthr >= 64); fin: smem double[64]; coefA/coefB: smem float[C].
__device__ __forceinline__ void stats_coef_prologue(const StatsSrc& S, int b, int plane, int C, double n_per_group, int tid, int nthr,
                                                    double* fin, float* coefA, float* coefB) {
    const float* film = nullptr;
    if (S.film) film = S.film + static_cast<size_t>(S.film_row ? S.film_row[b] : b) * S.film_dim + S.film_off;
    if (tid < 64) {
        const unsigned long long* p = S.acc + (static_cast<size_t>(b) * 3 + plane) * kGnRep * 64 + tid;
        unsigned long long v[kGnRep];
#pragma unroll
        for (int r = 0; r < kGnRep; ++r) v[r] = __ldcg(p + r * 64);
        unsigned long long tot = 0ull;                  // integer sum of the replicas: exact, order-free
#pragma unroll
        for (int r = 0; r < kGnRep; ++r) tot += v[r];
        fin[tid] = static_cast<double>(static_cast<long long>(tot)) * kGnFixInv;
    }
    if (S.zero && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
        for (int i = tid; i < S.zero_n; i += nthr) S.zero[i] = 0ull;
    __syncthreads();
    const int cpg = C / kGroups;
    const double inv_n = 1.0 / n_per_group;
    for (int c = tid; c < C; c += nthr) {
        const int g = c / cpg;
        // E[x^2] - mean^2 is formed in fp64 (cancellation), the rest in fp32 like torch's GroupNorm
        const double mean = fin[g * 2] * inv_n;
        const float var = fmaxf(static_cast<float>(fin[g * 2 + 1] * inv_n - mean * mean), 0.f);
        const float rstd = rsqrtf(var + kGnEps);
        float ga = __ldg(S.gamma.p[plane] + c) * rstd;
        float be = __ldg(S.beta.p[plane] + c) - static_cast<float>(mean) * ga;
        if (film) {
            const float sc = 1.f + __ldg(film + c), sh = __ldg(film + C + c);
            ga *= sc;
            be = fmaf(be, sc, sh);
        }
        coefA[c] = ga;
        coefB[c] = be;
    }
    __syncthreads();
}

// Block-wide reduction of per-thread (sum, sum-sq) float4 pairs laid out as block (C/4, NY), added to the accumulators.
// red: smem [(NY*2 + 2) * C] floats.
__device__ __forceinline__ void stats_block_add(const StatsSink& S, float4 s, float4 q, int b, int plane, float* red) {
    const int C = S.C, tx = threadIdx.x, ty = threadIdx.y, NY = blockDim.y;
    const int tid = ty * blockDim.x + tx, nthr = blockDim.x * NY;
    float* rs = red + (ty * 2 + 0) * C + tx * 4;
    float* rq = red + (ty * 2 + 1) * C + tx * 4;
    rs[0] = s.x; rs[1] = s.y; rs[2] = s.z; rs[3] = s.w;
    rq[0] = q.x; rq[1] = q.y; rq[2] = q.z; rq[3] = q.w;
    __syncthreads();
    float* tot = red + NY * 2 * C;
    for (int i = tid; i < 2 * C; i += nthr) {
        const int which = i / C, c = i - which * C;
        float acc = 0.f;
        for (int y = 0; y < NY; ++y) acc += red[(y * 2 + which) * C + c];
        tot[i] = acc;
    }
    __syncthreads();
    const int cpg = C / kGroups;
    if (tid < 2 * kGroups) {
        const int g = tid >> 1, which = tid & 1;
        double acc = 0.0;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) acc += static_cast<double>(tot[which * C + c]);
        gn_fix_add(gn_acc(S.acc, b, plane, blockIdx.x) + g * 2 + which, acc);
    }
}

// Stand-alone statistics pass (used where the producer kernel does not emit partials itself).
// grid (nslots, 3, B), block (C/4, NY)
__global__ void __launch_bounds__(1024) k_gn_stats(TriCF x, TriDims d, StatsSink S, int nslots) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float red[];   // [(NY*2 + 2) * C]
    const int plane = blockIdx.y, b = blockIdx.z, slot = blockIdx.x;
    const int C = S.C, tx = threadIdx.x, ty = threadIdx.y, NY = blockDim.y;
    const int npx = d.rows[plane] * d.cols[plane];
    const int ppc = (npx + nslots - 1) / nslots;
    const int p0 = slot * ppc, p1 = min(npx, p0 + ppc);
    const float* xp = x.p[plane] + static_cast<size_t>(b) * npx * C;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    for (int px = p0 + ty; px < p1; px += 4 * NY) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            v[k] = px + k * NY < p1 ? __ldg(reinterpret_cast<const float4*>(xp + static_cast<size_t>(px + k * NY) * C) + tx)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w;
            q.x = fmaf(v[k].x, v[k].x, q.x); q.y = fmaf(v[k].y, v[k].y, q.y);
            q.z = fmaf(v[k].z, v[k].z, q.z); q.w = fmaf(v[k].w, v[k].w, q.w);
        }
    }
    stats_block_add(S, s, q, b, plane, red);
}

// -------------------------------------------------------------------------------------
// Element-wise producers.  All of them use block (C/4, NY), give each CTA ("slot") a contiguous pixel range of one
// plane of one sample, write the fp32 NHWC result and — when `S.partial` is set — also emit the GroupNorm partial sums
// of what they wrote, so the consumer's statistics need no extra pass over the tensor.
// grid (nslots, 3, B)
// -------------------------------------------------------------------------------------
#define S3D_PRODUCER_TAIL(S, s, q, npx, red) \
    if ((S).acc) stats_block_add(S, s, q, b, plane, red);

__device__ __forceinline__ void acc_sq(float4& s, float4& q, const float4& v) {
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
}

// 2x2 average pool, stride 2, floor on odd sizes.  reference unet_triplane.py:127-145.  smem: red[(NY*2+2)*C]
__global__ void __launch_bounds__(256) k_avgpool2(TriCF x, TriDims din, TriDims dout, int C, TriF out, StatsSink S, int nslots,
                                                  Trace tr) {
    pdl_wait();
    pdl_trigger();
    if (threadIdx.x + threadIdx.y == 0) trace_mark(tr, 0);
    extern __shared__ float smi[];
    const int plane = blockIdx.y, b = blockIdx.z, slot = blockIdx.x;
    const int tx = threadIdx.x, ty = threadIdx.y, NY = blockDim.y;
    const int orows = dout.rows[plane], ocols = dout.cols[plane], icols = din.cols[plane];
    const int npx = orows * ocols, c4 = C / 4;
    const int ppc = (npx + nslots - 1) / nslots;
    const int p0 = slot * ppc, p1 = min(npx, p0 + ppc);
    const float4* ip = reinterpret_cast<const float4*>(x.p[plane] + static_cast<size_t>(b) * din.rows[plane] * icols * C);
    float4* op = reinterpret_cast<float4*>(out.p[plane] + static_cast<size_t>(b) * npx * C);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    for (int px = p0 + ty; px < p1; px += NY) {
        const int r = px / ocols, c = px - r * ocols;
        const float4* base = ip + (static_cast<size_t>(2 * r) * icols + 2 * c) * c4 + tx;
        const float4 a = __ldg(base), b4 = __ldg(base + c4), c0 = __ldg(base + static_cast<size_t>(icols) * c4),
                     d4 = __ldg(base + static_cast<size_t>(icols + 1) * c4);
        float4 o;
        o.x = (a.x + b4.x + c0.x + d4.x) * 0.25f;
        o.y = (a.y + b4.y + c0.y + d4.y) * 0.25f;
        o.z = (a.z + b4.z + c0.z + d4.z) * 0.25f;
        o.w = (a.w + b4.w + c0.w + d4.w) * 0.25f;
        op[static_cast<size_t>(px) * c4 + tx] = o;
        acc_sq(s, q, o);
    }
    if (threadIdx.x + threadIdx.y == 0) trace_mark(tr, 1);
    S3D_PRODUCER_TAIL(S, s, q, npx, smi)
    if (threadIdx.x + threadIdx.y == 0) trace_mark(tr, 2);
}

// Bilinear x2 upsample (align_corners=False) [+ bilinear resize to the skip's size when they differ]
// + channel concat with the skip.  reference unet_triplane.py:106-124, 494-503.
// out: (hi, lo) fp16 pair [2][B][rows][cols][Cu+Cs] — the only form its two readers need (k_gn_silu re-joins the halves,
// k_conv_tc's 1x1 skip GEMM reads them as they are), which saves writing the tensor a second time.
__device__ __forceinline__ void bilin_src(int dst, int in_size, float scale, int& i0, int& i1, float& l1) {
    // ATen area_pixel_compute_source_index(align_corners=false): scale*(dst+0.5)-0.5 clamped at 0
    float s = fmaxf(scale * (static_cast<float>(dst) + 0.5f) - 0.5f, 0.f);
    i0 = min(static_cast<int>(s), in_size - 1);
    i1 = min(i0 + 1, in_size - 1);
    l1 = s - static_cast<float>(i0);
}

__global__ void __launch_bounds__(256) k_upcat(TriCF low, TriDims dlow, int Cu, TriCF skip, int Cs, TriDims dout, TriH out, int B,
                                               int do_up, StatsSink S, int nslots, Trace tr) {
    pdl_wait();
    pdl_trigger();
    if (threadIdx.x + threadIdx.y == 0) trace_mark(tr, 0);
    extern __shared__ float smi[];
    const int plane = blockIdx.y, b = blockIdx.z, slot = blockIdx.x;
    const int tx = threadIdx.x, ty = threadIdx.y, NY = blockDim.y;
    const int orows = dout.rows[plane], ocols = dout.cols[plane];
    const int lrows = dlow.rows[plane], lcols = dlow.cols[plane];
    const int Ct = Cu + Cs, c4 = Ct / 4, u4 = Cu / 4;
    const int npx = orows * ocols;
    const int ppc = (npx + nslots - 1) / nslots;
    const int p0 = slot * ppc, p1 = min(npx, p0 + ppc);
    const float4* sp = Cs ? reinterpret_cast<const float4*>(skip.p[plane] + static_cast<size_t>(b) * npx * Cs) : nullptr;
    const float4* lp = reinterpret_cast<const float4*>(low.p[plane] + static_cast<size_t>(b) * lrows * lcols * Cu);
    __half* op = out.p[plane] + static_cast<size_t>(b) * npx * Ct;
    const size_t lo_off = static_cast<size_t>(B) * npx * Ct;
    const int urows = do_up ? 2 * lrows : lrows, ucols = do_up ? 2 * lcols : lcols;
    const bool resize = urows != orows || ucols != ocols;
    auto at = [&](int rr, int cc) { return __ldg(lp + (static_cast<size_t>(rr) * lcols + cc) * u4 + tx); };
    auto lerp4 = [](const float4& a, const float4& bb, const float4& cc2, const float4& dd, float lr, float lc) {
        const float w00 = (1.f - lr) * (1.f - lc), w01 = (1.f - lr) * lc, w10 = lr * (1.f - lc), w11 = lr * lc;
        float4 t;
        t.x = w00 * a.x + w01 * bb.x + w10 * cc2.x + w11 * dd.x;
        t.y = w00 * a.y + w01 * bb.y + w10 * cc2.y + w11 * dd.y;
        t.z = w00 * a.z + w01 * bb.z + w10 * cc2.z + w11 * dd.z;
        t.w = w00 * a.w + w01 * bb.w + w10 * cc2.w + w11 * dd.w;
        return t;
    };
    // value of the (virtual) x2-upsampled map at (ur, uc)
    auto up_at = [&](int ur, int uc) {
        if (!do_up) return at(ur, uc);
        int r0, r1, c0, c1;
        float lr, lc;
        bilin_src(ur, lrows, 0.5f, r0, r1, lr);
        bilin_src(uc, lcols, 0.5f, c0, c1, lc);
        return lerp4(at(r0, c0), at(r0, c1), at(r1, c0), at(r1, c1), lr, lc);
    };
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    // Four pixels per iteration with every gather of the batch issued before the first use (the loop is latency bound
    // otherwise).  The common case (plain x2 upsample, or plain copy) is written branch-free for that; the resize-to-skip
    // case (odd sizes) takes 16 gathers per pixel and goes one pixel at a time.
    constexpr int kUpBatch = 4;
    const bool is_skip = tx >= u4;
    for (int pb = p0 + ty; pb < p1; pb += kUpBatch * NY) {
        float4 o[kUpBatch];
        if (resize && !is_skip) {
#pragma unroll
            for (int j = 0; j < kUpBatch; ++j) {
                const int px = pb + j * NY;
                o[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (px < p1) {
                    const int r = px / ocols, c = px - r * ocols;
                    int r0, r1, c0, c1;
                    float lr, lc;
                    bilin_src(r, urows, static_cast<float>(urows) / static_cast<float>(orows), r0, r1, lr);
                    bilin_src(c, ucols, static_cast<float>(ucols) / static_cast<float>(ocols), c0, c1, lc);
                    o[j] = lerp4(up_at(r0, c0), up_at(r0, c1), up_at(r1, c0), up_at(r1, c1), lr, lc);
                }
            }
        } else {
            const float4* src[kUpBatch][4];
            float lr[kUpBatch], lc[kUpBatch];
#pragma unroll
            for (int j = 0; j < kUpBatch; ++j) {
                const int px = min(pb + j * NY, p1 - 1);          // clamped: the tail re-reads a valid pixel and drops it
                const int r = px / ocols, c = px - r * ocols;
                if (is_skip) {
                    src[j][0] = src[j][1] = src[j][2] = src[j][3] = sp + static_cast<size_t>(px) * (Cs / 4) + (tx - u4);
                    lr[j] = lc[j] = 0.f;
                } else {
                    int r0 = r, r1 = r, c0 = c, c1 = c;
                    lr[j] = lc[j] = 0.f;
                    if (do_up) {
                        bilin_src(r, lrows, 0.5f, r0, r1, lr[j]);
                        bilin_src(c, lcols, 0.5f, c0, c1, lc[j]);
                    }
                    src[j][0] = lp + (static_cast<size_t>(r0) * lcols + c0) * u4 + tx;
                    src[j][1] = lp + (static_cast<size_t>(r0) * lcols + c1) * u4 + tx;
                    src[j][2] = lp + (static_cast<size_t>(r1) * lcols + c0) * u4 + tx;
                    src[j][3] = lp + (static_cast<size_t>(r1) * lcols + c1) * u4 + tx;
                }
            }
            float4 g[kUpBatch][4];
            const bool one = is_skip || !do_up;                   // a single source element per pixel
#pragma unroll
            for (int j = 0; j < kUpBatch; ++j) {
                g[j][0] = __ldg(src[j][0]);
                if (!one) {
                    g[j][1] = __ldg(src[j][1]);
                    g[j][2] = __ldg(src[j][2]);
                    g[j][3] = __ldg(src[j][3]);
                }
            }
#pragma unroll
            for (int j = 0; j < kUpBatch; ++j) o[j] = one ? g[j][0] : lerp4(g[j][0], g[j][1], g[j][2], g[j][3], lr[j], lc[j]);
        }
#pragma unroll
        for (int j = 0; j < kUpBatch; ++j) {
            const int px = pb + j * NY;
            if (px < p1) {
                __half* dst = op + static_cast<size_t>(px) * Ct + tx * 4;
                store_split4(dst, dst + lo_off, o[j]);
                // the statistics are those of the values the consumer will see (the re-joined halves)
                acc_sq(s, q, roundtrip_split4(o[j]));
            }
        }
    }
    if (threadIdx.x + threadIdx.y == 0) trace_mark(tr, 1);
    S3D_PRODUCER_TAIL(S, s, q, npx, smi)
    if (threadIdx.x + threadIdx.y == 0) trace_mark(tr, 2);
}

// =====================================================================================
// Fused GroupNorm-apply (+FiLM) + SiLU -> fp16 (hi, lo) conv operand, plus the rollout axis means.
// reference unet_triplane.py:63-95 (norm, SiLU), :285-297 (FiLM), :37-46 (axis means)
// One CTA = an 8-row x (ny * ncg)-column tile of one plane of one sample, walked as ncg groups of ny columns (one column per
// thread per group, 8 independent loads in flight, the next group's loads issued under the current group's math; the per-CTA
// statistics prologue is amortised over the tile).  The host picks ncg so that the launch is a single wave when it can be.
// grid (max tiles, 3, B), block (C/4, ny).
// Axis sums are accumulated as 64-bit fixed point (value * 2^24) with integer atomics: exact, hence independent of
// tile order / batch composition / GPU count.  sums[b][seg_off[plane*2+kind] + pos][C], kind 0 = sum over columns
// (indexed by row), kind 1 = sum over rows (indexed by column).  The CTA that completes a row strip / a column tile
// (per-strip and per-column-tile tickets) turns those sums into fp16 (hi, lo) means [2][B][total_len][C] — the A operand
// of the rollout 1-D GEMM — and re-zeroes them, so no separate finalize / memset pass exists.
// =====================================================================================
constexpr int kGsRows = 8;
constexpr float kFixScale = 16777216.f;          // 2^24
constexpr double kFixInv = 1.0 / 16777216.0;

struct GnSiluArgs {
    TriCF x;          // fp32 [B][rows][cols][C]   (or nullptr, then:)
    TriCH xh;         // (hi, lo) fp16 [2][B][rows][cols][C] input, read as hi + lo / 2048
    TriDims d;
    int C;
    StatsSrc st;              // group sums of x + this layer's norm parameters
    TriH a;                   // out [2][B][rows][cols][C]
    TriH x16;                 // optional raw copy of x as (hi, lo) for the 1x1 skip GEMM
    unsigned long long* sums; // [B][total_len][C] fixed point, zero between launches; nullptr when rollout is off
    __half* means16;          // [2][B][total_len][C]
    unsigned int* ticket;     // [B][total_tickets]: per plane, `strips` row-strip tickets then `ctiles` column-tile tickets
    int tick_off[3];          // offset of plane p's tickets
    int total_tickets;
    int seg_off[6];
    int total_len;
    Trace tr;                 // slots: 0 entry, 1 coefficients ready, 2 column atomics issued, 3 row atomics + operand stores issued
    unsigned long long* zero_sums;   // axis sums of an earlier site whose reader (a conv's roll tiles) is done: cleared here
    long long zero_sums_n;
    int ncg;                  // column groups per CTA: the CTA's tile is kGsRows rows x (blockDim.y * ncg) columns
    int finalize;             // 1: the last CTA of a strip / column tile converts the sums to fp16 means itself (stand-alone roll
                              //    kernels); 0: k_conv_tc does it in its phase 0 and this kernel ends right after the atomics
};

__device__ __forceinline__ void fix_add(unsigned long long* p, float v) {
    atomicAdd(p, static_cast<unsigned long long>(__float2ll_rn(v * kFixScale)));
}

// sums -> fp16 (hi, lo) means for `n` consecutive elements, re-zeroing the accumulators
__device__ __forceinline__ void means_finalize(unsigned long long* sp, __half* mh, __half* ml, int n, double scale, int tid, int nthr) {
    for (int i = tid; i < n; i += nthr) {
        const long long sv = static_cast<long long>(__ldcg(sp + i));
        sp[i] = 0ull;
        const float m = static_cast<float>(static_cast<double>(sv) * scale);
        __half hi, lo;
        split_f16(m, hi, lo);
        mh[i] = hi;
        ml[i] = lo;
    }
}

__global__ void __launch_bounds__(256, 2) k_gn_silu(GnSiluArgs A, int B) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float gsm[];   // coefA[C], coefB[C], red[ny][8][C]
    __shared__ int last_row, last_col;
    __shared__ double fin[64];
    const int plane = blockIdx.y, b = blockIdx.z;
    const int rows = A.d.rows[plane], cols = A.d.cols[plane], C = A.C;
    const int tx = threadIdx.x, ty = threadIdx.y, ny = blockDim.y;
    const int tid = ty * blockDim.x + tx, nthr = blockDim.x * ny;
    const int ncg = A.ncg, tw = ny * ncg;                    // the CTA's tile is kGsRows x tw pixels, ncg column groups of ny
    const int ctiles = (cols + tw - 1) / tw, strips = (rows + kGsRows - 1) / kGsRows;
    if (A.zero_sums) {
        const long long cta = blockIdx.x + static_cast<long long>(gridDim.x) * (blockIdx.y + gridDim.y * blockIdx.z);
        const long long ncta = static_cast<long long>(gridDim.x) * gridDim.y * gridDim.z;
        for (long long i = cta * nthr + tid; i < A.zero_sums_n; i += ncta * nthr) A.zero_sums[i] = 0ull;
    }
    if (static_cast<int>(blockIdx.x) >= ctiles * strips) return;
    if (tid == 0) trace_mark(A.tr, 0);
    const int strip = blockIdx.x / ctiles, ct = blockIdx.x - strip * ctiles;
    const int r0 = strip * kGsRows;
    const int nr = min(kGsRows, rows - r0);
    const int cbase = ct * tw;
    float* coefA = gsm;
    float* coefB = gsm + C;
    float* red = gsm + 2 * C;
    const size_t plane_elems = static_cast<size_t>(rows) * cols * C;
    const size_t sample_off = static_cast<size_t>(b) * plane_elems;
    const size_t lo_off = static_cast<size_t>(B) * plane_elems;
    const float* xp = A.x.p[plane] ? A.x.p[plane] + sample_off : nullptr;
    const __half* xhp = A.xh.p[plane] ? A.xh.p[plane] + sample_off : nullptr;
    __half* ap = A.a.p[plane] + sample_off;
    __half* xq = A.x16.p[plane] ? A.x16.p[plane] + sample_off : nullptr;
    float4 v[kGsRows], y[kGsRows];
    // Addresses: one base per column group, then a 32-bit row stride (an element offset inside one sample's plane fits an int);
    // the 64-bit multiply per row and per stream used to be a third of this kernel's instructions, and it is issue bound.
    const int row_stride = cols * C;
    auto load_group = [&](int cg) {
        const int c = cbase + cg * ny + ty;
        const int base = (r0 * cols + min(c, cols - 1)) * C + tx * 4;
        if (xp) {
            const float* pb = xp + base;
#pragma unroll
            for (int r = 0; r < kGsRows; ++r)
                v[r] = (c < cols && r < nr) ? __ldg(reinterpret_cast<const float4*>(pb + r * row_stride)) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            uint2 hh[kGsRows], ll[kGsRows];
            const __half* ph = xhp + base;
            const __half* pl = ph + lo_off;
#pragma unroll
            for (int r = 0; r < kGsRows; ++r) {
                hh[r] = ll[r] = make_uint2(0u, 0u);
                if (c < cols && r < nr) {
                    hh[r] = __ldg(reinterpret_cast<const uint2*>(ph + r * row_stride));
                    ll[r] = __ldg(reinterpret_cast<const uint2*>(pl + r * row_stride));
                }
            }
#pragma unroll
            for (int r = 0; r < kGsRows; ++r) v[r] = join_halves4(hh[r], ll[r]);
        }
    };
    // operand stores of one column group (fire and forget: the kernel boundary publishes them)
    auto store_group = [&](int cg) {
        const int c = cbase + cg * ny + ty;
        if (c >= cols) return;
        const int base = (r0 * cols + c) * C + tx * 4;
        __half* ah = ap + base;
        __half* al = ah + lo_off;
        __half* qh = xq ? xq + base : nullptr;
        __half* ql = xq ? qh + lo_off : nullptr;
#pragma unroll
        for (int r = 0; r < kGsRows; ++r) {
            if (r < nr) {
                if (xq) store_split4(qh + r * row_stride, ql + r * row_stride, v[r]);
                store_split4(ah + r * row_stride, al + r * row_stride, y[r]);
            }
        }
    };
    load_group(0);
    // the group sums arrive while the activation loads above are in flight
    stats_coef_prologue(A.st, b, plane, C, static_cast<double>(rows) * cols * (C / kGroups), tid, nthr, fin, coefA, coefB);
    const float4 ca = *reinterpret_cast<const float4*>(coefA + tx * 4);
    const float4 cb = *reinterpret_cast<const float4*>(coefB + tx * 4);
    if (tid == 0) trace_mark(A.tr, 1);
    unsigned long long* sb = A.sums ? A.sums + static_cast<size_t>(b) * A.total_len * C : nullptr;
    unsigned long long* srow = A.sums ? sb + static_cast<size_t>(A.seg_off[plane * 2 + 0]) * C : nullptr;
    unsigned long long* scol = A.sums ? sb + static_cast<size_t>(A.seg_off[plane * 2 + 1]) * C : nullptr;
    int last_cg = 0;
    for (int cg = 0; cg < ncg; ++cg) {
        if (cbase + cg * ny >= cols) break;                  // uniform: no columns left for this CTA
        const int c = cbase + cg * ny + ty;
        const bool cvalid = c < cols;
        if (cg > 0) {
            store_group(cg - 1);                             // behind the previous group's atomics
            load_group(cg);
        }
        last_cg = cg;
        float4 cacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kGsRows; ++r) {
            y[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cvalid && r < nr) {
                y[r].x = silu_f(fmaf(v[r].x, ca.x, cb.x));
                y[r].y = silu_f(fmaf(v[r].y, ca.y, cb.y));
                y[r].z = silu_f(fmaf(v[r].z, ca.z, cb.z));
                y[r].w = silu_f(fmaf(v[r].w, ca.w, cb.w));
            }
            cacc.x += y[r].x; cacc.y += y[r].y; cacc.z += y[r].z; cacc.w += y[r].w;
        }
        if (!A.sums) continue;
        // the axis-sum atomics go out first: they are what the kernel's completion waits for
        if (cvalid) {
            unsigned long long* p = scol + static_cast<size_t>(c) * C + tx * 4;
            fix_add(p, cacc.x); fix_add(p + 1, cacc.y); fix_add(p + 2, cacc.z); fix_add(p + 3, cacc.w);
        }
        // row sums of this thread's column accumulate in its own shared-memory cell across the column groups
#pragma unroll
        for (int r = 0; r < kGsRows; ++r) {
            float4* cell = reinterpret_cast<float4*>(red + (static_cast<size_t>(ty) * kGsRows + r) * C + tx * 4);
            if (cg == 0) {
                *cell = y[r];
            } else {
                float4 o = *cell;
                o.x += y[r].x; o.y += y[r].y; o.z += y[r].z; o.w += y[r].w;
                *cell = o;
            }
        }
    }
    if (tid == 0) trace_mark(A.tr, 2);
    if (!A.sums) {
        store_group(last_cg);
        return;
    }
    __syncthreads();
    for (int i = tid; i < nr * C; i += nthr) {
        int r = i / C, ch = i - r * C;
        float acc = 0.f;
        for (int yy = 0; yy < ny; ++yy) acc += red[(static_cast<size_t>(yy) * kGsRows + r) * C + ch];
        fix_add(srow + static_cast<size_t>(r0 + r) * C + ch, acc);
    }
    store_group(last_cg);
    if (tid == 0) trace_mark(A.tr, 3);
    if (!A.finalize) return;
    // ---- whoever completes this row strip / this column tile converts it to fp16 means and re-zeroes it
    __threadfence();
    __syncthreads();
    if (tid == 0 || tid == 32) {        // two lanes of different warps: the two round trips overlap
        unsigned int* tk = A.ticket + static_cast<size_t>(b) * A.total_tickets + A.tick_off[plane];
        if (tid == 0) {
            const unsigned int pr = atomicAdd(tk + strip, 1u);
            last_row = pr == static_cast<unsigned int>(ctiles - 1);
            if (last_row) tk[strip] = 0u;
        } else {
            const unsigned int pc = atomicAdd(tk + strips + ct, 1u);
            last_col = pc == static_cast<unsigned int>(strips - 1);
            if (last_col) tk[strips + ct] = 0u;
        }
    }
    __syncthreads();
    if (!last_row && !last_col) return;
    __threadfence();
    const size_t mlo = static_cast<size_t>(B) * A.total_len * C;
    __half* mb = A.means16 + static_cast<size_t>(b) * A.total_len * C;
    if (last_row) {
        const size_t o = (static_cast<size_t>(A.seg_off[plane * 2 + 0]) + r0) * C;
        means_finalize(sb + o, mb + o, mb + mlo + o, nr * C, kFixInv / static_cast<double>(cols), tid, nthr);
    }
    if (last_col) {
        const int nc = min(tw, cols - cbase);
        const size_t o = (static_cast<size_t>(A.seg_off[plane * 2 + 1]) + cbase) * C;
        means_finalize(sb + o, mb + o, mb + mlo + o, nc * C, kFixInv / static_cast<double>(rows), tid, nthr);
    }
}

// =====================================================================================
// Rollout 1-D terms, CUDA-core cross-check kernel (conv_impl = 1).  The tensor-core version is k_roll_tc.
// Two thirds of a rollout conv's input channels are constant along one image
// axis (unet_triplane.py:37-46), so their 3x3 conv collapses exactly to a 1-D conv along the other
// axis, with the zero padding only distinguishing first / interior / last position across:
//   T[b][cls][pos][co] = sum_{along, c} mean[pos+along-1][c] * wc[along*C + c][cls*Cout + co]
//   wc[..][cls] = sum of the taps `across` that class keeps: 0 interior {0,1,2}, 1 first {1,2}, 2 last {0,1}, 3 single {1}
// grid (ceil(Lmax/16), 6 * ntn, B): blockIdx.y = (plane*2 + group) * ntn + n_tile.  block 128.
// =====================================================================================
struct Roll1dSrc {
    int sum_off;        // segment offset (positions) of the source means inside a sample's block
    int L;
    int ncls;           // 3, or 4 when the axis across has length 1
    const float* wc;    // [3*C][4*Cout]
    float* T;           // [B][4][L][Cout]
};
struct Roll1dArgs {
    Roll1dSrc s[6];
    const __half* means16;   // [2][B][total_len][C]
    int total_len, B;
    int C, Cout, ntn;
};

__global__ void __launch_bounds__(128) k_roll1d(Roll1dArgs A) {
    pdl_wait();
    pdl_trigger();
    constexpr int POS = 16, NT = 64;
    extern __shared__ __align__(16) float sm1[];     // means[(POS+2)][C+4]
    const int src_id = blockIdx.y / A.ntn, nt = blockIdx.y - src_id * A.ntn;
    const Roll1dSrc S = A.s[src_id];
    const int b = blockIdx.z, C = A.C, Cout = A.Cout, N4 = 4 * Cout;
    const int p0 = blockIdx.x * POS, n0 = nt * NT;
    if (S.T == nullptr || p0 >= S.L || n0 >= S.ncls * Cout) return;
    const int CP = C + 4;
    float* means = sm1;
    const int tid = threadIdx.x;
    const __half* mh = A.means16 + (static_cast<size_t>(b) * A.total_len + S.sum_off) * C;
    const __half* ml = mh + static_cast<size_t>(A.B) * A.total_len * C;
    for (int i = tid; i < (POS + 2) * C; i += 128) {
        int j = i / C, c = i - j * C, pos = p0 + j - 1;
        float v = 0.f;
        if (pos >= 0 && pos < S.L)
            v = __half2float(mh[static_cast<size_t>(pos) * C + c]) + __half2float(ml[static_cast<size_t>(pos) * C + c]) * (1.f / kLoScale);
        means[j * CP + c] = v;
    }
    __syncthreads();
    const int tn = tid & 15, tp = tid >> 4;          // 4 outputs x 2 positions per thread
    float acc[2][4] = {};
    for (int kk = 0; kk < 3 * C; ++kk) {
        const int al = kk / C, cb = kk - al * C;
        const float4 w = __ldg(reinterpret_cast<const float4*>(S.wc + static_cast<size_t>(kk) * N4 + n0 + tn * 4));
        const float s0 = means[(tp * 2 + al) * CP + cb], s1 = means[(tp * 2 + al + 1) * CP + cb];
        acc[0][0] = fmaf(s0, w.x, acc[0][0]); acc[0][1] = fmaf(s0, w.y, acc[0][1]);
        acc[0][2] = fmaf(s0, w.z, acc[0][2]); acc[0][3] = fmaf(s0, w.w, acc[0][3]);
        acc[1][0] = fmaf(s1, w.x, acc[1][0]); acc[1][1] = fmaf(s1, w.y, acc[1][1]);
        acc[1][2] = fmaf(s1, w.z, acc[1][2]); acc[1][3] = fmaf(s1, w.w, acc[1][3]);
    }
    const int n = n0 + tn * 4, cls = n / Cout, co = n - cls * Cout;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int pos = p0 + tp * 2 + j;
        if (pos < S.L)
            *reinterpret_cast<float4*>(S.T + ((static_cast<size_t>(b) * 4 + cls) * S.L + pos) * Cout + co) =
                make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    }
}

__device__ __forceinline__ int edge_class(int i, int n) {
    return n == 1 ? 3 : (i == 0 ? 1 : (i == n - 1 ? 2 : 0));
}

// =====================================================================================
// CUDA-core 3x3 conv on the same operands as the tcgen05 kernel (bring-up / cross-check path,
// selected with conv_impl=1).  Uses the ORIGINAL fp32 weights, so it also checks the operand packing.
// grid (ceil(max_px/4), 3, B), block (64, 4)
// =====================================================================================
struct ConvEpi {
    TriCF bias;        // [Cout] (conv bias + skip-conv bias)
    TriCF Trow, Tcol;  // [B][4][rows|cols][Cout] or nullptr
    TriCF resid;       // fp32 [B][rows][cols][Cout] identity skip, or nullptr
    const float* embadd;   // additive embedding (use_scale_shift_norm = False): film base, or nullptr
    const int* film_row;
    int film_dim, film_off;
    TriF out;          // fp32 [B][rows][cols][Cout]
};

struct ConvFfmaArgs {
    TriCH a;       // [2][B][rows][cols][C]
    TriCH x16;     // [2][B][rows][cols][Cs] or nullptr
    TriDims d;
    int C, Cout, Cw, Cs;   // Cw = in-channels of the stored weight (3C with rollout, else C)
    TriCF w;       // original [Cout][Cw][3][3]
    TriCF wskip;   // original [Cout][Cs]
    ConvEpi e;
    int single;    // 1: ignore the lo halves (precision=1 emulation)
};

__global__ void __launch_bounds__(256) k_conv_ffma(ConvFfmaArgs A, int B) {
    pdl_wait();
    pdl_trigger();
    const int plane = blockIdx.y, b = blockIdx.z;
    const int rows = A.d.rows[plane], cols = A.d.cols[plane];
    const int npx = rows * cols;
    const int px = blockIdx.x * 4 + threadIdx.y;
    if (px >= npx) return;
    const int r = px / cols, c = px - r * cols;
    const size_t plane_elems = static_cast<size_t>(npx) * A.C;
    const __half* ah = A.a.p[plane] + static_cast<size_t>(b) * plane_elems;
    const __half* al = ah + static_cast<size_t>(B) * plane_elems;
    const float inv = A.single ? 0.f : 1.f / kLoScale;
    for (int co = threadIdx.x; co < A.Cout; co += 64) {
        float acc = A.e.bias.p[plane][co];
        const float* wp = A.w.p[plane] + static_cast<size_t>(co) * A.Cw * 9;
        for (int kh = 0; kh < 3; ++kh) {
            int rr = r + kh - 1;
            if (rr < 0 || rr >= rows) continue;
            for (int kw = 0; kw < 3; ++kw) {
                int cc = c + kw - 1;
                if (cc < 0 || cc >= cols) continue;
                const size_t off = (static_cast<size_t>(rr) * cols + cc) * A.C;
                for (int ci = 0; ci < A.C; ++ci) {
                    float av = __half2float(ah[off + ci]) + __half2float(al[off + ci]) * inv;
                    acc = fmaf(av, wp[(ci * 3 + kh) * 3 + kw], acc);
                }
            }
        }
        if (A.x16.p[plane]) {
            const size_t pe = static_cast<size_t>(npx) * A.Cs;
            const __half* xh = A.x16.p[plane] + static_cast<size_t>(b) * pe + static_cast<size_t>(px) * A.Cs;
            const __half* xl = xh + static_cast<size_t>(B) * pe;
            const float* ws = A.wskip.p[plane] + static_cast<size_t>(co) * A.Cs;
            for (int ci = 0; ci < A.Cs; ++ci)
                acc = fmaf(__half2float(xh[ci]) + __half2float(xl[ci]) * inv, ws[ci], acc);
        }
        if (A.e.Trow.p[plane]) {
            const size_t bo = static_cast<size_t>(b) * 4;
            acc += A.e.Trow.p[plane][((bo + edge_class(c, cols)) * rows + r) * A.Cout + co];
            acc += A.e.Tcol.p[plane][((bo + edge_class(r, rows)) * cols + c) * A.Cout + co];
        }
        if (A.e.embadd)
            acc += A.e.embadd[static_cast<size_t>(A.e.film_row ? A.e.film_row[b] : b) * A.e.film_dim + A.e.film_off + co];
        const size_t oo = (static_cast<size_t>(b) * npx + px) * A.Cout + co;
        if (A.e.resid.p[plane]) acc += A.e.resid.p[plane][oo];
        A.e.out.p[plane][oo] = acc;
    }
}

// =====================================================================================
// Timestep conditioning: sinusoid -> Linear -> SiLU -> Linear -> (SiLU -> every block's Linear).
// reference nn.py:103-121, unet_triplane.py:371-375, 232-238.   One warp per output scalar.
// =====================================================================================
__global__ void k_sinusoid(const float* __restrict__ t, const float* __restrict__ freqs, int half, float* __restrict__ out,
                           int n) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * half) return;
    const int r = i / half, k = i - r * half;
    const float ang = __fmul_rn(t[r], freqs[k]);
    out[static_cast<size_t>(r) * 2 * half + k] = cosf(ang);
    out[static_cast<size_t>(r) * 2 * half + half + k] = sinf(ang);
}

// y[r][n] = bias[n] + sum_k W[n][k] * f(x[r][k]),  f = SiLU if silu_in.  grid (ceil(N/8), rows), block 256
__global__ void __launch_bounds__(256) k_linear(const float* __restrict__ x, const float* __restrict__ W,
                                                const float* __restrict__ bias, float* __restrict__ y, int K, int N,
                                                int silu_in) {
    pdl_wait();
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + warp, r = blockIdx.y;
    if (n >= N) return;
    const float* xr = x + static_cast<size_t>(r) * K;
    const float* wr = W + static_cast<size_t>(n) * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) {
        float v = xr[k];
        if (silu_in) v = v / (1.f + expf(-v));
        acc = fmaf(v, wr[k], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[static_cast<size_t>(r) * N + n] = acc + bias[n];
}

// =====================================================================================
// Scheduler step: one vectorised pass over [B, n].  Op order and roundings mirror the reference's
// fp32 tensor expressions (no FMA contraction), so with equal inputs the result is bit-identical.
// reference gaussian_diffusion.py:294-315 (x0), :209-220 (posterior mean), :431-439 (DDPM sample),
// :567-599 (DDIM), :626-636 (DDIM reverse).  Last CTA decrements t_idx when `advance` is set.
// =====================================================================================
struct SchedArgs {
    int kind, mean_type, clip, is_mask_t0;
    int B;
    long long n;            // per sample = C * hw
    int C;                  // channels
    long long hw;           // (H+D)*(W+D)
    const float* model_out;
    const float* x;
    const float* noise;     // nullptr -> philox
    const float* y0;
    const float* mask;
    float* sample;
    float* x0_out;
    const float* coef;      // [T][12]
    int* t_idx;             // [B]
    unsigned long long seed;
    unsigned int sample_base;
    const unsigned long long* dyn;   // loop mode: {seed, sample_base} in device memory (the captured graph does not bake them)
    int advance;            // loop mode: t_idx[b] -= 1 after the step (by the last CTA)
    unsigned int* ticket;
    long long noise_step_stride;   // loop mode with a noise buffer: noise + t*stride
    Trace tr;               // slots: 0 entry, 1 element loop done
};

__device__ __forceinline__ float sched_one(const SchedArgs& A, const float* cf, float nz, float out, float x, float nzv,
                                           float y0, float mk, float& x0r) {
    float x0;
    if (A.mean_type == 0) {
        x0 = out;
    } else {
        x0 = __fsub_rn(__fmul_rn(cf[0], x), __fmul_rn(cf[1], out));
    }
    if (A.clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    float res;
    if (A.kind == 0) {
        float mean = __fadd_rn(__fmul_rn(cf[2], x0), __fmul_rn(cf[3], x));
        res = __fadd_rn(mean, __fmul_rn(__fmul_rn(nz, cf[4]), nzv));
    } else {
        if (A.y0) {
            float mix = __fadd_rn(__fmul_rn(mk, y0), __fmul_rn(__fsub_rn(1.f, mk), x0));
            if (A.is_mask_t0) x0 = mix;
            else x0 = __fadd_rn(__fmul_rn(mix, nz), __fmul_rn(x0, __fsub_rn(1.f, nz)));
        }
        float eps = __fdiv_rn(__fsub_rn(__fmul_rn(cf[0], x), x0), cf[1]);
        if (A.kind == 1) {
            float mean = __fadd_rn(__fmul_rn(x0, cf[5]), __fmul_rn(cf[6], eps));
            res = __fadd_rn(mean, __fmul_rn(__fmul_rn(nz, cf[7]), nzv));
        } else {
            res = __fadd_rn(__fmul_rn(x0, cf[8]), __fmul_rn(cf[9], eps));
        }
    }
    x0r = x0;
    return res;
}

// Noise layout (shared with k_boundary<MODE_FUSED>, k_philox_normal and oracle/philox_ref.py): the N(0,1) of element
// (c, pixel) is component c & 3 of the Philox block with counter (pixel * ceil(C/4) + c / 4, step, global sample index), so
// one block serves the four channels c..c+3 of a pixel — the unit both kernels naturally own.
// One thread = one (pixel, channel quad).  grid (ceil(hw*nq/256) capped, B)
__global__ void __launch_bounds__(256) k_sched_step(SchedArgs A) {
    pdl_wait();
    pdl_trigger();
    __shared__ bool is_last;
    if (threadIdx.x == 0) trace_mark(A.tr, 0);
    const int b = blockIdx.y;
    const int t = A.t_idx[b];
    float cf[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) cf[k] = __ldg(A.coef + static_cast<size_t>(t) * 12 + k);
    const float nz = t != 0 ? 1.f : 0.f;
    const size_t base = static_cast<size_t>(b) * A.n;
    const unsigned long long seed = A.dyn ? __ldg(A.dyn) : A.seed;
    const unsigned int sample_base = A.dyn ? static_cast<unsigned int>(__ldg(A.dyn + 1)) : A.sample_base;
    const float* noise = A.noise ? A.noise + static_cast<size_t>(t) * A.noise_step_stride + base : nullptr;
    const int nq = (A.C + 3) / 4;
    const long long items = A.hw * nq;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < items;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int quad = static_cast<int>(i / A.hw);           // pixel fastest: adjacent threads -> adjacent addresses
        const long long pix = i - quad * A.hw;
        const int c0 = quad * 4, cnt = min(4, A.C - c0);
        float o[4], xv[4], nv[4] = {0, 0, 0, 0}, yv[4] = {0, 0, 0, 0}, mv[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const size_t e = static_cast<size_t>(c0 + k) * A.hw + pix;
            o[k] = k < cnt ? A.model_out[base + e] : 0.f;
            xv[k] = k < cnt ? A.x[base + e] : 0.f;
            if (noise && k < cnt) nv[k] = noise[e];
            if (A.y0 && k < cnt) {
                yv[k] = A.y0[base + e];
                mv[k] = A.mask[base + e];
            }
        }
        if (!noise && A.kind != 2) {
            const float4 z = philox_normal4(seed, sample_base + b, static_cast<uint32_t>(t), static_cast<uint32_t>(pix * nq + quad));
            nv[0] = z.x; nv[1] = z.y; nv[2] = z.z; nv[3] = z.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < cnt) {
                float x0;
                const float rs = sched_one(A, cf, nz, o[k], xv[k], nv[k], yv[k], mv[k], x0);
                const size_t e = static_cast<size_t>(c0 + k) * A.hw + pix;
                A.sample[base + e] = rs;
                if (A.x0_out) A.x0_out[base + e] = x0;
            }
        }
    }
    if (threadIdx.x == 0) trace_mark(A.tr, 1);
    if (!A.advance) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev = atomicAdd(A.ticket, 1u);
        is_last = prev == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        for (int k = 0; k < A.B; ++k) A.t_idx[k] -= 1;
        *A.ticket = 0u;
    }
}

// float4 path when a sample's element count is a multiple of 4 (every sample then starts 16-byte aligned), scalar otherwise
__global__ void __launch_bounds__(256) k_q_sample(const float* __restrict__ x0, const float* __restrict__ noise,
                                                  float* __restrict__ out, const float* __restrict__ coef,
                                                  const int* __restrict__ t_idx, long long n) {
    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.y, t = t_idx[b];
    const float a = coef[static_cast<size_t>(t) * 12 + 10], s = coef[static_cast<size_t>(t) * 12 + 11];
    const size_t base = static_cast<size_t>(b) * n;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x, i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if ((n & 3) == 0) {
        const float4* x4 = reinterpret_cast<const float4*>(x0 + base);
        const float4* z4 = reinterpret_cast<const float4*>(noise + base);
        float4* o4 = reinterpret_cast<float4*>(out + base);
        for (long long i = i0; i < (n >> 2); i += stride) {
            const float4 x = __ldg(x4 + i), z = __ldg(z4 + i);
            o4[i] = make_float4(__fadd_rn(__fmul_rn(a, x.x), __fmul_rn(s, z.x)), __fadd_rn(__fmul_rn(a, x.y), __fmul_rn(s, z.y)),
                                __fadd_rn(__fmul_rn(a, x.z), __fmul_rn(s, z.z)), __fadd_rn(__fmul_rn(a, x.w), __fmul_rn(s, z.w)));
        }
        return;
    }
    for (long long i = i0; i < n; i += stride)
        out[base + i] = __fadd_rn(__fmul_rn(a, x0[base + i]), __fmul_rn(s, noise[base + i]));
}

// ---------------------------------------------------------------- variational-bound terms (bits per dimension)
//   reference: GaussianDiffusion._vb_terms_bpd, gaussian_diffusion.py:736-769 (+ the two MSEs calc_bpd_loop adds, :912-916),
//              normal_kl / discretized_gaussian_log_likelihood, src/diffusion/losses.py:12-77
// One pass over (x_start, x_t, model_out[, noise]): per element the posterior KL or — at t == 0 — the discretised decoder NLL,
// (pred_xstart - x_start)^2 and (eps - noise)^2, reduced per sample.  Block sums are fp64 and are added by k_vb_finalize in
// block order, so the result does not depend on scheduling.  grid (gx, B)
struct VbArgs {
    int mean_type, clip, B;
    long long n;
    const float *x_start, *x_t, *model_out, *noise;
    float* x0_out;              // pred_xstart (nullable)
    const float* coef;          // [T][12]
    const float* logvar;        // [T][2]: posterior_log_variance_clipped, model log-variance
    const int* t_idx;
    double* partial;            // [B][gx][3]
    float* out;                 // [B][3]: vb term (bits), mean (pred_xstart - x_start)^2, mean (eps - noise)^2
};
__device__ __forceinline__ float vb_cdf(float v) {       // losses.py:44-49
    return 0.5f * (1.0f + tanhf(0.7978845608028654f * (v + 0.044715f * (v * v * v))));
}
__device__ __forceinline__ void vb_one(const VbArgs& A, const float (&cf)[4], int t, float lv1, float lv2, float e12, float einv2, float inv_stdv,
                                       float xs, float xt, float mo, float nzv, bool has_noise, float& x0_out, double (&acc)[3]) {
    float x0 = A.mean_type == 0 ? mo : __fsub_rn(__fmul_rn(cf[0], xt), __fmul_rn(cf[1], mo));
    if (A.clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    x0_out = x0;
    const float true_mean = __fadd_rn(__fmul_rn(cf[2], xs), __fmul_rn(cf[3], xt));
    const float mean = __fadd_rn(__fmul_rn(cf[2], x0), __fmul_rn(cf[3], xt));
    float term;
    if (t != 0) {
        const float d = true_mean - mean;
        term = 0.5f * (-1.0f + lv2 - lv1 + e12 + (d * d) * einv2);
    } else {
        const float c = xs - mean;
        const float cdf_plus = vb_cdf(inv_stdv * (c + 1.0f / 255.0f)), cdf_min = vb_cdf(inv_stdv * (c - 1.0f / 255.0f));
        float lp;
        if (xs < -0.999f) lp = logf(fmaxf(cdf_plus, 1e-12f));
        else if (xs > 0.999f) lp = logf(fmaxf(1.0f - cdf_min, 1e-12f));
        else lp = logf(fmaxf(cdf_plus - cdf_min, 1e-12f));
        term = -lp;
    }
    acc[0] += static_cast<double>(term);
    const float dx = x0 - xs;
    acc[1] += static_cast<double>(dx * dx);
    if (has_noise) {
        const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(cf[0], xt), x0), cf[1]);
        const float de = eps - nzv;
        acc[2] += static_cast<double>(de * de);
    }
}
// Sums are fp64 from the first element on (fixed order above the thread level): 1e-7 parity with the reference's mean_flat.
__global__ void __launch_bounds__(256) k_vb_terms(VbArgs A) {
    const int b = blockIdx.y, t = A.t_idx[b];
    float cf[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cf[k] = __ldg(A.coef + static_cast<size_t>(t) * 12 + k);
    const float lv1 = __ldg(A.logvar + 2 * t), lv2 = __ldg(A.logvar + 2 * t + 1);
    const float e12 = expf(lv1 - lv2), einv2 = expf(-lv2), inv_stdv = expf(-(0.5f * lv2));
    const size_t base = static_cast<size_t>(b) * A.n;
    const bool has_noise = A.noise != nullptr;
    double acc[3] = {0.0, 0.0, 0.0};
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x, i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if ((A.n & 3) == 0) {
        const float4* xs4 = reinterpret_cast<const float4*>(A.x_start + base);
        const float4* xt4 = reinterpret_cast<const float4*>(A.x_t + base);
        const float4* mo4 = reinterpret_cast<const float4*>(A.model_out + base);
        const float4* nz4 = has_noise ? reinterpret_cast<const float4*>(A.noise + base) : nullptr;
        float4* o4 = A.x0_out ? reinterpret_cast<float4*>(A.x0_out + base) : nullptr;
        for (long long i = i0; i < (A.n >> 2); i += stride) {
            const float4 xs = __ldg(xs4 + i), xt = __ldg(xt4 + i), mo = __ldg(mo4 + i);
            const float4 nz = has_noise ? __ldg(nz4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 o;
            vb_one(A, cf, t, lv1, lv2, e12, einv2, inv_stdv, xs.x, xt.x, mo.x, nz.x, has_noise, o.x, acc);
            vb_one(A, cf, t, lv1, lv2, e12, einv2, inv_stdv, xs.y, xt.y, mo.y, nz.y, has_noise, o.y, acc);
            vb_one(A, cf, t, lv1, lv2, e12, einv2, inv_stdv, xs.z, xt.z, mo.z, nz.z, has_noise, o.z, acc);
            vb_one(A, cf, t, lv1, lv2, e12, einv2, inv_stdv, xs.w, xt.w, mo.w, nz.w, has_noise, o.w, acc);
            if (o4) o4[i] = o;
        }
    } else {
        for (long long i = i0; i < A.n; i += stride) {
            float o;
            vb_one(A, cf, t, lv1, lv2, e12, einv2, inv_stdv, A.x_start[base + i], A.x_t[base + i], A.model_out[base + i],
                   has_noise ? A.noise[base + i] : 0.f, has_noise, o, acc);
            if (A.x0_out) A.x0_out[base + i] = o;
        }
    }
    __shared__ double red[3][8];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        A.partial[(static_cast<size_t>(b) * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = s;
    }
}
// grid B, block 32: lane k < 3 adds the gx block sums of term k in order
// block sums of sample b, term k: lane l adds blocks l, l + 32, ... in order, then a fixed butterfly (same result on every run)
__device__ __forceinline__ double partial_sum(const double* partial, int b, int gx, int k) {
    double s = 0.0;
    for (int j = threadIdx.x & 31; j < gx; j += 32) s += partial[(static_cast<size_t>(b) * gx + j) * 3 + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}
__global__ void k_vb_finalize(VbArgs A, int gx) {       // grid B, block 96: warp k reduces term k
    const int b = blockIdx.x, k = threadIdx.x >> 5;
    const double s = partial_sum(A.partial, b, gx, k);
    if ((threadIdx.x & 31) != 0) return;
    double m = s / static_cast<double>(A.n);
    if (k == 0) m /= 0.6931471805599453;        // nats -> bits (np.log(2.0))
    A.out[b * 3 + k] = static_cast<float>(m);
}

// ---------------------------------------------------------------- per-plane MSE of the training objective
//   reference: GaussianDiffusion.training_losses, gaussian_diffusion.py:822-851: decompose_featmaps(target / output) and
//   mean_flat((target - output)^2) per plane (xy, xz, yz).  One pass over the two composed tensors [B, C, H+D, W+D]; the plane of
//   an element follows from its (row, col) (the D x D corner belongs to none); fp64 block sums, added in block order by
//   k_plane_mse_finalize.  grid (gx, B)
struct PlaneMseArgs {
    const float *target, *output;
    int C, H, W, D;
    long long n;               // C * (H+D) * (W+D)
    double* partial;           // [B][gx][3]
    float* out;                // [B][3]
};
__global__ void __launch_bounds__(256) k_plane_mse(PlaneMseArgs A) {
    const int b = blockIdx.y;
    const size_t base = static_cast<size_t>(b) * A.n;
    const int Wc = A.W + A.D, hw = (A.H + A.D) * Wc;
    double acc[3] = {0.0, 0.0, 0.0};
    float q[3] = {0.f, 0.f, 0.f};                            // partial sums of one quad (fp32), folded into fp64 per quad
    auto add = [&](int r, int c, float tg, float ou) {       // the plane of composed pixel (r, c); the D x D corner belongs to none
        const float d = tg - ou, d2 = d * d;
        if (r < A.H) q[c < A.W ? 0 : 1] += d2;
        else if (c < A.W) q[2] += d2;
    };
    auto fold = [&]() {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            acc[k] += static_cast<double>(q[k]);
            q[k] = 0.f;
        }
    };
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x, i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if ((A.n & 3) == 0 && A.n < (1LL << 31)) {
        // Quads with 32-bit index arithmetic (the kernel is issue bound: the 64-bit modulo per quad used to dominate it).  A quad
        // that stays inside one row and on one side of column W lies in ONE plane: four squared differences, one predicated add.
        const float4* t4 = reinterpret_cast<const float4*>(A.target + base);
        const float4* o4 = reinterpret_cast<const float4*>(A.output + base);
        const unsigned nq = static_cast<unsigned>(A.n >> 2), uhw = static_cast<unsigned>(hw), uWc = static_cast<unsigned>(Wc);
        const unsigned ustride = static_cast<unsigned>(stride);
        for (unsigned i = static_cast<unsigned>(i0); i < nq; i += ustride) {
            const float4 tg = __ldg(t4 + i), ou = __ldg(o4 + i);
            const unsigned pix = (i << 2) % uhw;
            int r = static_cast<int>(pix / uWc), c = static_cast<int>(pix - static_cast<unsigned>(r) * uWc);
            const float dx = tg.x - ou.x, dy = tg.y - ou.y, dz = tg.z - ou.z, dw = tg.w - ou.w;
            if (c + 3 < Wc && (c + 3 < A.W || c >= A.W)) {
                const float s4 = (dx * dx + dy * dy) + (dz * dz + dw * dw);
                const bool top = r < A.H, left = c < A.W;
                q[0] += (top && left) ? s4 : 0.f;
                q[1] += (top && !left) ? s4 : 0.f;
                q[2] += (!top && left) ? s4 : 0.f;
            } else {
                const float dv[4] = {dx, dy, dz, dw};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float d2 = dv[k] * dv[k];
                    const bool top = r < A.H, left = c < A.W;
                    q[0] += (top && left) ? d2 : 0.f;
                    q[1] += (top && !left) ? d2 : 0.f;
                    q[2] += (!top && left) ? d2 : 0.f;
                    if (++c == Wc) {
                        c = 0;
                        if (++r == A.H + A.D) r = 0;                  // next channel
                    }
                }
            }
            fold();
        }
    } else {
        for (long long i = i0; i < A.n; i += stride) {
            const int pix = static_cast<int>(i % hw), r = pix / Wc, c = pix - r * Wc;
            add(r, c, A.target[base + i], A.output[base + i]);
            fold();
        }
    }
    __shared__ double red[3][8];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        A.partial[(static_cast<size_t>(b) * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = t;
    }
}
__global__ void k_plane_mse_finalize(PlaneMseArgs A, int gx) {       // grid B, block 96: warp k reduces plane k
    const int b = blockIdx.x, k = threadIdx.x >> 5;
    const double t = partial_sum(A.partial, b, gx, k);
    if ((threadIdx.x & 31) != 0) return;
    const double cnt = static_cast<double>(A.C) * (k == 0 ? A.H * A.W : (k == 1 ? A.H * A.D : A.W * A.D));
    A.out[b * 3 + k] = static_cast<float>(t / cnt);
}

// ---------------------------------------------------------------- fused AdamW + EMA step over a flat parameter buffer
//   reference: TrainLoop.run_step -> torch.optim.AdamW(lr, weight_decay).step() (train_util.py:82-84, 160-167) followed by
//   update_ema for every EMA rate (nn.py:53-63: targ = targ * rate + src * (1 - rate)).
// One pass: 5 + n_ema reads and 3 + n_ema writes of 4 bytes per parameter (HBM bound) instead of ~10 element-wise launches
// per tensor x 138 tensors.  Arithmetic follows torch's single-tensor AdamW: decoupled decay, lerp, addcmul, sqrt / bias
// corrections (host fp64 -> fp32 scalars), addcdiv.
struct AdamWArgs {
    float* p;
    const float* g;
    float *m, *v;
    float* ema[4];
    float ema_rate[4];
    int n_ema;
    long long n;
    float decay;          // 1 - lr * weight_decay
    float w1;             // 1 - beta1 (lerp weight)
    float beta2, w2;      // beta2, 1 - beta2
    float bc2_sqrt;       // sqrt(1 - beta2^step)
    float step_size;      // lr / (1 - beta1^step)
    float eps;
};
__device__ __forceinline__ void adamw_one(const AdamWArgs& A, float& p, float g, float& m, float& v) {
    p = p * A.decay;
    m = m + A.w1 * (g - m);                                   // lerp, weight < 0.5
    v = fmaf(A.w2 * g, g, v * A.beta2);
    const float denom = sqrtf(v) / A.bc2_sqrt + A.eps;
    p = p - A.step_size * (m / denom);
}
__global__ void __launch_bounds__(256) k_adamw_ema(AdamWArgs A) {
    const long long n4 = A.n >> 2;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
        float4 p = reinterpret_cast<float4*>(A.p)[i], m = reinterpret_cast<float4*>(A.m)[i], v = reinterpret_cast<float4*>(A.v)[i];
        const float4 g = __ldg(reinterpret_cast<const float4*>(A.g) + i);
        adamw_one(A, p.x, g.x, m.x, v.x);
        adamw_one(A, p.y, g.y, m.y, v.y);
        adamw_one(A, p.z, g.z, m.z, v.z);
        adamw_one(A, p.w, g.w, m.w, v.w);
        reinterpret_cast<float4*>(A.p)[i] = p;
        reinterpret_cast<float4*>(A.m)[i] = m;
        reinterpret_cast<float4*>(A.v)[i] = v;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < A.n_ema) {
                float4 e = reinterpret_cast<float4*>(A.ema[k])[i];
                const float r = A.ema_rate[k], a = 1.f - r;
                e.x = fmaf(p.x, a, e.x * r); e.y = fmaf(p.y, a, e.y * r); e.z = fmaf(p.z, a, e.z * r); e.w = fmaf(p.w, a, e.w * r);
                reinterpret_cast<float4*>(A.ema[k])[i] = e;
            }
        }
    }
    // scalar tail (n not a multiple of 4)
    for (long long i = (n4 << 2) + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < A.n; i += stride) {
        float p = A.p[i], m = A.m[i], v = A.v[i];
        adamw_one(A, p, A.g[i], m, v);
        A.p[i] = p; A.m[i] = m; A.v[i] = v;
        for (int k = 0; k < A.n_ema; ++k) A.ema[k][i] = fmaf(p, 1.f - A.ema_rate[k], A.ema[k][i] * A.ema_rate[k]);
    }
}

__global__ void __launch_bounds__(256) k_philox_normal(float* __restrict__ out, int C, long long hw, unsigned long long seed,
                                                       unsigned int sample_base, unsigned int step) {
    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.y;
    const int nq = (C + 3) / 4;
    const long long items = hw * nq;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < items;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int quad = static_cast<int>(i / hw);
        const long long pix = i - quad * hw;
        const float4 z = philox_normal4(seed, sample_base + b, step, static_cast<uint32_t>(pix * nq + quad));
        const float zz[4] = {z.x, z.y, z.z, z.w};
        for (int k = 0; k < 4 && quad * 4 + k < C; ++k)
            out[(static_cast<size_t>(b) * C + quad * 4 + k) * hw + pix] = zz[k];
    }
}

}  // namespace s3d
