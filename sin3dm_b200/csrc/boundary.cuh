// Step boundary of the sampling loop: the UNet's output head, the scheduler update and the next step's input conv are all
// per-pixel maps over the composed triplane, so they share one tiling and can run as ONE kernel.
//   head   : GroupNorm + SiLU + 1x1 conv C0 -> Cf, composed NCHW output        reference unet_triplane.py:441-445
//   sched  : DDPM / DDIM update of x with the model output of this pixel        reference gaussian_diffusion.py:294-315,
//                                                                               431-439, 567-599 (see k_sched_step)
//   in_conv: 1x1 conv Cf -> C0 off the composed tensor + GroupNorm partials      reference unet_triplane.py:378
// MODE_HEAD / MODE_INCONV are the two ends of a stand-alone forward; MODE_FUSED is head -> sched -> in_conv of the next
// step (same device functions, so both paths produce identical values).
//
// Mapping.  A CTA owns a 128-pixel tile of one plane: 4 x 32 pixels with the 32 along the axis that is contiguous in the
// composed tensor (plane columns for xy / xz and the dead corner, plane ROWS for yz, which is stored transposed,
// triplane_util.py:14,24), so composed loads / stores are 128-byte segments while the NHWC side moves whole 256 / 512-byte
// pixels.  Composed channels are staged through shared memory ([channel][pixel] tiles).  A thread is (pixel, channel quad):
// NQ = C0/4 threads per pixel, each with its 4 hidden channels' 1x1 weights in registers (no shared-memory weight traffic).
// The head's Cf dot products are reduced across a pixel's NQ lanes with a halving butterfly (16 shuffles, fixed order).
// grid (tiles of the 3 planes [+ dead-corner tiles], B), block 256.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace s3d {

enum { MODE_HEAD = 0, MODE_INCONV = 1, MODE_FUSED = 2 };
constexpr int kMaxCf = 16;
constexpr int kBndPx = 128, kBndFast = 32, kBndSlow = 4;

struct BoundaryArgs {
    TriCF h;              // last activation [B][rows][cols][C0]                  (HEAD / FUSED)
    TriDims d;
    int C0, Cf, H, W, Dd;
    int tile_start[5];    // prefix sums of the tiles of xy, xz, yz, dead corner
    int tiles_fast[4];    // tiles along the fast axis of each region
    StatsSrc st;          // statistics of h + out-norm parameters
    TriCF w_out, b_out;   // [Cf][C0], [Cf]
    float* model_out;     // composed [B][Cf][H+D][W+D]                           (HEAD)
    const float* x_in;    // composed input                                        (INCONV)
    TriCF w_in, b_in;     // [C0][Cf], [C0]
    TriF h0;              // in_conv output [B][rows][cols][C0]                    (INCONV / FUSED)
    StatsSink sink;       // GroupNorm group sums of h0
    SchedArgs sch;        // x, sample (in place), coefficient table, step index   (FUSED)
    Trace tr;             // slots: 0 entry, 1 head done, 2 scheduler done, 3 in_conv done
};

// Sum v[0..15] over the NQ lanes of a pixel; afterwards output `co` sits in v[0] of the lane the function returns true for.
template <int NQ>
__device__ __forceinline__ bool reduce_scatter16(float (&v)[16], int ql, int& co) {
    constexpr unsigned kFull = 0xffffffffu;
#define S3D_HALVE(MASK, N)                                        \
    {                                                             \
        const bool up = (ql & (MASK)) != 0;                       \
        _Pragma("unroll") for (int j = 0; j < (N); ++j) {         \
            const float send = up ? v[j] : v[j + (N)];            \
            const float keep = up ? v[j + (N)] : v[j];            \
            v[j] = keep + __shfl_xor_sync(kFull, send, (MASK));   \
        }                                                         \
    }
    if (NQ == 16) {
        S3D_HALVE(8, 8) S3D_HALVE(4, 4) S3D_HALVE(2, 2) S3D_HALVE(1, 1)
        co = ql;
        return true;
    } else {
        S3D_HALVE(16, 8) S3D_HALVE(8, 4) S3D_HALVE(4, 2) S3D_HALVE(2, 1)
        v[0] += __shfl_xor_sync(kFull, v[0], 1);
        co = ql >> 1;
        return (ql & 1) == 0;
    }
#undef S3D_HALVE
}

template <int MODE, int NQ>      // NQ = C0 / 4 lanes per pixel (16: C0 = 64, 32: C0 = 128)
__global__ void __launch_bounds__(256, 2) k_boundary(BoundaryArgs A) {
    constexpr int PPP = 256 / NQ;               // pixels per pass
    constexpr int NPASS = kBndPx / PPP;
    constexpr int C0 = NQ * 4;
    pdl_wait();
    pdl_trigger();
    __shared__ __align__(16) float xs[kMaxCf][kBndPx];     // composed input / updated x of the tile, [channel][pixel]
    __shared__ __align__(16) float os[kMaxCf][kBndPx];     // model output of the tile
    __shared__ __align__(16) float coefA[C0], coefB[C0];
    __shared__ __align__(16) float red[8][2][C0];          // per-warp (sum, sum-sq) per channel
    __shared__ double fin[64];
    __shared__ bool is_last;
    const int tid = threadIdx.x, b = blockIdx.y;
    if (tid == 0) trace_mark(A.tr, 0);
    const int Cf = A.Cf, nq = (Cf + 3) / 4;
    const int Hc = A.H + A.Dd, Wc = A.W + A.Dd;
    const long long hw = static_cast<long long>(Hc) * Wc;
    const long long nper = static_cast<long long>(Cf) * hw;
    // ---- tile of this CTA
    const int t = blockIdx.x;
    const int plane = t >= A.tile_start[3] ? 3 : (t >= A.tile_start[2] ? 2 : (t >= A.tile_start[1] ? 1 : 0));
    const int ip = t - A.tile_start[plane];
    const int rows = plane < 3 ? A.d.rows[plane] : A.Dd, cols = plane < 3 ? A.d.cols[plane] : A.Dd;
    const bool fast_rows = plane == 2;
    const int tf = ip % A.tiles_fast[plane], ts = ip / A.tiles_fast[plane];
    const int r0 = fast_rows ? tf * kBndFast : ts * kBndSlow, c0 = fast_rows ? ts * kBndSlow : tf * kBndFast;
    // pixel i of the tile -> plane coordinates, validity, composed offset
    auto pix_r = [&](int i) { return r0 + (fast_rows ? (i & (kBndFast - 1)) : (i >> 5)); };
    auto pix_c = [&](int i) { return c0 + (fast_rows ? (i >> 5) : (i & (kBndFast - 1))); };
    auto comp = [&](int r, int c) -> long long {
        return plane == 3 ? static_cast<long long>(A.H + r) * Wc + A.W + c : composed_offset(plane, r, c, A.H, A.W, Wc);
    };
    const int ql = tid & (NQ - 1), pslot = tid / NQ;
    const size_t npx = static_cast<size_t>(rows) * cols;
    for (int e = Cf * kBndPx + tid; e < kMaxCf * kBndPx; e += 256) xs[e >> 7][e & (kBndPx - 1)] = 0.f;   // unused channel rows

    // scheduler scalars of this sample (FUSED)
    int tstep = 0;
    float cf[12];
    float nz = 0.f;
    unsigned long long seed = A.sch.seed;
    unsigned int sample_base = A.sch.sample_base;
    if (MODE == MODE_FUSED) {
        tstep = A.sch.t_idx[b];
        if (A.sch.dyn) {
            seed = __ldg(A.sch.dyn);
            sample_base = static_cast<unsigned int>(__ldg(A.sch.dyn + 1));
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) cf[k] = __ldg(A.sch.coef + static_cast<size_t>(tstep) * 12 + k);
        nz = tstep != 0 ? 1.f : 0.f;
    }

    // ================= head: os[co][i] = b_out[co] + sum_c w_out[co][c] * silu(GN(h))[i][c] =================
    if (MODE != MODE_INCONV && plane < 3) {
        const float* hp = A.h.p[plane] + static_cast<size_t>(b) * npx * C0 + ql * 4;
        // activation loads of the first passes are in flight while the statistics are reduced
        constexpr int NB = NPASS < 4 ? NPASS : 4;
        float4 hv[NB];
        auto load_h = [&](int pass) {
            const int i = pass * PPP + pslot, r = pix_r(i), c = pix_c(i);
            return (r < rows && c < cols) ? __ldg(reinterpret_cast<const float4*>(hp + (static_cast<size_t>(r) * cols + c) * C0))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        };
#pragma unroll
        for (int k = 0; k < NB; ++k) hv[k] = load_h(k);
        float4 wq[kMaxCf];
#pragma unroll
        for (int co = 0; co < kMaxCf; ++co)
            wq[co] = co < Cf ? __ldg(reinterpret_cast<const float4*>(A.w_out.p[plane] + static_cast<size_t>(co) * C0 + ql * 4))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        stats_coef_prologue(A.st, b, plane, C0, static_cast<double>(npx) * (C0 / kGroups), tid, 256, fin, coefA, coefB);
        const float4 ca = *reinterpret_cast<const float4*>(coefA + ql * 4);
        const float4 cb = *reinterpret_cast<const float4*>(coefB + ql * 4);
#pragma unroll
        for (int p0 = 0; p0 < NPASS; p0 += NB) {
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const int pass = p0 + k;
                float4 y;
                y.x = silu_f(fmaf(hv[k].x, ca.x, cb.x));
                y.y = silu_f(fmaf(hv[k].y, ca.y, cb.y));
                y.z = silu_f(fmaf(hv[k].z, ca.z, cb.z));
                y.w = silu_f(fmaf(hv[k].w, ca.w, cb.w));
                if (pass + NB < NPASS) hv[k] = load_h(pass + NB);       // next batch, behind this one's math
                float v[16];
#pragma unroll
                for (int co = 0; co < kMaxCf; ++co)
                    v[co] = fmaf(y.x, wq[co].x, fmaf(y.y, wq[co].y, fmaf(y.z, wq[co].z, y.w * wq[co].w)));
                int co;
                const bool writer = reduce_scatter16<NQ>(v, ql, co);
                if (writer && co < Cf) os[co][pass * PPP + pslot] = v[0] + __ldg(A.b_out.p[plane] + co);
            }
        }
        __syncthreads();
    }
    if (tid == 0) trace_mark(A.tr, 1);
    if (MODE == MODE_HEAD) {
        for (int e = tid; e < Cf * kBndPx; e += 256) {
            const int co = e >> 7, i = e & (kBndPx - 1), r = pix_r(i), c = pix_c(i);
            if (r < rows && c < cols)
                A.model_out[static_cast<size_t>(b) * nper + static_cast<long long>(co) * hw + comp(r, c)] = plane < 3 ? os[co][i] : 0.f;
        }
        return;
    }

    // ================= scheduler: x <- step(x, model output), one (pixel, channel quad) per item =================
    if (MODE == MODE_FUSED) {
        for (int e = tid; e < nq * kBndPx; e += 256) {
            const int quad = e >> 7, i = e & (kBndPx - 1), r = pix_r(i), c = pix_c(i);
            const int ch0 = quad * 4, cnt = min(4, Cf - ch0);
            float xn[4] = {0.f, 0.f, 0.f, 0.f};
            if (r < rows && c < cols) {
                const long long pix = comp(r, c);
                float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!A.sch.noise && A.sch.kind != 2)
                    z = philox_normal4(seed, sample_base + b, static_cast<uint32_t>(tstep), static_cast<uint32_t>(pix * nq + quad));
                const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < cnt) {
                        const size_t gi = static_cast<size_t>(b) * nper + static_cast<long long>(ch0 + k) * hw + pix;
                        const float mo = plane < 3 ? os[ch0 + k][i] : 0.f;     // the network never sees the dead corner
                        const float nv = A.sch.noise ? A.sch.noise[static_cast<size_t>(tstep) * A.sch.noise_step_stride + gi] : zz[k];
                        const float y0 = A.sch.y0 ? A.sch.y0[gi] : 0.f, mk = A.sch.y0 ? A.sch.mask[gi] : 0.f;
                        float x0;
                        xn[k] = sched_one(A.sch, cf, nz, mo, A.sch.x[gi], nv, y0, mk, x0);
                        A.sch.sample[gi] = xn[k];
                        if (A.sch.x0_out) A.sch.x0_out[gi] = x0;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < cnt) xs[ch0 + k][i] = xn[k];
        }
    } else if (plane < 3) {
        for (int e = tid; e < Cf * kBndPx; e += 256) {
            const int ch = e >> 7, i = e & (kBndPx - 1), r = pix_r(i), c = pix_c(i);
            xs[ch][i] = (r < rows && c < cols) ? __ldg(A.x_in + static_cast<size_t>(b) * nper + static_cast<long long>(ch) * hw + comp(r, c)) : 0.f;
        }
    }
    if (tid == 0) trace_mark(A.tr, 2);

    // ================= in_conv: h0[i][co] = b_in[co] + sum_c w_in[co][c] * x[c][i], + GroupNorm partials =================
    if (plane < 3) {
        float4 wi[kMaxCf];
#pragma unroll
        for (int ch = 0; ch < kMaxCf; ++ch) {
            wi[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ch < Cf) {
                const float* wp = A.w_in.p[plane] + static_cast<size_t>(ql * 4) * Cf + ch;
                wi[ch] = make_float4(__ldg(wp), __ldg(wp + Cf), __ldg(wp + 2 * Cf), __ldg(wp + 3 * Cf));
            }
        }
        const float4 bi = __ldg(reinterpret_cast<const float4*>(A.b_in.p[plane] + ql * 4));
        float* h0p = A.h0.p[plane] + static_cast<size_t>(b) * npx * C0 + ql * 4;
        __syncthreads();          // xs complete
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
#pragma unroll 4
        for (int pass = 0; pass < NPASS; ++pass) {
            const int i = pass * PPP + pslot, r = pix_r(i), c = pix_c(i);
            float4 a = bi;
            float xv[kMaxCf];
#pragma unroll
            for (int ch = 0; ch < kMaxCf; ++ch) xv[ch] = xs[ch][i];       // rows >= Cf are zero (and so are their weights)
#pragma unroll
            for (int ch = 0; ch < kMaxCf; ++ch) {
                a.x = fmaf(xv[ch], wi[ch].x, a.x); a.y = fmaf(xv[ch], wi[ch].y, a.y);
                a.z = fmaf(xv[ch], wi[ch].z, a.z); a.w = fmaf(xv[ch], wi[ch].w, a.w);
            }
            if (r < rows && c < cols) {
                *reinterpret_cast<float4*>(h0p + (static_cast<size_t>(r) * cols + c) * C0) = a;
                acc_sq(s, q, a);
            }
        }
        if (A.sink.acc) {
            // (sum, sum-sq) per channel over the tile: lanes holding the same channel quad inside a warp first, then the 8 warps
            constexpr unsigned kFull = 0xffffffffu;
            if (NQ == 16) {
                s.x += __shfl_xor_sync(kFull, s.x, 16); s.y += __shfl_xor_sync(kFull, s.y, 16);
                s.z += __shfl_xor_sync(kFull, s.z, 16); s.w += __shfl_xor_sync(kFull, s.w, 16);
                q.x += __shfl_xor_sync(kFull, q.x, 16); q.y += __shfl_xor_sync(kFull, q.y, 16);
                q.z += __shfl_xor_sync(kFull, q.z, 16); q.w += __shfl_xor_sync(kFull, q.w, 16);
            }
            const int warp = tid >> 5, lane = tid & 31;
            if (lane < NQ) {
                *reinterpret_cast<float4*>(&red[warp][0][ql * 4]) = s;
                *reinterpret_cast<float4*>(&red[warp][1][ql * 4]) = q;
            }
            __syncthreads();
            constexpr int cpg = C0 / kGroups;
            if (tid < 2 * kGroups) {
                const int g = tid >> 1, which = tid & 1;
                double acc = 0.0;
                for (int ch = g * cpg; ch < (g + 1) * cpg; ++ch) {
                    float cs = 0.f;
#pragma unroll
                    for (int w = 0; w < 8; ++w) cs += red[w][which][ch];
                    acc += static_cast<double>(cs);
                }
                gn_fix_add(gn_acc(A.sink.acc, b, plane, ip) + g * 2 + which, acc);
            }
        }
    }
    if (tid == 0) trace_mark(A.tr, 3);
    if (MODE == MODE_FUSED && A.sch.advance) {
        __syncthreads();
        if (tid == 0) {
            const unsigned int prev = atomicAdd(A.sch.ticket, 1u);
            is_last = prev == gridDim.x * gridDim.y - 1;
        }
        __syncthreads();
        if (is_last && tid == 0) {
            for (int k = 0; k < A.sch.B; ++k) A.sch.t_idx[k] -= 1;
            *A.sch.ticket = 0u;
        }
    }
}

}  // namespace s3d
