// Backward kernels of the triplane UNet (training row).
//   reference: loss.backward() through TriplaneUNetModelSmall inside TrainLoop.forward_backward
//              (src/diffusion/train_util.py:198-235; forward being differentiated: src/diffusion/unet_triplane.py:21-145, 175-311,
//              465-510).  The adjoints are restated op by op in oracle/backward_ref.py (checked against torch.autograd), and these
//              kernels follow that structure.
// Every gradient tensor is carried multiplied by the loss scale S = 2^k (chosen on the device from max|dL/dout| so that the fp16
// (hi, lo) operands of the tensor-core dgrad / wgrad stay in range); parameter gradients are multiplied by 1/S at the very end.
// Layouts are the forward's: fp32 NHWC [B][rows][cols][C] per plane, conv operands as fp16 (hi, lo) pairs [2][B][rows][cols][C].
#pragma once
#include "boundary.cuh"
#include "common.cuh"
#include "kernels.cuh"

namespace s3d {

// ---------------------------------------------------------------- loss scale
// amax[0] = max |grad_out| as float bits (non-negative floats order like unsigned ints)
__global__ void __launch_bounds__(256) k_grad_amax(const float* __restrict__ g, long long n, unsigned int* amax) {
    float m = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
        m = fmaxf(m, fabsf(g[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax, __float_as_uint(m));
}
// S = 2^k with amax * S in [32, 64): every kernel derives it from the same word, so all of them agree bit for bit
__device__ __forceinline__ float loss_scale(const unsigned int* amax) {
    const float a = fmaxf(__uint_as_float(__ldg(amax)), 1e-30f);
    return exp2f(floorf(log2f(64.f / a)));
}

// ---------------------------------------------------------------- GroupNorm statistics -> (mean, rstd) per group
// same arithmetic as stats_coef_prologue (kernels.cuh), kept for the backward: fin[64] -> mean[g], rstd[g]
__device__ __forceinline__ void gn_mean_rstd(const unsigned long long* acc, int b, int plane, double n_per_group, int tid, float* mean, float* rstd) {
    if (tid < kGroups) {
        const unsigned long long* p = acc + (static_cast<size_t>(b) * 3 + plane) * kGnRep * 64;
        unsigned long long s = 0ull, q = 0ull;
#pragma unroll
        for (int r = 0; r < kGnRep; ++r) {
            s += __ldcg(p + r * 64 + tid * 2);
            q += __ldcg(p + r * 64 + tid * 2 + 1);
        }
        const double inv_n = 1.0 / n_per_group;
        const double mu = static_cast<double>(static_cast<long long>(s)) * kGnFixInv * inv_n;
        const float var = fmaxf(static_cast<float>(static_cast<double>(static_cast<long long>(q)) * kGnFixInv * inv_n - mu * mu), 0.f);
        mean[tid] = static_cast<float>(mu);
        rstd[tid] = rsqrtf(var + kGnEps);
    }
}

__device__ __forceinline__ float silu_grad(float f) {
    const float sig = 1.f / (1.f + __expf(-f));
    return sig * (1.f + f * (1.f - sig));
}

// =====================================================================================
// GroupNorm (+FiLM) + SiLU backward, two passes (oracle/backward_ref.py::gn_film_silu_backward).
//   forward:  xhat = (x - mean_g) rstd_g ; n = xhat gamma + beta ; f = n (1 + sc) + sh ; y = silu(f)
//   pass A :  P1[b][plane][c] = sum_px df,  P2 = sum_px df xhat,   df = dy silu'(f)
//   pass B :  dx = rstd_g ( df (1+sc) gamma - S1_g - xhat S2_g ),  S1_g = sum_{c in g} (1+sc) gamma P1 / n,  S2_g likewise with P2
// x: fp32 [B][rows][cols][C], or the (hi, lo) pair written by k_upcat (xh).  dy: fp32.  grid (slots, 3, B), block (C/4, NY).
// =====================================================================================
struct GnBwdArgs {
    TriCF x;
    TriCH xh;
    TriCF dy;
    TriDims d;
    int C, B;
    const unsigned long long* acc;      // forward group sums of x [B][3][kGnRep][64]
    TriCF gamma, beta;
    const float* film;                  // [rows][film_dim] or nullptr
    const int* film_row;
    int film_dim, film_off;
    double* psum;                       // [B][3][C][2] (P1, P2), zeroed before pass A
    TriCF add;                          // pass B: optional fp32 addend (identity-skip gradient), same shape as dx
    TriF dx;                            // pass B output
    int nslots;
};

__device__ __forceinline__ float4 gnb_load_x(const GnBwdArgs& A, int plane, size_t sample_off, size_t lo_off, size_t e) {
    if (A.x.p[plane]) return __ldg(reinterpret_cast<const float4*>(A.x.p[plane] + sample_off + e));
    const __half* ph = A.xh.p[plane] + sample_off + e;
    return join_halves4(__ldg(reinterpret_cast<const uint2*>(ph)), __ldg(reinterpret_cast<const uint2*>(ph + lo_off)));
}

// per-CTA prologue shared by both passes: per-channel forward coefficients in shared memory
//   ca, cb : f = x ca + cb        xm, xr : xhat = (x - xm) xr
__device__ __forceinline__ void gnb_prologue(const GnBwdArgs& A, int b, int plane, int tid, int nthr, float* mean, float* rstd, float* ca,
                                             float* cb, float* xm, float* xr) {
    const int C = A.C, cpg = C / kGroups;
    const double n = static_cast<double>(A.d.rows[plane]) * A.d.cols[plane] * cpg;
    gn_mean_rstd(A.acc, b, plane, n, tid, mean, rstd);
    __syncthreads();
    const float* film = A.film ? A.film + static_cast<size_t>(A.film_row ? A.film_row[b] : b) * A.film_dim + A.film_off : nullptr;
    for (int c = tid; c < C; c += nthr) {
        const int g = c / cpg;
        float ga = __ldg(A.gamma.p[plane] + c) * rstd[g];
        float be = __ldg(A.beta.p[plane] + c) - mean[g] * ga;
        if (film) {
            const float sc = 1.f + __ldg(film + c), sh = __ldg(film + C + c);
            ga *= sc;
            be = fmaf(be, sc, sh);
        }
        ca[c] = ga;
        cb[c] = be;
        xm[c] = mean[g];
        xr[c] = rstd[g];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256, 4) k_gn_bwd_a(GnBwdArgs A) {
    extern __shared__ float sm[];      // ca[C] cb[C] xm[C] xr[C] red[NY][2][C]
    __shared__ float mean[kGroups], rstd[kGroups];
    const int plane = blockIdx.y, b = blockIdx.z, C = A.C;
    const int tx = threadIdx.x, ty = threadIdx.y, NY = blockDim.y, tid = ty * blockDim.x + tx, nthr = blockDim.x * NY;
    float *ca = sm, *cb = sm + C, *xm = sm + 2 * C, *xr = sm + 3 * C, *red = sm + 4 * C;
    gnb_prologue(A, b, plane, tid, nthr, mean, rstd, ca, cb, xm, xr);
    const int npx = A.d.rows[plane] * A.d.cols[plane];
    const int ppc = (npx + A.nslots - 1) / A.nslots;
    const int p0 = blockIdx.x * ppc, p1 = min(npx, p0 + ppc);
    const size_t sample_off = static_cast<size_t>(b) * npx * C, lo_off = static_cast<size_t>(A.B) * npx * C;
    const float4 a4 = *reinterpret_cast<const float4*>(ca + tx * 4), b4 = *reinterpret_cast<const float4*>(cb + tx * 4);
    const float4 m4 = *reinterpret_cast<const float4*>(xm + tx * 4), r4 = *reinterpret_cast<const float4*>(xr + tx * 4);
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    // (unrolling this loop by four costs registers / occupancy: 1.43 -> 1.82 ms per step, measured; instead the NEXT pixel's two
    // loads are requested before the current pixel is used: two pixels in flight per thread for eight more registers)
    float4 xn = make_float4(0.f, 0.f, 0.f, 0.f), dyn = xn;
    if (p0 + ty < p1) {
        const size_t e = static_cast<size_t>(p0 + ty) * C + tx * 4;
        xn = gnb_load_x(A, plane, sample_off, lo_off, e);
        dyn = __ldg(reinterpret_cast<const float4*>(A.dy.p[plane] + sample_off + e));
    }
    for (int px = p0 + ty; px < p1; px += NY) {
        const float4 x = xn, dy = dyn;
        if (px + NY < p1) {
            const size_t e = static_cast<size_t>(px + NY) * C + tx * 4;
            xn = gnb_load_x(A, plane, sample_off, lo_off, e);
            dyn = __ldg(reinterpret_cast<const float4*>(A.dy.p[plane] + sample_off + e));
        }
        const float d0 = dy.x * silu_grad(fmaf(x.x, a4.x, b4.x)), d1 = dy.y * silu_grad(fmaf(x.y, a4.y, b4.y));
        const float d2 = dy.z * silu_grad(fmaf(x.z, a4.z, b4.z)), d3 = dy.w * silu_grad(fmaf(x.w, a4.w, b4.w));
        s1.x += d0; s1.y += d1; s1.z += d2; s1.w += d3;
        s2.x = fmaf(d0, (x.x - m4.x) * r4.x, s2.x); s2.y = fmaf(d1, (x.y - m4.y) * r4.y, s2.y);
        s2.z = fmaf(d2, (x.z - m4.z) * r4.z, s2.z); s2.w = fmaf(d3, (x.w - m4.w) * r4.w, s2.w);
    }
    float* r1 = red + (ty * 2 + 0) * C + tx * 4;
    float* r2 = red + (ty * 2 + 1) * C + tx * 4;
    r1[0] = s1.x; r1[1] = s1.y; r1[2] = s1.z; r1[3] = s1.w;
    r2[0] = s2.x; r2[1] = s2.y; r2[2] = s2.z; r2[3] = s2.w;
    __syncthreads();
    for (int i = tid; i < 2 * C; i += nthr) {
        const int which = i / C, c = i - which * C;
        double acc = 0.0;
        for (int y = 0; y < NY; ++y) acc += static_cast<double>(red[(y * 2 + which) * C + c]);
        atomicAdd(A.psum + ((static_cast<size_t>(b) * 3 + plane) * C + c) * 2 + which, acc);
    }
}

__global__ void __launch_bounds__(256, 4) k_gn_bwd_b(GnBwdArgs A) {
    extern __shared__ float sm[];      // ca cb xm xr k1[C] k2[C] k3[C] gs[2*32]
    __shared__ float mean[kGroups], rstd[kGroups];
    const int plane = blockIdx.y, b = blockIdx.z, C = A.C, cpg = C / kGroups;
    const int tx = threadIdx.x, ty = threadIdx.y, NY = blockDim.y, tid = ty * blockDim.x + tx, nthr = blockDim.x * NY;
    float *ca = sm, *cb = sm + C, *xm = sm + 2 * C, *xr = sm + 3 * C, *k2 = sm + 4 * C, *k3 = sm + 5 * C, *gs = sm + 6 * C;
    gnb_prologue(A, b, plane, tid, nthr, mean, rstd, ca, cb, xm, xr);
    const int npx = A.d.rows[plane] * A.d.cols[plane];
    // group sums S1_g, S2_g from the per-channel sums of pass A:  dxhat = df * (1+sc) gamma = df * ca / rstd
    if (tid < 2 * kGroups) {
        const int g = tid >> 1, which = tid & 1;
        double acc = 0.0;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c)
            acc += static_cast<double>(ca[c] / xr[c]) * A.psum[((static_cast<size_t>(b) * 3 + plane) * C + c) * 2 + which];
        gs[tid] = static_cast<float>(acc / (static_cast<double>(npx) * cpg));
    }
    __syncthreads();
    for (int c = tid; c < C; c += nthr) {
        const int g = c / cpg;
        const float r = xr[c], S1 = gs[g * 2], S2 = gs[g * 2 + 1];
        // dx = df ca - rstd S1 - (x - mean) rstd^2 S2  =  df ca - x k3 - k2
        k3[c] = r * r * S2;
        k2[c] = r * S1 - xm[c] * r * r * S2;
    }
    __syncthreads();
    const int ppc = (npx + A.nslots - 1) / A.nslots;
    const int p0 = blockIdx.x * ppc, p1 = min(npx, p0 + ppc);
    const size_t sample_off = static_cast<size_t>(b) * npx * C, lo_off = static_cast<size_t>(A.B) * npx * C;
    const float4 a4 = *reinterpret_cast<const float4*>(ca + tx * 4), b4 = *reinterpret_cast<const float4*>(cb + tx * 4);
    const float4 q2 = *reinterpret_cast<const float4*>(k2 + tx * 4), q3 = *reinterpret_cast<const float4*>(k3 + tx * 4);
    const float* addp = A.add.p[plane];
    float* out = A.dx.p[plane];
    float4 xn = make_float4(0.f, 0.f, 0.f, 0.f), dyn = xn;
    if (p0 + ty < p1) {
        const size_t e = static_cast<size_t>(p0 + ty) * C + tx * 4;
        xn = gnb_load_x(A, plane, sample_off, lo_off, e);
        dyn = __ldg(reinterpret_cast<const float4*>(A.dy.p[plane] + sample_off + e));
    }
    for (int px = p0 + ty; px < p1; px += NY) {
        const size_t e = static_cast<size_t>(px) * C + tx * 4;
        const float4 x = xn, dy = dyn;
        if (px + NY < p1) {                              // the next pixel's loads are in flight while this one is processed
            const size_t en = static_cast<size_t>(px + NY) * C + tx * 4;
            xn = gnb_load_x(A, plane, sample_off, lo_off, en);
            dyn = __ldg(reinterpret_cast<const float4*>(A.dy.p[plane] + sample_off + en));
        }
        float4 o;
        o.x = dy.x * silu_grad(fmaf(x.x, a4.x, b4.x)) * a4.x - x.x * q3.x - q2.x;
        o.y = dy.y * silu_grad(fmaf(x.y, a4.y, b4.y)) * a4.y - x.y * q3.y - q2.y;
        o.z = dy.z * silu_grad(fmaf(x.z, a4.z, b4.z)) * a4.z - x.z * q3.z - q2.z;
        o.w = dy.w * silu_grad(fmaf(x.w, a4.w, b4.w)) * a4.w - x.w * q3.w - q2.w;
        if (addp) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(addp + sample_off + e));
            o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
        }
        *reinterpret_cast<float4*>(out + sample_off + e) = o;
    }
}

// Parameter / FiLM gradients of one norm site from the pass-A sums (still multiplied by the loss scale).
//   dgamma[plane][c] = sum_b (1+sc) P2 ; dbeta = sum_b (1+sc) P1 ; dscale[b][c] += gamma P2 + beta P1 ; dshift[b][c] += P1 (over planes)
// grid ceil(B*C / 256), block 256 (grid-stride loops)
struct GnFinArgs {
    const double* psum;      // [B][3][C][2]
    TriCF gamma, beta;
    const float* film;
    const int* film_row;
    int film_dim, film_off, C, B;
    float* dgamma[3];
    float* dbeta[3];
    float* dfilm;            // [B][film_dim] accumulated (+=), or nullptr
};
__global__ void __launch_bounds__(256) k_gn_bwd_fin(GnFinArgs A) {
    const int C = A.C;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gstride = gridDim.x * blockDim.x;      // grid-stride: a few CTAs instead of one
    for (int i = gtid; i < 3 * C; i += gstride) {
        const int plane = i / C, c = i - plane * C;
        double dg = 0.0, db = 0.0;
#pragma unroll 8
        for (int b = 0; b < A.B; ++b) {
            const double* p = A.psum + ((static_cast<size_t>(b) * 3 + plane) * C + c) * 2;
            double sc = 1.0;
            if (A.film) sc += static_cast<double>(A.film[static_cast<size_t>(A.film_row ? A.film_row[b] : b) * A.film_dim + A.film_off + c]);
            db += sc * p[0];
            dg += sc * p[1];
        }
        A.dgamma[plane][c] = static_cast<float>(dg);
        A.dbeta[plane][c] = static_cast<float>(db);
    }
    if (A.film && A.dfilm) {
        for (int i = gtid; i < A.B * C; i += gstride) {
            const int b = i / C, c = i - b * C;
            double dsc = 0.0, dsh = 0.0;
            for (int plane = 0; plane < 3; ++plane) {
                const double* p = A.psum + ((static_cast<size_t>(b) * 3 + plane) * C + c) * 2;
                dsc += static_cast<double>(A.gamma.p[plane][c]) * p[1] + static_cast<double>(A.beta.p[plane][c]) * p[0];
                dsh += p[0];
            }
            A.dfilm[static_cast<size_t>(b) * A.film_dim + A.film_off + c] += static_cast<float>(dsc);
            A.dfilm[static_cast<size_t>(b) * A.film_dim + A.film_off + C + c] += static_cast<float>(dsh);
        }
    }
}

// =====================================================================================
// Gradient staging: the fp32 gradient of a conv OUTPUT -> what the conv's backward consumes:
//   * the (hi, lo) fp16 pair (operand of the tensor-core dgrad / wgrad),
//   * its axis sums (rollout adjoint; oracle/backward_ref.py::_axis_sums): rs[b][seg(plane,0)+row][C] = sum over columns,
//     cs[b][seg(plane,1)+col][C] = sum over rows  (fp64 atomics, zeroed before),
//   * the per-channel total (bias gradient) [3][C] summed over the batch, and per (b, c) (additive-embedding gradient).
// grid (row strips of 8, 3, B), block (C/4, NY)
// =====================================================================================
struct StageArgs {
    TriCF g;              // fp32 [B][rows][cols][C]
    TriDims d;
    int C, B;
    TriH pair;            // out [2][B][rows][cols][C]
    double* sums;         // [B][total_len][C]
    int seg_off[6];
    int total_len;
    double* bias_sum;     // [3][C]  (+=)
    double* bc_sum;       // [B][C] (+= over planes) or nullptr
};
__global__ void __launch_bounds__(256) k_grad_stage(StageArgs A) {
    extern __shared__ float sm[];      // red[NY][kGsRows][C]
    const int plane = blockIdx.y, b = blockIdx.z, C = A.C;
    const int rows = A.d.rows[plane], cols = A.d.cols[plane];
    const int tx = threadIdx.x, ty = threadIdx.y, NY = blockDim.y, tid = ty * blockDim.x + tx, nthr = blockDim.x * NY;
    const int r0 = blockIdx.x * kGsRows;
    if (r0 >= rows) return;
    const int nr = min(kGsRows, rows - r0);
    const size_t plane_elems = static_cast<size_t>(rows) * cols * C;
    const float* gp = A.g.p[plane] + static_cast<size_t>(b) * plane_elems;
    __half* ph = A.pair.p[plane] + static_cast<size_t>(b) * plane_elems;
    const size_t lo_off = static_cast<size_t>(A.B) * plane_elems;
    double* sb = A.sums + static_cast<size_t>(b) * A.total_len * C;
    double* srow = sb + static_cast<size_t>(A.seg_off[plane * 2 + 0]) * C;
    double* scol = sb + static_cast<size_t>(A.seg_off[plane * 2 + 1]) * C;
    // One pass over the strip: a thread takes a column at a time with all of its (<= 8) rows' loads in flight, splits and stores
    // them, and keeps the row sums in registers; the column sum falls out of the same values (the strip used to be read twice, one
    // row at a time with a block-wide reduction per row).  Same summation orders as before: results are bit-identical.
    float4 racc[kGsRows];
#pragma unroll
    for (int r = 0; r < kGsRows; ++r) racc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = ty; c < cols; c += NY) {
        float4 v[kGsRows];
#pragma unroll
        for (int r = 0; r < kGsRows; ++r)
            v[r] = r < nr ? __ldg(reinterpret_cast<const float4*>(gp + (static_cast<size_t>(r0 + r) * cols + c) * C + tx * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 cacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kGsRows; ++r)
            if (r < nr) {
                const size_t e = (static_cast<size_t>(r0 + r) * cols + c) * C + tx * 4;
                store_split4(ph + e, ph + lo_off + e, v[r]);
                racc[r].x += v[r].x; racc[r].y += v[r].y; racc[r].z += v[r].z; racc[r].w += v[r].w;
                cacc.x += v[r].x; cacc.y += v[r].y; cacc.z += v[r].z; cacc.w += v[r].w;
            }
        double* p = scol + static_cast<size_t>(c) * C + tx * 4;
        atomicAdd(p, static_cast<double>(cacc.x)); atomicAdd(p + 1, static_cast<double>(cacc.y));
        atomicAdd(p + 2, static_cast<double>(cacc.z)); atomicAdd(p + 3, static_cast<double>(cacc.w));
    }
    // row sums: sm[ty][r][C], added over ty in a fixed order
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kGsRows; ++r) {
        float* cellr = sm + (static_cast<size_t>(ty) * kGsRows + r) * C + tx * 4;
        cellr[0] = racc[r].x; cellr[1] = racc[r].y; cellr[2] = racc[r].z; cellr[3] = racc[r].w;
        if (r < nr) { tot.x += racc[r].x; tot.y += racc[r].y; tot.z += racc[r].z; tot.w += racc[r].w; }
    }
    __syncthreads();
    for (int i = tid; i < nr * C; i += nthr) {
        const int r = i / C, ch = i - r * C;
        double acc = 0.0;
        for (int y = 0; y < NY; ++y) acc += static_cast<double>(sm[(static_cast<size_t>(y) * kGsRows + r) * C + ch]);
        atomicAdd(srow + static_cast<size_t>(r0 + r) * C + ch, acc);
    }
    __syncthreads();
    // per-channel total of the strip
    float* cell = sm + ty * C + tx * 4;
    cell[0] = tot.x; cell[1] = tot.y; cell[2] = tot.z; cell[3] = tot.w;
    __syncthreads();
    for (int i = tid; i < C; i += nthr) {
        double acc = 0.0;
        for (int y = 0; y < NY; ++y) acc += static_cast<double>(sm[y * C + i]);
        atomicAdd(A.bias_sum + plane * C + i, acc);
        if (A.bc_sum) atomicAdd(A.bc_sum + static_cast<size_t>(b) * C + i, acc);
    }
}

// =====================================================================================
// Rollout adjoint (oracle/backward_ref.py::tri_conv_backward_folded).  For the source s = (plane p, group g) of a conv with
// stored weight W[Cout][3C][3][3]: the broadcast vector vec_s[L][C] (an axis mean of another plane) entered plane p's conv along
// rows (row_varying) or columns.  With Sy_a[pos][co] = the axis sum of dY over the lines the tap `across = a` reaches
// (a = 0: all but the first line, 1: all, 2: all but the last),
//   dvec[j][c]             = sum_{along, a, co} Sy_a[j - along + 1][co] W[co][gC + c][along, a]
//   dW[co][gC + c][along,a] = sum_{b, pos} Sy_a[pos][co] vec[pos + along - 1][c]
// dvec / n_avg is what every pixel of the source plane's line j receives: written as the Trow / Tcol addend of the dgrad conv
// (all four edge classes get the same values).
// =====================================================================================
struct RollBwdSrc {
    int plane, g, row_varying;   // plane p whose conv consumed the vector; channel group 1 | 2; vector indexed by p's row or column
    int L, across_len;           // vector length; number of lines across (cols if row_varying else rows)
    int sy_off;                  // segment (positions) of dY's axis sums: rs (row_varying) or cs of plane p
    int vec_off;                 // segment of the forward axis sums of the SOURCE plane (raw fixed point, k_gn_silu)
    float vec_scale;             // 2^-24 / averaged length: fixed-point sum -> mean
    float inv_navg;              // 1 / averaged length
    float* T;                    // dgrad addend of the source plane [B][4][L][C]
    const float* wv;             // the group's weights as [along*3 + across][Cout][C] (k_pack_rollv)
    float* dw;                   // gradient of W of plane p [Cout][3C][3][3] (+=)
};
struct RollBwdArgs {
    RollBwdSrc s[6];
    TriCF dy;                    // fp32 dY [B][rows][cols][Cout] (first / last lines)
    TriDims d;
    const double* sy;            // [B][total_len][Cout]
    const unsigned long long* fsums;   // forward axis sums [B][total_len][C]
    int total_len, C, Cout, B;
};
// line `which` (0 first, 1 last) across of dY at position pos, channel co
__device__ __forceinline__ float roll_edge(const RollBwdArgs& A, const RollBwdSrc& S, int b, int pos, int co, int which) {
    const int rows = A.d.rows[S.plane], cols = A.d.cols[S.plane];
    const float* p = A.dy.p[S.plane] + static_cast<size_t>(b) * rows * cols * A.Cout;
    const int line = which ? S.across_len - 1 : 0;
    const int r = S.row_varying ? pos : line, c = S.row_varying ? line : pos;
    return __ldg(p + (static_cast<size_t>(r) * cols + c) * A.Cout + co);
}
// Sy_a[pos][co] for the three taps across, from the full axis sum and the first / last line of dY
__device__ __forceinline__ void roll_sy3(const RollBwdArgs& A, const RollBwdSrc& S, int b, int pos, int co, float (&out)[3]) {
    out[0] = out[1] = out[2] = 0.f;
    if (pos >= 0 && pos < S.L) {
        const float full = static_cast<float>(A.sy[(static_cast<size_t>(b) * A.total_len + S.sy_off + pos) * A.Cout + co]);
        out[0] = full - roll_edge(A, S, b, pos, co, 0);
        out[1] = full;
        out[2] = full - roll_edge(A, S, b, pos, co, 1);
    }
}
// dvec for 32 positions x 64 channels of one source and sample: a small tiled fp32 GEMM, K = (along, across, co).
//   wrv: weights re-packed as [along*3 + across][Cout][C] (k_pack_rollv) so that weight tiles load coalesced
// grid (ceil(Lmax/32), 6 * (C/64), B), block 256; dynamic smem: sy3[3][34][Cout] + wt[32][64]
constexpr int kRvPos = 32;
__global__ void __launch_bounds__(256) k_roll_bwd_vec(RollBwdArgs A) {
    extern __shared__ __align__(16) float sm_rv[];
    float* sm = sm_rv;
    const int nct = A.C / 64;
    const RollBwdSrc S = A.s[blockIdx.y / nct];
    const int c0 = (blockIdx.y % nct) * 64;
    const int b = blockIdx.z, p0 = blockIdx.x * kRvPos, Cout = A.Cout, C = A.C;
    if (p0 >= S.L) return;
    float* sy3 = sm;                                   // [3][kRvPos + 2][Cout], slot j <-> position p0 + j - 1
    float* wt = sm + 3 * (kRvPos + 2) * Cout;          // [32][64]
    for (int i = threadIdx.x; i < (kRvPos + 2) * Cout; i += blockDim.x) {
        const int j = i / Cout, co = i - j * Cout;
        float v[3];
        roll_sy3(A, S, b, p0 + j - 1, co, v);
        sy3[(0 * (kRvPos + 2) + j) * Cout + co] = v[0];
        sy3[(1 * (kRvPos + 2) + j) * Cout + co] = v[1];
        sy3[(2 * (kRvPos + 2) + j) * Cout + co] = v[2];
    }
    const int cq = threadIdx.x & 15, pp = threadIdx.x >> 4;        // 4 channels x 2 positions per thread
    float acc[2][4] = {};
    // weight chunks of 32 (co) x 64 (c), software pipelined: the next chunk is requested into registers before the current one is
    // multiplied (two float4 per thread), so its L2 latency hides behind the FMAs
    const int nco = Cout / 32, nchunk = 9 * nco;
    float4 wn[2];
    auto request = [&](int ch) {
        const int t = ch / nco, co0 = (ch - t * nco) * 32;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int i = threadIdx.x + j * 256, k = i >> 4, q = i & 15;
            wn[j] = __ldg(reinterpret_cast<const float4*>(S.wv + (static_cast<size_t>(t) * Cout + co0 + k) * C + c0) + q);
        }
    };
    request(0);
    for (int ch = 0; ch < nchunk; ++ch) {
        const int t = ch / nco, co0 = (ch - t * nco) * 32;
        const int al = t / 3, ac = t - al * 3;
        const float* s0 = sy3 + (ac * (kRvPos + 2) + (2 * pp + 2 - al)) * Cout;      // position j - al + 1 -> slot jl + 2 - al
        const float* s1 = s0 + Cout;
        __syncthreads();                               // the previous chunk has been multiplied (first pass: sy3 is complete)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int i = threadIdx.x + j * 256, k = i >> 4, q = i & 15;
            *reinterpret_cast<float4*>(wt + k * 64 + q * 4) = wn[j];
        }
        __syncthreads();
        if (ch + 1 < nchunk) request(ch + 1);
        // k in steps of four: the two positions' operands come as one 16-byte broadcast load each (the loop was bound by
        // shared-memory instructions: 12 per 32 FMAs before, 6 now); same accumulation order as the scalar loop
#pragma unroll 2
        for (int k = 0; k < 32; k += 4) {
            const float4 a0 = *reinterpret_cast<const float4*>(s0 + co0 + k), a1 = *reinterpret_cast<const float4*>(s1 + co0 + k);
            const float av0[4] = {a0.x, a0.y, a0.z, a0.w}, av1[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 w = *reinterpret_cast<const float4*>(wt + (k + kk) * 64 + cq * 4);
                acc[0][0] = fmaf(av0[kk], w.x, acc[0][0]); acc[0][1] = fmaf(av0[kk], w.y, acc[0][1]);
                acc[0][2] = fmaf(av0[kk], w.z, acc[0][2]); acc[0][3] = fmaf(av0[kk], w.w, acc[0][3]);
                acc[1][0] = fmaf(av1[kk], w.x, acc[1][0]); acc[1][1] = fmaf(av1[kk], w.y, acc[1][1]);
                acc[1][2] = fmaf(av1[kk], w.z, acc[1][2]); acc[1][3] = fmaf(av1[kk], w.w, acc[1][3]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int j = p0 + 2 * pp + r;
        if (j < S.L) {
            const float4 o = make_float4(acc[r][0] * S.inv_navg, acc[r][1] * S.inv_navg, acc[r][2] * S.inv_navg, acc[r][3] * S.inv_navg);
#pragma unroll
            for (int cls = 0; cls < 4; ++cls)
                *reinterpret_cast<float4*>(S.T + ((static_cast<size_t>(b) * 4 + cls) * S.L + j) * C + c0 + cq * 4) = o;
        }
    }
}
// dW of the broadcast channels: a 64 (co) x 64 (c) tile of one (source, along) for the three taps across, K = (b, pos) over a
// slice of the batch; partial[split][source*3 + along][across][co][c] is reduced in a fixed order by k_roll_bwd_w_reduce.
// grid ((Cout/64) * (C/64), 18, nsplit), block 256
__global__ void __launch_bounds__(256, 2) k_roll_bwd_w(RollBwdArgs A, float* __restrict__ partial, int nsplit) {
    __shared__ __align__(16) float sy3[3][16][64];
    __shared__ __align__(16) float vv[16][64];
    const int Cout = A.Cout, C = A.C, nct = C / 64;
    const int co0 = (blockIdx.x / nct) * 64, c0 = (blockIdx.x % nct) * 64;
    const int src = blockIdx.y / 3, al = blockIdx.y - src * 3;
    const RollBwdSrc S = A.s[src];
    const int b0 = static_cast<int>(static_cast<long long>(A.B) * blockIdx.z / nsplit), b1 = static_cast<int>(static_cast<long long>(A.B) * (blockIdx.z + 1) / nsplit);
    const int tc = threadIdx.x & 15, tco = threadIdx.x >> 4;
    float acc[3][4][4] = {};
    // K chunks of 16 positions, software pipelined: the NEXT chunk's operands (axis sum of dY, its first / last line, the forward axis
    // sum) are requested into registers before the current chunk is multiplied, so their L2 latency hides behind the FMAs
    const int nper = (S.L + 15) / 16, nchunk = (b1 - b0) * nper;
    double pfull[4];
    float pe0[4], pe1[4];
    unsigned long long pfv[4];
    bool pok[4], pvok[4];
    auto request = [&](int ch) {
        const int b = b0 + ch / nper, p0 = (ch % nper) * 16;
        const unsigned long long* fv = A.fsums + (static_cast<size_t>(b) * A.total_len + S.vec_off) * C;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = threadIdx.x + j * 256, k = i >> 6, x = i & 63, pos = p0 + k, q = pos + al - 1;
            pok[j] = pos < S.L;
            pvok[j] = pok[j] && q >= 0 && q < S.L;
            pfull[j] = 0.0;
            pe0[j] = pe1[j] = 0.f;
            pfv[j] = 0ull;
            if (pok[j]) {
                pfull[j] = A.sy[(static_cast<size_t>(b) * A.total_len + S.sy_off + pos) * A.Cout + co0 + x];
                pe0[j] = roll_edge(A, S, b, pos, co0 + x, 0);
                pe1[j] = roll_edge(A, S, b, pos, co0 + x, 1);
            }
            if (pvok[j]) pfv[j] = __ldg(fv + static_cast<size_t>(q) * C + c0 + x);
        }
    };
    if (nchunk > 0) request(0);
    for (int ch = 0; ch < nchunk; ++ch) {
        __syncthreads();                                     // the previous chunk has been multiplied
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = threadIdx.x + j * 256, k = i >> 6, x = i & 63;
            const float full = static_cast<float>(pfull[j]);       // same arithmetic as roll_sy3
            sy3[0][k][x] = pok[j] ? full - pe0[j] : 0.f;
            sy3[1][k][x] = pok[j] ? full : 0.f;
            sy3[2][k][x] = pok[j] ? full - pe1[j] : 0.f;
            vv[k][x] = pvok[j] ? __ll2float_rn(static_cast<long long>(pfv[j])) * S.vec_scale : 0.f;
        }
        __syncthreads();
        if (ch + 1 < nchunk) request(ch + 1);
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            const float4 v = *reinterpret_cast<const float4*>(&vv[k][tc * 4]);
#pragma unroll
            for (int ac = 0; ac < 3; ++ac) {
                const float4 y = *reinterpret_cast<const float4*>(&sy3[ac][k][tco * 4]);
                const float yy[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[ac][i][0] = fmaf(yy[i], v.x, acc[ac][i][0]); acc[ac][i][1] = fmaf(yy[i], v.y, acc[ac][i][1]);
                    acc[ac][i][2] = fmaf(yy[i], v.z, acc[ac][i][2]); acc[ac][i][3] = fmaf(yy[i], v.w, acc[ac][i][3]);
                }
            }
        }
    }
    float* out = partial + (static_cast<size_t>(blockIdx.z) * 18 + blockIdx.y) * 3 * Cout * C;
#pragma unroll
    for (int ac = 0; ac < 3; ++ac)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(out + (static_cast<size_t>(ac) * Cout + co0 + tco * 4 + i) * C + c0 + tc * 4) =
                make_float4(acc[ac][i][0], acc[ac][i][1], acc[ac][i][2], acc[ac][i][3]);
}
// grid (ceil(3*Cout*C / 256), 18), block 256
__global__ void __launch_bounds__(256) k_roll_bwd_w_reduce(RollBwdArgs A, const float* __restrict__ partial, int nsplit) {
    const int Cout = A.Cout, C = A.C;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * Cout * C) return;
    const int src = blockIdx.y / 3, al = blockIdx.y - src * 3;
    const RollBwdSrc S = A.s[src];
    const int c = i % C, co = (i / C) % Cout, ac = i / (C * Cout);
    float acc = 0.f;
    for (int s = 0; s < nsplit; ++s) acc += partial[(static_cast<size_t>(s) * 18 + blockIdx.y) * 3 * Cout * C + i];
    S.dw[(static_cast<size_t>(co) * 3 * C + S.g * C + c) * 9 + (S.row_varying ? al * 3 + ac : ac * 3 + al)] += acc;
}
// wv[(along*3 + across)][co][c] = W[co][g*C + c][kh][kw], (kh, kw) = (along, across) for a row-indexed vector, else (across, along)
__global__ void __launch_bounds__(256) k_pack_rollv(const float* __restrict__ w, int Cout, int C, int g, int row_varying, float* __restrict__ out) {
    const long long n = static_cast<long long>(9) * Cout * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C), co = static_cast<int>((i / C) % Cout), t = static_cast<int>(i / (static_cast<long long>(C) * Cout));
        const int al = t / 3, ac = t - al * 3;
        out[i] = w[(static_cast<size_t>(co) * 3 * C + g * C + c) * 9 + (row_varying ? al * 3 + ac : ac * 3 + al)];
    }
}

// =====================================================================================
// Weight gradient of the 3x3 conv's own channels (and of the 1x1 skip conv), CUDA-core version (cross-check / bring-up):
//   dW[co][c][kh][kw] += sum_{b, px} dY[b][px][co] A[b][px + (kh-1, kw-1)][c]       (zero outside the plane)
// operands are the (hi, lo) pairs the tensor-core path uses, joined to fp32.  One CTA = a 32 x 32 (co, c) tile of one tap over a
// chunk of pixels; fp32 atomics at the end.  grid (chunks, ceil(Cout/32) * ceil(C/32) * ntap, 3), block (32, 8).
// =====================================================================================
struct WgradArgs {
    TriCH dy;          // [2][B][rows][cols][Cout]
    TriCH a;           // [2][B][rows][cols][C]
    TriDims d;
    int C, Cout, B, ntap;      // ntap 9 (3x3, pad 1) or 1 (1x1)
    int Cw;                    // in-channels of the stored weight (3C / C for the 3x3, Cs for the 1x1)
    float* dw[3];              // [Cout][Cw][ntap]
    int chunks;
};
__global__ void __launch_bounds__(256) k_wgrad_ffma(WgradArgs A) {
    __shared__ float sy[32][33], sa[32][33];      // [pixel][co] / [pixel][c]
    const int plane = blockIdx.z, C = A.C, Cout = A.Cout;
    const int rows = A.d.rows[plane], cols = A.d.cols[plane], npx = rows * cols;
    const int nco = (Cout + 31) / 32, nc = (C + 31) / 32;
    int t = blockIdx.y;
    const int tap = t % A.ntap;
    t /= A.ntap;
    const int c0 = (t % nc) * 32, co0 = (t / nc) * 32;
    const int dh = A.ntap == 9 ? tap / 3 - 1 : 0, dw_ = A.ntap == 9 ? tap % 3 - 1 : 0;
    const long long total = static_cast<long long>(A.B) * npx;
    const long long per = (total + A.chunks - 1) / A.chunks;
    const long long q0 = blockIdx.x * per, q1 = min(total, q0 + per);
    const int tx = threadIdx.x, ty = threadIdx.y;      // outputs: co = co0 + ty*4 + {0..3}, c = c0 + tx
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const size_t lo_y = static_cast<size_t>(A.B) * npx * Cout, lo_a = static_cast<size_t>(A.B) * npx * C;
    for (long long base = q0; base < q1; base += 32) {
        // stage 32 pixels: dY[px][co0..+31], A[px + tap][c0..+31]
        for (int i = ty; i < 32; i += 8) {
            const long long q = base + i;
            float vy = 0.f, va = 0.f;
            if (q < q1) {
                const int b = static_cast<int>(q / npx), px = static_cast<int>(q - static_cast<long long>(b) * npx);
                const int r = px / cols, c = px - r * cols;
                if (co0 + tx < Cout) {
                    const size_t e = (static_cast<size_t>(b) * npx + px) * Cout + co0 + tx;
                    vy = __half2float(A.dy.p[plane][e]) + __half2float(A.dy.p[plane][lo_y + e]) * (1.f / kLoScale);
                }
                const int rr = r + dh, cc = c + dw_;
                if (rr >= 0 && rr < rows && cc >= 0 && cc < cols && c0 + tx < C) {
                    const size_t e = (static_cast<size_t>(b) * npx + static_cast<size_t>(rr) * cols + cc) * C + c0 + tx;
                    va = __half2float(A.a.p[plane][e]) + __half2float(A.a.p[plane][lo_a + e]) * (1.f / kLoScale);
                }
            }
            sy[i][tx] = vy;
            sa[i][tx] = va;
        }
        __syncthreads();
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
            const float av = sa[i][tx];
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] = fmaf(sy[i][ty * 4 + k], av, acc[k]);
        }
        __syncthreads();
    }
    if (c0 + tx < C)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int co = co0 + ty * 4 + k;
            if (co < Cout) atomicAdd(A.dw[plane] + (static_cast<size_t>(co) * A.Cw + c0 + tx) * A.ntap + tap, acc[k]);
        }
}

// =====================================================================================
// Boundary 1x1 convs.
//   head  (unet_triplane.py:441-445): out[co][px] = b[co] + sum_c w[co][c] y[px][c],  y = silu(GN(h))
//     k_head_bwd: dy[px][c] = S * sum_co g[co][px] w[co][c] (fp32 NHWC, then the generic GroupNorm backward), and
//                 dw[co][c] += S * sum_px g y,  db[co] += S * sum_px g             (g = dL/dout in the composed layout)
//   in_conv (:378): h0[px][co] = b[co] + sum_c w[co][c] x[c][px]
//     k_inconv_wgrad: dw[co][c] += sum_px dh0[px][co] x[c][px],  db[co] += sum_px dh0[px][co]
// =====================================================================================
struct HeadBwdArgs {
    const float* g;            // composed [B][Cf][H+D][W+D]
    const unsigned int* amax;
    TriCF h;                   // [B][rows][cols][C0]
    TriDims d;
    int C0, Cf, H, W, Dd, B;
    int tile_start[4], tiles_fast[3];      // the boundary kernels' 4 x 32-pixel tiling of the three planes
    const unsigned long long* acc;
    TriCF gamma, beta;
    TriCF w_out;               // [Cf][C0]
    TriF dy;                   // out [B][rows][cols][C0]
    float* dw[3];              // [Cf][C0] (+=)
    float* db[3];              // [Cf] (+=)
};
// Same mapping as k_boundary: a CTA owns a 128-pixel tile (4 x 32 with the 32 along the axis that is contiguous in the composed
// tensor), the composed gradient tile is staged through shared memory, a thread is (pixel, channel quad) with its 4 channels'
// rows of w_out in registers; dw partials stay in registers over the tile's 8 passes and are reduced once per CTA.
// A CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... of ONE plane and reduces its dw / db partials once at the end (the
// global atomics all land on the same Cf x C0 addresses: as few CTAs as fill the machine).  grid (gx, B, 3), block 256
// PART 0: dy only (needs g and w_out: no activation read); PART 1: dw / db only (needs g and y = silu(GN(h))).  One kernel doing both
// held 96 weight + accumulator registers per thread (236 in all, one CTA per SM); the two halves run two CTAs per SM each.
template <int NQ, int PART>
__global__ void __launch_bounds__(256, 2) k_head_bwd(HeadBwdArgs A) {
    constexpr int PPP = 256 / NQ, NPASS = kBndPx / PPP, C0 = NQ * 4;
    __shared__ __align__(16) float gs[kMaxCf][kBndPx];      // S * dL/dout of the tile, [channel][pixel]
    __shared__ __align__(16) float coefA[C0], coefB[C0];
    __shared__ float dwacc[kMaxCf][C0];
    __shared__ float dbacc[kMaxCf];
    __shared__ float mean[kGroups], rstd[kGroups];
    const int tid = threadIdx.x, b = blockIdx.y, Cf = A.Cf, plane = blockIdx.z;
    const int ntile = A.tile_start[plane + 1] - A.tile_start[plane];
    if (static_cast<int>(blockIdx.x) >= ntile) return;
    const int rows = A.d.rows[plane], cols = A.d.cols[plane], npx = rows * cols, cpg = C0 / kGroups;
    const bool fast_rows = plane == 2;
    const int Hc = A.H + A.Dd, Wc = A.W + A.Dd;
    const long long hw = static_cast<long long>(Hc) * Wc;
    const float S = loss_scale(A.amax);
    const int ql = tid & (NQ - 1), pslot = tid / NQ;
    float4 wq[kMaxCf];                                       // PART 0: this lane's four columns of w_out
    float4 dwr[kMaxCf];                                      // PART 1: this lane's dw partials
    float dbr = 0.f;
    float4 ca = make_float4(0.f, 0.f, 0.f, 0.f), cb = ca;
    if (PART == 1) {
        for (int e = tid; e < kMaxCf * C0; e += 256) (&dwacc[0][0])[e] = 0.f;
        if (tid < kMaxCf) dbacc[tid] = 0.f;
        gn_mean_rstd(A.acc, b, plane, static_cast<double>(npx) * cpg, tid, mean, rstd);
        __syncthreads();
        for (int c = tid; c < C0; c += 256) {
            const int g = c / cpg;
            const float ga = __ldg(A.gamma.p[plane] + c) * rstd[g];
            coefA[c] = ga;
            coefB[c] = __ldg(A.beta.p[plane] + c) - mean[g] * ga;
        }
        __syncthreads();
        ca = *reinterpret_cast<const float4*>(coefA + ql * 4);
        cb = *reinterpret_cast<const float4*>(coefB + ql * 4);
#pragma unroll
        for (int co = 0; co < kMaxCf; ++co) dwr[co] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
        for (int co = 0; co < kMaxCf; ++co)
            wq[co] = co < Cf ? __ldg(reinterpret_cast<const float4*>(A.w_out.p[plane] + static_cast<size_t>(co) * C0 + ql * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float* hp = A.h.p[plane] + static_cast<size_t>(b) * npx * C0 + ql * 4;
    float* dyp = A.dy.p[plane] + static_cast<size_t>(b) * npx * C0 + ql * 4;
    for (int ip = blockIdx.x; ip < ntile; ip += gridDim.x) {
        const int tf = ip % A.tiles_fast[plane], ts = ip / A.tiles_fast[plane];
        const int r0 = fast_rows ? tf * kBndFast : ts * kBndSlow, c0 = fast_rows ? ts * kBndSlow : tf * kBndFast;
        auto pix_r = [&](int i) { return r0 + (fast_rows ? (i & (kBndFast - 1)) : (i >> 5)); };
        auto pix_c = [&](int i) { return c0 + (fast_rows ? (i >> 5) : (i & (kBndFast - 1))); };
        __syncthreads();                                     // the previous tile's gs has been read
        for (int e = tid; e < Cf * kBndPx; e += 256) {
            const int co = e >> 7, i = e & (kBndPx - 1), r = pix_r(i), c = pix_c(i);
            gs[co][i] = (r < rows && c < cols) ? S * __ldg(A.g + (static_cast<size_t>(b) * Cf + co) * hw + composed_offset(plane, r, c, A.H, A.W, Wc)) : 0.f;
        }
        __syncthreads();
        // four pixels in flight per thread: the activation loads of a batch are issued before the first use
        constexpr int NB = NPASS < 4 ? NPASS : 4;
#pragma unroll
        for (int p0 = 0; p0 < NPASS; p0 += NB) {
            float4 hv[NB];
            size_t off[NB];
            bool ok[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const int i = (p0 + k) * PPP + pslot, r = pix_r(i), c = pix_c(i);
                ok[k] = r < rows && c < cols;
                off[k] = ok[k] ? (static_cast<size_t>(r) * cols + c) * C0 : 0;
                if (PART == 1) hv[k] = ok[k] ? __ldg(reinterpret_cast<const float4*>(hp + off[k])) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const int i = (p0 + k) * PPP + pslot;
                if (PART == 1) {
                    float4 y;
                    y.x = silu_f(fmaf(hv[k].x, ca.x, cb.x)); y.y = silu_f(fmaf(hv[k].y, ca.y, cb.y));
                    y.z = silu_f(fmaf(hv[k].z, ca.z, cb.z)); y.w = silu_f(fmaf(hv[k].w, ca.w, cb.w));
#pragma unroll
                    for (int co = 0; co < kMaxCf; ++co)
                        if (co < Cf) {
                            const float gv = gs[co][i];          // zero for pixels outside the plane
                            dwr[co].x = fmaf(gv, y.x, dwr[co].x); dwr[co].y = fmaf(gv, y.y, dwr[co].y);
                            dwr[co].z = fmaf(gv, y.z, dwr[co].z); dwr[co].w = fmaf(gv, y.w, dwr[co].w);
                        }
                } else {
                    float4 dy = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int co = 0; co < kMaxCf; ++co)
                        if (co < Cf) {
                            const float gv = gs[co][i];
                            dy.x = fmaf(gv, wq[co].x, dy.x); dy.y = fmaf(gv, wq[co].y, dy.y); dy.z = fmaf(gv, wq[co].z, dy.z); dy.w = fmaf(gv, wq[co].w, dy.w);
                        }
                    if (ok[k]) *reinterpret_cast<float4*>(dyp + off[k]) = dy;
                }
            }
        }
        if (PART == 1 && tid < Cf)
            for (int i = 0; i < kBndPx; ++i) dbr += gs[tid][i];
    }
    if (PART == 1) {
#pragma unroll
        for (int co = 0; co < kMaxCf; ++co)
            if (co < Cf) {
                atomicAdd(&dwacc[co][ql * 4 + 0], dwr[co].x); atomicAdd(&dwacc[co][ql * 4 + 1], dwr[co].y);
                atomicAdd(&dwacc[co][ql * 4 + 2], dwr[co].z); atomicAdd(&dwacc[co][ql * 4 + 3], dwr[co].w);
            }
        if (tid < Cf) dbacc[tid] = dbr;
        __syncthreads();
        for (int e = tid; e < Cf * C0; e += 256) atomicAdd(A.dw[plane] + e, dwacc[e / C0][e % C0]);
        if (tid < Cf) atomicAdd(A.db[plane] + tid, dbacc[tid]);
    }
}

struct InconvBwdArgs {
    const float* x;            // composed input [B][Cf][H+D][W+D]
    TriCF dh0;                 // [B][rows][cols][C0]
    TriDims d;
    int C0, Cf, H, W, Dd, B;
    int tile_start[4], tiles_fast[3];
    float* dw[3];              // [C0][Cf] (+=)
    float* db[3];              // [C0] (+=)
};
// same tiling and tile walk (grid (gx, B, 3)); thread = (pixel, output-channel quad): dw[co][c] += dh0[px][co] x[c][px]
template <int NQ>
__global__ void __launch_bounds__(256, 2) k_inconv_wgrad(InconvBwdArgs A) {
    constexpr int PPP = 256 / NQ, NPASS = kBndPx / PPP, C0 = NQ * 4;
    __shared__ __align__(16) float xs[kMaxCf][kBndPx];
    __shared__ float dwacc[C0][kMaxCf];
    __shared__ float dbacc[C0];
    const int tid = threadIdx.x, b = blockIdx.y, Cf = A.Cf, plane = blockIdx.z;
    const int ntile = A.tile_start[plane + 1] - A.tile_start[plane];
    if (static_cast<int>(blockIdx.x) >= ntile) return;
    const int rows = A.d.rows[plane], cols = A.d.cols[plane], npx = rows * cols;
    const bool fast_rows = plane == 2;
    const int Hc = A.H + A.Dd, Wc = A.W + A.Dd;
    const long long hw = static_cast<long long>(Hc) * Wc;
    for (int e = tid; e < C0 * kMaxCf; e += 256) (&dwacc[0][0])[e] = 0.f;
    for (int e = tid; e < C0; e += 256) dbacc[e] = 0.f;
    const int ql = tid & (NQ - 1), pslot = tid / NQ;
    const float* dp = A.dh0.p[plane] + static_cast<size_t>(b) * npx * C0 + ql * 4;
    float4 dwr[kMaxCf];
#pragma unroll
    for (int ch = 0; ch < kMaxCf; ++ch) dwr[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 dbr = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ip = blockIdx.x; ip < ntile; ip += gridDim.x) {
        const int tf = ip % A.tiles_fast[plane], ts = ip / A.tiles_fast[plane];
        const int r0 = fast_rows ? tf * kBndFast : ts * kBndSlow, c0 = fast_rows ? ts * kBndSlow : tf * kBndFast;
        auto pix_r = [&](int i) { return r0 + (fast_rows ? (i & (kBndFast - 1)) : (i >> 5)); };
        auto pix_c = [&](int i) { return c0 + (fast_rows ? (i >> 5) : (i & (kBndFast - 1))); };
        __syncthreads();
        for (int e = tid; e < kMaxCf * kBndPx; e += 256) {
            const int ch = e >> 7, i = e & (kBndPx - 1), r = pix_r(i), c = pix_c(i);
            xs[ch][i] = (ch < Cf && r < rows && c < cols) ? __ldg(A.x + (static_cast<size_t>(b) * Cf + ch) * hw + composed_offset(plane, r, c, A.H, A.W, Wc)) : 0.f;
        }
        __syncthreads();
        // four pixels' gradient loads in flight per thread (a pixel outside the plane contributes zeros)
        constexpr int NB = NPASS < 4 ? NPASS : 4;
#pragma unroll
        for (int p0 = 0; p0 < NPASS; p0 += NB) {
            float4 dv[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const int i = (p0 + k) * PPP + pslot, r = pix_r(i), c = pix_c(i);
                dv[k] = (r < rows && c < cols) ? __ldg(reinterpret_cast<const float4*>(dp + (static_cast<size_t>(r) * cols + c) * C0))
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const int i = (p0 + k) * PPP + pslot;
                dbr.x += dv[k].x; dbr.y += dv[k].y; dbr.z += dv[k].z; dbr.w += dv[k].w;
#pragma unroll
                for (int ch = 0; ch < kMaxCf; ++ch) {
                    const float xv = xs[ch][i];
                    dwr[ch].x = fmaf(dv[k].x, xv, dwr[ch].x); dwr[ch].y = fmaf(dv[k].y, xv, dwr[ch].y);
                    dwr[ch].z = fmaf(dv[k].z, xv, dwr[ch].z); dwr[ch].w = fmaf(dv[k].w, xv, dwr[ch].w);
                }
            }
        }
    }
#pragma unroll
    for (int ch = 0; ch < kMaxCf; ++ch)
        if (ch < Cf) {
            atomicAdd(&dwacc[ql * 4 + 0][ch], dwr[ch].x); atomicAdd(&dwacc[ql * 4 + 1][ch], dwr[ch].y);
            atomicAdd(&dwacc[ql * 4 + 2][ch], dwr[ch].z); atomicAdd(&dwacc[ql * 4 + 3][ch], dwr[ch].w);
        }
    atomicAdd(&dbacc[ql * 4 + 0], dbr.x); atomicAdd(&dbacc[ql * 4 + 1], dbr.y);
    atomicAdd(&dbacc[ql * 4 + 2], dbr.z); atomicAdd(&dbacc[ql * 4 + 3], dbr.w);
    __syncthreads();
    for (int e = tid; e < C0 * Cf; e += 256) atomicAdd(A.dw[plane] + e, dwacc[e / Cf][e % Cf]);
    for (int e = tid; e < C0; e += 256) atomicAdd(A.db[plane] + e, dbacc[e]);
}

// =====================================================================================
// Resampling adjoints (oracle/backward_ref.py::avgpool2_backward, bilinear_resize_backward).
// k_upcat_bwd: the concat input of a decoder block was cat[resize(up2(low)) (Cu), skip (Cs)].  Given its gradient dcat
//   [B][rows][cols][Cu+Cs]: scatters the first Cu channels back through the SAME bilinear weights the forward used (fp32 atomics
//   into dlow, zeroed before) and writes the skip part (+ the 2x2-average-pool adjoint of `dpool`, when the skip tensor was also
//   pooled into the next level) as dskip.  grid (slots, 3, B), block (Ct/4, NY)
// =====================================================================================
struct UpcatBwdArgs {
    TriCF dcat;
    TriDims dout, dlow;
    int Cu, Cs, B, do_up;
    int skip_only;             // 1: only the skip / pool-adjoint half (the up half is done by k_up2_bwd_gather)
    TriF dlow_g;               // [B][lrows][lcols][Cu] (+=, atomics)
    TriF dskip;                // [B][rows][cols][Cs]
    TriCF dpool;               // gradient of the pooled copy of the skip tensor [B][rows/2][cols/2][Cs], or nullptr
    int nslots;
};
__global__ void __launch_bounds__(256) k_upcat_bwd(UpcatBwdArgs A) {
    const int plane = blockIdx.y, b = blockIdx.z;
    const int tx = threadIdx.x, ty = threadIdx.y, NY = blockDim.y;
    const int orows = A.dout.rows[plane], ocols = A.dout.cols[plane], lrows = A.dlow.rows[plane], lcols = A.dlow.cols[plane];
    const int Ct = A.Cu + A.Cs, u4 = A.Cu / 4;
    const int npx = orows * ocols;
    const int ppc = (npx + A.nslots - 1) / A.nslots;
    const int p0 = blockIdx.x * ppc, p1 = min(npx, p0 + ppc);
    const float* gp = A.dcat.p[plane] + static_cast<size_t>(b) * npx * Ct;
    float* lp = A.dlow_g.p[plane] + static_cast<size_t>(b) * lrows * lcols * A.Cu;
    const int urows = A.do_up ? 2 * lrows : lrows, ucols = A.do_up ? 2 * lcols : lcols;
    const bool resize = urows != orows || ucols != ocols;
    auto scatter_low = [&](int rr, int cc, float wgt, const float4& v) {       // (rr, cc) on the low-resolution grid
        float* p = lp + (static_cast<size_t>(rr) * lcols + cc) * A.Cu + tx * 4;
        atomicAdd(p, wgt * v.x); atomicAdd(p + 1, wgt * v.y); atomicAdd(p + 2, wgt * v.z); atomicAdd(p + 3, wgt * v.w);
    };
    auto scatter_up = [&](int ur, int uc, float wgt, const float4& v) {       // (ur, uc) on the (virtual) x2 grid
        if (!A.do_up) {
            scatter_low(ur, uc, wgt, v);
            return;
        }
        int r0, r1, c0, c1;
        float lr, lc;
        bilin_src(ur, lrows, 0.5f, r0, r1, lr);
        bilin_src(uc, lcols, 0.5f, c0, c1, lc);
        scatter_low(r0, c0, wgt * (1.f - lr) * (1.f - lc), v);
        scatter_low(r0, c1, wgt * (1.f - lr) * lc, v);
        scatter_low(r1, c0, wgt * lr * (1.f - lc), v);
        scatter_low(r1, c1, wgt * lr * lc, v);
    };
    for (int px = p0 + ty; px < p1; px += NY) {
        const int r = px / ocols, c = px - r * ocols;
        if (tx < u4 && A.skip_only) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(gp + static_cast<size_t>(px) * Ct + tx * 4));
        if (tx < u4) {
            if (resize) {
                int r0, r1, c0, c1;
                float lr, lc;
                bilin_src(r, urows, static_cast<float>(urows) / static_cast<float>(orows), r0, r1, lr);
                bilin_src(c, ucols, static_cast<float>(ucols) / static_cast<float>(ocols), c0, c1, lc);
                scatter_up(r0, c0, (1.f - lr) * (1.f - lc), v);
                scatter_up(r0, c1, (1.f - lr) * lc, v);
                scatter_up(r1, c0, lr * (1.f - lc), v);
                scatter_up(r1, c1, lr * lc, v);
            } else {
                scatter_up(r, c, 1.f, v);
            }
        } else {
            float4 o = v;
            const float* dpp = A.dpool.p[plane];
            if (dpp) {
                const int pr = orows >> 1, pc = ocols >> 1;
                if ((r >> 1) < pr && (c >> 1) < pc) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(dpp + ((static_cast<size_t>(b) * pr + (r >> 1)) * pc + (c >> 1)) * A.Cs) + (tx - u4));
                    o.x = fmaf(0.25f, q.x, o.x); o.y = fmaf(0.25f, q.y, o.y); o.z = fmaf(0.25f, q.z, o.z); o.w = fmaf(0.25f, q.w, o.w);
                }
            }
            *(reinterpret_cast<float4*>(A.dskip.p[plane] + (static_cast<size_t>(b) * npx + px) * A.Cs) + (tx - u4)) = o;
        }
    }
}

// Gather form of the plain x2 bilinear adjoint (no resize step, i.e. even plane sizes): every LOW-resolution pixel sums the <= 4 x 4
// high-resolution pixels whose bilinear footprint contains it, with the weights bilin_src() gives the forward (so the clamped edges
// come out right by construction).  No atomics, deterministic.  dcat: [B][rows][cols][Ct] (first Cu channels), dlow: [B][lrows][lcols][Cu].
// grid (slots, 3, B), block (Cu/4, NY)
__global__ void __launch_bounds__(256) k_up2_bwd_gather(UpcatBwdArgs A) {
    const int plane = blockIdx.y, b = blockIdx.z;
    const int tx = threadIdx.x, ty = threadIdx.y, NY = blockDim.y;
    const int orows = A.dout.rows[plane], ocols = A.dout.cols[plane], lrows = A.dlow.rows[plane], lcols = A.dlow.cols[plane];
    const int Ct = A.Cu + A.Cs;
    const int npx = lrows * lcols;
    const int ppc = (npx + A.nslots - 1) / A.nslots;
    const int p0 = blockIdx.x * ppc, p1 = min(npx, p0 + ppc);
    const float* gp = A.dcat.p[plane] + static_cast<size_t>(b) * orows * ocols * Ct + tx * 4;
    float* lp = A.dlow_g.p[plane] + static_cast<size_t>(b) * npx * A.Cu + tx * 4;
    for (int px = p0 + ty; px < p1; px += NY) {
        const int r = px / lcols, c = px - r * lcols;
        float wr[4], wc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                 // candidate high-res rows 2r-1 .. 2r+2 / columns 2c-1 .. 2c+2
            const int hr = 2 * r - 1 + k, hc = 2 * c - 1 + k;
            int i0, i1;
            float l1;
            wr[k] = wc[k] = 0.f;
            if (hr >= 0 && hr < orows) {
                bilin_src(hr, lrows, 0.5f, i0, i1, l1);
                wr[k] = (i0 == r ? 1.f - l1 : 0.f) + (i1 == r ? l1 : 0.f);
            }
            if (hc >= 0 && hc < ocols) {
                bilin_src(hc, lcols, 0.5f, i0, i1, l1);
                wc[k] = (i0 == c ? 1.f - l1 : 0.f) + (i1 == c ? l1 : 0.f);
            }
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kr = 0; kr < 4; ++kr) {
            if (wr[kr] == 0.f) continue;
            const int hr = 2 * r - 1 + kr;
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                if (wc[kc] == 0.f) continue;
                const int hc = 2 * c - 1 + kc;
                const float w = wr[kr] * wc[kc];
                const float4 v = __ldg(reinterpret_cast<const float4*>(gp + (static_cast<size_t>(hr) * ocols + hc) * Ct));
                acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
            }
        }
        *reinterpret_cast<float4*>(lp + static_cast<size_t>(px) * A.Cu) = acc;
    }
}

// out = a (+ b) element-wise, fp32 (plain gradient joins).  n4 = elements / 4
__global__ void __launch_bounds__(256) k_grad_add(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out, long long n4) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float4 v = a[i];
        if (b) {
            const float4 w = b[i];
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        out[i] = v;
    }
}

// bias gradients: db[plane][c] = bias_sum[plane][c] (fp64 -> fp32); additive-embedding gradient dfilm[b][off + c] += bc_sum[b][c]
__global__ void __launch_bounds__(256) k_bias_fin(const double* __restrict__ bias_sum, int C, float* d0, float* d1, float* d2, const double* bc_sum, int B,
                                                  float* dfilm, int film_dim, int film_off, float* s0, float* s1, float* s2) {
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
        const int plane = i / C, c = i - plane * C;
        const float v = static_cast<float>(bias_sum[i]);
        float* d = plane == 0 ? d0 : (plane == 1 ? d1 : d2);
        d[c] = v;
        float* s = plane == 0 ? s0 : (plane == 1 ? s1 : s2);      // the skip conv's bias sees the same output gradient
        if (s) s[c] = v;
    }
    if (bc_sum && dfilm)
        for (int i = threadIdx.x; i < B * C; i += blockDim.x) {
            const int b = i / C, c = i - b * C;
            dfilm[static_cast<size_t>(b) * film_dim + film_off + c] += static_cast<float>(bc_sum[i]);
        }
}

// Operands of the backward GEMMs, re-packed on the device whenever the weights change.
//   dgrad of a 3x3 conv (pad 1) = the same 3x3 conv of dY with the taps flipped and the channel roles swapped:
//     wd[half][c][tap' * Cout + co] = split(W[co][c][2 - kh'][2 - kw'])        (own channels c < C only; Cw = stored in-channels)
//   dgrad of the 1x1 skip conv:  wsd[half][cs][co] = split(Ws[co][cs])
__global__ void __launch_bounds__(256) k_pack_dgrad(const float* __restrict__ w, int Cout, int Cw, int C, __half* __restrict__ out) {
    const int K = 9 * Cout;
    const long long n = static_cast<long long>(C) * K;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i / K), k = static_cast<int>(i - static_cast<long long>(c) * K);
        const int tap = k / Cout, co = k - tap * Cout;
        const int kh = 2 - tap / 3, kw = 2 - tap % 3;
        __half hi, lo;
        split_f16(w[(static_cast<size_t>(co) * Cw + c) * 9 + kh * 3 + kw], hi, lo);
        out[i] = hi;
        out[n + i] = lo;
    }
}
__global__ void __launch_bounds__(256) k_pack_dgrad_1x1(const float* __restrict__ ws, int Cout, int Cs, __half* __restrict__ out) {
    const long long n = static_cast<long long>(Cs) * Cout;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int cs = static_cast<int>(i / Cout), co = static_cast<int>(i - static_cast<long long>(cs) * Cout);
        __half hi, lo;
        split_f16(ws[static_cast<size_t>(co) * Cs + cs], hi, lo);
        out[i] = hi;
        out[n + i] = lo;
    }
}

// Forward operands re-packed on the device (same values, bit for bit, as the host packers pack_conv / pack_roll in s3d.cu): a
// training step changes every weight, and a host round trip per step would cost more than the step.
//   w_pack[half][co][tap*C + c] (+ [9C + cs] from the 1x1 skip) = split(W[co][c][tap])
__global__ void __launch_bounds__(256) k_pack_conv(const float* __restrict__ w, const float* __restrict__ wskip, int Cout, int Cw, int C, int Cs,
                                                   __half* __restrict__ out) {
    const int Ktot = 9 * C + Cs;
    const long long n = static_cast<long long>(Cout) * Ktot;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i / Ktot), k = static_cast<int>(i - static_cast<long long>(co) * Ktot);
        float v;
        if (k < 9 * C) {
            const int tap = k / C, c = k - tap * C;
            v = w[(static_cast<size_t>(co) * Cw + c) * 9 + tap];
        } else {
            v = wskip[static_cast<size_t>(co) * Cs + (k - 9 * C)];
        }
        __half hi, lo;
        split_f16(v, hi, lo);
        out[i] = hi;
        out[n + i] = lo;
    }
}
//   wc[(along*C + c)][cls*Cout + co] = sum over the taps `across` class cls keeps of W[co][g*C + c][kh][kw]   (fp32, pack_roll)
//   w16[half][cls*Cout + co][along*C + c] = split(wc)
__global__ void __launch_bounds__(256) k_pack_roll(const float* __restrict__ w, int Cout, int C, int g, int row_varying, float* __restrict__ wc,
                                                   __half* __restrict__ w16) {
    const int K = 3 * C, N = 4 * Cout, Cw = 3 * C;
    const long long n = static_cast<long long>(K) * N;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i / N), nn = static_cast<int>(i - static_cast<long long>(k) * N);
        const int along = k / C, c = k - along * C, cls = nn / Cout, co = nn - cls * Cout;
        float acc = 0.f;
#pragma unroll
        for (int across = 0; across < 3; ++across) {
            const bool keep = cls == 0 || (cls == 1 && across >= 1) || (cls == 2 && across <= 1) || (cls == 3 && across == 1);
            if (keep) {
                const int kh = row_varying ? along : across, kw = row_varying ? across : along;
                acc += w[((static_cast<size_t>(co) * Cw + g * C + c) * 3 + kh) * 3 + kw];
            }
        }
        wc[i] = acc;
        __half hi, lo;
        split_f16(acc, hi, lo);
        w16[static_cast<size_t>(nn) * K + k] = hi;
        w16[(static_cast<size_t>(N) + nn) * K + k] = lo;
    }
}
__global__ void k_vec_add(const float* a, const float* b, float* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + (b ? b[i] : 0.f);
}

// One launch for a whole refresh: a table of jobs (plain copies, bias sums and the packers above), blockIdx.y = job.  Every job
// reads the CALLER's tensors, so there is no ordering between the jobs of a launch.
struct PackJob {
    int kind;                  // 0 copy, 1 dst = src + src2, 2 pack_conv, 3 pack_roll, 4 pack_dgrad, 5 pack_dgrad_1x1, 6 pack_rollv
    int Cout, Cw, C, Cs, g, rowv;
    long long n;               // elements of the job's index space
    const float* src;
    const float* src2;
    void* dst;
    void* dst2;
};
__global__ void __launch_bounds__(256) k_pack_jobs(const PackJob* __restrict__ jobs) {
    const PackJob J = jobs[blockIdx.y];
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < J.n; i += stride) {
        if (J.kind == 0) {
            static_cast<float*>(J.dst)[i] = J.src[i];
        } else if (J.kind == 1) {
            static_cast<float*>(J.dst)[i] = J.src[i] + (J.src2 ? J.src2[i] : 0.f);
        } else if (J.kind == 2) {                      // k_pack_conv
            const int Ktot = 9 * J.C + J.Cs;
            const int co = static_cast<int>(i / Ktot), k = static_cast<int>(i - static_cast<long long>(co) * Ktot);
            float v;
            if (k < 9 * J.C) {
                const int tap = k / J.C, c = k - tap * J.C;
                v = J.src[(static_cast<size_t>(co) * J.Cw + c) * 9 + tap];
            } else {
                v = J.src2[static_cast<size_t>(co) * J.Cs + (k - 9 * J.C)];
            }
            __half hi, lo;
            split_f16(v, hi, lo);
            static_cast<__half*>(J.dst)[i] = hi;
            static_cast<__half*>(J.dst)[J.n + i] = lo;
        } else if (J.kind == 3) {                      // k_pack_roll
            const int K = 3 * J.C, N = 4 * J.Cout, Cw = 3 * J.C;
            const int k = static_cast<int>(i / N), nn = static_cast<int>(i - static_cast<long long>(k) * N);
            const int along = k / J.C, c = k - along * J.C, cls = nn / J.Cout, co = nn - cls * J.Cout;
            float acc = 0.f;
#pragma unroll
            for (int across = 0; across < 3; ++across) {
                const bool keep = cls == 0 || (cls == 1 && across >= 1) || (cls == 2 && across <= 1) || (cls == 3 && across == 1);
                if (keep) {
                    const int kh = J.rowv ? along : across, kw = J.rowv ? across : along;
                    acc += J.src[((static_cast<size_t>(co) * Cw + J.g * J.C + c) * 3 + kh) * 3 + kw];
                }
            }
            static_cast<float*>(J.dst)[i] = acc;
            __half hi, lo;
            split_f16(acc, hi, lo);
            __half* w16 = static_cast<__half*>(J.dst2);
            w16[static_cast<size_t>(nn) * K + k] = hi;
            w16[(static_cast<size_t>(N) + nn) * K + k] = lo;
        } else if (J.kind == 4) {                      // k_pack_dgrad
            const int K = 9 * J.Cout;
            const int c = static_cast<int>(i / K), k = static_cast<int>(i - static_cast<long long>(c) * K);
            const int tap = k / J.Cout, co = k - tap * J.Cout;
            const int kh = 2 - tap / 3, kw = 2 - tap % 3;
            __half hi, lo;
            split_f16(J.src[(static_cast<size_t>(co) * J.Cw + c) * 9 + kh * 3 + kw], hi, lo);
            static_cast<__half*>(J.dst)[i] = hi;
            static_cast<__half*>(J.dst)[J.n + i] = lo;
        } else if (J.kind == 5) {                      // k_pack_dgrad_1x1
            const int cs = static_cast<int>(i / J.Cout), co = static_cast<int>(i - static_cast<long long>(cs) * J.Cout);
            __half hi, lo;
            split_f16(J.src[static_cast<size_t>(co) * J.Cs + cs], hi, lo);
            static_cast<__half*>(J.dst)[i] = hi;
            static_cast<__half*>(J.dst)[J.n + i] = lo;
        } else {                                       // k_pack_rollv
            const int c = static_cast<int>(i % J.C), co = static_cast<int>((i / J.C) % J.Cout), t = static_cast<int>(i / (static_cast<long long>(J.C) * J.Cout));
            const int al = t / 3, ac = t - al * 3;
            static_cast<float*>(J.dst)[i] = J.src[(static_cast<size_t>(co) * 3 * J.C + J.g * J.C + c) * 9 + (J.rowv ? al * 3 + ac : ac * 3 + al)];
        }
    }
}

// final pass: every gradient was carried times the loss scale
__global__ void __launch_bounds__(256) k_unscale(float* __restrict__ p, long long n, const unsigned int* amax) {
    const float inv = 1.f / loss_scale(amax);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) p[i] *= inv;
}

}  // namespace s3d
