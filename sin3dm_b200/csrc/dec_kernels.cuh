// Triplane decoder kernels: latent planes -> up-convolved feature planes (once per latent), then
// points -> bilinear gather from the three planes -> skip-concat MLP -> SDF + texture.
//
//   reference: AutoEncoderGroupSkip.decode            src/encoding/networks.py:192-220
//              sample_feature_plane2D                 src/encoding/networks.py:182-190 (F.grid_sample bilinear / border)
//              TriplaneGroupResnetBlock.forward       src/encoding/blocks.py:232-256
//              DecoderMLPSkipConcat.forward           src/encoding/blocks.py:84-91
//              ShapeAutoEncoder.decode_batch / _grid  src/encoding/model.py:319-349
//
// The grouped convolutions of the reference run on a zero-padded channel-wise composition of the three planes
// (blocks.py:164-177); a group never mixes planes and the padding is zero before and after the norm + SiLU, so each plane
// is simply convolved on its own with zero padding — which is what these kernels do.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace s3d {

constexpr int kDecUp = 64;          // feat_channel_up (parser_util.py:24); the kernels are specialised for it
constexpr int kDecHid = 256;        // mlp hidden width of the tensor-core kernel (parser_util.py:25)
constexpr float kInEps = 1e-6f;     // blocks.py:213-215  InstanceNorm2d(eps=1e-6)
constexpr int kDecTile = 16;        // conv tile edge (pixels)
constexpr int kDecMaxKs = 7;

struct DecPlanes {
    const float* F[3];   // [rows][cols][CF] fp32: channels [0,64) geo, [64,128) tex
    int rows[3], cols[3];
    int CF;
};

// ------------------------------------------------------------------------------------------------------------------
// k_dec_conv: ks x ks convolution of one plane, 64 output channels, CUDA cores (fp32).  Runs twice per branch and latent.
//   MODE 0: input = `cin` channels of the NCHW latent plane (no activation); writes h1 = conv + bias (NHWC, 64 ch) and the
//           1x1 shortcut (+ bias) into the feature plane F.
//   MODE 1: input = h1 (NHWC, 64 ch) with the InstanceNorm affine + SiLU applied on load; F += conv + bias.
// grid (tiles of the three planes), block 256 = one 16 x 16 pixel tile, thread = pixel, 64 accumulators.
// ------------------------------------------------------------------------------------------------------------------
struct DecConvArgs {
    const float* x[3];       // MODE 0: latent plane [c_total][rows][cols]; MODE 1 / 2: [rows][cols][64]
    int c0, cin;             // MODE 0: first channel / channel count of this branch inside the latent
    const float* w;          // [3 planes][ks*ks][cin][64]
    const float* bias;       // [3][64]
    const float* ws;         // MODE 0: shortcut [3][cin][64]   (nullptr: identity shortcut, cin == 64)
    const float* bs;         // MODE 0: shortcut bias [3][64]
    const float* coef;       // MODE 1 / 2: [3][64][2] = (gamma * rstd, beta - mean * gamma * rstd) of the input
    float* h1[3];            // MODE 0 / 2 output
    float* F[3];             // feature plane [rows][cols][CF]
    int CF, foff;            // channel stride / channel offset of this branch in F
    int rows[3], cols[3];
    int tiles_x[3], tile_start[4];
    int ks;
};

// MODE 0: in_layers conv of a block without input norm / activation + its 1x1 shortcut (raw latent in).
// MODE 1: InstanceNorm + SiLU + out_layers conv, added onto the shortcut already in F.
// MODE 2: in_layers of a block WITH input norm and activation (blocks.py:199-204, 233-236; AutoEncoderGroupPBR's second texture
//         block): h1 = conv(SiLU(norm(x))), and the identity shortcut norm(x) goes to F.
template <int MODE>
__global__ void __launch_bounds__(256) k_dec_conv(const DecConvArgs A) {
    extern __shared__ float dsm[];
    int plane = 0;
#pragma unroll
    for (int p = 1; p < 3; ++p)
        if (static_cast<int>(blockIdx.x) >= A.tile_start[p]) plane = p;
    const int ip = blockIdx.x - A.tile_start[plane];
    const int ty0 = (ip / A.tiles_x[plane]) * kDecTile, tx0 = (ip % A.tiles_x[plane]) * kDecTile;
    const int rows = A.rows[plane], cols = A.cols[plane];
    const int ks = A.ks, pad = (ks - 1) / 2, hw = kDecTile + ks - 1, hpix = hw * hw, hstride = hpix | 1;
    const int cin = MODE == 0 ? A.cin : kDecUp;
    float* sIn = dsm;                          // [cin][hstride]
    float* sW = dsm + cin * hstride;           // [cin][64] of the current tap
    const int tid = threadIdx.x;

    // ---- input halo tile
    if (MODE == 0) {
        const float* xp = A.x[plane] + static_cast<size_t>(A.c0) * rows * cols;
        for (int i = tid; i < cin * hpix; i += 256) {
            const int ci = i / hpix, q = i - ci * hpix;
            const int r = ty0 - pad + q / hw, c = tx0 - pad + q % hw;
            sIn[ci * hstride + q] = (r >= 0 && r < rows && c >= 0 && c < cols) ? __ldg(xp + (static_cast<size_t>(ci) * rows + r) * cols + c) : 0.f;
        }
    } else {
        const int ch = tid & 63;
        const float ca = A.coef[(plane * 64 + ch) * 2], cb = A.coef[(plane * 64 + ch) * 2 + 1];
        const float* hp = A.x[plane];
        for (int q = tid >> 6; q < hpix; q += 4) {
            const int r = ty0 - pad + q / hw, c = tx0 - pad + q % hw;
            float v = 0.f;                     // zero padding applies to the activated tensor (blocks.py:248-251)
            if (r >= 0 && r < rows && c >= 0 && c < cols) {
                const float h = __ldg(hp + (static_cast<size_t>(r) * cols + c) * 64 + ch);
                const float n = fmaf(h, ca, cb);
                v = n / (1.f + expf(-n));
            }
            sIn[ch * hstride + q] = v;
        }
    }
    const int py = tid >> 4, px = tid & 15;
    const int r = ty0 + py, c = tx0 + px;
    const bool valid = r < rows && c < cols;
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;

    const float* wp = A.w + static_cast<size_t>(plane) * ks * ks * cin * 64;
    for (int tap = 0; tap < ks * ks; ++tap) {
        __syncthreads();                       // previous tap's weights consumed (first pass: halo tile complete)
        for (int i = tid; i < cin * 16; i += 256)
            reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(wp + static_cast<size_t>(tap) * cin * 64) + i);
        __syncthreads();
        const int kh = tap / ks, kw = tap - kh * ks;
        const float* ip0 = sIn + (py + kh) * hw + px + kw;
        for (int ci = 0; ci < cin; ++ci) {
            const float a = ip0[ci * hstride];
            const float4* w4 = reinterpret_cast<const float4*>(sW + ci * 64);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 w = w4[j];
                acc[4 * j + 0] = fmaf(a, w.x, acc[4 * j + 0]);
                acc[4 * j + 1] = fmaf(a, w.y, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(a, w.z, acc[4 * j + 2]);
                acc[4 * j + 3] = fmaf(a, w.w, acc[4 * j + 3]);
            }
        }
    }
    const float4* b4 = reinterpret_cast<const float4*>(A.bias + plane * 64);
    if (MODE == 0) {
        if (valid) {
            float4* o = reinterpret_cast<float4*>(A.h1[plane] + (static_cast<size_t>(r) * cols + c) * 64);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 b = __ldg(b4 + j);
                o[j] = make_float4(acc[4 * j] + b.x, acc[4 * j + 1] + b.y, acc[4 * j + 2] + b.z, acc[4 * j + 3] + b.w);
            }
        }
        // 1x1 shortcut on the raw input (blocks.py:227-230, 254)
        __syncthreads();
        for (int i = tid; i < cin * 16; i += 256)
            reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(A.ws + static_cast<size_t>(plane) * cin * 64) + i);
        __syncthreads();
        if (valid) {
#pragma unroll
            for (int i = 0; i < 64; ++i) acc[i] = 0.f;
            const float* ip0 = sIn + (py + pad) * hw + px + pad;
            for (int ci = 0; ci < cin; ++ci) {
                const float a = ip0[ci * hstride];
                const float4* w4 = reinterpret_cast<const float4*>(sW + ci * 64);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 w = w4[j];
                    acc[4 * j + 0] = fmaf(a, w.x, acc[4 * j + 0]);
                    acc[4 * j + 1] = fmaf(a, w.y, acc[4 * j + 1]);
                    acc[4 * j + 2] = fmaf(a, w.z, acc[4 * j + 2]);
                    acc[4 * j + 3] = fmaf(a, w.w, acc[4 * j + 3]);
                }
            }
            const float4* s4 = reinterpret_cast<const float4*>(A.bs + plane * 64);
            float4* o = reinterpret_cast<float4*>(A.F[plane] + (static_cast<size_t>(r) * cols + c) * A.CF + A.foff);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 b = __ldg(s4 + j);
                o[j] = make_float4(acc[4 * j] + b.x, acc[4 * j + 1] + b.y, acc[4 * j + 2] + b.z, acc[4 * j + 3] + b.w);
            }
        }
    } else if (MODE == 2) {
        if (valid) {
            float4* o = reinterpret_cast<float4*>(A.h1[plane] + (static_cast<size_t>(r) * cols + c) * 64);
            const float4* xi = reinterpret_cast<const float4*>(A.x[plane] + (static_cast<size_t>(r) * cols + c) * 64);
            const float4* cf = reinterpret_cast<const float4*>(A.coef + plane * 64 * 2);      // (a, b) pairs: two channels per float4
            float4* f = reinterpret_cast<float4*>(A.F[plane] + (static_cast<size_t>(r) * cols + c) * A.CF + A.foff);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 b = __ldg(b4 + j);
                o[j] = make_float4(acc[4 * j] + b.x, acc[4 * j + 1] + b.y, acc[4 * j + 2] + b.z, acc[4 * j + 3] + b.w);
                const float4 xv = __ldg(xi + j), c0 = __ldg(cf + 2 * j), c1 = __ldg(cf + 2 * j + 1);
                f[j] = make_float4(fmaf(xv.x, c0.x, c0.y), fmaf(xv.y, c0.z, c0.w), fmaf(xv.z, c1.x, c1.y), fmaf(xv.w, c1.z, c1.w));
            }
        }
    } else if (valid) {
        float4* o = reinterpret_cast<float4*>(A.F[plane] + (static_cast<size_t>(r) * cols + c) * A.CF + A.foff);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float4 b = __ldg(b4 + j);
            float4 v = o[j];
            // h + shortcut(x): the conv result (with its bias) is formed first, then added (blocks.py:253-254)
            v.x = (acc[4 * j] + b.x) + v.x;
            v.y = (acc[4 * j + 1] + b.y) + v.y;
            v.z = (acc[4 * j + 2] + b.z) + v.z;
            v.w = (acc[4 * j + 3] + b.w) + v.w;
            o[j] = v;
        }
    }
}

// InstanceNorm statistics of h1 (per plane, per channel, biased variance): deterministic two-level reduction in fp64.
// grid (strips, 3), block 256 = 64 channels x 4 pixel phases; partial[plane][strip][64][2]
constexpr int kDecStripPix = 256;
__global__ void __launch_bounds__(256) k_dec_in_stats(const float* h0, const float* h1, const float* h2, int n0, int n1, int n2,
                                                      double* __restrict__ partial, int max_strips) {
    const int plane = blockIdx.y;
    const float* h = plane == 0 ? h0 : (plane == 1 ? h1 : h2);
    const int n = plane == 0 ? n0 : (plane == 1 ? n1 : n2);
    const int p0 = blockIdx.x * kDecStripPix;
    if (p0 >= n) return;
    const int ch = threadIdx.x & 63, ph = threadIdx.x >> 6;
    double s = 0.0, q = 0.0;
    for (int p = p0 + ph; p < min(p0 + kDecStripPix, n); p += 4) {
        const double v = static_cast<double>(__ldg(h + static_cast<size_t>(p) * 64 + ch));
        s += v;
        q += v * v;
    }
    __shared__ double sm[2][4][64];
    sm[0][ph][ch] = s;
    sm[1][ph][ch] = q;
    __syncthreads();
    if (ph == 0) {
        double* o = partial + ((static_cast<size_t>(plane) * max_strips + blockIdx.x) * 64 + ch) * 2;
        o[0] = (sm[0][0][ch] + sm[0][1][ch]) + (sm[0][2][ch] + sm[0][3][ch]);
        o[1] = (sm[1][0][ch] + sm[1][1][ch]) + (sm[1][2][ch] + sm[1][3][ch]);
    }
}
// grid 3, block 64: coef[plane][ch] = (gamma * rstd, beta - mean * gamma * rstd)
__global__ void k_dec_in_finalize(const double* __restrict__ partial, int max_strips, int n0, int n1, int n2,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ coef) {
    const int plane = blockIdx.x, ch = threadIdx.x;
    const int n = plane == 0 ? n0 : (plane == 1 ? n1 : n2);
    const int strips = (n + kDecStripPix - 1) / kDecStripPix;
    double s = 0.0, q = 0.0;
    for (int i = 0; i < strips; ++i) {
        const double* p = partial + ((static_cast<size_t>(plane) * max_strips + i) * 64 + ch) * 2;
        s += p[0];
        q += p[1];
    }
    const double mean = s / n, var = fmax(q / n - mean * mean, 0.0);
    const double a = static_cast<double>(gamma[plane * 64 + ch]) / sqrt(var + static_cast<double>(kInEps));
    coef[(plane * 64 + ch) * 2] = static_cast<float>(a);
    coef[(plane * 64 + ch) * 2 + 1] = static_cast<float>(static_cast<double>(beta[plane * 64 + ch]) - mean * a);
}

// ------------------------------------------------------------------------------------------------------------------
// Point sampling shared by both MLP kernels.
// ------------------------------------------------------------------------------------------------------------------
struct DecPoints {
    const float* pts;              // [n][3], or nullptr: grid mode
    const float *xs, *ys, *zs;     // grid mode: coordinate vectors (utils3d.py:21-23), point g = (ix*ny + iy)*nz + iz
    int ny, nz;
    long long n;
    float amin[3], asize[3];       // aabb min, (aabb max - aabb min)
};
__device__ __forceinline__ void dec_point(const DecPoints& P, long long g, float (&xn)[3]) {
    float p[3];
    if (P.pts) {
        p[0] = __ldg(P.pts + g * 3);
        p[1] = __ldg(P.pts + g * 3 + 1);
        p[2] = __ldg(P.pts + g * 3 + 2);
    } else {
        const int iz = static_cast<int>(g % P.nz);
        const long long t = g / P.nz;
        p[0] = __ldg(P.xs + t / P.ny);
        p[1] = __ldg(P.ys + t % P.ny);
        p[2] = __ldg(P.zs + iz);
    }
    // networks.py:196  x = 2 * (x - aabb[:3]) / (aabb[3:] - aabb[:3]) - 1
#pragma unroll
    for (int i = 0; i < 3; ++i) xn[i] = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fsub_rn(p[i], P.amin[i])), P.asize[i]), 1.f);
}
// F.grid_sample(bilinear, padding_mode="border", align_corners=False) source index + weights along one axis
__device__ __forceinline__ void dec_axis(float coord, int size, int& i0, int& i1, float& w0, float& w1) {
    float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(coord, 1.f), static_cast<float>(size)), 1.f), 0.5f);
    ix = fminf(static_cast<float>(size - 1), fmaxf(ix, 0.f));
    const float f = floorf(ix);
    i0 = static_cast<int>(f);
    i1 = min(i0 + 1, size - 1);          // weight of the clamped neighbour is exactly 0 when it would fall outside
    w1 = ix - f;
    w0 = (f + 1.f) - ix;
}
// plane p is sampled with rows <- coordinate kRowAxis[p], cols <- coordinate kColAxis[p] (networks.py:201 + .flip(-1))
__device__ __constant__ int kDecRowAxis[3] = {0, 0, 1};
__device__ __constant__ int kDecColAxis[3] = {1, 2, 2};

struct DecCorners {
    int off[12];       // pixel index (r * cols + c) of the 4 corners of the 3 planes
    float w[12];
};
__device__ __forceinline__ void dec_corners(const DecPlanes& G, const float (&xn)[3], DecCorners& K) {
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        int r0, r1, c0, c1;
        float wr0, wr1, wc0, wc1;
        dec_axis(xn[p == 2 ? 1 : 0], G.rows[p], r0, r1, wr0, wr1);
        dec_axis(xn[p == 0 ? 1 : 2], G.cols[p], c0, c1, wc0, wc1);
        K.off[4 * p + 0] = r0 * G.cols[p] + c0;  K.w[4 * p + 0] = wc0 * wr0;     // nw
        K.off[4 * p + 1] = r0 * G.cols[p] + c1;  K.w[4 * p + 1] = wc1 * wr0;     // ne
        K.off[4 * p + 2] = r1 * G.cols[p] + c0;  K.w[4 * p + 2] = wc0 * wr1;     // sw
        K.off[4 * p + 3] = r1 * G.cols[p] + c1;  K.w[4 * p + 3] = wc1 * wr1;     // se
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_dec_mlp_ffma: CUDA-core fp32 decoder (mlp_impl = 1): exact-precision cross-check of the tensor-core kernel and the
// path for MLP shapes the tensor-core kernel is not specialised for.  block 256 = 32 points; thread = hidden unit.
// ------------------------------------------------------------------------------------------------------------------
struct DecMlpF32 {
    const float* wt[8];    // layer l: transposed weights [K_l][N_l]
    const float* b[8];
    int K[8], N[8];
    int n_first;           // layers before the skip concat (first_layers)
    int n_layers;          // all linear layers
};
struct DecArgs {
    DecPlanes G;
    DecPoints P;
    int nb;                // MLP heads evaluated by this launch: 1 or 2
    int oc;                // output channels per point (row stride of `out`)
    int tex_channels;
    int clamp_tex;         // decode_batch's clamp of every channel but the first (model.py:332)
    float* out;            // [n][oc]
    // per head: first feature channel it reads in F (0 = geo planes, 64 = tex planes), first output column, sigmoid on the output
    // (AutoEncoderGroupSkip / V3: {0, 64}, {0, 1}, {0, 1}; AutoEncoderGroupPBR runs (geo, rgb) then (mr, normal), no sigmoid)
    int feat_off[2], out_col[2], sigmoid[2];
};

__global__ void __launch_bounds__(256) k_dec_mlp_ffma(const DecArgs A, const DecMlpF32 M0, const DecMlpF32 M1, int hid) {
    extern __shared__ float dsm[];
    float* sX = dsm;                    // [32][64]
    float* sA = sX + 32 * kDecUp;       // [32][hid]
    float* sB = sA + 32 * hid;          // [32][hid]
    const int tid = threadIdx.x;
    const long long g0 = static_cast<long long>(blockIdx.x) * 32;
    for (int br = 0; br < A.nb; ++br) {
        const DecMlpF32& M = br == 0 ? M0 : M1;
        __syncthreads();
        {   // gather: 8 lanes per point, lane j owns channels [8j, 8j+8)
            const int pt = tid >> 3, j = tid & 7;
            const long long g = min(g0 + pt, A.P.n - 1);
            float xn[3];
            dec_point(A.P, g, xn);
            DecCorners K;
            dec_corners(A.G, xn, K);
            float h[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) h[e] = 0.f;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                float s[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) s[e] = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4* f = reinterpret_cast<const float4*>(A.G.F[p] + static_cast<size_t>(K.off[4 * p + k]) * A.G.CF + A.feat_off[br] + j * 8);
                    const float4 a = __ldg(f), b = __ldg(f + 1);
                    const float w = K.w[4 * p + k];
                    s[0] = fmaf(a.x, w, s[0]); s[1] = fmaf(a.y, w, s[1]); s[2] = fmaf(a.z, w, s[2]); s[3] = fmaf(a.w, w, s[3]);
                    s[4] = fmaf(b.x, w, s[4]); s[5] = fmaf(b.y, w, s[5]); s[6] = fmaf(b.z, w, s[6]); s[7] = fmaf(b.w, w, s[7]);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] += s[e];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) sX[pt * kDecUp + j * 8 + e] = h[e];
        }
        __syncthreads();
        const float* in = sX;
        int in_stride = kDecUp;
        float* bufs[2] = {sA, sB};
        int cur = 0;
        for (int l = 0; l < M.n_layers; ++l) {
            const int K = M.K[l], N = M.N[l];
            const bool last = l == M.n_layers - 1;
            const bool concat = l == M.n_first;          // input = cat[x, h]  (blocks.py:88)
            float* outb = bufs[cur];
            if (tid < N) {
                float acc[32];
#pragma unroll
                for (int p = 0; p < 32; ++p) acc[p] = 0.f;
                const float* wt = M.wt[l];
                int k = 0;
                if (concat) {
                    for (; k < kDecUp; ++k) {
                        const float w = __ldg(wt + static_cast<size_t>(k) * N + tid);
#pragma unroll
                        for (int p = 0; p < 32; ++p) acc[p] = fmaf(sX[p * kDecUp + k], w, acc[p]);
                    }
                }
                const int kofs = concat ? kDecUp : 0;
                for (; k < K; ++k) {
                    const float w = __ldg(wt + static_cast<size_t>(k) * N + tid);
#pragma unroll
                    for (int p = 0; p < 32; ++p) acc[p] = fmaf(in[p * in_stride + k - kofs], w, acc[p]);
                }
                const float bias = __ldg(M.b[l] + tid);
                if (!last) {
#pragma unroll
                    for (int p = 0; p < 32; ++p) outb[p * hid + tid] = fmaxf(acc[p] + bias, 0.f);
                } else {
#pragma unroll
                    for (int p = 0; p < 32; ++p) {
                        float v = acc[p] + bias;
                        if (A.sigmoid[br]) v = 1.f / (1.f + expf(-v));      // networks.py:216  .sigmoid()
                        if (A.clamp_tex && A.out_col[br] + tid > 0) v = fminf(fmaxf(v, 0.f), 1.f);
                        if (g0 + p < A.P.n) A.out[(g0 + p) * A.oc + A.out_col[br] + tid] = v;
                    }
                }
            }
            __syncthreads();
            in = outb;
            in_stride = hid;
            cur ^= 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// k_dec_mlp_tc: the decoder on the sm_100a tensor cores.  Persistent, one CTA per SM, tile = 128 points.
//
//   per tile and branch:  gather (X: 128 x 64)  ->  L1: X*W1 -> ReLU -> L2 -> ReLU -> L3 -> ReLU
//                         -> L4: [X | H3]*W4 -> ReLU -> L5 -> ReLU -> 256 -> out (CUDA cores, in L5's epilogue)
//
// Activations never leave the SM: every layer's epilogue (TMEM -> registers -> +bias, ReLU -> fp16 (hi, lo) -> shared
// memory, written directly in the 128-byte-swizzled K-major layout tcgen05.mma reads) produces the next layer's A operand in
// place.  Weights ([256 x K] per layer, fp16 (hi, lo), pre-scaled by a power of two so that lo stays in fp16 normal range
// unscaled) stream through a TMA ring as [128 N x 64 K] tiles; they are the same for every tile and stay L2-resident.
// NSPLIT == 3:  D += Ah*Bh + Al*Bh + Ah*Bl  (fp32 accumulate in TMEM; error ~2^-21 per operand, i.e. fp32-grade).
// NSPLIT == 1:  D += Ah*Bh.
// Warps: 0 = TMA producer (weights), 1 = TMEM owner + MMA issuer, 2..9 = gather + epilogues (thread = point = TMEM lane; the two
// warps of a TMEM lane quarter take alternate 32-column blocks, so both work on the chunk the MMA warp is waiting for).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kDecPts = 128;
constexpr int kDecTcThreads = 320;               // TMA warp + MMA warp + 8 gather/epilogue warps
constexpr int kDecChunkBytes = kDecPts * 128;       // one 64-wide K chunk of an A operand: 128 rows x 128 B
constexpr int kDecBSlotBytes = 128 * 128;           // [128 N rows][64 K] fp16
constexpr int kDecLayers = 5;                       // tensor-core layers per branch (the 256 -> out layer runs in L5's epilogue)

template <int NSPLIT>
struct DecTcCfg {
    static constexpr int kParts = NSPLIT == 3 ? 2 : 1;
    static constexpr int kXBytes = kParts * kDecChunkBytes;
    static constexpr int kActBytes = kParts * 4 * kDecChunkBytes;
    static constexpr int kBSlots = NSPLIT == 3 ? 4 : 8;
    static constexpr int kRingBytes = kBSlots * kDecBSlotBytes;
    static constexpr int kSmemBytes = kXBytes + kActBytes + kRingBytes + 1024 /*align*/ + 256 /*barriers*/;
};

struct DecTcMaps {
    CUtensorMap w[2][kDecLayers];     // [branch][layer]: (K, 256, 2) fp16, box {64, 128, 1}
};
struct DecTcArgs {
    DecArgs D;
    float inv_scale[2][kDecLayers];       // 1 / (power-of-two weight scale)
    int n_out[2];
    int skip;                             // 1: the fourth layer reads cat[X, H3] (DecoderMLPSkipConcat); 0: plain DecoderMLP
    long long n_tiles;
};
// Epilogue constants travel as a kernel parameter (18.5 KB of the 32 KB parameter space): every thread of a warp reads the same
// element, so they come out of the constant cache instead of stalling the epilogue on global loads (ncu r1: 38 % of the
// epilogue warps' samples sat behind the bias __ldg).
struct DecTcConst {
    float bias[2][kDecLayers][kDecHid];   // [branch][layer][channel]
    float w_last[2][4][kDecHid];          // second_layers' last Linear, [branch][out][channel]
    float b_last[2][4];
};

__device__ __forceinline__ void split_f16_plain(float v, __half& hi, __half& lo) {
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

template <int NSPLIT>
__global__ void __launch_bounds__(kDecTcThreads, 1) k_dec_mlp_tc(const __grid_constant__ DecTcMaps M, const DecTcArgs A,
                                                                  const __grid_constant__ DecTcConst C) {
    using Cfg = DecTcCfg<NSPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sx = smem;                                  // X hi | X lo
    uint8_t* sact = smem + Cfg::kXBytes;                 // ACT hi (4 chunks) | ACT lo (4 chunks)
    uint8_t* sb = sact + Cfg::kActBytes;                 // weight ring
    uint64_t* fullB = reinterpret_cast<uint64_t*>(sb + Cfg::kRingBytes);
    uint64_t* emptyB = fullB + Cfg::kBSlots;
    uint64_t* x_ready = emptyB + Cfg::kBSlots;           // gather -> MMA warp: X (layer 0 / the concat chunk) is in shared memory
    uint64_t* chunk_ready = x_ready + 1;                 // [4] epilogue -> MMA warp: 64-channel chunk c of the next A operand is written
    uint64_t* d_full = chunk_ready + 4;                  // MMA warp -> epilogue warps: the layer's accumulator is complete
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(d_full + 1);
    constexpr uint32_t kXLo = kDecChunkBytes, kActLo = 4 * kDecChunkBytes;

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int nb = A.D.nb;

    if (warp == 0 && lane == 0) {
        for (int br = 0; br < nb; ++br)
            for (int l = 0; l < kDecLayers; ++l) ptx::prefetch_tmap(&M.w[br][l]);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < Cfg::kBSlots; ++s) {
                ptx::mbar_init(&fullB[s], 1);
                ptx::mbar_init(&emptyB[s], 1);
            }
            ptx::mbar_init(x_ready, 8);
            for (int c = 0; c < 4; ++c) ptx::mbar_init(&chunk_ready[c], 8);
            ptx::mbar_init(d_full, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc<512>(tmem_ptr_smem);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // K chunks of layer l: layer 0 reads X, layer 3 reads [X | ACT], the others ACT
    const bool skip = A.skip != 0;
    auto n_chunks = [skip](int l) { return l == 0 ? 1 : ((l == 3 && skip) ? 5 : 4); };

    if (warp == 0) {
        // ===================== TMA producer: weight tiles, in the order the MMA warp consumes them =====================
        int gb = 0;
        for (long long t = blockIdx.x; t < A.n_tiles; t += gridDim.x)
            for (int br = 0; br < nb; ++br)
                for (int l = 0; l < kDecLayers; ++l) {
                    const int nk = n_chunks(l);
                    for (int kc = 0; kc < nk; ++kc)
                        for (int nh = 0; nh < 2; ++nh)
                            for (int part = 0; part < Cfg::kParts; ++part, ++gb) {
                                const int s = gb % Cfg::kBSlots;
                                ptx::mbar_wait(&emptyB[s], ((gb / Cfg::kBSlots) & 1) ^ 1);
                                if (ptx::elect_one()) {
                                    ptx::mbar_arrive_expect_tx(&fullB[s], kDecBSlotBytes);
                                    ptx::tma_load_3d(sb + s * kDecBSlotBytes, &M.w[br][l], &fullB[s], kc * 64, nh * 128, part);
                                }
                                __syncwarp();
                            }
                }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = ptx::make_idesc_f16(128, 128);
        // The A operand is handed over per 64-channel chunk: layer l+1's MMAs over chunk c start as soon as layer l's epilogue
        // has written it, so the tensor pipe works under the rest of that epilogue instead of after it.
        int gb = 0;
        uint32_t xit = 0;               // gathers consumed so far (phase of x_ready)
        for (long long t = blockIdx.x; t < A.n_tiles; t += gridDim.x)
            for (int br = 0; br < nb; ++br, ++xit)
                for (int l = 0; l < kDecLayers; ++l) {
                    const uint32_t d = tmem_base + (l & 1) * 256;
                    const int nk = n_chunks(l);
                    for (int kc = 0; kc < nk; ++kc) {
                        // A chunk: X for the first chunk of layers 0 and 3, else ACT chunk
                        const bool from_x = (l == 0) || (skip && l == 3 && kc == 0);
                        const int ac = (skip && l == 3) ? kc - 1 : kc;
                        if (l == 0) ptx::mbar_wait(x_ready, xit & 1);
                        else if (!from_x) ptx::mbar_wait(&chunk_ready[ac], (l - 1) & 1);     // 4 producing epilogues per branch
                        ptx::tc_fence_after();
                        const uint32_t a_hi = ptx::smem_u32(from_x ? sx : sact + ac * kDecChunkBytes);
                        const uint32_t a_lo = ptx::smem_u32(from_x ? sx + kXLo : sact + kActLo + ac * kDecChunkBytes);
                        for (int nh = 0; nh < 2; ++nh)
                            for (int part = 0; part < Cfg::kParts; ++part, ++gb) {
                                const int s = gb % Cfg::kBSlots;
                                ptx::mbar_wait(&fullB[s], (gb / Cfg::kBSlots) & 1);
                                ptx::tc_fence_after();
                                const uint64_t da_hi = ptx::make_sw128_desc1024(a_hi), da_lo = ptx::make_sw128_desc1024(a_lo);
                                const uint64_t db = ptx::make_sw128_desc1024(ptx::smem_u32(sb + s * kDecBSlotBytes));
                                const uint32_t dn = d + nh * 128;
                                if (ptx::elect_one()) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const uint64_t ko = static_cast<uint64_t>((k * 32) >> 4);
                                        if (part == 0) {
                                            ptx::umma_f16(dn, da_hi + ko, db + ko, idesc, (kc == 0 && k == 0) ? 0u : 1u);
                                            if (NSPLIT == 3) ptx::umma_f16(dn, da_lo + ko, db + ko, idesc, 1u);
                                        } else {
                                            ptx::umma_f16(dn, da_hi + ko, db + ko, idesc, 1u);
                                        }
                                    }
                                    ptx::umma_commit(&emptyB[s]);
                                }
                                __syncwarp();
                            }
                    }
                    if (ptx::elect_one()) ptx::umma_commit(d_full);
                    __syncwarp();
                }
    } else {
        // ===================== gather + epilogues (warps 2..9) =====================
        const int quarter = warp & 3;                      // TMEM lane quarter this warp may read (warp id % 4)
        const int half = (warp - 2) >> 2;                  // which of the quarter's two warps
        const int m = quarter * 32 + lane;                 // point of the tile == TMEM lane == A operand row
        uint32_t dphase = 0;                               // d_full completions consumed so far
        const uint32_t sw = static_cast<uint32_t>(m & 7);
        const uint32_t sact_u32 = ptx::smem_u32(sact);
        for (long long t = blockIdx.x; t < A.n_tiles; t += gridDim.x) {
            const long long g0 = t * kDecPts;
            for (int br = 0; br < nb; ++br) {
                // ---- gather: 8 lanes per point (lane j owns channels [8j, 8j+8) of the branch), 4 points per warp pass
                {
                    const int j = lane & 7;
#pragma unroll 1
                    for (int pass = half * 4; pass < half * 4 + 4; ++pass) {
                        const int pm = quarter * 32 + pass * 4 + (lane >> 3);
                        const long long g = min(g0 + pm, A.D.P.n - 1);
                        float xn[3];
                        dec_point(A.D.P, g, xn);
                        DecCorners K;
                        dec_corners(A.D.G, xn, K);
                        float4 fa[12], fb[12];
#pragma unroll
                        for (int c = 0; c < 12; ++c) {
                            const float4* f = reinterpret_cast<const float4*>(A.D.G.F[c >> 2] + static_cast<size_t>(K.off[c]) * A.D.G.CF + A.D.feat_off[br] + j * 8);
                            fa[c] = __ldg(f);
                            fb[c] = __ldg(f + 1);
                        }
                        float h[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) h[e] = 0.f;
#pragma unroll
                        for (int p = 0; p < 3; ++p) {
                            float s[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) s[e] = 0.f;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float4 a = fa[4 * p + k], b = fb[4 * p + k];
                                const float w = K.w[4 * p + k];
                                s[0] = fmaf(a.x, w, s[0]); s[1] = fmaf(a.y, w, s[1]); s[2] = fmaf(a.z, w, s[2]); s[3] = fmaf(a.w, w, s[3]);
                                s[4] = fmaf(b.x, w, s[4]); s[5] = fmaf(b.y, w, s[5]); s[6] = fmaf(b.z, w, s[6]); s[7] = fmaf(b.w, w, s[7]);
                            }
#pragma unroll
                            for (int e = 0; e < 8; ++e) h[e] += s[e];
                        }
                        __half hh[8], hl[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) split_f16_plain(h[e], hh[e], hl[e]);
                        const uint32_t off = static_cast<uint32_t>(pm) * 128u + (static_cast<uint32_t>(j ^ (pm & 7)) << 4);
                        *reinterpret_cast<uint4*>(sx + off) = *reinterpret_cast<const uint4*>(hh);
                        if (NSPLIT == 3) *reinterpret_cast<uint4*>(sx + kXLo + off) = *reinterpret_cast<const uint4*>(hl);
                    }
                    ptx::fence_proxy_async();           // a warp writes only rows of its own quarter: per-warp arrival
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(x_ready);
                }
                // ---- layer epilogues
                float o_acc[4] = {0.f, 0.f, 0.f, 0.f};
                const int n_out = A.n_out[br];
                for (int l = 0; l < kDecLayers; ++l, ++dphase) {
                    ptx::mbar_wait(d_full, dphase & 1);
                    __syncwarp();
                    ptx::tc_fence_after();
                    const uint32_t lane_addr = tmem_base + (l & 1) * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
                    const float inv = A.inv_scale[br][l];
                    const float* bias = C.bias[br][l];                 // constant bank (kernel parameter)
                    const bool last = l == kDecLayers - 1;
                    // 32 accumulator columns -> +bias, ReLU -> next operand (or the 256 -> n_out layer on the CUDA cores)
                    auto consume = [&](const uint32_t (&v)[32], int cb) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int n0 = cb * 32 + q * 8;
                            float r[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) r[e] = fmaxf(fmaf(__uint_as_float(v[q * 8 + e]), inv, bias[n0 + e]), 0.f);
                            if (!last) {
                                __half hh[8], hl[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) {          // r >= 0 after the ReLU: only the upper clamp
                                    const float c = fminf(r[e], 65504.f);
                                    hh[e] = __float2half_rn(c);
                                    hl[e] = __float2half_rn(c - __half2float(hh[e]));
                                }
                                // chunk (n0 / 64), row m, 16-byte unit ((n0 % 64) / 8) ^ (m & 7)
                                const uint32_t off = static_cast<uint32_t>(n0 >> 6) * kDecChunkBytes + static_cast<uint32_t>(m) * 128u +
                                                     ((static_cast<uint32_t>((n0 & 63) >> 3) ^ sw) << 4);
                                ptx::st_shared_v4(sact_u32 + off, *reinterpret_cast<const uint4*>(hh));
                                if (NSPLIT == 3) ptx::st_shared_v4(sact_u32 + kActLo + off, *reinterpret_cast<const uint4*>(hl));
                            } else {
                                // 256 -> n_out on the CUDA cores (second_layers' last Linear, blocks.py:80)
#pragma unroll
                                for (int o = 0; o < 4; ++o) {
                                    if (o < n_out) {
                                        const float* w = C.w_last[br][o] + n0;
                                        float a = o_acc[o];
#pragma unroll
                                        for (int e = 0; e < 8; ++e) a = fmaf(r[e], w[e], a);
                                        o_acc[o] = a;
                                    }
                                }
                            }
                        }
                        if (!last) {
                            // this warp's half of chunk cb/2 of the next layer's A operand is written (32 rows x 32 channels)
                            ptx::fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) ptx::mbar_arrive(&chunk_ready[cb >> 1]);
                        }
                    };
                    // TMEM loads run one 32-column block ahead of the arithmetic; this warp owns blocks half, half + 2, ...
                    uint32_t va[32], vb[32];
                    ptx::tmem_ld_32x32b_x32(lane_addr + half * 32, va);
#pragma unroll 1
                    for (int cb = half; cb < 8; cb += 4) {
                        ptx::tmem_ld_wait();
                        ptx::tmem_ld_32x32b_x32(lane_addr + (cb + 2) * 32, vb);
                        consume(va, cb);
                        ptx::tmem_ld_wait();
                        if (cb + 4 < 8) ptx::tmem_ld_32x32b_x32(lane_addr + (cb + 4) * 32, va);
                        consume(vb, cb + 2);
                    }
                    ptx::tc_fence_before();
                    if (last) {
                        // the two warps of a quarter hold partial sums over alternate column blocks: combine through the (now idle)
                        // activation buffer; nothing writes it again before every epilogue warp has arrived on x_ready
                        float4* xch = reinterpret_cast<float4*>(sact) + m;
                        if (half == 1) *xch = make_float4(o_acc[0], o_acc[1], o_acc[2], o_acc[3]);
                        asm volatile("bar.sync 1, 256;" ::: "memory");
                        if (half == 0) {
                            const float4 o = *xch;
                            o_acc[0] += o.x; o_acc[1] += o.y; o_acc[2] += o.z; o_acc[3] += o.w;
                        }
                    }
                    if (last && half == 0 && g0 + m < A.D.P.n) {
                        const int col0 = A.D.out_col[br];
                        float* op = A.D.out + (g0 + m) * A.D.oc + col0;
#pragma unroll
                        for (int o = 0; o < 4; ++o) {
                            if (o < n_out) {
                                float v = o_acc[o] + C.b_last[br][o];
                                if (A.D.sigmoid[br]) v = 1.f / (1.f + expf(-v));
                                if (A.D.clamp_tex && col0 + o > 0) v = fminf(fmaxf(v, 0.f), 1.f);
                                op[o] = v;
                            }
                        }
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Encoder half of the auto-encoder: volume -> latent planes.
//   reference: AutoEncoderGroupSkip.encode, src/encoding/networks.py:164-180
//     vol_feat = cat[Conv3d(1 -> GEO, k4 s2 p1)(vol[:, :1]), Conv3d(CT -> TEX, k4 s2 p1)(vol)]     [C, H, W, D]
//     plane    = tanh(InstanceNorm2d(mean over one volume axis) / 2)                               xy / xz / yz
// k_enc_conv3d never writes vol_feat: each thread forms the C conv outputs of one voxel (weights are FFMA constant-bank
// operands: they travel as a kernel parameter and every index is a compile-time constant), and the CTA adds its tile's three
// axis sums to 64-bit fixed-point accumulators (value * 2^40, integer atomics: exact, hence independent of the tile order).
// Tile = 2 x 4 x 32 output voxels (thread = voxel, lane = z: the smem reads of a warp are unit-stride because the input tile
// is stored de-interleaved in z), input tile 6 x 10 x 66 per volume channel staged in shared memory.
// k_enc_finalize: one CTA per (plane, channel): axis mean, instance statistics in fp64, tanh.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kEncTH = 2, kEncTW = 4, kEncTD = 32;
constexpr int kEncIH = 2 * kEncTH + 2, kEncIW = 2 * kEncTW + 2, kEncIZ = 2 * kEncTD + 2;       // 6, 10, 66
constexpr int kEncZP = kEncIZ / 2;                                                               // 33 even + 33 odd
constexpr double kEncFix = 1099511627776.0;                                                      // 2^40

template <int GEO, int TEX, int CT>
struct EncW {
    float wg[GEO][64];                                  // [co][kx*16 + ky*4 + kz]
    float wt[TEX > 0 ? TEX : 1][CT > 0 ? CT : 1][64];   // [co][ci][tap]
    float bias[GEO + TEX];
};
struct EncArgs {
    const float* vol;                 // [CV][X][Y][Z]
    int X, Y, Z, H, W, D;             // volume and conv-output sizes
    unsigned long long* sums[3];      // xy [C][H][W], xz [C][H][D], yz [C][W][D]
    float* out[3];
};

template <int GEO, int TEX, int CT>
__global__ void __launch_bounds__(256) k_enc_conv3d(const __grid_constant__ EncW<GEO, TEX, CT> Wt, const EncArgs A) {
    constexpr int CV = CT > 0 ? CT : 1, C = GEO + TEX;
    extern __shared__ float enc_smem[];
    float* tile = enc_smem;                                        // [CV][IH][IW][2][ZP]
    float* red = enc_smem + CV * kEncIH * kEncIW * 2 * kEncZP;     // [C][8][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d0 = blockIdx.x * kEncTD, w0 = blockIdx.y * kEncTW, h0 = blockIdx.z * kEncTH;
    // ---- stage the input tile: one warp per (channel, x, y) row of 66 z values, zero padding outside the volume
    const int iz0 = 2 * d0 - 1;
    for (int row = warp; row < CV * kEncIH * kEncIW; row += 8) {
        const int c = row / (kEncIH * kEncIW), rem = row - c * (kEncIH * kEncIW);
        const int ix = 2 * h0 - 1 + rem / kEncIW, iy = 2 * w0 - 1 + rem % kEncIW;
        const bool in = ix >= 0 && ix < A.X && iy >= 0 && iy < A.Y;
        const float* src = A.vol + ((static_cast<size_t>(c) * A.X + (in ? ix : 0)) * A.Y + (in ? iy : 0)) * A.Z;
        float* dst = tile + static_cast<size_t>(row) * 2 * kEncZP;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int j = lane + 32 * k;
            if (j < kEncIZ) {
                // asynchronous 4-byte copies (zero-filled outside the volume): every row of the tile is in flight at once
                const int iz = iz0 + j;
                const bool ok = in && iz >= 0 && iz < A.Z;
                const unsigned sz = ok ? 4u : 0u;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(ptx::smem_u32(dst + (j & 1) * kEncZP + (j >> 1))),
                             "l"(ok ? src + iz : A.vol), "r"(sz) : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- one voxel per thread
    const int hl = warp >> 2, wl = warp & 3;
    const int h = h0 + hl, w = w0 + wl, d = d0 + lane;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = Wt.bias[c];
#pragma unroll
    for (int c = 0; c < CV; ++c)
#pragma unroll
        for (int kx = 0; kx < 4; ++kx)
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
                const float* rowp = tile + ((static_cast<size_t>(c) * kEncIH + 2 * hl + kx) * kEncIW + 2 * wl + ky) * 2 * kEncZP + lane;
#pragma unroll
                for (int kz = 0; kz < 4; ++kz) {
                    const float v = rowp[(kz & 1) * kEncZP + (kz >> 1)];
                    const int tap = kx * 16 + ky * 4 + kz;
                    if (c == 0) {
#pragma unroll
                        for (int g = 0; g < GEO; ++g) acc[g] = fmaf(v, Wt.wg[g][tap], acc[g]);
                    }
                    if (TEX > 0) {
#pragma unroll
                        for (int t = 0; t < TEX; ++t) acc[GEO + t] = fmaf(v, Wt.wt[t][c][tap], acc[GEO + t]);
                    }
                }
            }
    const bool valid = h < A.H && w < A.W && d < A.D;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        if (!valid) acc[c] = 0.f;
        red[(c * 8 + warp) * 32 + lane] = acc[c];
    }
    // ---- xy: sum over z inside the warp (fixed butterfly), one atomic per (channel, h, w)
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float s = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && h < A.H && w < A.W)
            atomicAdd(A.sums[0] + (static_cast<size_t>(c) * A.H + h) * A.W + w,
                      static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(s) * kEncFix)));
    }
    __syncthreads();
    // ---- xz: sum over the tile's 4 y;  yz: over its 2 x  (fixed order), atomics coalesced along z
    for (int i = threadIdx.x; i < C * kEncTH * 32; i += 256) {
        const int c = i / (kEncTH * 32), r = i - c * (kEncTH * 32), th = r >> 5, z = r & 31;
        const float* b = red + (c * 8 + th * 4) * 32 + z;
        const float s = (b[0] + b[32]) + (b[64] + b[96]);
        if (h0 + th < A.H && d0 + z < A.D)
            atomicAdd(A.sums[1] + (static_cast<size_t>(c) * A.H + h0 + th) * A.D + d0 + z,
                      static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(s) * kEncFix)));
    }
    for (int i = threadIdx.x; i < C * kEncTW * 32; i += 256) {
        const int c = i / (kEncTW * 32), r = i - c * (kEncTW * 32), tw = r >> 5, z = r & 31;
        const float* b = red + (c * 8 + tw) * 32 + z;
        const float s = b[0] + b[4 * 32];
        if (w0 + tw < A.W && d0 + z < A.D)
            atomicAdd(A.sums[2] + (static_cast<size_t>(c) * A.W + w0 + tw) * A.D + d0 + z,
                      static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(s) * kEncFix)));
    }
}

// k_enc_conv3d_tma: the same tile, staged by ONE 4-D TMA box {72 z, 10 y, 6 x, CV} (out-of-volume elements are the TMA unit's
// zero fill == the conv's padding, no index arithmetic in the kernel), natural z order in shared memory.  128 threads, each
// forms two z-adjacent voxels from three shared loads per (channel, kx, ky): 4608 FFMA per 192 loads.
// Needs Z % 4 == 0 (16-byte global strides); other volumes take k_enc_conv3d.  Measured on B200 (tools/probe/tma_f32_probe.cu):
// with INTERLEAVE_NONE the innermost TMA coordinate must be a multiple of 16 bytes (z0 = -1 raises "illegal instruction"), so
// the box starts at z = 2*d0 - 4 and the window of a thread sits 3 floats into it.
constexpr int kEncZPitch = 72;
template <int GEO, int TEX, int CT>
__global__ void __launch_bounds__(128) k_enc_conv3d_tma(const __grid_constant__ CUtensorMap vmap,
                                                        const __grid_constant__ EncW<GEO, TEX, CT> Wt, const EncArgs A) {
    constexpr int CV = CT > 0 ? CT : 1, C = GEO + TEX;
    constexpr uint32_t kTileBytes = CV * kEncIH * kEncIW * kEncZPitch * 4;
    extern __shared__ __align__(128) float enc_smem_t[];
    float* tile = enc_smem_t;                                  // [CV][IH][IW][72]
    float* red = enc_smem_t + kTileBytes / 4;                  // [C][8][32]
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + C * 8 * 32);
    const int d0 = blockIdx.x * kEncTD, w0 = blockIdx.y * kEncTW, h0 = blockIdx.z * kEncTH;
    if (threadIdx.x == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_barrier_init();
        ptx::mbar_arrive_expect_tx(bar, kTileBytes);
        ptx::tma_load_4d(tile, &vmap, bar, 2 * d0 - 4, 2 * w0 - 1, 2 * h0 - 1, 0);
    }
    __syncthreads();
    ptx::mbar_wait(bar, 0);
    const int combo = threadIdx.x >> 4, L = threadIdx.x & 15;
    const int hl = combo >> 2, wl = combo & 3;
    const int h = h0 + hl, w = w0 + wl, d = d0 + 2 * L;
    float acc0[C], acc1[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc0[c] = acc1[c] = Wt.bias[c];
#pragma unroll
    for (int c = 0; c < CV; ++c)
#pragma unroll
        for (int kx = 0; kx < 4; ++kx)
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) {
                // inputs z = 2*d - 1 .. 2*d + 4 of the pair = floats 4L+3 .. 4L+8 of the row
                const float* rowp = tile + ((c * kEncIH + 2 * hl + kx) * kEncIW + 2 * wl + ky) * kEncZPitch + 4 * L;
                const float a0 = rowp[3];
                const float4 a = *reinterpret_cast<const float4*>(rowp + 4);
                const float a5 = rowp[8];
                const float in[6] = {a0, a.x, a.y, a.z, a.w, a5};
#pragma unroll
                for (int kz = 0; kz < 4; ++kz) {
                    const int tap = kx * 16 + ky * 4 + kz;
                    if (c == 0) {
#pragma unroll
                        for (int g = 0; g < GEO; ++g) {
                            acc0[g] = fmaf(in[kz], Wt.wg[g][tap], acc0[g]);
                            acc1[g] = fmaf(in[kz + 2], Wt.wg[g][tap], acc1[g]);
                        }
                    }
                    if (TEX > 0) {
#pragma unroll
                        for (int t = 0; t < TEX; ++t) {
                            acc0[GEO + t] = fmaf(in[kz], Wt.wt[t][c][tap], acc0[GEO + t]);
                            acc1[GEO + t] = fmaf(in[kz + 2], Wt.wt[t][c][tap], acc1[GEO + t]);
                        }
                    }
                }
            }
    const bool hw_ok = h < A.H && w < A.W;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        if (!(hw_ok && d < A.D)) acc0[c] = 0.f;
        if (!(hw_ok && d + 1 < A.D)) acc1[c] = 0.f;
        *reinterpret_cast<float2*>(red + (c * 8 + combo) * 32 + 2 * L) = make_float2(acc0[c], acc1[c]);
    }
    // ---- xy: sum over the tile's 32 z (pair, then a fixed butterfly over the 16 lanes of the (h, w) group)
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float s = acc0[c] + acc1[c];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (L == 0 && hw_ok)
            atomicAdd(A.sums[0] + (static_cast<size_t>(c) * A.H + h) * A.W + w,
                      static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(s) * kEncFix)));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * kEncTH * 32; i += 128) {
        const int c = i / (kEncTH * 32), r = i - c * (kEncTH * 32), th = r >> 5, z = r & 31;
        const float* b = red + (c * 8 + th * 4) * 32 + z;
        const float s = (b[0] + b[32]) + (b[64] + b[96]);
        if (h0 + th < A.H && d0 + z < A.D)
            atomicAdd(A.sums[1] + (static_cast<size_t>(c) * A.H + h0 + th) * A.D + d0 + z,
                      static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(s) * kEncFix)));
    }
    for (int i = threadIdx.x; i < C * kEncTW * 32; i += 128) {
        const int c = i / (kEncTW * 32), r = i - c * (kEncTW * 32), tw = r >> 5, z = r & 31;
        const float* b = red + (c * 8 + tw) * 32 + z;
        const float s = b[0] + b[4 * 32];
        if (w0 + tw < A.W && d0 + z < A.D)
            atomicAdd(A.sums[2] + (static_cast<size_t>(c) * A.W + w0 + tw) * A.D + d0 + z,
                      static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(s) * kEncFix)));
    }
}

// Any channel configuration (fdim_geo / fdim_tex / tex_channels other than the reference defaults): straightforward direct
// convolution, one thread per output voxel, weights read through the read-only cache.  Same fixed-point axis sums, so
// k_enc_finalize is shared.  grid (ceil(D/128), W, H), block 128 (lane = z).
constexpr int kEncMaxC = 32;
struct EncGenericW {
    const float *wg, *bg;      // [GEO][1][64], [GEO]
    const float *wt, *bt;      // [TEX][CT][64], [TEX]   (nullptr without colour)
    int GEO, TEX, CT;
};
__global__ void __launch_bounds__(128) k_enc_conv3d_generic(const EncGenericW Wt, const EncArgs A) {
    const int d = blockIdx.x * 128 + threadIdx.x, w = blockIdx.y, h = blockIdx.z;
    const int C = Wt.GEO + Wt.TEX;
    const bool valid = d < A.D;
    float acc[kEncMaxC];
#pragma unroll
    for (int c = 0; c < kEncMaxC; ++c) acc[c] = 0.f;
    if (valid) {
        for (int c = 0; c < Wt.GEO; ++c) acc[c] = __ldg(Wt.bg + c);
        for (int c = 0; c < Wt.TEX; ++c) acc[Wt.GEO + c] = __ldg(Wt.bt + c);
        const int nin = Wt.TEX > 0 ? Wt.CT : 1;
        for (int ci = 0; ci < nin; ++ci)
            for (int kx = 0; kx < 4; ++kx) {
                const int ix = 2 * h - 1 + kx;
                if (ix < 0 || ix >= A.X) continue;
                for (int ky = 0; ky < 4; ++ky) {
                    const int iy = 2 * w - 1 + ky;
                    if (iy < 0 || iy >= A.Y) continue;
                    const float* row = A.vol + ((static_cast<size_t>(ci) * A.X + ix) * A.Y + iy) * A.Z;
                    for (int kz = 0; kz < 4; ++kz) {
                        const int iz = 2 * d - 1 + kz;
                        if (iz < 0 || iz >= A.Z) continue;
                        const float v = __ldg(row + iz);
                        const int tap = kx * 16 + ky * 4 + kz;
                        if (ci == 0)
                            for (int c = 0; c < Wt.GEO; ++c) acc[c] = fmaf(v, __ldg(Wt.wg + c * 64 + tap), acc[c]);
                        for (int c = 0; c < Wt.TEX; ++c) acc[Wt.GEO + c] = fmaf(v, __ldg(Wt.wt + (c * Wt.CT + ci) * 64 + tap), acc[Wt.GEO + c]);
                    }
                }
            }
    }
#pragma unroll
    for (int c = 0; c < kEncMaxC; ++c) {
        if (c < C) {                                           // uniform
            const float v = valid ? acc[c] : 0.f;
            float s = v;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) == 0)
                atomicAdd(A.sums[0] + (static_cast<size_t>(c) * A.H + h) * A.W + w,
                          static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(s) * kEncFix)));
            if (valid) {
                const unsigned long long f = static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(v) * kEncFix));
                atomicAdd(A.sums[1] + (static_cast<size_t>(c) * A.H + h) * A.D + d, f);
                atomicAdd(A.sums[2] + (static_cast<size_t>(c) * A.W + w) * A.D + d, f);
            }
        }
    }
}

// grid (C, 3), block 1024: the plane's values are formed once and kept in registers (<= 16 per thread covers 128 x 128)
constexpr int kEncFinThreads = 1024, kEncFinKeep = 16;
__global__ void __launch_bounds__(kEncFinThreads) k_enc_finalize(const EncArgs A) {
    const int c = blockIdx.x, p = blockIdx.y;
    const int rows = p == 2 ? A.W : A.H, cols = p == 0 ? A.W : A.D;
    const int n = rows * cols;
    const float len = static_cast<float>(p == 0 ? A.D : (p == 1 ? A.W : A.H));     // length of the averaged axis
    const unsigned long long* s = A.sums[p] + static_cast<size_t>(c) * n;
    float* o = A.out[p] + static_cast<size_t>(c) * n;
    __shared__ double sh[32];
    __shared__ double stat[2];
    auto val = [&](int i) { return static_cast<float>(static_cast<double>(static_cast<long long>(s[i])) * (1.0 / kEncFix)) / len; };
    auto block_sum = [&](double v) {
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) v += __shfl_xor_sync(0xffffffffu, v, k);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
        __syncthreads();
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < kEncFinThreads / 32; ++k) t += sh[k];
        return t;
    };
    float keep[kEncFinKeep];
#pragma unroll
    for (int k = 0; k < kEncFinKeep; ++k) {
        const int i = threadIdx.x + k * kEncFinThreads;
        keep[k] = i < n ? val(i) : 0.f;
    }
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < kEncFinKeep; ++k) a += static_cast<double>(keep[k]);            // zeros beyond n
    for (int i = threadIdx.x + kEncFinKeep * kEncFinThreads; i < n; i += kEncFinThreads) a += static_cast<double>(val(i));
    const double mean = block_sum(a) / n;
    a = 0.0;
#pragma unroll
    for (int k = 0; k < kEncFinKeep; ++k) {
        const double dv = static_cast<double>(keep[k]) - mean;
        if (threadIdx.x + k * kEncFinThreads < n) a += dv * dv;
    }
    for (int i = threadIdx.x + kEncFinKeep * kEncFinThreads; i < n; i += kEncFinThreads) {
        const double dv = static_cast<double>(val(i)) - mean;
        a += dv * dv;
    }
    const double var = block_sum(a) / n;                       // biased (InstanceNorm2d)
    if (threadIdx.x == 0) {
        stat[0] = mean;
        stat[1] = 1.0 / sqrt(var + 1e-5);
    }
    __syncthreads();
    const float mu = static_cast<float>(stat[0]), rstd = static_cast<float>(stat[1]);
#pragma unroll
    for (int k = 0; k < kEncFinKeep; ++k) {
        const int i = threadIdx.x + k * kEncFinThreads;
        if (i < n) o[i] = tanhf(((keep[k] - mu) * rstd) * 0.5f);
    }
    for (int i = threadIdx.x + kEncFinKeep * kEncFinThreads; i < n; i += kEncFinThreads) o[i] = tanhf(((val(i) - mu) * rstd) * 0.5f);
}

}  // namespace s3d
