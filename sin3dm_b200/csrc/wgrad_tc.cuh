// Weight gradient of the 3x3 (and 1x1 skip) triplane convolution on the sm_100a tensor cores.
//
//   reference: autograd of nn.Conv2d inside TriplaneConv (src/diffusion/unet_triplane.py:21-60) under TrainLoop.forward_backward
//              (src/diffusion/train_util.py:198-235); restated in oracle/backward_ref.py (torch.nn.grad.conv2d_weight).
//
//   dW[co][c][kh][kw] = sum_{b, px} dY[b][px][co] * A[b][px + (kh-1, kw-1)][c]            (zero outside the plane)
//
// GEMM view: the contraction index K is the PIXEL, which is the slow index of both NHWC operands — i.e. both operands are
// "MN-major" (64 channels contiguous per pixel row).  tcgen05.mma takes such operands directly (a_major = b_major = 1 in the
// instruction descriptor; canonical SWIZZLE_128B MN-major layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units), and the TMA
// boxes the forward already uses ARE that layout: one pixel = one 128-byte row of 64 channels, 8-pixel groups SBO bytes apart.
//   M side (128 rows) = the activations: 64 channels x TWO taps — the second 64-row block of the descriptor is reached through the
//           leading byte offset, and LBO = 128 B / 2048 B is exactly "one pixel to the right" / "one image row down" inside the
//           18 x 16-pixel halo patch, so a tap pair costs one MMA and the A operand is fetched ONCE per tile for all nine taps;
//   N side = dY, 64 or 128 output channels (two 64-channel tiles LBO = 16 KiB apart);
//   K step = 16 pixels = two image rows of the 16 x 8-pixel tile (SBO = 2048 B in the patch, 1024 B in the dY tile).
// Accumulators: one [128 x Cout] fp32 block in TMEM per tap pair (kw 0|1 of each kh, then kw 2 of kh 0|1, then kw 2 of kh 2 with
// an unused second half) — 5 blocks for the 3x3; when 5*Cout exceeds the 512 TMEM columns the pairs are split over two CTAs.
// A CTA owns (plane, 64-channel block of A, accumulator group, a contiguous range of pixel tiles) and streams its tiles through
// a 3-stage TMA ring; split-K partials go to a scratch buffer with plain coalesced stores and k_wgrad_reduce adds them in a
// fixed order into the reference layout (deterministic, no atomics).
// Precision: fp16 hi halves only (single MMA per term): the products are rounded to ~2^-11 and averaged over thousands of pixels;
// measured against the fp32 CUDA-core kernel k_wgrad_ffma in tests/test_gpu_backward.py.
#pragma once
#include "common.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace s3d {

constexpr int kWgThreads = 192;                 // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int kWgStages = 3;
constexpr int kWgMaxAcc = 5;
// The wgrad keeps its own halo patch with a 16-pixel row pitch (2048 B between image rows): with the forward's 10-pixel pitch the
// MN-major reads came out wrong on B200 (measured: 4 % error in dW), while every K-major read of the forward is fine with it.
constexpr int kWgHaloW = 16;
constexpr int kWgABytes = kHaloH * kWgHaloW * kBK * 2;          // 36 KiB halo patch (hi half)
constexpr int kWgYBytes = kBM * kBK * 2;        // 16 KiB: 128 pixels x 64 output channels

struct WgradTcMaps {
    CUtensorMap a[3];     // activations (C, cols, rows, B, 2) box {64, kWgHaloW, 18}
    CUtensorMap y[3];     // dY          (Cout, cols, rows, B, 2) box {64, 8, 16}
};
struct WgUnit {
    int plane, cb;        // plane, 64-channel block of the activations
    int acc0, nacc;       // accumulators [acc0, acc0 + nacc) of the table below
    int tile0, tile1;     // range of the plane's pixel tiles (b major, then tile row, then tile column)
};
struct WgradTcArgs {
    int Cout;
    int tiles_x[3], tiles_per_sample[3];
    uint32_t a_off[kWgMaxAcc], lbo[kWgMaxAcc];     // start offset of the first tap inside the patch, offset of the paired tap
    const WgUnit* units;
    float* partial;       // [unit][nacc of the unit <= kWgMaxAcc][128][Cout]: unit u starts at u * kWgMaxAcc * 128 * Cout
};

// MN-major operand, 128-byte swizzle (cute/arch/mma_sm100_desc.hpp SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_f16_mn(int M, int N) {
    return (1u << 4) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1) k_wgrad_tc(const __grid_constant__ WgradTcMaps M, const WgradTcArgs A) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nco = A.Cout / kBK;                                   // 64-channel tiles of dY
    const uint32_t stage_bytes = kWgABytes + nco * kWgYBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWgStages * stage_bytes);
    uint64_t* empty = full + kWgStages;
    uint64_t* done = empty + kWgStages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const WgUnit U = A.units[blockIdx.x];
    uint32_t cols = 32;
    while (cols < static_cast<uint32_t>(U.nacc * A.Cout)) cols <<= 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&M.a[U.plane]);
        ptx::prefetch_tmap(&M.y[U.plane]);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kWgStages; ++s) {
                ptx::mbar_init(&full[s], 1);
                ptx::mbar_init(&empty[s], 1);
            }
            ptx::mbar_init(done, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc_dyn(tmem_ptr_smem, cols);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int tps = A.tiles_per_sample[U.plane], tx_n = A.tiles_x[U.plane];

    if (warp == 0) {
        // ===================== TMA producer =====================
        int g = 0;
        for (int t = U.tile0; t < U.tile1; ++t, ++g) {
            const int b = t / tps, ip = t - b * tps, ty = ip / tx_n, tx = ip - ty * tx_n;
            const int h0 = ty * kTileH, w0 = tx * kTileW;
            const int s = g % kWgStages;
            ptx::mbar_wait(&empty[s], ((g / kWgStages) & 1) ^ 1);
            uint8_t* st = smem + s * stage_bytes;
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&full[s], stage_bytes);
                ptx::tma_load_5d(st, &M.a[U.plane], &full[s], U.cb * kBK, w0 - 1, h0 - 1, b, 0);
                for (int j = 0; j < nco; ++j) ptx::tma_load_5d(st + kWgABytes + j * kWgYBytes, &M.y[U.plane], &full[s], j * kBK, w0, h0, b, 0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_f16_mn(kBM, A.Cout);
        int g = 0;
        for (int t = U.tile0; t < U.tile1; ++t, ++g) {
            const int s = g % kWgStages;
            ptx::mbar_wait(&full[s], (g / kWgStages) & 1);
            ptx::tc_fence_after();
            const uint32_t a_base = ptx::smem_u32(smem + s * stage_bytes), y_base = a_base + kWgABytes;
            if (ptx::elect_one()) {
#pragma unroll 1
                for (int i = 0; i < kTileH / 2; ++i) {                    // K step: tile rows 2i, 2i + 1
                    const uint64_t yd = make_mn_sw128_desc(y_base + i * 2048u, kWgYBytes, 1024u);
                    for (int a = 0; a < U.nacc; ++a) {
                        const uint64_t ad = make_mn_sw128_desc(a_base + A.a_off[U.acc0 + a] + i * (2u * kWgHaloW * 128u), A.lbo[U.acc0 + a], kWgHaloW * 128u);
                        ptx::umma_f16(tmem_base + a * A.Cout, ad, yd, idesc, (g > 0 || i > 0) ? 1u : 0u);
                    }
                }
                ptx::umma_commit(&empty[s]);
            }
            __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit(done);
        __syncwarp();
    } else {
        // ===================== epilogue: TMEM -> split-K partial =====================
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        ptx::mbar_wait(done, 0);
        __syncwarp();
        ptx::tc_fence_after();
        float* outp = A.partial + (static_cast<size_t>(blockIdx.x) * kWgMaxAcc * kBM + m) * A.Cout;
        for (int a = 0; a < U.nacc; ++a) {
            for (int c0 = 0; c0 < A.Cout; c0 += 32) {
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + a * A.Cout + c0, v);
                ptx::tmem_ld_wait();
                float4* dst = reinterpret_cast<float4*>(outp + static_cast<size_t>(a) * kBM * A.Cout + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc_dyn(tmem_base, cols);
    }
}

// dW[co][cb*64 + c][tap] (+)= sum over the split-K units of one (plane, channel block, accumulator group), fixed order.
struct WgReduceGroup {
    int unit0, nsplit, plane, cb, acc0, nacc;
};
struct WgReduceArgs {
    const WgReduceGroup* groups;
    const float* partial;
    int Cout, Cw, ntap, C;
    int tap[kWgMaxAcc][2];      // reference tap index (kh*3 + kw) of the two 64-row halves of every accumulator, -1: unused
    float* dw[3];               // [Cout][Cw][ntap]
};
// grid (ceil(nacc_max * 128 * Cout / 256), groups), block 256
__global__ void __launch_bounds__(256) k_wgrad_reduce(WgReduceArgs A) {
    const WgReduceGroup G = A.groups[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G.nacc * kBM * A.Cout) return;
    const int co = i % A.Cout, m = (i / A.Cout) % kBM, a = i / (A.Cout * kBM);
    const int tap = A.tap[G.acc0 + a][m >> 6], c = G.cb * kBK + (m & 63);
    if (tap < 0 || c >= A.C) return;
    float acc = 0.f;
    for (int s = 0; s < G.nsplit; ++s)
        acc += A.partial[((static_cast<size_t>(G.unit0 + s) * kWgMaxAcc + a) * kBM + m) * A.Cout + co];
    A.dw[G.plane][(static_cast<size_t>(co) * A.Cw + c) * A.ntap + tap] += acc;
}

}  // namespace s3d
