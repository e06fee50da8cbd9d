// Microbenchmark: what does TMA tile delivery L2 -> shared memory sustain on a B200, per SM and chip-wide?
// Each of G CTAs (1 per SM, 200 KB of shared memory so that nothing else co-resides) streams `iters` 2-D tiles
// (64 x ROWS fp16, SWIZZLE_128B) through a ring of DEPTH slots; a consumer lane releases a slot as soon as it lands.
// Variants: all CTAs read the SAME `hot` bytes (like conv weights) or each CTA its own region (like activations).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_bw tools/tma_bw.cu -lcuda && /tmp/tma_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../sin3dm_b200/csrc/ptx.cuh"
using namespace s3d;

template <int DEPTH, int ROWS>
__global__ void __launch_bounds__(64, 1) k_stream(const __grid_constant__ CUtensorMap M, int iters, int tiles_per_region, int shared_region,
                                                  unsigned long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kTile = 128 * ROWS;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + DEPTH * kTile);
    uint64_t* empty = full + DEPTH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < DEPTH; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 1);
        }
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&M);
    }
    __syncthreads();
    const int region = shared_region ? 0 : blockIdx.x;
    const long long t0 = clock64();
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < iters; ++i) {
            const int s = i % DEPTH;
            ptx::mbar_wait(&empty[s], ((i / DEPTH) & 1) ^ 1);
            ptx::mbar_arrive_expect_tx(&full[s], kTile);
            const int tile = region * tiles_per_region + (i % tiles_per_region);
            ptx::tma_load_3d(smem + s * kTile, &M, &full[s], 0, tile * ROWS, 0);
        }
    } else if (warp == 1 && lane == 0) {
        for (int i = 0; i < iters; ++i) {
            const int s = i % DEPTH;
            ptx::mbar_wait(&full[s], (i / DEPTH) & 1);
            ptx::mbar_arrive(&empty[s]);
        }
        cyc[blockIdx.x] = clock64() - t0;
    }
}

// Cluster-of-2 variant: each CTA fetches HALF of every tile (ROWS/2 rows) and multicasts it to both CTAs, so every CTA still
// receives the full tile but the pair reads it from L2 once.  A slot is refilled only after BOTH consumers released it
// (empty barriers count 2, remote arrive through mapa).
template <int DEPTH, int ROWS>
__global__ void __launch_bounds__(64, 1) k_stream_mc(const __grid_constant__ CUtensorMap M, int iters, int tiles_per_region, int shared_region,
                                                     unsigned long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kTile = 128 * ROWS;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + DEPTH * kTile);
    uint64_t* empty = full + DEPTH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int s = 0; s < DEPTH; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], 2);
        }
        ptx::fence_barrier_init();
        ptx::prefetch_tmap(&M);
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const int region = shared_region ? 0 : (blockIdx.x >> 1);
    const long long t0 = clock64();
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < iters; ++i) {
            const int s = i % DEPTH;
            ptx::mbar_wait(&empty[s], ((i / DEPTH) & 1) ^ 1);
            ptx::mbar_arrive_expect_tx(&full[s], kTile);
            const int tile = region * tiles_per_region + (i % tiles_per_region);
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
                " [%0], [%1, {%3, %4, %5}], [%2], %6;"
                ::"r"(ptx::smem_u32(smem + s * kTile + rank * (kTile / 2))), "l"(reinterpret_cast<uint64_t>(&M)), "r"(ptx::smem_u32(&full[s])),
                "r"(0), "r"(tile * ROWS + static_cast<int>(rank) * (ROWS / 2)), "r"(0), "h"(static_cast<uint16_t>(3))
                : "memory");
        }
    } else if (warp == 1 && lane == 0) {
        for (int i = 0; i < iters; ++i) {
            const int s = i % DEPTH;
            ptx::mbar_wait(&full[s], (i / DEPTH) & 1);
#pragma unroll
            for (uint32_t r = 0; r < 2; ++r) {
                uint32_t ra;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(ptx::smem_u32(&empty[s])), "r"(r));
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
            }
        }
        cyc[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    return reinterpret_cast<PFN_encodeTiled>(p);
}

template <int DEPTH, int ROWS>
static void run(const char* name, void* buf, size_t total_rows, int G, int iters, int tiles_per_region, int shared_region, int mc = 0) {
    CUtensorMap m;
    cuuint64_t gdim[3] = {64, total_rows, 1}, gstr[2] = {128, 128 * total_rows};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(mc ? ROWS / 2 : ROWS), 1}, es[3] = {1, 1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("encode failed %d\n", (int)r);
        return;
    }
    unsigned long long* cyc;
    cudaMalloc(&cyc, G * sizeof(unsigned long long));
    const size_t smem = 200 * 1024;
    cudaFuncSetAttribute(k_stream<DEPTH, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_stream_mc<DEPTH, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (mc) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(G);
            cfg.blockDim = dim3(64);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            if (rep == 0) {
                int nc = 0;
                cudaOccupancyMaxActiveClusters(&nc, k_stream_mc<DEPTH, ROWS>, &cfg);
                printf("[max active clusters of 2: %d] ", nc);
            }
            cudaLaunchKernelEx(&cfg, k_stream_mc<DEPTH, ROWS>, m, iters, tiles_per_region, shared_region, cyc);
        } else {
            k_stream<DEPTH, ROWS><<<G, 64, smem>>>(m, iters, tiles_per_region, shared_region, cyc);
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> h(G);
    cudaMemcpy(h.data(), cyc, G * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0, sum = 0;
    for (auto v : h) {
        mx = v > mx ? v : mx;
        sum += v;
    }
    const double bytes = (double)iters * 128 * ROWS;
    printf("%-44s depth %d tile %5d B  G %3d: %6.1f B/clk/SM (mean) %6.1f (slowest)  chip %7.0f B/clk  kernel %.1f us -> %.2f TB/s  err=%s\n", name,
           DEPTH, 128 * ROWS, G, bytes / ((double)sum / G), bytes / (double)mx, bytes * G / (double)mx, ms * 1e3, bytes * G / (ms * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(cyc);
}

int main() {
    const int G = 148, iters = 512;
    // 148 regions of 27 tiles x 16 KB (= one conv's weight slab per region) = 64 MB: stays inside the L2 after the first pass
    const int tiles_per_region = 27;
    const size_t rows128 = (size_t)G * tiles_per_region * 128;
    void* buf;
    cudaMalloc(&buf, rows128 * 128);
    cudaMemset(buf, 1, rows128 * 128);
    run<3, 128>("own region, 16 KB tiles", buf, rows128, G, iters, tiles_per_region, 0);
    run<4, 128>("own region, 16 KB tiles", buf, rows128, G, iters, tiles_per_region, 0);
    run<8, 128>("own region, 16 KB tiles", buf, rows128, G, iters, tiles_per_region, 0);
    run<3, 128>("same region (hot 432 KB), 16 KB tiles", buf, rows128, G, iters, tiles_per_region, 1);
    run<4, 128>("same region (hot 432 KB), 16 KB tiles", buf, rows128, G, iters, tiles_per_region, 1);
    run<8, 128>("same region (hot 432 KB), 16 KB tiles", buf, rows128, G, iters, tiles_per_region, 1);
    run<4, 128>("MC2 same region (hot), 16 KB tiles", buf, rows128, G, iters, tiles_per_region, 1, 1);
    run<8, 128>("MC2 same region (hot), 16 KB tiles", buf, rows128, G, iters, tiles_per_region, 1, 1);
    run<4, 128>("MC2 per-cluster region, 16 KB tiles", buf, rows128, G, iters, tiles_per_region, 0, 1);
    run<8, 64>("own region, 8 KB tiles", buf, rows128, G, iters, tiles_per_region * 2, 0);
    run<8, 64>("same region, 8 KB tiles", buf, rows128, G, iters, tiles_per_region * 2, 1);
    run<3, 128>("own region, 16 KB tiles, 32 CTAs", buf, rows128, 32, iters, tiles_per_region, 0);
    run<3, 128>("same region, 16 KB tiles, 32 CTAs", buf, rows128, 32, iters, tiles_per_region, 1);
    run<3, 128>("own region, 16 KB tiles, 1 CTA", buf, rows128, 1, iters, tiles_per_region, 0);
    run<8, 128>("own region, 16 KB tiles, 1 CTA", buf, rows128, 1, iters, tiles_per_region, 0);
    return 0;
}
