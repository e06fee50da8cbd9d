// Probe: which fp32 / SWIZZLE_NONE TMA box shapes does sm_100a accept?  nvcc -arch=sm_100a -o tma_f32_probe tma_f32_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__global__ void k(const __grid_constant__ CUtensorMap m, float* out, int n, int c0, int c1, int c2, int c3) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint64_t bar;
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n * 4));
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(d), "l"((uint64_t)&m), "r"(b), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    __syncthreads();
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.b32 %0,1,0,P;}" : "=r"(ok) : "r"(b));
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
    int bz = argc > 1 ? atoi(argv[1]) : 68, Z = argc > 2 ? atoi(argv[2]) : 184, cz = argc > 3 ? atoi(argv[3]) : -1;
    int Y = 12, X = 8, C = 2, by = 10, bx = 6;
    std::vector<float> h((size_t)C * X * Y * Z);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    int n = bz * by * bx * C;
    cudaMalloc(&o, n * 4);
    CUtensorMap m;
    cuuint64_t gd[4] = {(cuuint64_t)Z, (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)C}, gs[3] = {(cuuint64_t)Z * 4, (cuuint64_t)Z * Y * 4, (cuuint64_t)Z * Y * X * 4};
    cuuint32_t box[4] = {(cuuint32_t)bz, (cuuint32_t)by, (cuuint32_t)bx, (cuuint32_t)C}, es[4] = {1, 1, 1, 1};
    cuInit(0);
    CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box z %d Z %d coord z %d: encode %d; ", bz, Z, cz, (int)r);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 4);
    k<<<1, 128, n * 4>>>(m, o, n, cz, -1, 3, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s; ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> g(n);
        cudaMemcpy(g.data(), o, n * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int c = 0; c < C; ++c) for (int x = 0; x < bx; ++x) for (int y = 0; y < by; ++y) for (int z = 0; z < bz; ++z) {
            int gx = 3 + x, gy = -1 + y, gz = cz + z;
            float want = (gx < X && gy >= 0 && gy < Y && gz >= 0 && gz < Z) ? h[(((size_t)c * X + gx) * Y + gy) * Z + gz] : 0.f;
            if (g[((c * bx + x) * by + y) * bz + z] != want) ++bad;
        }
        printf("mismatches %d", bad);
    }
    printf("\n");
    return 0;
}
