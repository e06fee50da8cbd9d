// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) issued back to back by one thread, by N, with a commit every
// `per_commit` instructions and a wait on it (what a tap of the conv main loop does).  Operands: whatever is in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mma_issue tools/mma_issue.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../sin3dm_b200/csrc/ptx.cuh"
using namespace s3d;

template <int N>
__global__ void __launch_bounds__(128, 1) k_issue(int n_mma, int per_commit, int wait_each, int n_acc, int conv_pattern, int converged, unsigned long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, ring[8];
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (warp == 0) {
        if (lane == 0) {
            ptx::mbar_init(&bar, 1);
            for (int i = 0; i < 8; ++i) ptx::mbar_init(&ring[i], 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc<512>(&tmem_ptr);
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    if (converged ? warp == 0 : threadIdx.x == 0) {
        const bool issuer = lane == 0;
        const uint32_t idesc = ptx::make_idesc_f16(128, N);
        const uint64_t a = ptx::make_sw128_desc1024(ptx::smem_u32(smem));
        const uint64_t b = ptx::make_sw128_desc1024(ptx::smem_u32(smem + 32 * 1024));
        uint32_t phase = 0;
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            if (!issuer) {
                // the other lanes only walk the loop (converged mode)
            } else if (conv_pattern) {
                // the conv's k-step: [D1|D2] += Ah*[Bh|Bl] (N = 2*N'), then D2 += Al*Bh (N = N')   (N here is N' = 64)
                if ((i & 1) == 0) ptx::umma_f16(tmem, a + ((i & 6)), b + ((i & 6)), ptx::make_idesc_f16(128, 2 * N), i > 1 ? 1u : 0u);
                else ptx::umma_f16(tmem + N, a + ((i & 6)), b + ((i & 6)), idesc, 1u);
            } else {
                ptx::umma_f16(tmem + (i % n_acc) * N, a + ((i & 3) * 2), b + ((i & 3) * 2), idesc, i >= n_acc ? 1u : 0u);
            }
            if (issuer && (i + 1) % per_commit == 0) {
                if (wait_each) {
                    ptx::umma_commit(&bar);
                    ptx::mbar_wait(&bar, phase);
                    phase ^= 1;
                } else {
                    ptx::umma_commit(&ring[(i / per_commit) & 7]);      // like the conv's slot-release commits: nobody waits here
                }
            }
        }
        __syncwarp(converged ? 0xffffffffu : 1u);
        const long long t1 = clock64();
        if (issuer) ptx::umma_commit(&bar);
        // bounded wait: never hang the box
        for (long long spin = 0; !ptx::mbar_try_wait(&bar, phase); ++spin)
            if (spin > (1LL << 24)) __trap();
        const long long t2 = clock64();
        if (issuer) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<512>(tmem);
}

template <int N>
static void run(int n_mma, int per_commit, int wait_each, int n_acc = 1, int conv_pattern = 0, int converged = 0) {
    unsigned long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k_issue<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int r = 0; r < 2; ++r) k_issue<N><<<1, 128, 100 * 1024>>>(n_mma, per_commit, wait_each, n_acc, conv_pattern, converged, d);
    cudaDeviceSynchronize();
    unsigned long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("acc %d conv %d converged %d ", n_acc, conv_pattern, converged);
    printf("N=%3d  %4d MMAs, commit every %2d, %s: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (tensor floor %d)  %s\n", N, n_mma, per_commit,
           wait_each ? "wait after each commit" : "no waits              ", (double)h[0] / n_mma, (double)h[1] / n_mma, 128 * N / 256,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    run<64>(512, 8, 0, 1, 0, 0);
    run<64>(512, 8, 0, 1, 0, 1);
    run<128>(512, 8, 0, 1, 0, 1);
    run<256>(512, 8, 0, 1, 0, 1);
    run<64>(512, 8, 0, 4, 0, 1);
    run<64>(512, 8, 0, 1, 1, 1);
    run<64>(512, 8, 1, 1, 1, 1);
    return 0;
}
