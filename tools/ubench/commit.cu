// Micro-benchmark: cost of tcgen05.mma + tcgen05.commit (round trip and pipelined), kind::i8 M=128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o commit commit.cu && ./commit
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../alpha_yolo_quant_b200/csrc/conv_tma.cuh"
using namespace ayq::tc;

__global__ void __launch_bounds__(128, 1) k(int N, int nmma, int iters, int mode, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[16];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += 128) ((uint32_t*)smem)[i] = 0x01010101u;
    if (tid == 0) {
        for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp == 0) {
        const uint32_t idesc = make_idesc_i8(N);
        const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem) + 32768;
        uint32_t phase[16] = {0};
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int b = it & 15;
            if (mode == 1 && it >= 16) { mbar_wait(smem_u32(&bars[b]), phase[b]); phase[b] ^= 1; }
            if (elect_one()) {
            for (int j = 0; j < nmma; ++j)
                mma_i8(tmem + (uint32_t)(((mode == 5 || mode == 6) ? 0 : (it & 3)) * N), make_desc(abase + (j & 3) * 4096, 2048, 128), make_desc(bbase + (j & 3) * 2 * N * 16, N * 16, 128), idesc, (mode == 5 || mode == 7) ? 1u : (uint32_t)(j > 0));
            if (mode <= 1) mma_commit(smem_u32(&bars[b]));
            if (mode == 3 && (it & 3) == 3) mma_commit(smem_u32(&bars[15]));     // never waited on
            if (mode == 4) { mma_commit(smem_u32(&bars[14])); mma_commit(smem_u32(&bars[15])); }
            }
            __syncwarp();
            if (mode == 0) { mbar_wait(smem_u32(&bars[b]), phase[b]); phase[b] ^= 1; }
        }
        long long t1 = clock64();
        if (tid == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 2000;
    for (int mode = 0; mode < 5; ++mode)
        for (int N : {16, 128})
            for (int nmma : {1, 2, 8}) {
                k<<<1, 128, 72 * 1024>>>(N, nmma, iters, mode, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                printf("mode %d (%s) N=%3d mma/iter=%d: %.1f cycles/iter  (%s)\n", mode, mode == 0 ? "commit+wait round trip" : mode == 1 ? "commit, wait 16 later" : mode == 2 ? "no commit" : mode == 3 ? "commit every 4th iter" : mode == 4 ? "two commits/iter, no wait" : mode == 5 ? "no commit, acc=1, fixed D" : mode == 6 ? "no commit, acc=0 first, fixed D" : "no commit, acc=1, rotating D",
                       N, nmma, (double)c / iters, cudaGetErrorString(e));
            }
    return 0;
}
