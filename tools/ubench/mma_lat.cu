// Where do the ~400 cycles per tcgen05.mma-issuing loop iteration go?  Stamps clock64 around the pieces.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../alpha_yolo_quant_b200/csrc/conv_tma.cuh"
using namespace ayq::tc;

__global__ void __launch_bounds__(128, 1) k(int N, int iters, int mode, long long* out) {
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
        const uint64_t ad = make_desc(abase, 2048, 128), bd = make_desc(bbase, N * 16, 128);
        long long tA = 0, tB = 0, tC = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            long long a0 = clock64();
            if (mode == 0) {                      // elect + mma + syncwarp
                if (elect_one()) mma_i8(tmem, ad, bd, idesc, 1);
                long long a1 = clock64();
                __syncwarp();
                long long a2 = clock64();
                tA += a1 - a0; tB += a2 - a1;
            } else if (mode == 1) {               // elect + mma, no syncwarp
                if (elect_one()) mma_i8(tmem, ad, bd, idesc, 1);
                long long a1 = clock64();
                tA += a1 - a0;
            } else if (mode == 2) {               // elect + 4 mma
                if (elect_one()) { mma_i8(tmem, ad, bd, idesc, 1); mma_i8(tmem, ad, bd, idesc, 1); mma_i8(tmem, ad, bd, idesc, 1); mma_i8(tmem, ad, bd, idesc, 1); }
                long long a1 = clock64();
                tA += a1 - a0;
            } else if (mode == 3) {               // elect + mma + commit
                if (elect_one()) { mma_i8(tmem, ad, bd, idesc, 1); mma_commit(smem_u32(&bars[15])); }
                long long a1 = clock64();
                tA += a1 - a0;
            } else if (mode == 4) {               // commit only
                if (elect_one()) mma_commit(smem_u32(&bars[15]));
                long long a1 = clock64();
                tA += a1 - a0;
            } else {                              // nothing: loop + clock overhead
                long long a1 = clock64();
                tA += a1 - a0;
            }
            tC += clock64() - a0;
        }
        long long t1 = clock64();
        if (tid == 0) { out[0] = t1 - t0; out[1] = tA; out[2] = tB; out[3] = tC; }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 2000;
    const char* names[] = {"elect+mma+syncwarp", "elect+mma", "elect+4mma", "elect+mma+commit", "elect+commit", "empty"};
    for (int mode = 0; mode < 6; ++mode)
        for (int N : {16, 128}) {
            k<<<1, 128, 72 * 1024>>>(N, iters, mode, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long c[4]; cudaMemcpy(c, d, 32, cudaMemcpyDeviceToHost);
            printf("%-20s N=%3d: total %.1f/iter  issue-part %.1f  syncwarp-part %.1f  (%s)\n", names[mode], N, (double)c[0] / iters, (double)c[1] / iters,
                   (double)c[2] / iters, cudaGetErrorString(e));
        }
    return 0;
}
