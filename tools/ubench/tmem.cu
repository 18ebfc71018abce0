// tmem.cu -- tcgen05.ld (TMEM -> registers) throughput and whether it competes with instruction issue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem tmem.cu && ./tmem
// One CTA per SM, 8 or 16 warps.  "ld" warps (one per SM sub-partition x G groups) run back-to-back tcgen05.ld.32x32b.x16 + wait::ld
// on an allocated (uninitialised) TMEM block; "fma" warps run a dependent-free FFMA stream.  Three runs: ld only, fma only, both.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__global__ void __launch_bounds__(1024, 1) k(int iters, int ld_warps, int fma_warps, float fa, unsigned* sink, long long* cyc) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
    unsigned x = 0;
    long long t0 = clock64(), t1 = t0;
    if (warp < ld_warps) {
        int r[32];
        for (int it = 0; it < iters; ++it) {
            const uint32_t col = (uint32_t)((it & 7) * 32);
            if (X == 16) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                               "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(tb + col) : "memory");
            } else {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                               "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                               "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                               "=r"(r[30]), "=r"(r[31]) : "r"(tb + col) : "memory");
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            x ^= (unsigned)r[0] ^ (unsigned)r[X - 1];
        }
        t1 = clock64();
        if (threadIdx.x == 0) cyc[blockIdx.x * 2] = t1 - t0;
    } else if (warp >= 16 && warp < 16 + fma_warps) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = fa + j;
        for (int it = 0; it < iters * 2; ++it) {
#pragma unroll
            for (int j = 0; j < 16; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[j]) : "f"(fa));
        }
        t1 = clock64();
#pragma unroll
        for (int j = 0; j < 16; ++j) x ^= __float_as_uint(f[j]);
        if (threadIdx.x == 16 * 32) cyc[blockIdx.x * 2 + 1] = t1 - t0;
    }
    if (x == 0x12345678u) sink[0] = x;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512u) : "memory");
}

template <int X>
static void run(const char* name, int ld_warps, int fma_warps, unsigned* sink, long long* d_cyc) {
    const int iters = 4000;
    CK(cudaMemset(d_cyc, 0, 148 * 16));
    k<X><<<148, 1024>>>(iters, ld_warps, fma_warps, 1.0001f, sink, d_cyc);
    CK(cudaDeviceSynchronize());
    std::vector<long long> c(296);
    cudaMemcpy(c.data(), d_cyc, 296 * 8, cudaMemcpyDeviceToHost);
    double ld = 0, fm = 0;
    for (int i = 0; i < 148; ++i) { ld += (double)c[2 * i] / 148; fm += (double)c[2 * i + 1] / 148; }
    printf("%-34s ld warps %2d (x%d), fma warps %2d:", name, ld_warps, X, fma_warps);
    if (ld_warps) printf("  %7.1f cycles per tcgen05.ld per warp = %5.1f B/clk/SM TMEM read,", ld / iters, (double)ld_warps * 32 * X * 4 * iters / ld);
    if (fma_warps) printf("  %5.2f cycles per FFMA per sub-partition", fm / (iters * 2.0 * 16 * (fma_warps / 4)));
    printf("\n");
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    unsigned* sink; long long* d_cyc;
    CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&d_cyc, 148 * 16));
    run<16>("tcgen05.ld alone", 4, 0, sink, d_cyc);
    run<16>("tcgen05.ld alone", 8, 0, sink, d_cyc);
    run<16>("tcgen05.ld alone", 16, 0, sink, d_cyc);
    run<32>("tcgen05.ld alone", 4, 0, sink, d_cyc);
    run<32>("tcgen05.ld alone", 16, 0, sink, d_cyc);
    run<16>("FFMA alone", 0, 4, sink, d_cyc);
    run<16>("FFMA alone", 0, 8, sink, d_cyc);
    run<16>("both", 4, 4, sink, d_cyc);
    run<16>("both", 8, 8, sink, d_cyc);
    run<16>("both", 16, 8, sink, d_cyc);
    run<32>("both", 16, 8, sink, d_cyc);
    return 0;
}
