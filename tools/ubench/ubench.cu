// ubench.cu -- B200 micro-benchmarks behind the design decisions of the conv epilogue and the roofline denominators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o ubench ubench.cu && ./ubench
// 1. issue / pipe throughput of the instructions the fixed-point epilogue is made of (F2I.S8.FLOOR, FADD.RM, FFMA, LDS.64 gather)
// 2. the whole silu_magic epilogue from registers (no MMA, no global traffic) at 2..6 warps per SMSP
// 3. dense int8 tcgen05.mma peak (kind::i8, M = 128, N = 64 / 128 / 256, K = 32) from resident shared-memory operands
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../alpha_yolo_quant_b200/csrc/fixedpoint.cuh"
using namespace ayq;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Coef { float k1[16], c1[16], k2[16]; int bias[16]; };

// mode 0: full silu_magic; 1: F2I only; 2: FADD.RM only; 3: FFMA only; 4: LDS.64 gather only; 5: silu_magic with the table lookup
// replaced by a constant; 6: silu_magic with both conversions replaced by magic adds + clamps (no XU)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) epi_kernel(const __grid_constant__ Coef cf, int iters, int seed, unsigned* sink, long long* cyc) {
    __shared__ float2 lut2[256];
    extern __shared__ __align__(128) float lut_rep[];
    if (MODE == 7) {
        for (int i = threadIdx.x; i < AYQ_LUTREP_N * 32; i += blockDim.x) lut_rep[i] = (float)((((i >> 5) * 3) & 127));
    }
    const uint32_t lut_thr = (uint32_t)__cvta_generic_to_shared(lut_rep) + ((threadIdx.x & 31u) << 2) + 0x80000000u;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const float l = (float)(i < 128 ? (i * 3) & 127 : 127 - ((i * 5) & 63));
        lut2[i] = make_float2(l, -__fmul_rn(l, AYQ_MAGIC_F));
    }
    __syncthreads();
    int acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = (int)((threadIdx.x * 2654435761u + j * 40503u + seed) >> 12) - (1 << 19);
    unsigned x = 0;
    const float half = 0.5f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        int r[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int v = acc[j] + cf.bias[j];
            if (MODE == 0) r[j] = silu_magic(v, cf.k1[j], cf.c1[j], cf.k2[j], lut2, half);
            else if (MODE == 7) r[j] = silu_magic2(v, cf.k1[j] * 0.00390625f, cf.k2[j], lut_thr, half);
            else if (MODE == 1) r[j] = floor_sat_s8(__int_as_float(v));
            else if (MODE == 2) r[j] = __float_as_int(__fadd_rd(__int_as_float(v), half));
            else if (MODE == 3) r[j] = __float_as_int(__fmaf_rn(cf.k1[j], __int_as_float(v), cf.c1[j]));
            else if (MODE == 4) { const float2 l = lut2[(v >> 3) & 255]; r[j] = __float_as_int(l.x) ^ __float_as_int(l.y); }
            else if (MODE == 5) {
                const float m = __int_as_float(v);
                const int r1 = floor_sat_s8(__fadd_rd(__fmaf_rn(cf.k1[j], m, cf.c1[j]), half));
                const float pr = __fmaf_rn((float)1.0f + cf.k2[j], m, (float)r1);
                r[j] = floor_sat_s8(__fadd_rd(__fmul_rn(cf.k2[j], pr), half));
            } else {
                const float m = __int_as_float(v);
                const float t = fminf(fmaxf(__fmaf_rn(cf.k1[j], m, cf.c1[j]), -128.f), 127.25f);
                const int i1 = __float_as_int(__fadd_rd(t, 6291456.5f));           // 1.5 * 2^22 + 0.5: ulp 0.5, low bits = 2 * floor(t + 0.5) (+ half bit)
                const float2 l = *(const float2*)((const char*)lut2 + ((i1 << 2) & 0x7f8));
                const float pr = __fmaf_rn(l.x, m, l.y);
                const float z = fminf(fmaxf(__fmul_rn(cf.k2[j], pr), -128.f), 127.25f);
                r[j] = __float_as_int(__fadd_rd(z, 6291456.5f)) >> 1;
            }
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4) x ^= pack4_sat(r[j], r[j + 1], r[j + 2], r[j + 3]);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += (int)(x & 7) + it;
    }
    const long long t1 = clock64();
    if (x == 0x12345678u) sink[0] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run_epi(const char* name, int warps_per_smsp, const Coef& cf, unsigned* sink, long long* d_cyc) {
    const int threads = 128 * warps_per_smsp, iters = 2000;
    const size_t dyn = MODE == 7 ? AYQ_LUTREP_BYTES : 0;
    CK(cudaFuncSetAttribute(epi_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    epi_kernel<MODE><<<148, threads, dyn>>>(cf, 10, 1, sink, d_cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    epi_kernel<MODE><<<148, threads, dyn>>>(cf, iters, 2, sink, d_cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> c(148);
    cudaMemcpy(c.data(), d_cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (auto v : c) mean += (double)v / 148;
    // cycles per 32-element row per SMSP = cycles / (iters * 16 elements * warps_per_smsp)
    printf("%-28s %d warps/SMSP: %8.0f cycles, %6.2f cycles per element-row per SMSP, %.3f ms, %.1f Gelem/s chip\n", name, warps_per_smsp, mean,
           mean / ((double)iters * 16 * warps_per_smsp), ms, 148.0 * threads * 16 * iters / ms * 1e-6);
}

// ---- tcgen05 int8 peak -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc_i8(int N) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int N, int iters, int nacc, long long* cyc) {
    extern __shared__ __align__(1024) unsigned char smem[];      // A: 8 K-steps x 4 KB, B: 8 x N*32
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (8 * 4096 + 8 * N * 32) / 4; i += 128) ((unsigned*)smem)[i] = 0x01010101u * (i & 3);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(1u) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[1])), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_i8(N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 8 * 4096);
        // Two barriers used alternately, each with at most ONE phase outstanding: batch `it` commits to bars[it & 1]; before batch
        // it + 1 is issued, batch it - 1 (the previous user of that barrier) is waited for.  Bounded spin: a protocol bug traps.
        auto wait_bar = [&](int it) {
            const uint32_t bar = smem_u32(&bars[it & 1]), parity = (uint32_t)(it >> 1) & 1u;
            for (unsigned spin = 0;; ++spin) {
                uint32_t done;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                if (done) break;
                if (spin > (1u << 24)) __trap();
            }
        };
        uint64_t ad[8], bd[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) { ad[ks] = make_desc(a0 + ks * 4096, 2048, 128); bd[ks] = make_desc(b0 + ks * N * 32, N * 16, 128); }
        const uint32_t d0 = tmem_base, d1 = tmem_base + (uint32_t)(nacc > 1 ? N : 0);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
            for (int k8 = 0; k8 < 8; ++k8) {                       // 64 MMAs per batch; descriptors precomputed, no index arithmetic
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) mma_i8((ks & 1) ? d1 : d0, ad[ks], bd[ks], idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[it & 1])) : "memory");
            if (it >= 1) wait_bar(it - 1);
        }
        wait_bar(iters - 1);
        cyc[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

static void run_mma(int N, int nacc, long long* d_cyc) {
    const int iters = 400;
    const size_t smem = 8 * 4096 + 8 * (size_t)N * 32;
    CK(cudaFuncSetAttribute(mma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    mma_peak_kernel<<<148, 128, smem>>>(N, 10, nacc, d_cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        mma_peak_kernel<<<148, 128, smem>>>(N, iters, nacc, d_cyc);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    std::vector<long long> c(148);
    cudaMemcpy(c.data(), d_cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (auto v : c) mean += (double)v / 148;
    const double ops = 2.0 * 128 * N * 32 * 64 * iters * 148;
    printf("tcgen05.mma kind::i8 M=128 N=%3d K=32, %d accumulators: %.3f ms, %7.1f TOP/s dense int8, %.1f cycles per MMA (SM clock)\n", N, nacc, best, ops / best * 1e-9,
           mean / (64.0 * iters));
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IONBF, 0);
    unsigned* sink; long long* d_cyc;
    CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&d_cyc, 148 * 8));
    Coef cf;
    for (int j = 0; j < 16; ++j) { cf.k1[j] = 0.0001f * (j + 3); cf.c1[j] = -cf.k1[j] * AYQ_MAGIC_F; cf.k2[j] = 0.00002f * (j + 5); cf.bias[j] = AYQ_MAGIC_I + 17 * j; }
    for (int w : {2, 4, 6, 8}) run_epi<0>("silu_magic (full)", w, cf, sink, d_cyc);
    for (int w : {2, 4, 6, 8}) run_epi<7>("silu_magic2 (product)", w, cf, sink, d_cyc);
    for (int w : {4, 6}) run_epi<5>("silu_magic, no table", w, cf, sink, d_cyc);
    for (int w : {4, 6, 8}) run_epi<6>("silu, no XU (magic adds)", w, cf, sink, d_cyc);
    for (int w : {4, 8}) run_epi<1>("F2I.S8.FLOOR only", w, cf, sink, d_cyc);
    for (int w : {4, 8}) run_epi<2>("FADD.RM only", w, cf, sink, d_cyc);
    for (int w : {4, 8}) run_epi<3>("FFMA only", w, cf, sink, d_cyc);
    for (int w : {4, 8}) run_epi<4>("LDS.64 gather only", w, cf, sink, d_cyc);
    for (int N : {16, 32, 64, 128, 256}) run_mma(N, 512 / N >= 2 ? 2 : 1, d_cyc);
    run_mma(256, 1, d_cyc);
    run_mma(128, 4, d_cyc);
    return 0;
}
