// pipes.cu -- issue rate of the instructions the fixed-point epilogue is made of, alone and in pairs (which share a pipe?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Every test: 8 warps per SM sub-partition, 16 independent dependency chains per thread, inline PTX so that nothing folds.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

enum { FADD_RN, FADD_RM, FADD_RM_SAT, FMUL_RN, FFMA_RN, IADD, LEA_, F2I_S8, I2IP_, LDS32, MIX_FADD_FMUL, MIX_FFMA_FADDRM, MIX_FMUL_IADD, MIX_FADDRM_FADDRM_SAT, MIX_F2I_LDS,
       MIX_FMUL_FADDRM, MIX_FFMA_FMUL, EPI_NEW, NTEST };
static const char* NAMES[] = {"FADD.rn", "FADD.rm", "FADD.rm.sat", "FMUL", "FFMA", "IADD3", "LEA(shl+add)", "F2I.S8.floor", "I2IP pack", "LDS.32 (conflict-free)",
                              "FADD + FMUL", "FFMA + FADD.rm", "FMUL + IADD3", "FADD.rm + FADD.rm.sat", "F2I + LDS", "FMUL + FADD.rm", "FFMA + FMUL", "epilogue mix (7 fp, 2.5 int, LDS, F2I)"};
static const int PER_ITER[] = {16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 32, 32, 32, 32, 32, 32, 32, 16};

template <int T>
__global__ void __launch_bounds__(1024, 1) k(int iters, float fa, float fb, int ia, unsigned* sink, long long* cyc) {
    __shared__ float tab[257 * 32];
    for (int i = threadIdx.x; i < 257 * 32; i += blockDim.x) tab[i] = (float)(i & 127);
    __syncthreads();
    float f[16]; int v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { f[j] = fa + j * 0.001f + threadIdx.x * 1e-5f; v[j] = ia + j + threadIdx.x; }
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(tab) + (threadIdx.x & 31) * 4;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (T == FADD_RN) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(fb));
            if (T == FADD_RM) asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(fb));
            if (T == FADD_RM_SAT) asm volatile("add.rm.sat.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(fb));
            if (T == FMUL_RN) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(fb));
            if (T == FFMA_RN) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(fb), "f"(fa));
            if (T == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(v[j]) : "r"(ia));
            if (T == LEA_) asm volatile("{.reg .b32 t; shl.b32 t, %0, 7; add.s32 %0, t, %1;}" : "+r"(v[j]) : "r"(ia));
            if (T == F2I_S8) asm volatile("cvt.rmi.sat.s8.f32 %0, %1;" : "=r"(v[j]) : "f"(f[j]));
            if (T == I2IP_) asm volatile("cvt.pack.sat.s8.s32.b32 %0, %0, %1, %2;" : "+r"(v[j]) : "r"(ia), "r"(v[(j + 1) & 15]));
            if (T == LDS32) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f[j]) : "r"(sbase + ((unsigned)(v[j] & 255) << 7)));
            if (T == MIX_FADD_FMUL) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(fb)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[(j + 8) & 15]) : "f"(fa)); }
            if (T == MIX_FFMA_FADDRM) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(fb), "f"(fa)); asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(f[(j + 8) & 15]) : "f"(fa)); }
            if (T == MIX_FMUL_IADD) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(fb)); asm volatile("add.s32 %0, %0, %1;" : "+r"(v[j]) : "r"(ia)); }
            if (T == MIX_FADDRM_FADDRM_SAT) { asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(fb)); asm volatile("add.rm.sat.f32 %0, %0, %1;" : "+f"(f[(j + 8) & 15]) : "f"(fa)); }
            if (T == MIX_F2I_LDS) { asm volatile("cvt.rmi.sat.s8.f32 %0, %1;" : "=r"(v[j]) : "f"(f[j])); asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f[(j + 8) & 15]) : "r"(sbase + ((unsigned)(v[(j + 4) & 15] & 255) << 7))); }
            if (T == MIX_FMUL_FADDRM) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(fb)); asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(f[(j + 8) & 15]) : "f"(fa)); }
            if (T == MIX_FFMA_FMUL) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[j]) : "f"(fb), "f"(fa)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[(j + 8) & 15]) : "f"(fa)); }
            if (T == EPI_NEW) {   // the MAGIC2 instruction mix on one element
                int a = v[j]; float af, t, y, w, l, pr, z; int r;
                asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "r"(ia));
                asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(af) : "f"(__int_as_float(a)), "f"(-12582912.f));
                asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(t) : "f"(af), "f"(fb));
                asm volatile("add.rm.sat.f32 %0, %1, %2;" : "=f"(y) : "f"(t), "f"(0.501953125f));
                asm volatile("add.rm.f32 %0, %1, %2;" : "=f"(w) : "f"(y), "f"(32768.f));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l) : "r"((__float_as_uint(w) << 7) + sbase + 0x80000000u));
                asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(pr) : "f"(l), "f"(af));
                asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(z) : "f"(pr), "f"(fa));
                asm volatile("add.rm.f32 %0, %1, %2;" : "=f"(z) : "f"(z), "f"(0.5f));
                asm volatile("cvt.rmi.sat.s8.f32 %0, %1;" : "=r"(r) : "f"(z));
                if (j & 1) asm volatile("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(v[j]) : "r"(r), "r"(v[j - 1]), "r"(0));
                else v[j] = r + it;
            }
        }
    }
    const long long t1 = clock64();
    unsigned x = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) x ^= __float_as_uint(f[j]) ^ (unsigned)v[j];
    if (x == 0x12345678u) sink[0] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int T>
static void run(unsigned* sink, long long* d_cyc) {
    const int iters = 1000, warps = 8;
    k<T><<<148, 128 * warps>>>(10, 1.0f, 1.0001f, 3, sink, d_cyc);
    CK(cudaDeviceSynchronize());
    k<T><<<148, 128 * warps>>>(iters, 1.0f, 1.0001f, 3, sink, d_cyc);
    CK(cudaDeviceSynchronize());
    std::vector<long long> c(148);
    cudaMemcpy(c.data(), d_cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (auto v : c) mean += (double)v / 148;
    printf("%-42s %6.2f cycles per warp instruction per SM sub-partition (%d instr / iteration / thread)\n", NAMES[T], mean / ((double)iters * PER_ITER[T] * warps), PER_ITER[T]);
}
template <int T> struct Runner { static void go(unsigned* s, long long* c) { run<T>(s, c); Runner<T + 1>::go(s, c); } };
template <> struct Runner<NTEST> { static void go(unsigned*, long long*) {} };

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    unsigned* sink; long long* d_cyc;
    CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&d_cyc, 148 * 8));
    Runner<0>::go(sink, d_cyc);
    return 0;
}
