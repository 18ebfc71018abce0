// Which operation of the conv kernel's main loop makes an SM slow to take the next kernel's CTA?  The chain of pdlgap.cu (148 CTAs
// of 1024 threads, 200 KB shared memory, 512 TMEM columns, early launch_dependents), each CTA additionally doing ONE of:
//   bit 0: one tcgen05.mma (i8, M128 N32 K32) + commit + mbarrier wait       bit 1: one bulk copy global -> shared (4 KB) + wait
//   bit 2: one tcgen05.ld of the accumulator                                  bit 3: one nanosleep-backed mbarrier wait loop
//   bit 4: one tensor-map TMA load (2-D box)                                  bit 5: prefetch.tensormap only
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../alpha_yolo_quant_b200/csrc -o pdlgap2 pdlgap2.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <map>
#include <algorithm>
#include "plan_format.h"
#include "kernels.cuh"
#include "conv_tc.cuh"
#include "conv_tma.cuh"
using namespace ayq::tc;

#define CKX(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void __launch_bounds__(1024, 1) chain2(long long* out, int idx, int spin_ns, int feat, const unsigned char* src, const __grid_constant__ CUtensorMap tm) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[4];
    __shared__ uint32_t tmem_base_s;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    unsigned long long t0 = gtime();
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (feat & 32) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    unsigned long long t1 = gtime();
    // feat: 1 try_wait loop (default hint), arrival 2 us late   2 explicit nanosleep(256) + try_wait loop   3 one lane nanosleep(2000)
    //       4 try_wait with a 200 ns time hint                   5 test_wait polling (no suspend)          6 whole warp nanosleep(100) once
    //       +16: do it BEFORE the 20 us spin instead of right before the exit
    auto waits = [&](int v) {
        const uint32_t bar = smem_u32(&bars[2]);
        const int lane = tid & 31;
        if (v == 1) { if (warp == 1) { if (lane == 0) { const unsigned long long u = gtime() + 2000ull; while (gtime() < u) { } } __syncwarp(); } }        // A: one lane spins, no barrier
        else if (v == 2) { if (warp == 1) { const unsigned long long u = gtime() + 2000ull; while (gtime() < u) { } } }                                    // B: whole warp spins
        else if (v == 3) { if (warp == 1) { if (lane == 0) mbar_arrive(bar); __syncwarp(); uint32_t done = 0;
                           while (!done) asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(0u) : "memory"); } }   // C
        else if (v == 4 || v == 5) {                                                                                                                       // D / E: arrive from another warp
            if (warp == 2) { if (lane == 0) { const unsigned long long u = gtime() + 2000ull; while (gtime() < u) { } mbar_arrive(bar); } __syncwarp(); }
            if (warp == 1) { if (v == 4) { uint32_t done = 0;
                             while (!done) asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(0u) : "memory"); }
                             else mbar_wait(bar, 0); }
        } else if (v == 6) { if (warp == 2 && lane == 0) mbar_arrive(bar); }                                                                               // F: an arrive nobody waits for
    };
    if (feat & 16) waits(feat & 15);
    const unsigned long long until = (feat & 16 ? gtime() : t1) + (unsigned long long)spin_ns + (unsigned long long)((blockIdx.x * 37) % 16) * 200ull;
    while (gtime() < until) { }
    if (!(feat & 16)) waits(feat & 15);
    tc_fence_before();
    __syncthreads();
    unsigned long long t2 = gtime();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
    if (tid == 0) {
        long long* r = out + 8 + ((size_t)idx * gridDim.x + blockIdx.x) * 4;
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        r[0] = (long long)t0; r[1] = (long long)t1; r[2] = (long long)t2; r[3] = (long long)smid;
    }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const int NK = 24, G = 148;
    long long* d; unsigned char* src;
    CKX(cudaMalloc(&d, sizeof(long long) * (8 + NK * G * 4)));
    CKX(cudaMalloc(&src, 1 << 20)); CKX(cudaMemset(src, 1, 1 << 20));
    CKX(cudaFuncSetAttribute(chain2, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    CKX(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    CUtensorMap tm;
    cuuint64_t dims[2] = {64, 2048}; cuuint64_t strides[1] = {64}; cuuint32_t box[2] = {64, 64}; cuuint32_t es[2] = {1, 1};
    CUresult cr = ((PFN_enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("tensor map encode failed %d\n", (int)cr); return 1; }
    cudaStream_t st; CKX(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const int feats[] = {0, 1, 2, 3, 4, 5, 6, 0, 4, 0};
    const char* names[] = {"nothing", "A one lane spins 2us", "B whole warp spins 2us", "C arrive + test_wait (no delay)", "D other warp arrives late, test_wait", "E same, try_wait", "F arrive only", "nothing (again)", "D again", "nothing (again)"};
    for (int f = 0; f < 10; ++f) {
        CKX(cudaMemset(d, 0, sizeof(long long) * (8 + NK * G * 4)));
        cudaGraph_t g; cudaGraphExec_t ge;
        CKX(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < NK; ++i) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 200 * 1024; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CKX(cudaLaunchKernelEx(&cfg, chain2, d, i, 20000, feats[f], (const unsigned char*)src, tm));
        }
        CKX(cudaStreamEndCapture(st, &g));
        CKX(cudaGraphInstantiate(&ge, g, 0));
        CKX(cudaGraphLaunch(ge, st)); CKX(cudaStreamSynchronize(st));
        CKX(cudaGraphLaunch(ge, st)); CKX(cudaStreamSynchronize(st));
        std::vector<long long> h((size_t)8 + NK * G * 4);
        CKX(cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        std::vector<double> a, b;
        for (int i = 1; i < NK; ++i) {
            long long pfe = 0, ple = 0, fi = 0, fg = 0;
            for (int k = 0; k < G; ++k) {
                const long long* p = &h[8 + ((size_t)(i - 1) * G + k) * 4];
                const long long* q = &h[8 + ((size_t)i * G + k) * 4];
                if (!k || p[2] < pfe) pfe = p[2];
                if (!k || p[2] > ple) ple = p[2];
                if (!k || q[0] < fi) fi = q[0];
                if (!k || q[1] < fg) fg = q[1];
            }
            a.push_back((fi - pfe) / 1e3); b.push_back((fg - ple) / 1e3);
        }
        std::vector<double> lag;
        for (int i = 1; i < NK; ++i) {
            std::map<long long, long long> ex;
            for (int k = 0; k < G; ++k) { const long long* p = &h[8 + ((size_t)(i - 1) * G + k) * 4]; ex[p[3]] = p[2]; }
            for (int k = 0; k < G; ++k) { const long long* q = &h[8 + ((size_t)i * G + k) * 4]; if (ex.count(q[3])) lag.push_back((q[0] - ex[q[3]]) / 1e3); }
        }
        std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end()); std::sort(lag.begin(), lag.end());
        printf("%-34s | entry - prev first exit %6.2f us | go - prev last exit %6.2f us | same-SM exit -> entry: min %5.2f median %5.2f p90 %5.2f max %5.2f\n", names[f], a[a.size() / 2], b[b.size() / 2],
               lag[0], lag[lag.size() / 2], lag[lag.size() * 9 / 10], lag.back());
        CKX(cudaGraphExecDestroy(ge)); CKX(cudaGraphDestroy(g));
    }
    return 0;
}
