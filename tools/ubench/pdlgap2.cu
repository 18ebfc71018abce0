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
    const unsigned long long until = t1 + (unsigned long long)spin_ns + (unsigned long long)((blockIdx.x * 37) % 16) * 200ull;
    while (gtime() < until) { }
    if ((feat & 2) && tid == 32) {                                  // bulk copy
        mbar_arrive_expect_tx(smem_u32(&bars[1]), 4096u);
        bulk_g2s(smem_u32(smem) + 65536, src + (size_t)blockIdx.x * 4096, 4096u, smem_u32(&bars[1]));
        mbar_wait(smem_u32(&bars[1]), 0);
    }
    if ((feat & 16) && tid == 64) {                                 // tensor-map load: 64 x 64 bytes box
        mbar_arrive_expect_tx(smem_u32(&bars[3]), 4096u);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(smem) + 98304), "l"(&tm), "r"(0), "r"((int)(blockIdx.x % 32) * 64), "r"(smem_u32(&bars[3])) : "memory");
        mbar_wait(smem_u32(&bars[3]), 0);
    }
    if ((feat & 1) && warp == 0) {                                  // one MMA on whatever the shared memory holds
        const uint32_t idesc = make_idesc_i8(32);
        const uint64_t ad = make_desc(smem_u32(smem), 2048, 128), bd = make_desc(smem_u32(smem) + 32768, 32 * 16, 128);
        if (elect_one()) { mma_i8(tmem, ad, bd, idesc, 0); mma_commit(smem_u32(&bars[0])); }
        __syncwarp();
        mbar_wait(smem_u32(&bars[0]), 0);
        tc_fence_after();
    }
    if ((feat & 4) && warp == 0) {
        int acc[16];
        tmem_ld16(tmem, acc);
        tmem_ld_wait16(acc);
        if (acc[0] == 0x12345678) out[0] = 1;
    }
    if ((feat & 8) && warp == 1) {                                  // a wait that goes through nanosleep: lane 0 arrives late
        if ((tid & 31) == 0) { __nanosleep(2000); mbar_arrive(smem_u32(&bars[2])); }
        mbar_wait_relaxed<256>(smem_u32(&bars[2]), 0);
    }
    tc_fence_before();
    __syncthreads();
    unsigned long long t2 = gtime();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
    if (tid == 0) {
        long long* r = out + 8 + ((size_t)idx * gridDim.x + blockIdx.x) * 3;
        r[0] = (long long)t0; r[1] = (long long)t1; r[2] = (long long)t2;
    }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const int NK = 24, G = 148;
    long long* d; unsigned char* src;
    CKX(cudaMalloc(&d, sizeof(long long) * (8 + NK * G * 3)));
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
    const int feats[] = {0, 1, 2, 4 | 1, 8, 16, 32, 1 | 2 | 4 | 8 | 16 | 32};
    const char* names[] = {"nothing", "mma+commit", "bulk copy", "mma+commit+tcgen05.ld", "nanosleep wait", "tensor-map load", "prefetch.tensormap", "all"};
    for (int f = 0; f < 8; ++f) {
        CKX(cudaMemset(d, 0, sizeof(long long) * (8 + NK * G * 3)));
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
        std::vector<long long> h((size_t)8 + NK * G * 3);
        CKX(cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        std::vector<double> a, b;
        for (int i = 1; i < NK; ++i) {
            long long pfe = 0, ple = 0, fi = 0, fg = 0;
            for (int k = 0; k < G; ++k) {
                const long long* p = &h[8 + ((size_t)(i - 1) * G + k) * 3];
                const long long* q = &h[8 + ((size_t)i * G + k) * 3];
                if (!k || p[2] < pfe) pfe = p[2];
                if (!k || p[2] > ple) ple = p[2];
                if (!k || q[0] < fi) fi = q[0];
                if (!k || q[1] < fg) fg = q[1];
            }
            a.push_back((fi - pfe) / 1e3); b.push_back((fg - ple) / 1e3);
        }
        std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
        printf("%-24s | entry - prev first exit %6.2f us | go - prev last exit %6.2f us\n", names[f], a[a.size() / 2], b[b.size() / 2]);
        CKX(cudaGraphExecDestroy(ge)); CKX(cudaGraphDestroy(g));
    }
    return 0;
}
