// Launch-to-launch gap between consecutive dependent kernels (programmatic dependent launch) as a function of the CTA's resource
// footprint: dynamic shared memory, threads, kernel parameter bytes, TMEM allocation.  Every CTA stamps %globaltimer at entry, after
// griddepcontrol.wait and at exit; the host prints, per configuration, the medians over the kernel boundaries of
//   entry(i+1).first - exit(i).first   (how soon a freed SM runs the next kernel's CTA)
//   go(i+1).first - exit(i).last       (dead time between the kernels' useful parts)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pdlgap pdlgap.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int PB> struct Pad { char b[PB]; };

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <int PB, bool TMEM, int VAR = 0>
__global__ void __launch_bounds__(1024, 1) chain_kernel(long long* out, int idx, int spin_ns, int early, const __grid_constant__ Pad<PB> pad, uint4* sink_g) {
    extern __shared__ unsigned char sm[];
    __shared__ uint32_t tbase;
    __shared__ unsigned char sink;
    if (early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    unsigned long long t0 = gtime();
    if (TMEM && threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) sink = pad.b[idx % PB];
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    unsigned long long t1 = gtime();
    const unsigned long long until = t1 + (unsigned long long)spin_ns + (unsigned long long)((blockIdx.x * 37) % 16) * 200ull;   // exits spread over 3 us
    while (gtime() < until) { }
    if (sink_g) {                                                  // a burst of stores right before the exit (what a conv epilogue leaves in flight)
        uint4* dst = sink_g + ((size_t)blockIdx.x * 64 + VAR) * blockDim.x + threadIdx.x;
#pragma unroll 8
        for (int r = 0; r < 64; ++r) dst[(size_t)r * blockDim.x] = make_uint4(r, idx, VAR, 0);
    }
    __syncthreads();
    unsigned long long t2 = gtime();
    if (TMEM && threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
    if (!early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        long long* r = out + ((size_t)idx * gridDim.x + blockIdx.x) * 3;
        r[0] = (long long)t0; r[1] = (long long)t1; r[2] = (long long)t2;
    }
}

template <int PB, bool TMEM>
static void run(const char* name, int threads, int smem, int pdl, int early, int graph) {
    const int NK = 24, G = 148;
    long long* d;
    CK(cudaMalloc(&d, sizeof(long long) * NK * G * 3));
    CK(cudaMemset(d, 0, sizeof(long long) * NK * G * 3));
    CK(cudaFuncSetAttribute(chain_kernel<PB, TMEM, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    Pad<PB> pad{};
    auto enqueue = [&]() {
        for (int i = 0; i < NK; ++i) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
            CK(cudaLaunchKernelEx(&cfg, chain_kernel<PB, TMEM, 0>, d, i, 20000, early, pad, (uint4*)nullptr));
        }
    };
    if (graph) {
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        enqueue();
        CK(cudaStreamEndCapture(st, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
        CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
    } else {
        enqueue(); CK(cudaStreamSynchronize(st));
        enqueue(); CK(cudaStreamSynchronize(st));
    }
    std::vector<long long> h((size_t)NK * G * 3);
    CK(cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    std::vector<double> a, b, c;
    for (int i = 1; i < NK; ++i) {
        long long pfe = 0, ple = 0, fi = 0, fg = 0;
        for (int k = 0; k < G; ++k) {
            const long long* p = &h[((size_t)(i - 1) * G + k) * 3];
            const long long* q = &h[((size_t)i * G + k) * 3];
            if (!k || p[2] < pfe) pfe = p[2];
            if (!k || p[2] > ple) ple = p[2];
            if (!k || q[0] < fi) fi = q[0];
            if (!k || q[1] < fg) fg = q[1];
        }
        a.push_back((fi - pfe) / 1e3); b.push_back((fg - ple) / 1e3); c.push_back((fi - ple) / 1e3);
    }
    std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end()); std::sort(c.begin(), c.end());
    printf("%-34s threads %4d smem %3d KB params %5d B tmem %d pdl %d early %d graph %d | entry - prev first exit %6.2f us | entry - prev last exit %6.2f us | go - prev last exit %6.2f us\n",
           name, threads, smem / 1024, PB, (int)TMEM, pdl, early, graph, a[a.size() / 2], c[c.size() / 2], b[b.size() / 2]);
    CK(cudaFree(d)); CK(cudaStreamDestroy(st));
}

template <int PB>
static void run_alt(const char* name, int smem_a, int smem_b, int stores) {
    const int NK = 24, G = 148, threads = 1024;
    long long* d; uint4* sink = nullptr;
    CK(cudaMalloc(&d, sizeof(long long) * NK * G * 3));
    CK(cudaMemset(d, 0, sizeof(long long) * NK * G * 3));
    if (stores) CK(cudaMalloc(&sink, sizeof(uint4) * (size_t)G * 64 * 1024 + 4096));
    CK(cudaFuncSetAttribute(chain_kernel<PB, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    CK(cudaFuncSetAttribute(chain_kernel<PB, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    Pad<PB> pad{};
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < NK; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(G); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = (i & 1) ? smem_b : smem_a; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (i & 1) CK(cudaLaunchKernelEx(&cfg, chain_kernel<PB, true, 1>, d, i, 20000, 1, pad, sink));
        else CK(cudaLaunchKernelEx(&cfg, chain_kernel<PB, true, 0>, d, i, 20000, 1, pad, sink));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
    CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
    std::vector<long long> h((size_t)NK * G * 3);
    CK(cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    std::vector<double> a, b;
    for (int i = 1; i < NK; ++i) {
        long long pfe = 0, ple = 0, fi = 0, fg = 0;
        for (int k = 0; k < G; ++k) {
            const long long* p = &h[((size_t)(i - 1) * G + k) * 3];
            const long long* q = &h[((size_t)i * G + k) * 3];
            if (!k || p[2] < pfe) pfe = p[2];
            if (!k || p[2] > ple) ple = p[2];
            if (!k || q[0] < fi) fi = q[0];
            if (!k || q[1] < fg) fg = q[1];
        }
        a.push_back((fi - pfe) / 1e3); b.push_back((fg - ple) / 1e3);
    }
    std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
    printf("%-34s alternating functions, smem %d / %d KB, params %d B, tmem, stores %d | entry - prev first exit %6.2f us | go - prev last exit %6.2f us\n",
           name, smem_a / 1024, smem_b / 1024, PB, stores, a[a.size() / 2], b[b.size() / 2]);
    CK(cudaFree(d)); if (sink) CK(cudaFree(sink)); CK(cudaStreamDestroy(st));
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    run<16, false>("small CTA, no pdl", 128, 0, 0, 1, 1);
    run<16, false>("small CTA", 128, 0, 1, 1, 1);
    run<16, false>("small CTA, stream", 128, 0, 1, 1, 0);
    run<16, false>("1024 threads", 1024, 0, 1, 1, 1);
    run<16, false>("128 threads, 200 KB", 128, 200 * 1024, 1, 1, 1);
    run<16, false>("1024 threads, 120 KB", 1024, 120 * 1024, 1, 1, 1);
    run<16, false>("1024 threads, 200 KB", 1024, 200 * 1024, 1, 1, 1);
    run<16, false>("1024 threads, 200 KB, no pdl", 1024, 200 * 1024, 0, 1, 1);
    run<16, false>("1024 threads, 200 KB, late trigger", 1024, 200 * 1024, 1, 0, 1);
    run<4000, false>("1024 thr, 200 KB, 4 KB params", 1024, 200 * 1024, 1, 1, 1);
    run<16, true>("1024 thr, 200 KB, tmem", 1024, 200 * 1024, 1, 1, 1);
    run<4000, true>("1024 thr, 200 KB, 4 KB, tmem", 1024, 200 * 1024, 1, 1, 1);
    run<4000, true>("same, stream launches", 1024, 200 * 1024, 1, 1, 0);
    run<6000, false>("1024 thr, 200 KB, 6 KB params", 1024, 200 * 1024, 1, 1, 1);
    run<6000, true>("1024 thr, 200 KB, 6 KB, tmem", 1024, 200 * 1024, 1, 1, 1);
    run<6000, true>("same, stream launches", 1024, 200 * 1024, 1, 1, 0);
    run<6000, true>("same, no pdl", 1024, 200 * 1024, 0, 1, 1);
    run<12000, true>("1024 thr, 200 KB, 12 KB, tmem", 1024, 200 * 1024, 1, 1, 1);
    run_alt<6000>("two functions, same smem", 200 * 1024, 200 * 1024, 0);
    run_alt<6000>("two functions, 200/140 KB", 200 * 1024, 140 * 1024, 0);
    run_alt<6000>("two functions, 200/84 KB", 200 * 1024, 84 * 1024, 0);
    run_alt<6000>("two functions, stores", 200 * 1024, 200 * 1024, 1);
    run_alt<6000>("two functions, 200/140, stores", 200 * 1024, 140 * 1024, 1);
    return 0;
}
