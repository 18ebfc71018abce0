// conv_tc.cuh -- tcgen05 / TMEM implicit-GEMM quantised convolution for sm_100a.
//
// GEMM view of one conv: D[M = 128 output pixels, N = cout] (int32, TMEM) += A[M, K] * B[N, K]^T with
// K = 16-channel chunks x taps (the plan's K-chunk list, which also encodes concat and residual adds).
//   A  (activations)  gathered from the 16-channel plane buffers by 128 producer threads with 16-byte
//      cp.async (zero fill = conv padding): one pixel's 16 channels = one 16-byte row of a K-major core matrix.
//   B  (weights)      packed by plan.py as [K-chunk][cout][16] int8, which IS the canonical no-swizzle K-major
//      layout (8 rows x 16 bytes per core matrix), so a stage is ONE bulk-TMA copy (cp.async.bulk -> UBLKCP).
//   D  accumulators in TMEM, read back with tcgen05.ld 32x32b.x16: thread = output pixel, 16 registers =
//      16 consecutive output channels = exactly one 16-byte plane row after the fixed-point epilogue.
// Warp roles: warps 0-3 producers then epilogue, warp 4 = MMA issuer (one thread) + TMEM allocator,
// warp 5 = weight loader (one thread).  smem ring of NS stages, each KS K-chunks (KS*16 of K) deep.
// One tile per CTA; several CTAs per SM overlap one tile's epilogue with the next tile's loads and MMAs.
#pragma once
#include "kernels.cuh"

namespace ayq {

struct TcState { int ready = 0; int num_sms = 148; };

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // try_wait suspends the thread for a bounded time in hardware; the iteration cap turns a protocol bug into a
    // trap (reported as a launch failure) instead of hanging the GPU.
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22); ++it) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no swizzle (INTERLEAVE): rows of 16 bytes, 8-row core matrices; SBO = stride between 8-row groups,
// LBO = stride between the two 16-byte K halves of one K=32 MMA.  (cute::UMMA::SmemDescriptor, version 1.)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;                 // descriptor version for sm_100
    return d;                               // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// kind::i8, A/B signed 8-bit K-major, D int32, M = 128  (cute::UMMA::InstrDescriptor)
__host__ __device__ __forceinline__ uint32_t make_idesc_i8(int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcParams {
    int KS;            // K chunks per stage (even)
    int NS;            // stages in the ring
    int nst;           // number of stages to run = ceil(nkc_pad / KS)
    int nkc_pad;       // nkc rounded up to even
    int tmem_cols;     // power of two >= max(32, cout)
};

constexpr int TC_THREADS = 192;
constexpr int TC_MAX_NS = 4;
constexpr int TC_LAG = 2;      // producer stages in flight before the oldest is published (< NS)

// dynamic smem: [A ring NS*KS*2048][B ring NS*KS*cout*16][kc table nkc*24][lut 2M+1 floats]
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const ConvArgs a, const TcParams tp) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2 * TC_MAX_NS + 1];   // full[NS], empty[NS], tmem_full
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int N = a.cout, KS = tp.KS, NS = tp.NS, nst = tp.nst;
    const uint32_t a_stage_bytes = (uint32_t)KS * 2048u, b_stage_bytes = (uint32_t)KS * N * 16u;
    unsigned char* sA = smem;
    unsigned char* sB = smem + (size_t)NS * a_stage_bytes;
    KChunk* skc = (KChunk*)(sB + (size_t)NS * b_stage_bytes);
    float* lut_s = (float*)(skc + a.nkc);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[TC_MAX_NS]), tfull = smem_u32(&bars[2 * TC_MAX_NS]);

    for (int i = tid; i < a.nkc; i += TC_THREADS) skc[i] = a.kc[i];
    if (a.epi == 0)
        for (int i = tid; i < 2 * a.M + 1; i += TC_THREADS) lut_s[i] = a.lut[i];
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(full0 + 8 * s, 128 + 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)tp.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 4) {
        // ===== producers: im2col gather of this thread's output pixel, then epilogue of the same pixel =====
        const size_t npix = (size_t)a.n * a.Hout * a.Wout;
        const size_t p = (size_t)blockIdx.x * 128 + tid;
        const bool valid = p < npix;
        const int ox = valid ? (int)(p % a.Wout) : 0;
        const int oy = valid ? (int)((p / a.Wout) % a.Hout) : 0;
        const int img = valid ? (int)(p / ((size_t)a.Wout * a.Hout)) : 0;
        const int iy0 = oy * a.stride, ix0 = ox * a.stride;
        const int8_t* img_base = a.ws + (size_t)img * a.Hin * a.Win * 16;
        for (int st = 0; st < nst; ++st) {
            const int slot = st % NS;
            if (st >= NS) mbar_wait(empty0 + 8 * slot, ((st / NS) - 1) & 1);
            const uint32_t dst0 = smem_u32(sA) + slot * a_stage_bytes + tid * 16;
            const int kc0 = st * KS;
            for (int c = 0; c < KS; ++c) {
                const int kc = kc0 + c;
                if (kc >= a.nkc) break;                       // odd tail chunk: its weights are zero
                const KChunk k = skc[kc];
                const int iy = iy0 + k.dy, ix = ix0 + k.dx;
                const bool ok = valid && (unsigned)iy < (unsigned)a.Hin && (unsigned)ix < (unsigned)a.Win;
                const int8_t* src = img_base + k.off + (size_t)k.plane * a.in_plane_bytes + ((size_t)(ok ? iy : 0) * a.Win + (ok ? ix : 0)) * 16;
                cp_async16(dst0 + c * 2048, src, ok ? 16u : 0u);
            }
            cp_async_commit();
            if (st >= TC_LAG) {
                cp_async_wait<TC_LAG>();
                fence_proxy_async();
                mbar_arrive(full0 + 8 * ((st - TC_LAG) % NS));
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        for (int st = (nst > TC_LAG ? nst - TC_LAG : 0); st < nst; ++st) mbar_arrive(full0 + 8 * (st % NS));

        // ===== epilogue: TMEM -> registers -> fixed-point SiLU / requant -> 16-byte plane rows =====
        mbar_wait(tfull, 0);
        tc_fence_after();
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int g = 0; g < N / 16; ++g) {
            int acc[16];
            tmem_ld16(lane_base + (uint32_t)(g * 16), acc);
            if (valid) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += __ldg(a.bias + g * 16 + j);
                epilogue16(a, acc, g * 16, img, oy, ox, lut_s);
            }
        }
        tc_fence_before();
    } else if (warp == 4) {
        // ===== MMA issuer =====
        if ((tid & 31) == 0) {
            const uint32_t idesc = make_idesc_i8(N);
            uint32_t accum = 0;
            for (int st = 0; st < nst; ++st) {
                const int slot = st % NS;
                mbar_wait(full0 + 8 * slot, (st / NS) & 1);
                tc_fence_after();
                const uint32_t abase = smem_u32(sA) + slot * a_stage_bytes, bbase = smem_u32(sB) + slot * b_stage_bytes;
                const int pairs = min(KS, tp.nkc_pad - st * KS) / 2;
                for (int j = 0; j < pairs; ++j) {
                    const uint64_t ad = make_desc(abase + j * 4096, 2048, 128);
                    const uint64_t bd = make_desc(bbase + j * 2 * N * 16, N * 16, 128);
                    mma_i8(tmem_base, ad, bd, idesc, accum);
                    accum = 1;
                }
                mma_commit(empty0 + 8 * slot);            // frees the smem slot when these MMAs retire
            }
            mma_commit(tfull);                             // accumulator complete -> epilogue
        }
        __syncwarp();
    } else {
        // ===== weight loader: one bulk-TMA copy per stage =====
        if ((tid & 31) == 0) {
            for (int st = 0; st < nst; ++st) {
                const int slot = st % NS;
                if (st >= NS) mbar_wait(empty0 + 8 * slot, ((st / NS) - 1) & 1);
                const int chunks = min(KS, tp.nkc_pad - st * KS);
                const uint32_t bytes = (uint32_t)chunks * N * 16u;
                mbar_arrive_expect_tx(full0 + 8 * slot, bytes);
                bulk_g2s(smem_u32(sB) + slot * b_stage_bytes, a.w + (size_t)st * KS * N * 16, bytes, full0 + 8 * slot);
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tp.tmem_cols) : "memory");
    }
}

}  // namespace tc

static inline void tc_init(TcState& s) {
    cudaFuncSetAttribute(tc::conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&s.num_sms, cudaDevAttrMultiProcessorCount, dev);
    s.ready = 1;
}
static inline void tc_release(TcState&) {}

// returns 0 = launched, 1 = shape not covered (caller uses the CUDA-core kernel), <0 = error
static inline int tc_launch_conv(TcState& s, const ConvArgs& a, const int32_t* /*op fields*/, cudaStream_t st) {
    if (!s.ready) return 1;
    const int N = a.cout;
    if (N % 16 != 0 || N < 16 || N > 256) return 1;
    tc::TcParams tp;
    tp.nkc_pad = (a.nkc + 1) & ~1;
    tp.KS = tp.nkc_pad < 8 ? tp.nkc_pad : 8;
    tp.nst = (tp.nkc_pad + tp.KS - 1) / tp.KS;
    tp.NS = tp.nst < 3 ? tp.nst : 3;
    if (tp.NS <= tc::TC_LAG && tp.nst > tp.NS) tp.NS = tc::TC_LAG + 1;
    int cols = 32;
    while (cols < N) cols <<= 1;
    tp.tmem_cols = cols;
    const size_t lut_bytes = a.epi == 0 ? (size_t)(2 * a.M + 1) * 4 : 0;
    const size_t smem = (size_t)tp.NS * tp.KS * 2048 + (size_t)tp.NS * tp.KS * N * 16 + (size_t)a.nkc * sizeof(KChunk) + lut_bytes + 16;
    if (smem > 220 * 1024) return 1;
    const size_t npix = (size_t)a.n * a.Hout * a.Wout;
    const unsigned grid = (unsigned)((npix + 127) / 128);
    tc::conv_tc_kernel<<<grid, tc::TC_THREADS, smem, st>>>(a, tp);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace ayq
