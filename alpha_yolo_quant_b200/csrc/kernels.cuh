// kernels.cuh -- CUDA-core kernels of the integer YOLOv8n engine (sm_100a).
//   conv_dp4a_kernel   generic quantised conv over 16-channel plane segments, dp4a, fused epilogue
//   conv_p1_kernel     Conv_P1 on the fp32 image with the fused per-image input quantiser
//   absmax_kernel      per-image max|x| (quant_matrix / save_max_a)
//   sppf_pool_kernel   three cascaded MaxPool2d(5,1,2)
//   head_kernel        DFL decode + 16-bit class scores (max / first argmax)
//   nms_kernel         coord_quant + nms_quant + clip_boxes, one CTA per image
// The tcgen05 convolution lives in conv_tc.cuh; both share ConvArgs and the epilogue below.
#pragma once
#include "fixedpoint.cuh"
#include <utility>

namespace ayq {

// Host launch helper: every kernel of a pass goes through here so that it can carry the programmatic-stream-serialization
// attribute (PDL, see fixedpoint.cuh); g_pdl is switched off by AYQ_NO_PDL=1.
static int g_pdl = 1;
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

struct KChunk { long long off; int plane, dy, dx, pad_; };   // workspace byte offset of the buffer, plane, tap offset minus padding
struct OutSpec { void* base; int mode; float k, inv; int up; };

struct ConvArgs {
    const KChunk* kc; int nkc;
    const int8_t* ws;           // activation workspace base
    size_t in_plane_bytes;      // n * Hin * Win * 16
    const int8_t* w;            // [nkc_pad][cout][16]
    const int* bias; const float* tab; const float* lut;
    const float* lut_rep;       // the sigmoid table replicated per lane ([257][32], fixedpoint.cuh), built once per engine: one bulk copy in the prologue
    int n, Hin, Win, Hout, Wout, stride, cout, epi, M;
    int nout; OutSpec out[3];
    int* acc_tap;               // NCHW int32 (n, cout, Hout, Wout) or nullptr
    float half;                 // 0.5f, kept in a register by the fast epilogues (see silu_q127)
    long long* dbg;             // AYQ_ROLE_PROF=1: per-CTA cycle counters of the warp roles [grid][16], else nullptr
    int gen_outs;               // MAGIC epilogue only: 1 = general output list (requantised copies through 256-byte tables, upsample)
    int dbg_mode;               // profiling build only (AYQ_EPI_SKIP=1): 1 = skip the fixed-point arithmetic, store the low accumulator bytes
};

// byte offset of the 16-byte row (channels [c0, c0+16) of output pixel (img, oy, ox)) in a phase-split buffer
// [(y&1)*2 + (x&1)][plane][n][Hout/2][Wout/2][16]
__device__ __forceinline__ size_t ps_offset(const ConvArgs& a, int c0, int img, int oy, int ox) {
    const int H2 = a.Hout >> 1, W2 = a.Wout >> 1;
    const size_t plane = (size_t)(((oy & 1) << 1) | (ox & 1)) * (a.cout >> 4) + (c0 >> 4);
    return (((plane * a.n + img) * H2 + (oy >> 1)) * W2 + (ox >> 1)) * 16;
}

// ---- shared epilogue: 16 consecutive output channels [c0, c0+16) of one output pixel -----------------
// acc[] already holds the bias.  pix = (img*Hout + oy)*Wout + ox.
__device__ __forceinline__ void epilogue16(const ConvArgs& a, const int* acc, int c0, int img, int oy, int ox,
                                           const float* __restrict__ lut_s) {
    const int M = a.M;
    const int cout = a.cout;
    const size_t npix = (size_t)a.n * a.Hout * a.Wout;
    const size_t pix = ((size_t)img * a.Hout + oy) * a.Wout + ox;
    if (a.acc_tap) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            a.acc_tap[(((size_t)img * cout + c0 + j) * a.Hout + oy) * a.Wout + ox] = acc[j];
    }
    const float* tab = a.tab;
    if (a.epi == 0) {                 // EPI_SILU
        int r[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int c = c0 + j;
            r[j] = silu_q(acc[j], __ldg(tab + c), __ldg(tab + cout + c), __ldg(tab + 2 * cout + c), __ldg(tab + 3 * cout + c), lut_s, M);
        }
        for (int o = 0; o < a.nout; ++o) {
            const OutSpec& os = a.out[o];
            uint32_t wd[4];
            if (os.mode == 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    wd[j] = pack4(requant8((float)r[4 * j], os.k, os.inv, M), requant8((float)r[4 * j + 1], os.k, os.inv, M),
                                  requant8((float)r[4 * j + 2], os.k, os.inv, M), requant8((float)r[4 * j + 3], os.k, os.inv, M));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) wd[j] = pack4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            }
            uint4 v = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            int8_t* base = (int8_t*)os.base;
            if (!os.up) {
                *(uint4*)(base + ((size_t)(c0 >> 4) * npix + pix) * 16) = v;
            } else if (os.up == 2) {   // phase-split copy for a stride-2 consumer
                *(uint4*)(base + ps_offset(a, c0, img, oy, ox)) = v;
            } else {             // nn.Upsample(None, 2, 'nearest') then requantize (:900-903): 2x2 replicate
                const int H2 = a.Hout * 2, W2 = a.Wout * 2;
                const size_t np2 = npix * 4;
                size_t p00 = ((size_t)img * H2 + 2 * oy) * W2 + 2 * ox;
                int8_t* pl = base + (size_t)(c0 >> 4) * np2 * 16;
                *(uint4*)(pl + p00 * 16) = v;
                *(uint4*)(pl + (p00 + 1) * 16) = v;
                *(uint4*)(pl + (p00 + W2) * 16) = v;
                *(uint4*)(pl + (p00 + W2 + 1) * 16) = v;
            }
        }
    } else if (a.epi == 1) {          // EPI_REQUANT8
        uint32_t wd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int q[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                int c = c0 + 4 * j + t;
                q[t] = requant8(__int2float_rn(acc[4 * j + t]), __ldg(tab + c), __ldg(tab + cout + c), M);
            }
            wd[j] = pack4(q[0], q[1], q[2], q[3]);
        }
        *(uint4*)((int8_t*)a.out[0].base + ((size_t)(c0 >> 4) * npix + pix) * 16) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    } else {                          // EPI_REQUANT16
        uint32_t wd[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int c = c0 + 2 * j;
            int q0 = requant16(__int2float_rn(acc[2 * j]), __ldg(tab + c), __ldg(tab + cout + c));
            int q1 = requant16(__int2float_rn(acc[2 * j + 1]), __ldg(tab + c + 1), __ldg(tab + cout + c + 1));
            wd[j] = (uint32_t)(q0 & 0xffff) | ((uint32_t)(q1 & 0xffff) << 16);
        }
        uint4* dst = (uint4*)((int16_t*)a.out[0].base + ((size_t)(c0 >> 4) * npix + pix) * 16);
        dst[0] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        dst[1] = make_uint4(wd[4], wd[5], wd[6], wd[7]);
    }
}

// ---- generic dp4a convolution (TEST BUILD ONLY: an independent CUDA-core implementation the parity tests cross-check against;
// the product library contains the TMA-fed tcgen05 convolution only) --------------------------------------
// grid (ceil(n*Hout*Wout / 128), cout / NC), block 128: one output pixel x NC output channels per thread.
#ifdef AYQ_TEST_BUILD
template <int NC>
__global__ void __launch_bounds__(128) conv_dp4a_kernel(const ConvArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* sW = (uint4*)smem_raw;                                 // [nkc][NC] 16-byte rows
    float* lut_s = (float*)(smem_raw + (size_t)a.nkc * NC * 16);  // [256]
    const int c0 = blockIdx.y * NC;
    pdl_trigger();
    for (int i = threadIdx.x; i < a.nkc * NC; i += 128) {
        int kc = i / NC, j = i % NC;
        sW[i] = *(const uint4*)(a.w + ((size_t)kc * a.cout + c0 + j) * 16);
    }
    if (a.epi == 0) fill_lut256(lut_s, a.lut, a.M, threadIdx.x, 128);
    __syncthreads();
    pdl_wait();
    const size_t npix = (size_t)a.n * a.Hout * a.Wout;
    const size_t p = (size_t)blockIdx.x * 128 + threadIdx.x;
    if (p >= npix) return;
    const int ox = (int)(p % a.Wout);
    const int oy = (int)((p / a.Wout) % a.Hout);
    const int img = (int)(p / ((size_t)a.Wout * a.Hout));
    int acc[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = __ldg(a.bias + c0 + j);
    const int iy0 = oy * a.stride, ix0 = ox * a.stride;
    for (int kc = 0; kc < a.nkc; ++kc) {
        const KChunk k = a.kc[kc];
        const int iy = iy0 + k.dy, ix = ix0 + k.dx;
        if ((unsigned)iy >= (unsigned)a.Hin || (unsigned)ix >= (unsigned)a.Win) continue;   // zero padding
        const int4 v = __ldg((const int4*)(a.ws + k.off + (size_t)k.plane * a.in_plane_bytes + (((size_t)img * a.Hin + iy) * a.Win + ix) * 16));
        const uint4* wr = sW + kc * NC;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const uint4 wv = wr[j];
            int s = acc[j];
            s = __dp4a(v.x, (int)wv.x, s);
            s = __dp4a(v.y, (int)wv.y, s);
            s = __dp4a(v.z, (int)wv.z, s);
            s = __dp4a(v.w, (int)wv.w, s);
            acc[j] = s;
        }
    }
#pragma unroll
    for (int g = 0; g < NC / 16; ++g) epilogue16(a, acc + 16 * g, c0 + 16 * g, img, oy, ox, lut_s);
}
#endif  // AYQ_TEST_BUILD

// ---- Conv_P1 + quant_matrix ---------------------------------------------------------------------------
struct P1Args {
    const float* img;           // (n,3,H,W) fp32 (U8 == false)
    const uint8_t* img_u8;      // (n,3,H,W) uint8 (U8 == true): ToTensor (u8 / 255, stage_8_torch.py:985-990) happens in the kernel
    const float* amax;          // (n) per-image max|x|
    float* amax_rw;             // same array, written by the fused abs-max step of conv_p1_tc_kernel<., true>
    unsigned* sync;             // fused kernel: {ticket, band counter per image}, zeroed by the host
    int fuse_d;                 // fused kernel: images between the abs-max step and the convolution step of a ticket (0 or 1)
    const float* lut;           // sigmoid table [2M+1]
    const float* lut_rep8;      // the same, eight copies per entry ([257][8]), built once per engine
    int n, H, W, Hout, Wout, M;
    int8_t* out;                // plane buffer (1 plane) (n,Hout,Wout,16)
    int* acc_tap;
    float half;                 // 0.5f in a register (see silu_q127f)
    int img0;                   // first image of this launch (grid.z counts from here); n stays the pass size (phase-split strides)
    int ps;                     // 1: phase-split output [(y&1)*2+(x&1)][n][Hout/2][Wout/2][16] (input layout of the stride-2 Conv_P2)
};
// weights and per-channel epilogue coefficients as a __grid_constant__ parameter: every use below has a compile-time
// index, so they become constant-bank operands of IDP.4A / FMUL (no weight or coefficient loads in the kernel).
struct alignas(16) P1Const { unsigned w4[9][16]; float k1[16], i1[16], k2[16], i2[16]; int bias[16]; };

// grid (Wout/32, Hout/8, n), block 256: one 32x8 output tile.  The fp32 input patch (3 x 17 x 65) is read row-wise with
// coalesced loads, quantised once (quant_matrix) and kept in smem as one packed word (c0,c1,c2,0) per pixel.
#define P1_TW 32
#define P1_TH 8
template <bool U8>
__global__ void __launch_bounds__(256) conv_p1_kernel(const __grid_constant__ P1Args a, const __grid_constant__ P1Const pc) {
    __shared__ unsigned sQ[2 * P1_TH + 1][2 * P1_TW + 2];       // +1 pad word per row
    __shared__ unsigned char qlut[256];                          // U8: quantised value of every byte, q = rint(fl32(fl32(v / 255) * s))
    __shared__ float lut_s[AYQ_LUT256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * P1_TW, y0 = blockIdx.y * P1_TH, img = blockIdx.z + a.img0;
    pdl_trigger();
    fill_lut256(lut_s, a.lut, a.M, tid, 256);
    pdl_wait();                                                  // amax[] comes from the abs-max kernel
    // quant_matrix: a = max|x|, s = scale(a, k) = M / a evaluated by torch as reciprocal(a) * M (Tensor.__rtruediv__),
    // q = rint(fl32(clip(x) * s))   (utils/quant_matrix_torch.py:57-70, utils/scale.py:4-5).  new_clip(x, a) with
    // a = max|x| of the same image is the identity, so no clamp is needed here.
    const float amax = a.amax[img];
    const float s = __fmul_rn(__frcp_rn(amax), (float)a.M);
    const size_t cs = (size_t)a.H * a.W;
    if (U8) {
        qlut[tid] = amax > 0.f ? (unsigned char)__float2int_rn(__fmul_rn(__fdiv_rn((float)tid, 255.f), s)) : 0;
        __syncthreads();
        const uint8_t* base = a.img_u8 + (size_t)img * 3 * cs;
        for (int i = tid; i < (2 * P1_TH + 1) * (2 * P1_TW + 1); i += 256) {      // flat over the 17 x 65 patch: balanced warps
            const int r = i / (2 * P1_TW + 1), c = i - r * (2 * P1_TW + 1);
            const int iy = 2 * y0 - 1 + r, ix = 2 * x0 - 1 + c;
            unsigned wd = 0;
            if ((unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W) {
                const uint8_t* px = base + (unsigned)iy * (unsigned)a.W + (unsigned)ix;
                wd = (unsigned)qlut[__ldg(px)] | ((unsigned)qlut[__ldg(px + cs)] << 8) | ((unsigned)qlut[__ldg(px + 2 * cs)] << 16);
            }
            sQ[r][c] = wd;
        }
    } else {
    const float* base = a.img + (size_t)img * 3 * cs;
    const bool any = amax > 0.f;
    for (int i = tid; i < (2 * P1_TH + 1) * (2 * P1_TW + 1); i += 256) {          // flat over the 17 x 65 patch: balanced warps
        const int r = i / (2 * P1_TW + 1), c = i - r * (2 * P1_TW + 1);
        const int iy = 2 * y0 - 1 + r, ix = 2 * x0 - 1 + c;
        unsigned wd = 0;
        if (any && (unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W) {
            const float* px = base + (unsigned)iy * (unsigned)a.W + (unsigned)ix;
            const int q0 = __float2int_rn(__fmul_rn(__ldg(px), s));
            const int q1 = __float2int_rn(__fmul_rn(__ldg(px + cs), s));
            const int q2 = __float2int_rn(__fmul_rn(__ldg(px + 2 * cs), s));
            wd = pack4(q0, q1, q2, 0);
        }
        sQ[r][c] = wd;
    }
    }
    __syncthreads();
    const int tx = tid & (P1_TW - 1), ty = tid / P1_TW;
    const int ox = x0 + tx, oy = y0 + ty;
    if (ox >= a.Wout || oy >= a.Hout) return;
    int acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = pc.bias[j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int v = (int)sQ[2 * ty + ky][2 * tx + kx];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = __dp4a(v, (int)pc.w4[ky * 3 + kx][j], acc[j]);
        }
    if (a.acc_tap) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a.acc_tap[(((size_t)img * 16 + j) * a.Hout + oy) * a.Wout + ox] = acc[j];
    }
    int r[16];
    if (a.M == 127) {                 // K = 8: folded coefficients (pc.k1 / pc.k2 hold k * 2^-s, see fixedpoint.cuh)
        const float half = a.half;
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = silu_q127f(acc[j], pc.k1[j], pc.k2[j], lut_s, half);
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = silu_q(acc[j], pc.k1[j], pc.i1[j], pc.k2[j], pc.i2[j], lut_s, a.M);
    }
    const size_t p = a.ps ? ((size_t)(((oy & 1) << 1) | (ox & 1)) * a.n + img) * (size_t)(a.Hout >> 1) * (a.Wout >> 1) + (size_t)(oy >> 1) * (a.Wout >> 1) + (ox >> 1)
                          : ((size_t)img * a.Hout + oy) * a.Wout + ox;
    *(uint4*)(a.out + p * 16) = make_uint4(pack4(r[0], r[1], r[2], r[3]), pack4(r[4], r[5], r[6], r[7]),
                                           pack4(r[8], r[9], r[10], r[11]), pack4(r[12], r[13], r[14], r[15]));
}

// Conv_P1, lean variant for K = 8 when the MAGIC epilogue is exact (host check: ayq.cu / magic_coeffs_ok):
//   * the patch is read with 16-byte loads (4 pixels x 3 channels per work item) and quantised without conversions:
//     rint(RN(x * s)) = low byte of the bits of RN(RN(x * s) + 1.5 * 2^23)  (|x * s| <= 127);
//   * even and odd input columns live in separate shared arrays, so the stride-2 taps of a warp are stride-1 words
//     (no bank conflicts);
//   * the accumulators start at bias + 0x4B400000 and go straight into silu_magic (no I2F, no lower clamp).
// Geometry: H = 2 Hout, W = 2 Wout, Wout % 32 == 0, Hout % 8 == 0 (checked by the host).  pc.i1 holds -k1p * C.
#ifdef AYQ_TEST_BUILD   // CUDA-core lean Conv_P1: cross-check implementation of conv_p1_tc_kernel, test build only
template <bool U8>
__global__ void __launch_bounds__(256) conv_p1_fast_kernel(const __grid_constant__ P1Args a, const __grid_constant__ P1Const pc) {
    __shared__ __align__(16) unsigned sE[2 * P1_TH + 1][P1_TW + 2];    // [r][j]     input x = 2 x0 + 2 j
    __shared__ __align__(16) unsigned sO[2 * P1_TH + 1][P1_TW + 2];    // [r][j + 2] input x = 2 x0 + 2 j + 1   (j = -1: left halo)
    __shared__ float2 lut2[AYQ_LUT256];
    __shared__ unsigned qlut[U8 ? 256 : 1];
    const int tid = threadIdx.x;
    const int y0 = blockIdx.y * P1_TH, img = blockIdx.z + a.img0;
    pdl_trigger();
    fill_lut256_magic(lut2, a.lut, a.M, tid, 256);
    pdl_wait();                                                  // amax[] comes from the abs-max kernel
    const float amax = a.amax[img];
    const float s = __fmul_rn(__frcp_rn(amax), (float)a.M);      // quant_matrix: scale(a, k) evaluated as reciprocal(a) * M
    const bool any = amax > 0.f;
    const size_t cs = (size_t)a.H * a.W;
    const int iy0 = 2 * y0 - 1;
    if (U8) qlut[tid] = any ? ((unsigned)__float2int_rn(__fmul_rn(__fdiv_rn((float)tid, 255.f), s)) & 0xffu) : 0u;
    constexpr int ROWS = 2 * P1_TH + 1, GROUPS = P1_TW / 2;     // 17 rows x 16 groups of 4 input pixels
    // one CTA walks the whole band of x tiles: the prologue (tables, scale) is paid once per band
    for (int x0 = 0; x0 < a.Wout; x0 += P1_TW) {
    const int ix0 = 2 * x0;
    __syncthreads();                                             // previous tile's taps are consumed (first pass: qlut / lut2 visible)
    for (int i = tid; i < ROWS * GROUPS + ROWS; i += 256) {
        if (i < ROWS * GROUPS) {
            const int r = i / GROUPS, g = i - r * GROUPS;
            const int iy = iy0 + r;
            unsigned w[4] = {0u, 0u, 0u, 0u};
            if ((unsigned)iy < (unsigned)a.H && any) {
                const size_t off = (size_t)img * 3 * cs + (size_t)iy * a.W + ix0 + 4 * g;
                if (U8) {
                    const unsigned c0 = __ldg((const unsigned*)(a.img_u8 + off)), c1 = __ldg((const unsigned*)(a.img_u8 + off + cs)),
                                   c2 = __ldg((const unsigned*)(a.img_u8 + off + 2 * cs));
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                        w[p] = qlut[(c0 >> (8 * p)) & 0xffu] | (qlut[(c1 >> (8 * p)) & 0xffu] << 8) | (qlut[(c2 >> (8 * p)) & 0xffu] << 16);
                } else {
                    const float4 v0 = __ldg((const float4*)(a.img + off)), v1 = __ldg((const float4*)(a.img + off + cs)),
                                 v2 = __ldg((const float4*)(a.img + off + 2 * cs));
                    const float f0[4] = {v0.x, v0.y, v0.z, v0.w}, f1[4] = {v1.x, v1.y, v1.z, v1.w}, f2[4] = {v2.x, v2.y, v2.z, v2.w};
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const unsigned z0 = __float_as_uint(__fadd_rn(__fmul_rn(f0[p], s), AYQ_MAGIC_F));
                        const unsigned z1 = __float_as_uint(__fadd_rn(__fmul_rn(f1[p], s), AYQ_MAGIC_F));
                        const unsigned z2 = __float_as_uint(__fadd_rn(__fmul_rn(f2[p], s), AYQ_MAGIC_F));
                        w[p] = __byte_perm(__byte_perm(z0, z1, 0x0040), z2, 0x0410);   // (q0, q1, q2, q0): byte 3 meets a zero weight
                    }
                }
            }
            *(uint2*)&sE[r][2 * g] = make_uint2(w[0], w[2]);
            *(uint2*)&sO[r][2 * g + 2] = make_uint2(w[1], w[3]);
        } else {                                                  // left halo column x = 2 x0 - 1
            const int r = i - ROWS * GROUPS;
            const int iy = iy0 + r, ix = ix0 - 1;
            unsigned wd = 0u;
            if ((unsigned)iy < (unsigned)a.H && ix >= 0 && any) {
                const size_t off = (size_t)img * 3 * cs + (size_t)iy * a.W + ix;
                if (U8) {
                    wd = qlut[__ldg(a.img_u8 + off)] | (qlut[__ldg(a.img_u8 + off + cs)] << 8) | (qlut[__ldg(a.img_u8 + off + 2 * cs)] << 16);
                } else {
                    const unsigned z0 = __float_as_uint(__fadd_rn(__fmul_rn(__ldg(a.img + off), s), AYQ_MAGIC_F));
                    const unsigned z1 = __float_as_uint(__fadd_rn(__fmul_rn(__ldg(a.img + off + cs), s), AYQ_MAGIC_F));
                    const unsigned z2 = __float_as_uint(__fadd_rn(__fmul_rn(__ldg(a.img + off + 2 * cs), s), AYQ_MAGIC_F));
                    wd = __byte_perm(__byte_perm(z0, z1, 0x0040), z2, 0x0410);
                }
            }
            sO[r][1] = wd;
        }
    }
    __syncthreads();
    const int tx = tid & (P1_TW - 1), ty = tid / P1_TW;
    const int ox = x0 + tx, oy = y0 + ty;
    int acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = pc.bias[j];             // bias + 0x4B400000 (host)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int vL = (int)sO[2 * ty + ky][tx + 1], vC = (int)sE[2 * ty + ky][tx], vR = (int)sO[2 * ty + ky][tx + 2];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            acc[j] = __dp4a(vL, (int)pc.w4[ky * 3][j], acc[j]);
            acc[j] = __dp4a(vC, (int)pc.w4[ky * 3 + 1][j], acc[j]);
            acc[j] = __dp4a(vR, (int)pc.w4[ky * 3 + 2][j], acc[j]);
        }
    }
    const float half = a.half;
    int r[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = silu_magic(acc[j], pc.k1[j], pc.i1[j], pc.k2[j], lut2, half);
    const uint32_t p = a.ps ? ((uint32_t)(((oy & 1) << 1) | (ox & 1)) * (uint32_t)a.n + (uint32_t)img) * (uint32_t)((a.Hout >> 1) * (a.Wout >> 1)) +
                                  (uint32_t)(oy >> 1) * (uint32_t)(a.Wout >> 1) + (uint32_t)(ox >> 1)
                            : ((uint32_t)img * (uint32_t)a.Hout + (uint32_t)oy) * (uint32_t)a.Wout + (uint32_t)ox;
    *(uint4*)(a.out + (size_t)p * 16) = make_uint4(pack4_sat(r[0], r[1], r[2], r[3]), pack4_sat(r[4], r[5], r[6], r[7]),
                                                   pack4_sat(r[8], r[9], r[10], r[11]), pack4_sat(r[12], r[13], r[14], r[15]));
    }
}
#endif  // AYQ_TEST_BUILD

// ---- per-image abs-max --------------------------------------------------------------------------------
// grid (blocks, n); out[] must be zeroed first.  |x| >= 0 so the float bit pattern orders like an int.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, float* __restrict__ out, size_t per_image) {
    pdl_trigger();
    pdl_wait();
    const float* base = x + (size_t)blockIdx.y * per_image;
    float m = 0.f;
    const size_t nvec = per_image / 4;
    const bool aligned = (((uintptr_t)base) & 15) == 0;
    if (aligned) {
        const float4* b4 = (const float4*)base;
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (size_t)gridDim.x * 256) {
            float4 v = __ldg(b4 + i);
            m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
        for (size_t i = nvec * 4 + (size_t)blockIdx.x * 256 + threadIdx.x; i < per_image; i += (size_t)gridDim.x * 256)
            m = fmaxf(m, fabsf(base[i]));
    } else {
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < per_image; i += (size_t)gridDim.x * 256)
            m = fmaxf(m, fabsf(base[i]));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmaxf(m, sm[i]);
        atomicMax((int*)out + blockIdx.y, __float_as_int(m));
    }
}

// uint8 images: max|u8 / 255| = fl32(max(u8) / 255) (the division is monotone).  grid (blocks, n); out[] zeroed first.
__global__ void __launch_bounds__(256) absmax_u8_kernel(const uint8_t* __restrict__ x, float* __restrict__ out, size_t per_image) {
    pdl_trigger();
    pdl_wait();
    const uint8_t* base = x + (size_t)blockIdx.y * per_image;
    unsigned m = 0;
    const size_t nvec = per_image / 16;
    if ((((uintptr_t)base) & 15) == 0) {
        const uint4* b4 = (const uint4*)base;
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (size_t)gridDim.x * 256) {
            const uint4 v = __ldg(b4 + i);
            m = __vmaxu4(m, __vmaxu4(__vmaxu4(v.x, v.y), __vmaxu4(v.z, v.w)));
        }
        for (size_t i = nvec * 16 + (size_t)blockIdx.x * 256 + threadIdx.x; i < per_image; i += (size_t)gridDim.x * 256) m = __vmaxu4(m, base[i]);
    } else {
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < per_image; i += (size_t)gridDim.x * 256) m = __vmaxu4(m, base[i]);
    }
    unsigned mb = max(max(m & 0xff, (m >> 8) & 0xff), max((m >> 16) & 0xff, m >> 24));
#pragma unroll
    for (int o = 16; o; o >>= 1) mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, o));
    __shared__ unsigned sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mb;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) mb = max(mb, sm[i]);
        atomicMax((int*)out + blockIdx.y, __float_as_int(__fdiv_rn((float)mb, 255.f)));
    }
}

// ---- SPPF: three cascaded 5x5 stride-1 max pools (padding = -inf, i.e. window clipped to the map) -----
// grid (nplanes, n), block 256.  in/out: plane buffers (plane, n, H, W, 16) int8.
__device__ __forceinline__ uint4 vmaxs4x4(uint4 a, uint4 b) { return make_uint4(__vmaxs4(a.x, b.x), __vmaxs4(a.y, b.y), __vmaxs4(a.z, b.z), __vmaxs4(a.w, b.w)); }
// block = H * W threads rounded up to a warp (<= 1024): thread = pixel (its x, y are computed once), 16 channels = one uint4.
__global__ void __launch_bounds__(1024) sppf_pool_kernel(const int8_t* __restrict__ in, int8_t* __restrict__ out,
                                                         int n, int H, int W, int nplanes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* A = (uint4*)smem_raw;                        // [H*W]
    uint4* B = A + H * W;
    const int pl = blockIdx.x, img = blockIdx.y;
    pdl_trigger();
    pdl_wait();
    const size_t plane_px = (size_t)n * H * W;
    const int px = threadIdx.x, npx = H * W;
    const bool act = px < npx;
    const int y = px / W, x = px - y * W;
    const uint4* src = (const uint4*)(in + ((size_t)pl * plane_px + (size_t)img * npx) * 16);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (act) { v = src[px]; A[px] = v; }
    __syncthreads();
    for (int stage = 0; stage < 3; ++stage) {
        if (act) {                                      // row pass: max over x-2..x+2 (window clipped to the map = -inf padding)
            uint4 m = v;
#pragma unroll
            for (int d = -2; d <= 2; ++d)
                if (d != 0 && x + d >= 0 && x + d < W) m = vmaxs4x4(m, A[px + d]);
            B[px] = m;
            v = m;
        }
        __syncthreads();
        if (act) {                                      // column pass
            uint4 m = v;
#pragma unroll
            for (int d = -2; d <= 2; ++d)
                if (d != 0 && y + d >= 0 && y + d < H) m = vmaxs4x4(m, B[px + d * W]);
            A[px] = m;
            v = m;
            ((uint4*)(out + ((size_t)(stage * nplanes + pl) * plane_px + (size_t)img * npx) * 16))[px] = m;
        }
        __syncthreads();
    }
}

// e / S correctly rounded from the correctly rounded reciprocal r = RN(1 / S): q0 = RN(e r), rem = e - S q0 (exact in
// one FMA), q = RN(q0 + rem r).  For the DFL operands (integers 0 <= e <= 127, 1 <= S <= 2032) this equals the IEEE
// division bit for bit; div_selfcheck_kernel verifies ALL operand pairs on the device when the engine is created and the
// head falls back to __fdiv_rn if a single one differed.
__device__ __forceinline__ float div_by_rcp(float e, float S, float r) {
    const float q0 = __fmul_rn(e, r);
    const float rem = __fmaf_rn(-S, q0, e);
    return __fmaf_rn(rem, r, q0);
}
__global__ void div_selfcheck_kernel(int emax, int smax, int* __restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (emax + 1) * smax) return;
    const float e = (float)(i / smax), S = (float)(i % smax + 1);
    if (__float_as_int(div_by_rcp(e, S, __frcp_rn(S))) != __float_as_int(__fdiv_rn(e, S))) atomicAdd(bad, 1);
}

// ---- Detect head: DFL decode + class scores -----------------------------------------------------------
struct HeadArgs {
    const int8_t* box[3];       // (4 planes, n, H, W, 16) int8, channel = side*16 + bin
    const int16_t* cls[3];      // (5 planes, n, H, W, 16) int16
    const float* lut_exp;       // [2^K], index y + 2^K - 1
    const int16_t* lut16;       // [65535], index l + 32767
    const int16_t* lo16;        // [65535]: smallest logit with the same table value
    int mono;                   // table is monotone: class max / argmax need two gathers instead of 80
    int fast_div;               // div_by_rcp verified against __fdiv_rn on every DFL operand pair (ayq_create)
    const int* dflw;            // [16]
    const int* anchors;         // [A][2]
    float kd, id;
    int n, K, A;
    float4* dbox;               // (n, A) cx cy w h   (the first four rows of dbox_cls)
    int* conf; int* cls_id;     // (n, A)
    float* dbox_cls;            // optional (n, 84, A)
};

__global__ void __launch_bounds__(128, 8) head_kernel(const HeadArgs a) {
    __shared__ float lexp[512];
    __shared__ int dflw[16];
    const int top = (1 << a.K) - 1;
    pdl_trigger();
    for (int i = threadIdx.x; i <= top; i += 128) lexp[i] = a.lut_exp[i];
    if (threadIdx.x < 16) dflw[threadIdx.x] = a.dflw[threadIdx.x];
    __syncthreads();
    pdl_wait();
    const int idx = blockIdx.x * 128 + threadIdx.x;
    if (idx >= a.n * a.A) return;
    const int img = idx / a.A, an = idx % a.A;
    int lvl, hw, local;
    if (an < 6400) { lvl = 0; hw = 80; local = an; }
    else if (an < 8000) { lvl = 1; hw = 40; local = an - 6400; }
    else { lvl = 2; hw = 20; local = an - 8000; }
    const float stride = (float)(8 << lvl);
    const size_t plane_px = (size_t)a.n * hw * hw;
    const size_t pix = (size_t)img * hw * hw + local;
    float dq[4];
#pragma unroll
    for (int side = 0; side < 4; ++side) {
        const int4 raw = __ldg((const int4*)(a.box[lvl] + ((size_t)side * plane_px + pix) * 16));
        int b[16];
        const int wv[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int j = 0; j < 16; ++j) b[j] = (int)(int8_t)((wv[j >> 2] >> (8 * (j & 3))) & 0xff);
        int mx = b[0];
#pragma unroll
        for (int j = 1; j < 16; ++j) mx = max(mx, b[j]);
        float e[16], S = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { e[j] = lexp[b[j] - mx + top]; S += e[j]; }        // exact: integers <= 16*M
        int d = 0;
        if (a.fast_div) {
            const float rS = __frcp_rn(S);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int pj = (int)__fmul_rn(div_by_rcp(e[j], S, rS), 127.f);            // (y / ax_sum * 127).to(int64)  :1205
                d += pj * dflw[j];                                                        // self.dfl(p)  :1232
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int pj = (int)__fmul_rn(__fdiv_rn(e[j], S), 127.f);
                d += pj * dflw[j];
            }
        }
        dq[side] = (float)requant((float)d, a.kd, a.id, 32767);                           // requantize(dfl, ..., 16)  :1236
    }
    const float ax = (float)a.anchors[2 * an], ay = (float)a.anchors[2 * an + 1];
    const float x1 = ax - dq[0], y1 = ay - dq[1], x2 = ax + dq[2], y2 = ay + dq[3];       // dist2bbox :117-126
    const float cx = __fmul_rn(__fdiv_rn(x1 + x2, 2.f), stride), cy = __fmul_rn(__fdiv_rn(y1 + y2, 2.f), stride);
    const float w = __fmul_rn(x2 - x1, stride), h = __fmul_rn(y2 - y1, stride);
    a.dbox[idx] = make_float4(cx, cy, w, h);
    float* full = a.dbox_cls ? a.dbox_cls + (size_t)img * 84 * a.A + an : nullptr;
    if (full) { full[0] = cx; full[(size_t)a.A] = cy; full[(size_t)2 * a.A] = w; full[(size_t)3 * a.A] = h; }
    int best = -1, bj = 0;
    if (a.mono && !full) {
        // conf = max_c LUT[l_c] = LUT[max_c l_c]; first arg-max = first class with l_c >= lo16[max logit]   (torch.max :326)
        int4 r[10];
#pragma unroll
        for (int pl = 0; pl < 5; ++pl) {
            const int4* src = (const int4*)(a.cls[lvl] + ((size_t)pl * plane_px + pix) * 16);
            r[2 * pl] = __ldg(src); r[2 * pl + 1] = __ldg(src + 1);
        }
        const int* wv = (const int*)r;
        int m2 = (int)0x80008000;                                  // packed (int16, int16) running maxima
#pragma unroll
        for (int q = 0; q < 40; ++q) m2 = __vmaxs2(m2, wv[q]);
        const int mx = max((int)(short)(m2 & 0xffff), m2 >> 16);
        best = (int)__ldg(a.lut16 + mx + 32767);
        const int lo = (int)__ldg(a.lo16 + mx + 32767);
        bj = 80;
#pragma unroll
        for (int q = 39; q >= 0; --q) {
            if ((wv[q] >> 16) >= lo) bj = 2 * q + 1;
            if ((int)(short)(wv[q] & 0xffff) >= lo) bj = 2 * q;
        }
    } else {
#pragma unroll
        for (int pl = 0; pl < 5; ++pl) {
            const int4* src = (const int4*)(a.cls[lvl] + ((size_t)pl * plane_px + pix) * 16);
            const int4 r0 = __ldg(src), r1 = __ldg(src + 1);
            const int wv[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int l = (int)(int16_t)((wv[j >> 1] >> (16 * (j & 1))) & 0xffff);
                const int sc = (int)__ldg(a.lut16 + l + 32767);                          // sigmoid_quant(cls, lookup_final) :1250
                if (full) full[(size_t)(4 + pl * 16 + j) * a.A] = (float)sc;
                if (sc > best) { best = sc; bj = pl * 16 + j; }                           // first maximum, like torch.max :326
            }
        }
    }
    a.conf[idx] = best;
    a.cls_id[idx] = bj;
}

// per-anchor max / first argmax from a caller-provided (n,84,A) fp32 prediction tensor (ayq_nms entry)
__global__ void __launch_bounds__(128) pred_to_cand_kernel(const float* __restrict__ pred, int n, int A,
                                                           float4* dbox, int* conf, int* cls_id) {
    const int idx = blockIdx.x * 128 + threadIdx.x;
    if (idx >= n * A) return;
    const int img = idx / A, an = idx % A;
    const float* p = pred + (size_t)img * 84 * A + an;
    dbox[idx] = make_float4(p[0], p[(size_t)A], p[(size_t)2 * A], p[(size_t)3 * A]);
    float best = -1.f; int bj = 0;
    for (int c = 0; c < 80; ++c) {
        const float s = p[(size_t)(4 + c) * A];
        if (s > best) { best = s; bj = c; }
    }
    conf[idx] = (int)best;
    cls_id[idx] = bj;
}

// ---- q_NMS: coord_quant (:297-361) + nms_quant (:248-294) + clip_boxes (:403-423) ---------------------
// One CTA (1024 threads) per image.  Ordering pinned to (score desc, candidate index asc) = a stable
// descending argsort (SURVEY.md hard part 3).
// mode 0: candidates come from the head (dbox xywh, conf, cls_id per anchor); output = detection rows.
// mode 1: nms_quant() stand-alone on caller boxes (nb,4) xyxy (class offsets already added) and
//         integer-valued scores; output = kept indices in selection order (all of them, <= 1000).
#define NMS_THREADS 1024
#define NMS_TOPK 1000
#define NMS_MAXDET 300
#define NMS_SORT_N 16384
struct NmsArgs {
    const float4* dbox; const int* conf; const int* cls_id;   // mode 0: (n, A)
    const float* boxes; const float* scores;                  // mode 1: (A,4), (A)
    int n, A, mode, max_keep;
    float* dets;       // mode 0: (n, 300, 6)      mode 1: kept indices as float (max_keep)
    int* counts;       // (n)
};
// dynamic smem: keys[16384] | bx[5][1000] | diag[1024] (histogram scratch before the sort) | rem[32] | sh[8]
static constexpr size_t NMS_SMEM = (size_t)NMS_SORT_N * 4 + (size_t)NMS_TOPK * 5 * 4 + 1024 * 4 + 32 * 4 + 64;

// box i suppresses box j (:270-283); row i comes from shared memory (warp-uniform broadcast), column j from registers
__device__ __forceinline__ bool nms_suppresses(const float* __restrict__ bx, int i, float jx1, float jy1, float jx2, float jy2, float jarea) {
    const float xx1 = fmaxf(bx[i], jx1), yy1 = fmaxf(bx[NMS_TOPK + i], jy1);
    const float xx2 = fminf(bx[2 * NMS_TOPK + i], jx2), yy2 = fminf(bx[3 * NMS_TOPK + i], jy2);
    const float w = fmaxf(0.f, (xx2 - xx1) + 412.f), h = fmaxf(0.f, (yy2 - yy1) + 412.f);
    const float inter = __fmul_rn(__fmul_rn(w, h), 2.22f);
    const float rhs = __fadd_rn(bx[4 * NMS_TOPK + i], jarea) - inter;
    return !(inter <= rhs);
}

__global__ void __launch_bounds__(NMS_THREADS) nms_kernel(const NmsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* keys = (unsigned*)smem_raw;                          // [16384]
    float* bx = (float*)(keys + NMS_SORT_N);                       // [5][1000]: x1 y1 x2 y2 area (class-offset boxes)
    unsigned* diag = (unsigned*)(bx + 5 * NMS_TOPK);               // [1024] intra-chunk suppression words (row i: bits j > i of its chunk)
    unsigned* rem = diag + 1024;                                   // [32]   removed-set, one word per 32-candidate chunk
    int* sh = (int*)(rem + 32);                                    // [0]=nsorted [1]=nkeep [3..5] selection scratch [6]=keep word
    __shared__ int kept[NMS_TOPK];
    const int img = blockIdx.x, tid = threadIdx.x, A = a.A;
    const int lane = tid & 31, wid = tid >> 5;
    const int* conf = a.mode == 0 ? a.conf + (size_t)img * A : nullptr;
    pdl_trigger();
    if (tid < 8) sh[tid] = 0;
    pdl_wait();
    // candidates: conf > 8192 (:299,:302,:327).  Only the 1000 best survive argsort(...)[:1000] (:260), so first find
    // the score c* of the 1000th best with a two-level histogram (score >> 9, score & 511) and sort only candidates
    // with score >= c* (all ties at c* are kept: the index inside the key decides among them, as a stable sort would).
    int* hist = (int*)diag;                                         // [256] + [512]
    for (int i = tid; i < 768; i += NMS_THREADS) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < A; i += NMS_THREADS) {
        const int c = a.mode == 0 ? conf[i] : (int)a.scores[i];
        if (a.mode != 0 || c > 8192) atomicAdd(&hist[c >> 9], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int total = 0;
        for (int b = 0; b < 256; ++b) total += hist[b];
        if (total > NMS_TOPK) {
            int above = 0, b1 = 255;
            for (; b1 > 0 && above + hist[b1] < NMS_TOPK; --b1) above += hist[b1];
            sh[3] = b1; sh[4] = above;
        } else {
            sh[3] = -1;
        }
        sh[5] = 0;
    }
    __syncthreads();
    if (sh[3] >= 0) {
        const int b1 = sh[3];
        for (int i = tid; i < A; i += NMS_THREADS) {
            const int c = a.mode == 0 ? conf[i] : (int)a.scores[i];
            if ((a.mode != 0 || c > 8192) && (c >> 9) == b1) atomicAdd(&hist[256 + (c & 511)], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int above = sh[4], b2 = 511;
            for (; b2 > 0 && above + hist[256 + b2] < NMS_TOPK; --b2) above += hist[256 + b2];
            sh[5] = (b1 << 9) | b2;
        }
        __syncthreads();
    }
    const int cstar = sh[5];
    for (int i0 = 0; i0 < A; i0 += NMS_THREADS) {                  // warp-uniform trip count
        const int i = i0 + tid;
        unsigned key = 0;
        bool cand = false;
        if (i < A) {
            if (a.mode == 0) {
                const int c = conf[i];
                cand = c > 8192 && c >= cstar;
                key = ((unsigned)(32767 - c) << 14) | (unsigned)i;
            } else {
                const int c = (int)a.scores[i];                    // host wrapper guarantees 0 <= c <= 131071, integer
                cand = c >= cstar;
                key = ((unsigned)(131071 - c) << 14) | (unsigned)i;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, cand);
        if (bal) {
            const int leader = __ffs(bal) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&sh[0], __popc(bal));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (cand) keys[base + __popc(bal & ((1u << lane) - 1))] = key;
        }
    }
    __syncthreads();
    const int ncand = sh[0];                                       // candidates that entered the sort (>= min(total, 1000))
    if (ncand == 0) {                                              // reference: coord_quant returns None -> (None, None)
        if (tid == 0) a.counts[img] = 0;
        return;
    }
    // bitonic sort, ascending, over the smallest power of two that holds all candidates (padding keys sort last)
    int sort_n = 64;
    while (sort_n < ncand) sort_n <<= 1;
    for (int i = ncand + tid; i < sort_n; i += NMS_THREADS) keys[i] = 0xffffffffu;
    __syncthreads();
    for (int k = 2; k <= sort_n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < sort_n / 2; t += NMS_THREADS) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int hi = lo | j;
                const unsigned x = keys[lo], y = keys[hi];
                const bool up = (lo & k) == 0;
                if ((x > y) == up) { keys[lo] = y; keys[hi] = x; }
            }
            __syncthreads();
        }
    }
    const int T = min(ncand, NMS_TOPK);                            // argsort(descending)[:1000]  :260
    const float4* dbox = a.mode == 0 ? a.dbox + (size_t)img * A : nullptr;
    const int* cls_id = a.mode == 0 ? a.cls_id + (size_t)img * A : nullptr;
    for (int i = tid; i < T; i += NMS_THREADS) {
        const int an = (int)(keys[i] & 0x3fffu);
        float x1, y1, x2, y2;
        if (a.mode == 0) {
            const float4 d = dbox[an];
            const float dw = __fdiv_rn(d.z, 2.f), dh = __fdiv_rn(d.w, 2.f);      // xywh2xyxy :129-148, .to(int) :316
            const float off = __fmul_rn((float)cls_id[an], 7680.f);                // :340
            x1 = truncf(d.x - dw) + off; y1 = truncf(d.y - dh) + off;              // boxes = x[:, :4] + c  :344
            x2 = truncf(d.x + dw) + off; y2 = truncf(d.y + dh) + off;
        } else {
            x1 = a.boxes[4 * an]; y1 = a.boxes[4 * an + 1]; x2 = a.boxes[4 * an + 2]; y2 = a.boxes[4 * an + 3];
        }
        bx[i] = x1; bx[NMS_TOPK + i] = y1; bx[2 * NMS_TOPK + i] = x2; bx[3 * NMS_TOPK + i] = y2;
        bx[4 * NMS_TOPK + i] = __fmul_rn((x2 - x1) + 412.f, (y2 - y1) + 412.f);    // areas :258
    }
    if (tid < 32) rem[tid] = 0;
    __syncthreads();
    // Greedy suppression (:262-292) in chunks of 32 sorted candidates.  Only rows that are KEPT ever suppress anything,
    // so instead of the full T x T matrix: (1) the 32 x 32 diagonal blocks (all warps, in parallel), then per chunk
    // (2) one warp resolves the chunk serially from the removed-set word and the diagonal words, (3) every later chunk's
    // warp applies the rows kept in this chunk to its own 32 columns.  Work = T*32 + kept*T pair tests, kept <= max_keep.
    const int nchunk = (T + 31) >> 5;                              // <= 32 = number of warps
    if (wid < nchunk) {
        const int j = wid * 32 + lane;
        const bool jv = j < T;
        const int jj = jv ? j : 0;
        const float jx1 = bx[jj], jy1 = bx[NMS_TOPK + jj], jx2 = bx[2 * NMS_TOPK + jj], jy2 = bx[3 * NMS_TOPK + jj], ja = bx[4 * NMS_TOPK + jj];
        unsigned mine = 0;
        for (int ii = 0; ii < 32; ++ii) {
            const int i = wid * 32 + ii;
            const bool s = jv && i < T && j > i && nms_suppresses(bx, i, jx1, jy1, jx2, jy2, ja);
            const unsigned bits = __ballot_sync(0xffffffffu, s);
            if (lane == ii) mine = bits;
        }
        diag[wid * 32 + lane] = mine;
    }
    __syncthreads();
    int nk = 0;                                                    // block-uniform
    for (int c = 0; c < nchunk && nk < a.max_keep; ++c) {
        if (wid == 0) {
            unsigned cur = rem[c];
            const int nrow = min(32, T - c * 32);
            if (nrow < 32) cur |= 0xffffffffu << nrow;            // rows past T do not exist
            const unsigned d = diag[c * 32 + lane];
            unsigned keep = 0;
            int k = nk;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const unsigned di = __shfl_sync(0xffffffffu, d, i);
                if (!((cur >> i) & 1u) && k < a.max_keep) { keep |= 1u << i; cur |= di; ++k; }
            }
            if ((keep >> lane) & 1u) kept[nk + __popc(keep & ((1u << lane) - 1))] = c * 32 + lane;
            if (lane == 0) sh[6] = (int)keep;
        }
        __syncthreads();
        const unsigned keep = (unsigned)sh[6];
        nk += __popc(keep);
        if (wid > c && wid < nchunk && keep && nk < a.max_keep) {
            const int j = wid * 32 + lane;
            const bool jv = j < T;
            const int jj = jv ? j : 0;
            const float jx1 = bx[jj], jy1 = bx[NMS_TOPK + jj], jx2 = bx[2 * NMS_TOPK + jj], jy2 = bx[3 * NMS_TOPK + jj], ja = bx[4 * NMS_TOPK + jj];
            unsigned acc = 0;
            for (unsigned m = keep; m; m &= m - 1) {
                const int i = c * 32 + __ffs(m) - 1;
                acc |= __ballot_sync(0xffffffffu, jv && nms_suppresses(bx, i, jx1, jy1, jx2, jy2, ja));
            }
            if (lane == 0) rem[wid] |= acc;
        }
        __syncthreads();
    }
    if (tid == 0) a.counts[img] = nk;
    for (int r = tid; r < nk; r += NMS_THREADS) {
        const int i = kept[r];
        const int an = (int)(keys[i] & 0x3fffu);
        if (a.mode != 0) { a.dets[r] = (float)an; continue; }
        float* row = a.dets + ((size_t)img * NMS_MAXDET + r) * 6;
        // output rows hold the un-offset boxes x[:, :4] (:351-359), then clip to [0, 640] (:403-423)
        const float4 d = dbox[an];
        const float dw = __fdiv_rn(d.z, 2.f), dh = __fdiv_rn(d.w, 2.f);
        const float c[4] = {truncf(d.x - dw), truncf(d.y - dh), truncf(d.x + dw), truncf(d.y + dh)};
#pragma unroll
        for (int q = 0; q < 4; ++q) row[q] = fminf(fmaxf(__fdiv_rn(c[q], 412.1635f), 0.f), 640.f);
        row[4] = __fdiv_rn((float)conf[an], 32767.0f);
        row[5] = (float)cls_id[an];
    }
}

// ---- FLOAT Detect head of stage_8_torch.py ---------------------------------------
// The six raw head accumulators (NCHW int32, written by the convs' accumulator outputs) are dequantised and decoded in
// fp32: x / scale (:915-922), softmax over the 16 DFL bins + dfl conv (:930-933), dist2bbox * strides (:936), class
// sigmoid (:939-940).  Not bit-exact by construction (the reference runs torch's CPU softmax / sigmoid); the parity
// tests state the tolerance.  One thread per (image, anchor); every load is coalesced over consecutive anchors.
struct HeadFloatArgs {
    const int* box[3];          // (n, 64, H, W) int32 accumulators, channel = side*16 + bin
    const int* cls[3];          // (n, 80, H, W) int32
    const float* box_scale;     // [3][64]
    const float* cls_scale;     // [3][80]
    const float* dflw;          // [16] dfl.weight
    int n, A;
    float4* dbox;               // (n, A) cx cy w h
    float* conf; int* cls_id;   // (n, A) max class probability / first arg-max
    float* dbox_cls;            // optional (n, 84, A)
};

__global__ void __launch_bounds__(128) head_float_kernel(const HeadFloatArgs a) {
    __shared__ float bs[3 * 64], cs[3 * 80], dw[16];
    pdl_trigger();
    for (int i = threadIdx.x; i < 3 * 64; i += 128) bs[i] = a.box_scale[i];
    for (int i = threadIdx.x; i < 3 * 80; i += 128) cs[i] = a.cls_scale[i];
    if (threadIdx.x < 16) dw[threadIdx.x] = a.dflw[threadIdx.x];
    __syncthreads();
    pdl_wait();
    const int idx = blockIdx.x * 128 + threadIdx.x;
    if (idx >= a.n * a.A) return;
    const int img = idx / a.A, an = idx % a.A;
    int lvl, hw, local;
    if (an < 6400) { lvl = 0; hw = 80; local = an; }
    else if (an < 8000) { lvl = 1; hw = 40; local = an - 6400; }
    else { lvl = 2; hw = 20; local = an - 8000; }
    const float stride = (float)(8 << lvl);
    const size_t hw2 = (size_t)hw * hw;
    const int* bp = a.box[lvl] + (size_t)img * 64 * hw2 + local;
    float d[4];
#pragma unroll
    for (int side = 0; side < 4; ++side) {
        float x[16], mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            x[j] = __fdiv_rn((float)__ldg(bp + (size_t)(side * 16 + j) * hw2), bs[lvl * 64 + side * 16 + j]);
            mx = fmaxf(mx, x[j]);
        }
        float S = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { x[j] = expf(x[j] - mx); S += x[j]; }
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) acc = fmaf(__fdiv_rn(x[j], S), dw[j], acc);
        d[side] = acc;
    }
    const float ax = (float)(local % hw) + 0.5f, ay = (float)(local / hw) + 0.5f;          // make_anchors :97-109
    const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];            // dist2bbox :112-121
    const float cx = __fmul_rn(__fdiv_rn(x1 + x2, 2.f), stride), cy = __fmul_rn(__fdiv_rn(y1 + y2, 2.f), stride);
    const float w = __fmul_rn(x2 - x1, stride), h = __fmul_rn(y2 - y1, stride);
    a.dbox[idx] = make_float4(cx, cy, w, h);
    float* full = a.dbox_cls ? a.dbox_cls + (size_t)img * 84 * a.A + an : nullptr;
    if (full) { full[0] = cx; full[(size_t)a.A] = cy; full[(size_t)2 * a.A] = w; full[(size_t)3 * a.A] = h; }
    const int* cp = a.cls[lvl] + (size_t)img * 80 * hw2 + local;
    float best = -1.f; int bj = 0;
#pragma unroll 8
    for (int c = 0; c < 80; ++c) {
        const float z = __fdiv_rn((float)__ldg(cp + (size_t)c * hw2), cs[lvl * 80 + c]);
        const float p = __fdiv_rn(1.f, 1.f + expf(-z));                                     // sigmoid :940
        if (full) full[(size_t)(4 + c) * a.A] = p;
        if (p > best) { best = p; bj = c; }                                                 // first maximum, cls.max(1) :170
    }
    a.conf[idx] = best;
    a.cls_id[idx] = bj;
}

// per-anchor max / first argmax from a caller-provided (n,84,A) fp32 prediction tensor (ayq_coord_float entry)
__global__ void __launch_bounds__(128) pred_to_cand_float_kernel(const float* __restrict__ pred, int n, int A,
                                                                 float4* dbox, float* conf, int* cls_id) {
    const int idx = blockIdx.x * 128 + threadIdx.x;
    if (idx >= n * A) return;
    const int img = idx / A, an = idx % A;
    const float* p = pred + (size_t)img * 84 * A + an;
    dbox[idx] = make_float4(p[0], p[(size_t)A], p[(size_t)2 * A], p[(size_t)3 * A]);
    float best = -INFINITY; int bj = 0;
    for (int c = 0; c < 80; ++c) {
        const float s = p[(size_t)(4 + c) * A];
        if (s > best) { best = s; bj = c; }
    }
    conf[idx] = best;
    cls_id[idx] = bj;
}

// coord() of stage_8_torch.py:146-190 (+ clip_boxes :240-252): candidates conf > 1e-8 (in practice all A anchors), boxes
// offset by class * 7680, torchvision.ops.nms(boxes, scores, 0.45) = greedy over the stable descending score order,
// j suppressed by a kept i when inter / (area_i + area_j - inter) > 0.45 with every operation rounded to fp32 in that
// order, first 300 kept.  One CTA (1024 threads) per image; A <= 16384.
// dynamic smem: region0 = max(8 * sort_n, 16 * A) bytes: 64-bit sort keys, then x1 y1 x2 y2 [A] | order u16[A] | diag
// u32[32 * nchunk] | rem u32[nchunk] | kept int[300] | sh int[8]
#define NMSF_THR 0.45f
#define NMSF_CONF 0.00000001f
static inline size_t nmsf_smem_bytes(int A) {
    size_t sort_n = 64;
    while (sort_n < (size_t)A) sort_n <<= 1;
    const size_t r0 = 8 * sort_n > 16 * (size_t)A ? 8 * sort_n : 16 * (size_t)A;
    const size_t nchunk = ((size_t)A + 31) / 32;
    return r0 + (((size_t)A * 2 + 15) & ~(size_t)15) + nchunk * 32 * 4 + ((nchunk * 4 + 15) & ~(size_t)15) + NMS_MAXDET * 4 + 64;
}
struct NmsFloatArgs {
    const float4* dbox; const float* conf; const int* cls_id;   // (n, A)
    int n, A, max_keep;
    float* dets;       // (n, 300, 6)
    int* counts;       // (n)
};

__device__ __forceinline__ bool nmsf_suppresses(const float* __restrict__ bx, int A, int i, float jx1, float jy1, float jx2, float jy2, float jarea) {
    const float ix1 = bx[i], iy1 = bx[A + i], ix2 = bx[2 * A + i], iy2 = bx[3 * A + i];
    const float iarea = __fmul_rn(ix2 - ix1, iy2 - iy1);
    const float w = fmaxf(0.f, fminf(ix2, jx2) - fmaxf(ix1, jx1)), h = fmaxf(0.f, fminf(iy2, jy2) - fmaxf(iy1, jy1));
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, jarea), inter));
    return ovr > NMSF_THR;
}

__global__ void __launch_bounds__(NMS_THREADS) nms_float_kernel(const NmsFloatArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int img = blockIdx.x, tid = threadIdx.x, A = a.A;
    const int lane = tid & 31, wid = tid >> 5;
    int sort_cap = 64;
    while (sort_cap < A) sort_cap <<= 1;
    const size_t r0 = 8 * (size_t)sort_cap > 16 * (size_t)A ? 8 * (size_t)sort_cap : 16 * (size_t)A;
    const int nchunk_max = (A + 31) >> 5;
    unsigned long long* keys = (unsigned long long*)smem_raw;
    float* bx = (float*)smem_raw;                                   // reuses the key region after the sort
    unsigned short* order = (unsigned short*)(smem_raw + r0);
    unsigned* diag = (unsigned*)(smem_raw + r0 + (((size_t)A * 2 + 15) & ~(size_t)15));
    unsigned* rem = diag + (size_t)nchunk_max * 32;
    int* kept = (int*)((unsigned char*)rem + (((size_t)nchunk_max * 4 + 15) & ~(size_t)15));
    int* sh = kept + NMS_MAXDET;
    const float* conf = a.conf + (size_t)img * A;
    const float4* dbox = a.dbox + (size_t)img * A;
    const int* cls_id = a.cls_id + (size_t)img * A;
    pdl_trigger();
    if (tid < 8) sh[tid] = 0;
    pdl_wait();
    __syncthreads();
    for (int i0 = 0; i0 < A; i0 += NMS_THREADS) {                   // warp-uniform trip count
        const int i = i0 + tid;
        bool cand = false;
        unsigned long long key = 0;
        if (i < A) {
            const float c = conf[i];
            cand = c > NMSF_CONF;                                   // :150, :172 (positive, so the bit pattern orders like the value)
            key = ((unsigned long long)(0xffffffffu - __float_as_uint(c)) << 32) | (unsigned)i;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, cand);
        if (bal) {
            const int leader = __ffs(bal) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&sh[0], __popc(bal));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (cand) keys[base + __popc(bal & ((1u << lane) - 1))] = key;
        }
    }
    __syncthreads();
    const int T = sh[0];
    if (T == 0) {                                                   // coord() leaves the empty output -> (None, None) upstream
        if (tid == 0) a.counts[img] = 0;
        return;
    }
    int sort_n = 64;
    while (sort_n < T) sort_n <<= 1;
    for (int i = T + tid; i < sort_n; i += NMS_THREADS) keys[i] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= sort_n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < sort_n / 2; t += NMS_THREADS) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int hi = lo | j;
                const unsigned long long x = keys[lo], y = keys[hi];
                const bool up = (lo & k) == 0;
                if ((x > y) == up) { keys[lo] = y; keys[hi] = x; }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < T; i += NMS_THREADS) order[i] = (unsigned short)(keys[i] & 0xffffu);
    __syncthreads();                                                // keys are dead from here on: the region becomes bx
    for (int i = tid; i < T; i += NMS_THREADS) {
        const int an = order[i];
        const float4 d = dbox[an];
        const float dw = __fdiv_rn(d.z, 2.f), dh = __fdiv_rn(d.w, 2.f);            // xywh2xyxy :124-143
        const float off = __fmul_rn((float)cls_id[an], 7680.f);                     // :182
        bx[i] = (d.x - dw) + off; bx[A + i] = (d.y - dh) + off;                     // boxes = x[:, :4] + c  :186
        bx[2 * A + i] = (d.x + dw) + off; bx[3 * A + i] = (d.y + dh) + off;
    }
    const int nchunk = (T + 31) >> 5;
    for (int i = tid; i < nchunk; i += NMS_THREADS) rem[i] = 0;
    __syncthreads();
    // Same lazy greedy scheme as nms_kernel: diagonal 32x32 blocks first, then chunk by chunk resolve + apply kept rows.
    for (int c = wid; c < nchunk; c += NMS_THREADS / 32) {
        const int j = c * 32 + lane;
        const bool jv = j < T;
        const int jj = jv ? j : 0;
        const float jx1 = bx[jj], jy1 = bx[A + jj], jx2 = bx[2 * A + jj], jy2 = bx[3 * A + jj];
        const float ja = __fmul_rn(jx2 - jx1, jy2 - jy1);
        unsigned mine = 0;
        for (int ii = 0; ii < 32; ++ii) {
            const int i = c * 32 + ii;
            const bool s = jv && i < T && j > i && nmsf_suppresses(bx, A, i, jx1, jy1, jx2, jy2, ja);
            const unsigned bits = __ballot_sync(0xffffffffu, s);
            if (lane == ii) mine = bits;
        }
        diag[c * 32 + lane] = mine;
    }
    __syncthreads();
    int nk = 0;                                                     // block-uniform
    for (int c = 0; c < nchunk && nk < a.max_keep; ++c) {
        if (wid == 0) {
            unsigned cur = rem[c];
            const int nrow = min(32, T - c * 32);
            if (nrow < 32) cur |= 0xffffffffu << nrow;
            const unsigned d = diag[c * 32 + lane];
            unsigned keep = 0;
            int k = nk;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const unsigned di = __shfl_sync(0xffffffffu, d, i);
                if (!((cur >> i) & 1u) && k < a.max_keep) { keep |= 1u << i; cur |= di; ++k; }
            }
            if ((keep >> lane) & 1u) kept[nk + __popc(keep & ((1u << lane) - 1))] = c * 32 + lane;
            if (lane == 0) sh[6] = (int)keep;
        }
        __syncthreads();
        const unsigned keep = (unsigned)sh[6];
        nk += __popc(keep);
        if (keep && nk < a.max_keep) {
            for (int wc = c + 1 + wid; wc < nchunk; wc += NMS_THREADS / 32) {
                const int j = wc * 32 + lane;
                const bool jv = j < T;
                const int jj = jv ? j : 0;
                const float jx1 = bx[jj], jy1 = bx[A + jj], jx2 = bx[2 * A + jj], jy2 = bx[3 * A + jj];
                const float ja = __fmul_rn(jx2 - jx1, jy2 - jy1);
                unsigned acc = 0;
                for (unsigned m = keep; m; m &= m - 1) {
                    const int i = c * 32 + __ffs(m) - 1;
                    acc |= __ballot_sync(0xffffffffu, jv && nmsf_suppresses(bx, A, i, jx1, jy1, jx2, jy2, ja));
                }
                if (lane == 0) rem[wc] |= acc;
            }
        }
        __syncthreads();
    }
    if (tid == 0) a.counts[img] = nk;
    for (int r = tid; r < nk; r += NMS_THREADS) {
        const int an = order[kept[r]];
        float* row = a.dets + ((size_t)img * NMS_MAXDET + r) * 6;
        const float4 d = dbox[an];
        const float dw = __fdiv_rn(d.z, 2.f), dh = __fdiv_rn(d.w, 2.f);
        const float c4[4] = {d.x - dw, d.y - dh, d.x + dw, d.y + dh};
#pragma unroll
        for (int q = 0; q < 4; ++q) row[q] = fminf(fmaxf(c4[q], 0.f), 640.f);      // scale_boxes (gain 1, pad 0) + clip_boxes :203-252
        row[4] = conf[an];
        row[5] = (float)cls_id[an];
    }
}

// ---- weight quantiser: conv_quant() of stage_6_full_quant.py:89-126 ---------------------------
// grid (cout), block 256.  Per output channel: a = max|w|, s = fl32(M / a), q = rint(fl32(w * s)) (utils/quant_matrix.py:56-78,
// float32 arithmetic as numpy >= 2 evaluates it); bias_q = trunc(double(b) * (scale_input * double(s))) (utils/quant_bias.py:2-4,
// float64); scale_res = scale_input * double(s) (:93-96, :122).  An all-zero channel gives q = 0, scale = +inf.
__global__ void __launch_bounds__(256) quant_weights_kernel(const float* __restrict__ w, const float* __restrict__ bias, size_t per_channel,
                                                            int M, double scale_input, int8_t* __restrict__ qw, long long* __restrict__ qb,
                                                            double* __restrict__ scale_res) {
    __shared__ float red[8];
    __shared__ float s_sh;
    const int c = blockIdx.x;
    const float* wc = w + (size_t)c * per_channel;
    float m = 0.f;
    for (size_t i = threadIdx.x; i < per_channel; i += 256) m = fmaxf(m, fabsf(wc[i]));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
        const float s = __fdiv_rn((float)M, m);
        s_sh = s;
        const double bs = __dmul_rn(scale_input, (double)s);
        scale_res[c] = bs;
        qb[c] = m > 0.f ? (long long)__dmul_rn((double)bias[c], bs) : 0ll;
    }
    __syncthreads();
    const float s = s_sh;
    for (size_t i = threadIdx.x; i < per_channel; i += 256)
        qw[(size_t)c * per_channel + i] = isfinite(s) ? (int8_t)__float2int_rn(__fmul_rn(wc[i], s)) : (int8_t)0;
}

// ---- calibration forward (stage_4.py:475-946): the BN-fused FLOAT network with abs-max taps -------
// Not a hot path (a handful of calibration images per bit-width sweep); plain fp32 kernels, NCHW like the reference.
// conv: grid (ceil(Hout*Wout / 128), ceil(cout / 16), n), block 128: one output pixel x 16 output channels per thread, weights of
// 8 input channels at a time in shared memory, FMA accumulation; the per-image abs-max tap (save_max_a, utils/save_a.py:22) is
// fused: amax[img] = max(amax[img], max|y|) through an integer atomicMax on the (non-negative) float bits.
#define CALIB_CO 16
#define CALIB_CI 8
__global__ void __launch_bounds__(128) calib_conv_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                             float* __restrict__ y, float* __restrict__ amax, int cin, int H, int W,
                                                             int cout, int Hout, int Wout, int ks, int stride) {
    __shared__ float sw[CALIB_CI * 9 * CALIB_CO];                 // [ci][tap][co]
    const int pad = ks >> 1, taps = ks * ks;
    const int img = blockIdx.z, co0 = blockIdx.y * CALIB_CO;
    const int p = blockIdx.x * 128 + threadIdx.x;
    const bool live = p < Hout * Wout;
    const int oy = live ? p / Wout : 0, ox = live ? p - oy * Wout : 0;
    float acc[CALIB_CO];
#pragma unroll
    for (int j = 0; j < CALIB_CO; ++j) acc[j] = co0 + j < cout ? __ldg(b + co0 + j) : 0.f;
    const float* xi = x + (size_t)img * cin * H * W;
    for (int c0 = 0; c0 < cin; c0 += CALIB_CI) {
        const int nc = min(CALIB_CI, cin - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nc * taps * CALIB_CO; i += 128) {
            const int j = i % CALIB_CO, t = (i / CALIB_CO) % taps, c = i / (CALIB_CO * taps);
            sw[i] = co0 + j < cout ? __ldg(w + ((size_t)(co0 + j) * cin + c0 + c) * taps + t) : 0.f;
        }
        __syncthreads();
        if (live) {
            for (int c = 0; c < nc; ++c) {
                const float* xc = xi + (size_t)(c0 + c) * H * W;
                for (int t = 0; t < taps; ++t) {
                    const int iy = oy * stride + t / ks - pad, ix = ox * stride + t % ks - pad;
                    if ((unsigned)iy >= (unsigned)H || (unsigned)ix >= (unsigned)W) continue;
                    const float v = __ldg(xc + (size_t)iy * W + ix);
                    const float* wr = sw + (c * taps + t) * CALIB_CO;
#pragma unroll
                    for (int j = 0; j < CALIB_CO; ++j) acc[j] = __fmaf_rn(v, wr[j], acc[j]);
                }
            }
        }
    }
    float m = 0.f;
    if (live) {
#pragma unroll
        for (int j = 0; j < CALIB_CO; ++j)
            if (co0 + j < cout) {
                y[(((size_t)img * cout + co0 + j) * Hout + oy) * Wout + ox] = acc[j];
                m = fmaxf(m, fabsf(acc[j]));
            }
    }
    if (amax) {
#pragma unroll
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax((int*)amax + img, __float_as_int(m));
    }
}
__global__ void calib_silu_f32_kernel(float* __restrict__ x, size_t total) {                    // nn.SiLU: x * sigmoid(x)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        x[i] = __fdiv_rn(v, 1.f + expf(-v));
    }
}
__global__ void calib_maxpool5_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int planes, int H, int W) {   // MaxPool2d(5, 1, 2)
    const size_t total = (size_t)planes * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int xx = (int)(i % W), yy = (int)((i / W) % H);
        const float* pl = x + (i / ((size_t)H * W)) * H * W;
        float m = -INFINITY;
        for (int dy = -2; dy <= 2; ++dy)
            for (int dx = -2; dx <= 2; ++dx) {
                const int iy = yy + dy, ix = xx + dx;
                if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W) m = fmaxf(m, pl[(size_t)iy * W + ix]);
            }
        y[i] = m;
    }
}
__global__ void calib_upsample2_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int planes, int H, int W) { // nearest, x2
    const size_t total = (size_t)planes * H * W * 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int xx = (int)(i % (2 * W)), yy = (int)((i / (2 * W)) % (2 * H));
        const size_t pl = i / ((size_t)4 * H * W);
        y[i] = x[pl * H * W + (size_t)(yy >> 1) * W + (xx >> 1)];
    }
}

// ---- export a plane buffer as NCHW int32 (parity taps) ------------------------------------------------
__global__ void export_planes_kernel(const void* __restrict__ src, int elem_bytes, int nplanes, int n, int H, int W, int* __restrict__ dst) {
    const size_t total = (size_t)n * nplanes * 16 * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const int c = (int)((i / ((size_t)W * H)) % (nplanes * 16));
        const int img = (int)(i / ((size_t)W * H * nplanes * 16));
        const size_t s = (((size_t)(c >> 4) * n + img) * H * W + (size_t)y * W + x) * 16 + (c & 15);
        dst[i] = elem_bytes == 1 ? (int)((const int8_t*)src)[s] : (int)((const int16_t*)src)[s];
    }
}

// ---- quantised layer library on fp32-carried integer tensors (unit-level drop-ins) --------------------
__global__ void requantize_f32_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ k,
                                      const float* __restrict__ inv, int per_channel, int c, int hw, size_t total, int M) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ch = per_channel ? (int)((i / hw) % c) : 0;
        y[i] = (float)requant(x[i], k[ch], inv[ch], M);
    }
}
__global__ void silu_f32_kernel(const float* __restrict__ acc, float* __restrict__ y, const float* __restrict__ tab,
                                const float* __restrict__ lut, int c, int hw, size_t total, int M) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)((i / hw) % c);
        const float a = acc[i];
        const int r1 = rq_round(__fmul_rn(tab[ch], a), tab[c + ch], M);
        const float pr = __fmul_rn(lut[r1 + M], a);
        y[i] = (float)rq_round(__fmul_rn(tab[2 * c + ch], pr), tab[3 * c + ch], M);
    }
}
__global__ void lut_f32_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ lut,
                               int key_min, int key_max, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        const float r = rintf(v);
        y[i] = (r == v && r >= (float)key_min && r <= (float)key_max) ? lut[(int)r - key_min] : 0.f;
    }
}
__global__ void quant_input_f32_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ amax,
                                       float* __restrict__ scales, size_t per_image, int n, int M) {
    const size_t total = per_image * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int img = (int)(i / per_image);
        const float a = amax[img];
        const float s = __fmul_rn(__frcp_rn(a), (float)M);      // M / tensor == reciprocal(tensor) * M in torch
        if (i % per_image == 0) scales[img] = s;
        const float v = fminf(fmaxf(x[i], -a), a);
        y[i] = a > 0.f ? rintf(__fmul_rn(v, s)) : 0.f;
    }
}

}  // namespace ayq
