// fixedpoint.cuh -- the reference's fixed-point arithmetic as device functions.
//
// Restates utils/rescale_coeff_torch.py:42-46 (requantize) and stage_8_torch_full_quant.py:439-452
// (silu) exactly as they evaluate on fp32 tensors that carry integers:
//   t   = RN32(k * x)                      fp32 product (can exceed 2^31, so no int32 maths)
//   q   = floor(t / 2^(s-1));  q = floor(q / 2) + q mod 2        ==  floor((t + 2^(s-1)) / 2^s)
//   out = clamp(q, -M, M)
// floor(t*2^-s + 1/2) is evaluated as  cvt.rmi(fma_rd(t, 2^-s, 0.5)):  the round-DOWN fma can never
// step over the integer below the exact sum, so the floor is exact; the conversion saturates at the
// destination width (int32 / int16 / int8), which is harmless because the result is clamped to +-M
// (M <= 127 for activations, 32767 for the 16-bit logits) right after.  All products use __fmul_rn
// so that nvcc cannot contract them.
//
// The hot epilogues use the narrow conversions: cvt.rmi.sat.s8.f32 (SASS F2I.S8.FLOOR) gives
// clamp(floor(x), -128, 127) in one instruction, so an activation requant costs
// FMUL + FFMA.RM + F2I.S8 + max(-M) [+ min(M) when M < 127].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ayq {

// ---- programmatic dependent launch (PDL): kernels of one pass are launched with programmatic stream serialization, so a
// kernel's prologue (barrier init, TMEM allocation, weight / table loads) overlaps the tail of its predecessor.
// pdl_wait() blocks until every prerequisite grid has completed and flushed; it must precede the first access to data
// the predecessor wrote.  Both are no-ops for launches without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- generic (any clamp up to 2^23): used by the fp32 layer-library kernels ---------------------------
__device__ __forceinline__ int rq_round(float t, float inv2s, int M) {
    int q = __float2int_rd(__fmaf_rd(t, inv2s, 0.5f));
    return max(-M, min(M, q));
}
__device__ __forceinline__ int requant(float x, float k, float inv2s, int M) {
    return rq_round(__fmul_rn(k, x), inv2s, M);
}

// ---- narrow saturating floors ----------------------------------------------------------------------
__device__ __forceinline__ int floor_sat_s8(float x) {       // clamp(floor(x), -128, 127)
    int r;
    asm("cvt.rmi.sat.s8.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ int floor_sat_s16(float x) {      // clamp(floor(x), -32768, 32767)
    int r;
    asm("cvt.rmi.sat.s16.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// requantize() to an activation width (M <= 127)
__device__ __forceinline__ int requant8(float x, float k, float inv2s, int M) {
    return max(-M, min(M, floor_sat_s8(__fmaf_rd(__fmul_rn(k, x), inv2s, 0.5f))));
}
// requantize() to 16 bits (M = 32767)
__device__ __forceinline__ int requant16(float x, float k, float inv2s) {
    return max(-32767, floor_sat_s16(__fmaf_rd(__fmul_rn(k, x), inv2s, 0.5f)));
}

// ---- silu() -----------------------------------------------------------------------------------------
// 256-entry sigmoid table indexed by the SATURATED first requant:  lut256[i] = table[clamp(i - 128, -M, M)],
// so the first clamp of silu() is folded into the table (fill_lut256 below).
#define AYQ_LUT256 256
__device__ __forceinline__ void fill_lut256(float* __restrict__ dst, const float* __restrict__ table /*[2M+1]*/, int M,
                                            int tid, int nthreads) {
    for (int i = tid; i < AYQ_LUT256; i += nthreads) {
        const int r = max(-M, min(M, i - 128));
        dst[i] = table[r + M];
    }
}
// acc = conv accumulator (+bias).  Returns the K-bit activation in [-M, M].
__device__ __forceinline__ int silu_q(int acc, float k1, float i1, float k2, float i2,
                                      const float* __restrict__ lut256, int M) {
    const float a = __int2float_rn(acc);
    const int r1 = floor_sat_s8(__fmaf_rd(__fmul_rn(k1, a), i1, 0.5f));
    const float pr = __fmul_rn(lut256[r1 + 128], a);       // res_silu *= res_conv_copy (fp32), round() is a no-op
    return max(-M, min(M, floor_sat_s8(__fmaf_rd(__fmul_rn(k2, pr), i2, 0.5f))));
}
// K = 8 (M = 127): the saturating conversion already bounds the result above, only -128 needs the clamp.  `half` = 0.5f
// passed in a REGISTER (kernel parameter) so that the per-channel 2^-s can be the constant-bank operand of the FFMA.
__device__ __forceinline__ int silu_q127(int acc, float k1, float i1, float k2, float i2,
                                         const float* __restrict__ lut256, float half) {
    const float a = __int2float_rn(acc);
    const int r1 = floor_sat_s8(__fmaf_rd(__fmul_rn(k1, a), i1, half));
    const float pr = __fmul_rn(lut256[r1 + 128], a);
    return max(-127, floor_sat_s8(__fmaf_rd(__fmul_rn(k2, pr), i2, half)));
}
// Folded coefficients kp = k * 2^-s (an exact power-of-two scaling, so RN32(kp * x) == RN32(k * x) * 2^-s bit for bit):
// floor(RN32(k*x) * 2^-s + 1/2) = floor(RD(RN32(kp*x) + 1/2)).  One coefficient per requant instead of two.
__device__ __forceinline__ int silu_q127f(int acc, float k1p, float k2p, const float* __restrict__ lut256, float half) {
    const float a = __int2float_rn(acc);
    const int r1 = floor_sat_s8(__fadd_rd(__fmul_rn(k1p, a), half));
    const float pr = __fmul_rn(lut256[r1 + 128], a);
    return max(-127, floor_sat_s8(__fadd_rd(__fmul_rn(k2p, pr), half)));
}
// MAGIC variant (no int->float conversion, no lower clamp).  The accumulator arrives as m = 1.5 * 2^23 + acc, built by an
// INTEGER add of (bias + 0x4B400000) to the raw accumulator and reinterpreted as a float: exact for |acc| < 2^22 (the host
// proves the bound per layer from the weights).  Then
//   RN(kp * acc)      = fma_rn(kp, m, -kp * C)     C = 1.5 * 2^23; kp * C = k * 3 * 2^(22-s) is exact, so the FMA's single
//   RN(lut[r1] * acc) = fma_rn(l, m, -l * C)       rounding sees exactly kp * (m - C) = kp * acc (the table holds (l, -l*C))
// and the final saturating conversion needs no max(-127, .) when the host has shown that no accumulator can reach -128
// (SiLU is bounded below; see magic_epilogue_ok in conv_tma.cuh).
#define AYQ_MAGIC_F 12582912.0f
#define AYQ_MAGIC_I 0x4B400000
__device__ __forceinline__ int silu_magic(int acc_plus_bias_magic, float k1p, float c1, float k2p, const float2* __restrict__ lut2, float half) {
    const float m = __int_as_float(acc_plus_bias_magic);
    const int r1 = floor_sat_s8(__fadd_rd(__fmaf_rn(k1p, m, c1), half));
    const float2 l = lut2[r1 + 128];
    const float pr = __fmaf_rn(l.x, m, l.y);
    return floor_sat_s8(__fadd_rd(__fmul_rn(k2p, pr), half));
}
__device__ __forceinline__ void fill_lut256_magic(float2* __restrict__ dst, const float* __restrict__ table /*[2M+1]*/, int M,
                                                  int tid, int nthreads) {
    for (int i = tid; i < AYQ_LUT256; i += nthreads) {
        const int r = max(-M, min(M, i - 128));
        const float l = table[r + M];
        dst[i] = make_float2(l, -__fmul_rn(l, AYQ_MAGIC_F));
    }
}
// MAGIC2: the variant the product kernels run.  Measured on B200 (tools/ubench, profiles/ubench_r2.txt): silu_magic above costs
// 17.5 cycles per 32-element row per SM sub-partition however many warps run it, because F2I.S8 issues once every 8 cycles (two per
// element = 16.5 cycles of the XU pipe) and the 8-byte table gather averages 3.5 shared-memory wavefronts.  Here
//   * the first requant never becomes an integer: with the coefficient pre-scaled by 2^-8 (exact), y = RD_sat(t / 256 + (0.5 + 128) / 256)
//     is (clamp(floor(t + 1/2), -128, 128) + 128) / 256 (the float floor cannot step over j / 256, all representable; .sat is
//     the clamp), RD(y + 32768) has that index j in its low mantissa bits (ulp 2^-8), and ONE shift-add turns the bits into the
//     shared-memory address of table entry j: no conversion, no separate clamp, no index arithmetic;
//   * the table is replicated per lane (entry j of lane L at word j * 32 + L, 257 entries, 32.9 KB): every gather is one
//     conflict-free wavefront of 4-byte words;
//   * float(acc) = m - C is formed once (exact) and feeds both products as plain multiplies whose coefficient is a constant-bank
//     operand (no per-channel constants in registers).
// 11.5 issue slots per element with one XU instruction (the final saturating floor): bound by instruction issue.
#define AYQ_LUTREP_N 257
#define AYQ_LUTREP_BYTES (AYQ_LUTREP_N * 32 * 4)
// REP_LOG: log2 of the copies per entry (5 = one per lane, conflict-free; 3 = eight copies, 8 KB, at most 4-way conflicts: for
// kernels that are not bound by the gather and need the shared memory for occupancy, i.e. Conv_P1)
#define AYQ_LUTREP8_BYTES (AYQ_LUTREP_N * 8 * 4)
template <int REP_LOG = 5>
__device__ __forceinline__ void fill_lut_rep(float* __restrict__ dst, const float* __restrict__ table /*[2M+1]*/, int M, int tid, int nthreads) {
    for (int i = tid; i < (AYQ_LUTREP_N << REP_LOG); i += nthreads) {
        const int r = max(-M, min(M, (i >> REP_LOG) - 128));
        dst[i] = table[r + M];
    }
}
// per-thread part of the gather address: table base + 4 * (lane mod copies) + the wrapped bits of 32768.0f << (2 + REP_LOG)
template <int REP_LOG = 5>
__device__ __forceinline__ uint32_t lut_rep_thread_base(uint32_t table_smem_addr, uint32_t lane) {
    return table_smem_addr + ((lane & ((1u << REP_LOG) - 1u)) << 2) - (0x47000000u << (2 + REP_LOG));
}
// lut_thr = shared address of the table + 4 * lane + 0x80000000 (the bits of 32768.0f, shifted left by 7, wrap to 0x80000000)
// WIDE: for layers whose accumulator bound exceeds 2^22 or whose clamp is not 127 (K = 6 / 4): float(acc) by conversion (one more
// XU instruction) and explicit clamps to +-M on the result; everything else identical.
template <bool WIDE>
__device__ __forceinline__ int silu_magic2(int acc_plus_bias_magic, float k1s /* k1 * 2^-s1 * 2^-8 */, float k2p, uint32_t lut_thr, float half, int M = 127) {
    const float af = WIDE ? __int2float_rn(acc_plus_bias_magic)                          // plain accumulator + bias
                          : __fadd_rn(__int_as_float(acc_plus_bias_magic), -AYQ_MAGIC_F);  // exact: |acc + bias| < 2^22
    const float t = __fmul_rn(k1s, af);                                                  // RN32(k1 * acc) * 2^-(s1 + 8), bit for bit
    float y;
    asm("add.rm.sat.f32 %0, %1, %2;" : "=f"(y) : "f"(t), "f"(0.501953125f));
    const float w = __fadd_rd(y, 32768.0f);
    float l;
    asm("ld.shared.f32 %0, [%1];" : "=f"(l) : "r"((__float_as_uint(w) << 7) + lut_thr));
    const float pr = __fmul_rn(l, af);                                                   // RN32(sig * acc): res_silu *= res_conv_copy
    const int r = floor_sat_s8(__fadd_rd(__fmul_rn(k2p, pr), half));
    return WIDE ? max(-M, min(M, r)) : r;
}
// ---- two elements per instruction: sm_100 has packed FP32 pairs (add / mul / fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2, any
// rounding mode, no .sat).  The conv epilogue is bound by instruction ISSUE (ncu: 0.9 instructions per cycle and scheduler, of
// which two thirds are this arithmetic), not by the FP32 pipe, so pairing the four multiplies / adds that need no saturation
// saves 3 of every 11.5 issue slots per element.  Bit-exact: each half of a packed instruction is the same IEEE operation.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float a, float b) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2_t f2_add_rn(f32x2_t a, f32x2_t b) { f32x2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t f2_add_rm(f32x2_t a, f32x2_t b) { f32x2_t r; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t f2_mul_rn(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// The three constant pairs of the packed epilogue, materialised ONCE per thread in ordinary register pairs.  (As literals the
// compiler rebuilt them in uniform registers in front of every use: 37 UMOVs per 32 elements in the ncu source view.)
struct EpiPairs { f32x2_t neg_c, k32768, half; };
__device__ __forceinline__ f32x2_t f2_opaque(float v) { f32x2_t r; asm volatile("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v)); return r; }
__device__ __forceinline__ EpiPairs epi_pairs() { EpiPairs p; p.neg_c = f2_opaque(-AYQ_MAGIC_F); p.k32768 = f2_opaque(32768.0f); p.half = f2_opaque(0.5f); return p; }
// silu_magic2 on elements (v0, v1) of two adjacent channels; k1s / k2p hold the two channels' coefficients
// I2F: float(acc) by conversion (accumulators beyond 2^22) instead of the magic add; CLAMP: explicit clamp of the result to +-M
// (a clamp other than 127, or a layer for which the host cannot prove that -128 is unreachable)
template <bool I2F, bool CLAMP = I2F, int REP_LOG = 5>
__device__ __forceinline__ void silu_magic2_x2(int v0, int v1, f32x2_t k1s, f32x2_t k2p, uint32_t lut_thr, int M, int& r0, int& r1, const EpiPairs& cp) {
    const f32x2_t af = I2F ? f2_pack(__int2float_rn(v0), __int2float_rn(v1))
                            : f2_add_rn(f2_pack(__int_as_float(v0), __int_as_float(v1)), cp.neg_c);
    float t0, t1, y0, y1, w0, w1, l0, l1, z0, z1;
    f2_unpack(f2_mul_rn(k1s, af), t0, t1);
    asm("add.rm.sat.f32 %0, %1, %2;" : "=f"(y0) : "f"(t0), "f"(0.501953125f));
    asm("add.rm.sat.f32 %0, %1, %2;" : "=f"(y1) : "f"(t1), "f"(0.501953125f));
    f2_unpack(f2_add_rm(f2_pack(y0, y1), cp.k32768), w0, w1);
    asm("ld.shared.f32 %0, [%1];" : "=f"(l0) : "r"((__float_as_uint(w0) << (2 + REP_LOG)) + lut_thr));
    asm("ld.shared.f32 %0, [%1];" : "=f"(l1) : "r"((__float_as_uint(w1) << (2 + REP_LOG)) + lut_thr));
    const f32x2_t pr = f2_mul_rn(f2_pack(l0, l1), af);
    f2_unpack(f2_add_rm(f2_mul_rn(k2p, pr), cp.half), z0, z1);
    r0 = floor_sat_s8(z0); r1 = floor_sat_s8(z1);
    if (CLAMP) { r0 = max(-M, min(M, r0)); r1 = max(-M, min(M, r1)); }
}
// requantize() of two adjacent channels of a raw accumulator (requant_last_layers / exponent_requant epilogues) with the magic
// int -> float add and packed multiplies: v = acc + bias + 0x4B400000.  WIDE16: 16-bit result (clamp +-32767), else 8-bit (+-127).
template <bool WIDE16>
__device__ __forceinline__ void requant_magic_x2(int v0, int v1, f32x2_t kp, int& r0, int& r1, const EpiPairs& cp) {
    const f32x2_t af = f2_add_rn(f2_pack(__int_as_float(v0), __int_as_float(v1)), cp.neg_c);
    float z0, z1;
    f2_unpack(f2_add_rm(f2_mul_rn(kp, af), cp.half), z0, z1);
    if (WIDE16) { r0 = max(-32767, floor_sat_s16(z0)); r1 = max(-32767, floor_sat_s16(z1)); }
    else { r0 = max(-127, floor_sat_s8(z0)); r1 = max(-127, floor_sat_s8(z1)); }
}
// four values already in [-128, 127] -> one word (cvt.pack: two instructions instead of three logic ops)
__device__ __forceinline__ uint32_t pack4_sat(int a, int b, int c, int d) {
    // cvt.pack d, x, y, z:  d = (z << 16) | (sat8(x) << 8) | sat8(y)
    uint32_t hi, w;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(0));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(b), "r"(a), "r"(hi));
    return w;
}
__device__ __forceinline__ int requant8_127f(float x, float kp, float half) {
    return max(-127, floor_sat_s8(__fadd_rd(__fmul_rn(kp, x), half)));
}
__device__ __forceinline__ int requant16_f(float x, float kp, float half) {
    return max(-32767, floor_sat_s16(__fadd_rd(__fmul_rn(kp, x), half)));
}
__device__ __forceinline__ int requant8_127(float x, float k, float inv2s, float half) {
    return max(-127, floor_sat_s8(__fmaf_rd(__fmul_rn(k, x), inv2s, half)));
}
__device__ __forceinline__ int requant16_h(float x, float k, float inv2s, float half) {
    return max(-32767, floor_sat_s16(__fmaf_rd(__fmul_rn(k, x), inv2s, half)));
}

__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {
    return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) | ((uint32_t)(d & 0xff) << 24);
}

}  // namespace ayq
