// fixedpoint.cuh -- the reference's fixed-point arithmetic as device functions.
//
// Restates utils/rescale_coeff_torch.py:42-46 (requantize) and stage_8_torch_full_quant.py:439-452
// (silu) exactly as they evaluate on fp32 tensors that carry integers:
//   t   = RN32(k * x)                      fp32 product (can exceed 2^31, so no int32 maths)
//   q   = floor(t / 2^(s-1));  q = floor(q / 2) + q mod 2        ==  floor((t + 2^(s-1)) / 2^s)
//   out = clamp(q, -M, M)
// floor(t*2^-s + 1/2) is evaluated as  float2int_rd(fma_rd(t, 2^-s, 0.5)):  the round-DOWN fma can
// never step over the integer below the exact sum, so the result is exact whenever it lies inside
// the int32 range, and saturates (then clamps to +-M) outside it.  All products use __fmul_rn so
// that nvcc cannot contract them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ayq {

__device__ __forceinline__ int rq_round(float t, float inv2s, int M) {
    int q = __float2int_rd(__fmaf_rd(t, inv2s, 0.5f));
    return max(-M, min(M, q));
}

// requantize(): x is an integer carried in fp32 (exact below 2^24, like the reference)
__device__ __forceinline__ int requant(float x, float k, float inv2s, int M) {
    return rq_round(__fmul_rn(k, x), inv2s, M);
}

// silu(): acc = conv accumulator (+bias); lut[r + M] = sigmoid table entry as float
__device__ __forceinline__ int silu_q(int acc, float k1, float i1, float k2, float i2,
                                      const float* __restrict__ lut, int M) {
    float a = __int2float_rn(acc);
    int r1 = rq_round(__fmul_rn(k1, a), i1, M);
    float pr = __fmul_rn(lut[r1 + M], a);          // res_silu *= res_conv_copy (fp32), round() is a no-op
    return rq_round(__fmul_rn(k2, pr), i2, M);
}

__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {
    return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) | ((uint32_t)(d & 0xff) << 24);
}

}  // namespace ayq
