// plan_format.h -- binary layout of the plan blob produced by alpha_yolo_quant_b200/plan.py and
// consumed by ayq_create().  Keep the field indices in sync with plan.py (tests/test_plan.py checks
// the header constants by parsing this file).
//
// Activation buffers are stored as 16-channel planes:  [plane][N][H][W][16] elements (int8; int16
// for the class logits).  A pixel's 16 channels are 16 contiguous bytes = one row of a tcgen05
// K-major core matrix, so concatenation is a list of plane ranges ("segments") and never a copy.
#pragma once
#include <stdint.h>

#define AYQ_MAGIC 0x31515941u      // "AYQ1"
#define AYQ_PLAN_VERSION 7

struct PlanHeader {
    uint32_t magic, version;
    int32_t K;                     // activation / weight bit width (stage_0.py:7)
    int32_t n_bufs, n_ops;
    int32_t img_h, img_w, n_anchors;
    uint64_t bufs_off, ops_off, data_off, data_bytes;
};

struct BufDesc { int32_t nplanes, H, W, elem_bytes; };

#define AYQ_OP_FIELDS 64
struct OpDesc { int32_t f[AYQ_OP_FIELDS]; };

enum { OP_CONV = 1, OP_CONV_P1 = 2, OP_POOL = 3, OP_HEAD = 4, OP_NMS = 5,
       OP_HEAD_FLOAT = 6, OP_NMS_FLOAT = 7 };   // stage_8_torch.py: float Detect head + coord() (torchvision NMS semantics)
// conv epilogues
enum { EPI_SILU = 0,        // silu() then up to AYQ_MAX_OUT stores (identity or scalar requant, optional 2x upsample)
       EPI_REQUANT8 = 1,    // per-channel requantize of the raw accumulator to K bits   (requant_last_layers)
       EPI_REQUANT16 = 2 }; // per-channel requantize of the raw accumulator to 16 bits  (exponent_requant)
enum { OUT_IDENT = 0, OUT_REQUANT = 1 };

#define AYQ_MAX_OUT 3
// OP_CONV fields
enum {
    CF_KIND = 0, CF_KSIZE = 1, CF_STRIDE = 2, CF_HIN = 3, CF_WIN = 4, CF_HOUT = 5, CF_WOUT = 6, CF_COUT = 7,
    CF_NKC = 8,                        // number of 16-channel K chunks = sum over segments of taps * planes
    CF_KC_OFF = 9,                     // data offset: int32[nkc][4] = {buf, plane, ky, kx}
    CF_W_OFF = 10,                     // data offset: int8[nkc_pad][cout][16]  (nkc_pad = nkc rounded up to even, zero filled)
    CF_BIAS_OFF = 11,                  // data offset: int32[cout]
    CF_TAB_OFF = 12,                   // data offset: float[4][cout] = k1, 2^-s1, k2, 2^-s2   (EPI_REQUANT*: k, 2^-s in rows 0,1)
    CF_EPI = 13, CF_CLAMP = 14,        // clamp = 2^(bits-1)-1 of the epilogue result
    CF_LUT_OFF = 15,                   // data offset: float[2*M+1] sigmoid table (EPI_SILU)
    CF_NOUT = 16,
    CF_OUT0 = 17,                      // AYQ_MAX_OUT x {buf, plane0, mode, k (float bits), 2^-s (float bits), layout}; layout 0 = planes
                                       // [plane][n][H][W][16], 1 = 2x nearest upsample, 2 = phase-split [(y&1)*2+(x&1)][plane][n][H/2][W/2][16]
                                       // (the input layout of a stride-2 conv: every tap becomes a stride-1 box of one phase image)
    CF_OUT_STRIDE = 6,
    CF_LAYER = 40,                     // index into plan.LAYERS (all_scales key order)
    CF_ACC_TAP = 41,                   // -1 or index of the int32 accumulator tap (parity tests)
    CF_NAME_OFF = 42,                  // data offset of a NUL terminated layer name
    CF_ACC_BUF = 43                    // -1 or an int32 buffer (elem_bytes 4) that receives the raw accumulators as NCHW (n, cout, H, W):
                                       // the head inputs of stage_8_torch.py:915-922
};

// OP_CONV_P1: Conv_P1 reads the fp32 NCHW image, fuses the per-image input quantiser (quant_matrix)
enum { P1_HOUT = 1, P1_WOUT = 2, P1_OUT_BUF = 3, P1_W_OFF = 4,   // int8[16][32]: k = (ky*3+kx)*3 + c, zero padded to 32
       P1_BIAS_OFF = 5, P1_TAB_OFF = 6, P1_CLAMP = 7, P1_LUT_OFF = 8, P1_ACC_TAP = 9, P1_QTAP_BUF = 10,
       P1_OUT_PS = 11 };           // 1: the output buffer is phase-split (see CF_OUT0 'upsample' = 2)
// OP_POOL: SPPF cascade of three MaxPool2d(5,1,2); writes p1,p2,p3
enum { PL_IN_BUF = 1, PL_IN_PLANE0 = 2, PL_NPLANES = 3, PL_OUT_BUF = 4, PL_OUT_PLANE0 = 5, PL_H = 6, PL_W = 7 };
// OP_HEAD: DFL decode + class score max/argmax
enum { HD_BOX_BUF0 = 1,                // 3 x int8 buffers (4 planes each), P3 P4 P5
       HD_CLS_BUF0 = 4,                // 3 x int16 buffers (5 planes each)
       HD_LUT_EXP_OFF = 7,             // float[2^K] exponent table, index y + 2^K - 1
       HD_LUT16_OFF = 8,               // int16[65535] final sigmoid table, index l + 32767
       HD_DFLW_OFF = 9,                // int32[16] integer dfl.weight
       HD_ANCH_OFF = 10,               // int32[n_anchors][2] quantised anchor points
       HD_KD = 11, HD_ID = 12,         // float bits of the dfl requant coefficient and 2^-s
       HD_LO16_OFF = 13,               // int16[65535]: smallest logit with the same final-sigmoid value, index l + 32767
       HD_MONO = 14 };                 // 1 when the final sigmoid table is monotone non-decreasing
// OP_HEAD_FLOAT: dequantise + float softmax/DFL/sigmoid decode of the six raw head accumulators (stage_8_torch.py:915-947)
enum { HF_BOX_BUF0 = 1,                // 3 x int32 NCHW accumulator buffers (64 channels), P3 P4 P5
       HF_CLS_BUF0 = 4,                // 3 x int32 NCHW accumulator buffers (80 channels)
       HF_BOX_SCALE_OFF = 7,           // float[3][64] all_scales['*_up_2']
       HF_CLS_SCALE_OFF = 8,           // float[3][80] all_scales['*_down_2']
       HF_DFLW_OFF = 9 };              // float[16] dfl.weight

