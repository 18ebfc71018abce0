/* ayq.h -- C ABI of the B200-native integer YOLOv8n + q_NMS engine (libayq.so).
 *
 * The reference (Alpha-Chip/Alpha-Yolo-Quant) is pure Python and has no FFI; the boundary it
 * exposes is the set of Python names at the bottom of quantisation/stage_8_torch_full_quant.py.
 * Each entry point below names the reference code it replaces (paths relative to
 * /root/reference/quantisation/).  The Python shim alpha_yolo_quant_b200/ binds these with
 * ctypes and re-exports the reference's names.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a
 * negative errno-style code, ayq_last_error() gives the message of the last failure on the
 * calling thread; image / detection buffers are caller-owned (torch allocations), the engine owns
 * its packed weights, tables and activation workspace; a handle is bound to one GPU and is not thread-safe.
 *
 * Stream semantics.  Entry points that take a `stream` (a cudaStream_t passed as void*) launch all their kernels on it and
 * return without synchronising.  The host-buffer entries (ayq_forward_host*, ayq_forward_host_async) run on three internal
 * non-blocking streams (H2D | kernels | D2H).  Every entry shares the ONE engine-owned workspace, so the engine orders them
 * itself: each entry first makes its stream(s) wait for an event recorded at the end of the previous entry (whatever stream
 * that one ran on) and records the event again when it has enqueued its own work.  Calls on different streams are therefore
 * serialised on the device in call order and never overlap; the caller only has to order accesses to its OWN buffers.
 */
#ifndef AYQ_H
#define AYQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ayq_engine* ayq_handle;

#define AYQ_MAX_DET 300          /* max_det, stage_8_torch_full_quant.py:312 */
#define AYQ_ANCHORS 8400         /* 80*80 + 40*40 + 20*20 anchors at 640x640 */
#define AYQ_DET_STRIDE 6         /* x1 y1 x2 y2 conf class, stage_8_torch_full_quant.py:355-359 */

const char* ayq_last_error(void);
int ayq_version(void);

/* Yolov8() + load_state_dict() (stage_8_torch_full_quant.py:1278-1281) and the import-time
 * globals all_scales / max_a_dict / lookup* (:432-436): the plan blob is produced by
 * alpha_yolo_quant_b200.plan.compile_plan() from the same three artefacts. */
int ayq_create(const void* plan_blob, size_t nbytes, int device, ayq_handle* out);
int ayq_destroy(ayq_handle h);

/* images processed per internal pass (the activation workspace is sized for min(n, max_batch) images, ~17 MB per image);
 * default 512: every layer is its own launch with ~8.6 us of fixed cost, 512-image passes run 8 % faster per image than 256 */
int ayq_set_max_batch(ayq_handle h, int max_batch);
/* bytes of engine-owned device workspace currently allocated */
size_t ayq_workspace_bytes(ayq_handle h);

/* Yolov8.forward (stage_8_torch_full_quant.py:704-1275) for a batch: element i of the outputs
 * equals the reference's model(img[i:i+1]).
 *   img       device, float32 (n,3,640,640) in [0,1]            (read only)
 *   dbox_cls  device, float32 (n,84,8400) or NULL: the tensor passed to coord_quant (:1261)
 *   dets      device, float32 (n,300,6): rows [x1,y1,x2,y2,conf,class], first counts[i] valid
 *   counts    device, int32 (n): 0 <=> the reference returns (None, None)                        */
int ayq_forward(ayq_handle h, const float* img, int n, float* dbox_cls, float* dets, int32_t* counts, void* stream);
/* The same with the images as DEVICE uint8 (n,3,640,640) CHW, the loader's format before ToTensor (stage_8_torch.py:985-990; what
 * a GPU JPEG decoder delivers): u8 / 255 and the per-image input quantiser (utils/quant_matrix_torch.py:57-70) run inside
 * Conv_P1.  Bit-identical to ayq_forward on (u8 / 255).to(float32); a quarter of the input bytes. */
int ayq_forward_u8(ayq_handle h, const uint8_t* img_u8, int n, float* dbox_cls, float* dets, int32_t* counts, void* stream);

/* The same call with HOST buffers (validation-driver loop, stage_8_torch.py:1004-1013): H2D of the
 * images and D2H of the detections happen inside (pipelined against compute when the host buffers are
 * pinned); returns after the results are in dets_host / counts_host. */
int ayq_forward_host(ayq_handle h, const float* img_host, int n, float* dets_host, int32_t* counts_host);
/* Same, images given as uint8 (n,3,640,640) CHW: ToTensor (u8 / 255, stage_8_torch.py:985-990) runs on the GPU. */
int ayq_forward_host_u8(ayq_handle h, const uint8_t* img_host, int n, float* dets_host, int32_t* counts_host);
/* Asynchronous form of the two calls above (the validation loop of stage_8_torch.py:1004-1013 with the next batch's upload
 * overlapping the current batch's kernels): enqueues the whole call on the engine's internal streams and returns at once;
 * img_host (float32 when is_u8 == 0, uint8 otherwise), dets_host and counts_host must stay valid, and the results may only be
 * read, after the next ayq_wait().  Any number of calls may be queued before a wait (each with its own host buffers); they
 * run back to back, so only the first upload and the last pass of the whole sequence are exposed. */
int ayq_forward_host_async(ayq_handle h, const void* img_host, int is_u8, int n, float* dets_host, int32_t* counts_host);
/* blocks until every queued ayq_forward_host_async call has delivered its results; returns the first device error, if any */
int ayq_wait(ayq_handle h);

/* Per-layer taps for parity tests: copies activation buffer `buf` (plan.info['bufs']) of the last pass
 * into dst as int32 NCHW (n, 16*nplanes, H, W). */
int ayq_export_buffer(ayq_handle h, int buf, int n, int32_t* dst, void* stream);
int ayq_buffer_shape(ayq_handle h, int buf, int* channels, int* height, int* width);
/* raw int32 conv accumulators (n, cout, H, W) of the last pass; only for plans compiled with taps */
int ayq_export_acc_tap(ayq_handle h, int tap, int n, int32_t* dst, void* stream);
/* Own out-of-bounds check for the activation workspace: when the engine is created with AYQ_WS_GUARD=1 in the environment, a 4 KB
 * canary zone follows every activation buffer; returns how many canary bytes have been overwritten so far (0 = no kernel wrote past
 * the end of a buffer), or a negative error. */
int ayq_check_guards(ayq_handle h);
/* number of kernels one internal pass launches (bench.py gpu_launches) */
int ayq_launches_per_pass(ayq_handle h);
/* Convolution kernel family: 2 = TMA-fed tcgen05 / TMEM implicit GEMM (conv_tma_kernel) -- the ONLY family in the product
 * library libayq.so, where any other value returns an error and a conv shape the kernel does not cover is an error when the
 * workspace is built (no fallback).  The test build libayq_test.so (same sources, -DAYQ_TEST_BUILD) additionally carries two
 * independent implementations for cross-checking: 0 = CUDA-core dp4a kernel, 1 = tcgen05 kernel fed by cp.async. */
int ayq_set_conv_impl(ayq_handle h, int impl);
/* per conv op of the last pass, which implementation ran it (2 / 1 / 0 as above, -1 not run yet); returns the number of ops
 * written (ops that are not convolutions get -2).  Parity tests assert that every conv ran on the TMA kernel. */
int ayq_get_conv_impls(ayq_handle h, int32_t* impl, int cap);
/* Launch-plan variant of every conv op.  With AYQ_AUTOTUNE=1 in the environment of ayq_create a load-time tuner times the
 * variants of every sufficiently large layer when the workspace is built and keeps the fastest (default: off, variant 0 everywhere): 0 = default (two producer -> issuer chains per pipeline, two accumulators per
 * epilogue group), 1 = one chain per pipeline, 2 = one accumulator per epilogue group; -1 = workspace not built yet, -2 = not a
 * convolution.  All variants compute the same bits; layers with fewer than two tiles per CTA are not tuned. */
int ayq_get_conv_variants(ayq_handle h, int32_t* variant, int cap);
/* kernel time accounting: when enabled, every pass records CUDA events around each op;
 * ayq_get_op_times copies per-op accumulated milliseconds and launch counts (arrays of n_ops). */
int ayq_set_profiling(ayq_handle h, int enabled);
int ayq_get_op_times(ayq_handle h, float* ms, int32_t* calls, int cap);

/* ---- quantised layer library on fp32-carried integer tensors (unit-level drop-ins) ---- */

/* requantize(), utils/rescale_coeff_torch.py:14-46: y = clamp(rsh(RN32(k[c]*x), s[c]), +-(2^(bits-1)-1)).
 * x,y device float32 (n,c,hw); k, inv2s = 2^-s device float32 (c) or (1) when per_channel == 0. */
int ayq_requantize_f32(const float* x, float* y, const float* k, const float* inv2s, int per_channel,
                       int n, int c, int hw, int bits, void* stream);
/* silu(), stage_8_torch_full_quant.py:439-452 with a precomputed coefficient table tab = float[4][c]
 * (k1, 2^-s1, k2, 2^-s2) and the (2*M+1)-entry sigmoid LUT of create_sigmoid_lookup_table (utils/silu.py:32-50). */
int ayq_silu_f32(const float* acc, float* y, const float* tab, const float* lut, int n, int c, int hw, int bits, void* stream);
/* sigmoid_quant()/exponent_quant(), utils/silu_torch.py:4-18, utils/exp_torch.py:4-18:
 * y = lut[x - key_min] if key_min <= x <= key_max and x is an integer else 0. */
int ayq_lut_f32(const float* x, float* y, const float* lut, int key_min, int key_max, size_t count, void* stream);
/* quant_matrix(), utils/quant_matrix_torch.py:57-70: per-image a = max|x|, y = rint(fl32(x * fl32(M/a))).
 * amax: device float32 (n) scratch, scales: device float32 (n) out. */
int ayq_quant_input_f32(const float* x, float* y, float* amax, float* scales, int n, size_t per_image, int bits, void* stream);
/* save_max_a(), utils/save_a.py:11-26: out[i] = max|x[i, :]| over per_image floats (calibration taps). */
int ayq_absmax_f32(const float* x, float* out, int n, size_t per_image, void* stream);
/* Calibration forward (stage_4.py:475-946): the BN-fused float network, NCHW float32 like the reference.
 * ayq_calib_conv_f32: y = Conv2d(x; w, b, kernel ks in {1,3}, stride in {1,2}, padding ks/2); when amax != NULL the tap of
 * save_max_a() (utils/save_a.py:22) is fused: amax[i] = max(amax[i], max|y[i]|) per image (the caller zeroes amax).
 * x (n,cin,H,W), w (cout,cin,ks,ks), b (cout), y (n,cout,Hout,Wout), amax (n): device float32.
 * ayq_calib_silu_f32: nn.SiLU in place.  ayq_calib_maxpool5_f32: MaxPool2d(5,1,2) on `planes` = n*c planes of H x W.
 * ayq_calib_upsample2_f32: nn.Upsample(None, 2, 'nearest'), y has planes x 2H x 2W elements. */
int ayq_calib_conv_f32(const float* x, const float* w, const float* b, float* y, float* amax, int n, int cin, int H, int W,
                       int cout, int ks, int stride, void* stream);
int ayq_calib_silu_f32(float* x, size_t count, void* stream);
int ayq_calib_maxpool5_f32(const float* x, float* y, int planes, int H, int W, void* stream);
int ayq_calib_upsample2_f32(const float* x, float* y, int planes, int H, int W, void* stream);
/* conv_quant() of stage_6_full_quant.py:89-126 without its text dumps (the producer of the stage_7 weights):
 * per output channel a = max|w|, s = fl32((2^(bits-1)-1) / a), qw = rint(fl32(w * s)) (utils/quant_matrix.py:56-78);
 * qb = trunc(double(bias) * (scale_input * s)) (utils/quant_bias.py:2-4); scale_res = scale_input * s (:93-96,:122; the first
 * layer passes scale_input = 2^(bits-1)-1, `start=True`).  w device float32 (cout, per_channel), bias device float32 (cout),
 * qw device int8 (cout, per_channel), qb device int64 (cout), scale_res device float64 (cout). */
int ayq_quant_weights_f32(const float* w, const float* bias, int cout, size_t per_channel, int bits, double scale_input,
                          int8_t* qw, int64_t* qb, double* scale_res, void* stream);
/* coord_quant() + nms_quant() + scale_boxes/clip_boxes (stage_8_torch_full_quant.py:248-423) on a
 * caller-provided prediction tensor: dbox_cls device float32 (n,84,8400). */
int ayq_nms(ayq_handle h, const float* dbox_cls, int n, float* dets, int32_t* counts, void* stream);
/* coord() of stage_8_torch.py:146-190 (float path) + scale_boxes / clip_boxes / convert_res
 * (:203-258, :949-957) on a caller-provided prediction tensor dbox_cls device float32 (n,84,8400): candidates
 * conf > 1e-8, class offsets 7680, torchvision.ops.nms semantics at IoU 0.45 (stable descending score order, fp32
 * arithmetic in the same order), first 300.  Output layout as ayq_forward. */
int ayq_coord_float(ayq_handle h, const float* dbox_cls, int n, float* dets, int32_t* counts, void* stream);
/* nms_quant(dets, scores, thresh) (stage_8_torch_full_quant.py:248-294) stand-alone: boxes device float32 (nb,4)
 * xyxy, scores device float32 (nb) holding integers in [0, 131071], nb <= 16384.  keep: device float32 (1000)
 * receives the kept indices in selection order (stable tie-break), count: device int32 (1). */
int ayq_nms_boxes(const float* boxes, const float* scores, int nb, float* keep, int32_t* count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AYQ_H */
