// ayq.cu -- engine + C ABI (include/ayq.h) of the B200-native integer YOLOv8n + q_NMS path.
//
// The engine interprets the plan blob written by alpha_yolo_quant_b200/plan.py (plan_format.h): a list of
// ops over 16-channel plane buffers.  It owns the packed weights / tables and the activation workspace
// (sized for `max_batch` images); larger batches are processed as consecutive passes.
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

#include "../../include/ayq.h"
#include "plan_format.h"
#include "kernels.cuh"
#include "conv_tc.cuh"
#include "conv_tma.cuh"

using namespace ayq;

static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return fail(-5, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct ayq_engine {
    int device = 0;
    PlanHeader hdr{};
    std::vector<BufDesc> bufs;
    std::vector<OpDesc> ops;
    std::vector<unsigned char> host_data;
    unsigned char* d_data = nullptr;       // data section on the device
    std::vector<signed char> conv_nq1;     // per conv op: launch-plan variant picked by the load-time tuner (tune_variant(); 0 = default, -1 = not tuned yet)
    bool autotune = false;                 // AYQ_AUTOTUNE=1 switches the load-time tuner on
    float* d_lutrep = nullptr;             // replicated sigmoid tables (one [257][32] + one [257][8] block per distinct table of the plan)
    std::vector<const float*> op_lutrep;   // per op: its [257][32] block (convs with the SiLU epilogue) / [257][8] block (Conv_P1), else nullptr
    int max_batch = 512;
    int cap = 0;                           // images the workspace is sized for
    unsigned char* ws = nullptr;
    size_t ws_bytes = 0;
    std::vector<size_t> buf_off;           // per buffer, for `cap`
    bool guard = false;                    // AYQ_WS_GUARD=1: a 4 KB canary zone after every activation buffer (ayq_check_guards)
    std::vector<size_t> guard_off;
    size_t off_amax = 0, off_dbox = 0, off_conf = 0, off_cls = 0, off_stage = 0;
    std::vector<KChunk*> d_kc;             // per op (device), rebuilt when cap changes
    std::vector<std::vector<KChunk>> h_kc; // per op (host copy)
    std::vector<int*> acc_taps;            // per tap device buffers (cap images)
    std::vector<size_t> acc_tap_elems;     // per image
    int conv_impl = 2;                     // 2 = TMA-fed tcgen05 (falls back per layer to 1 = cp.async-fed tcgen05, then 0 = dp4a)
    bool debug_sync = false;
    int last_n = 0;
    // host-pipeline resources
    cudaStream_t s_copy = nullptr, s_comp = nullptr, s_d2h = nullptr, s_cap = nullptr;
    float* d_img[2] = {nullptr, nullptr};
    uint8_t* d_img_u8[2] = {nullptr, nullptr};
    float* d_dets[2] = {nullptr, nullptr};
    int* d_counts[2] = {nullptr, nullptr};
    cudaEvent_t ev_h2d[2]{}, ev_done[2]{}, ev_d2h[2]{}, ev_img[2]{};   // ev_img: Conv_P1 (the last reader of the staged images) of the pass has run
    int host_cap = 0;
    bool host_u8 = false;
    long long host_passes = 0;             // passes the host pipeline has enqueued so far (slot = host_passes & 1; persists across async calls)
    // cross-entry ordering: every entry point records `ev_busy` on the stream it launched on and waits for it first, so that
    // two calls on different streams (or a device entry followed by the host pipeline) never share the workspace concurrently
    cudaEvent_t ev_busy = nullptr;
    bool busy_recorded = false;
    bool busy_from_host = false;           // ev_busy was recorded by the host pipeline itself (its streams need not wait for it)
    // profiling
    bool profiling = false;
    std::vector<float> op_ms;
    std::vector<int> op_calls;
    std::vector<cudaEvent_t> prof_ev;
    std::map<int, cudaGraphExec_t> graphs; // per pass size: captured conv / pool / head section
    bool use_graph = true;
    int fast_div = 0;                      // DFL division shortcut verified on this device (div_selfcheck_kernel)
    std::vector<int> host_tail;            // AYQ_HOST_TAIL=a,b,..: sizes the last pass of a host call is split into (sum = pass size)
    int host_pass = 0, host_ramp = 0;      // AYQ_HOST_PASS: pass size of the host pipeline (default 64); AYQ_HOST_RAMP=1: smaller passes at both ends
    int p1_chunk = 0;                      // AYQ_P1_CHUNK: images per abs-max -> Conv_P1 chunk (fp32 device input), 0 = whole pass
    bool p1_fuse = false;                  // AYQ_P1_FUSE=1: abs-max inside Conv_P1 (one HBM read of the image; measured slower, off)
    bool last_fused = false;               // the last pass ran the fused abs-max + Conv_P1 kernel (one launch fewer)
    bool p1_dp4a = false;                  // AYQ_P1_DP4A=1: keep Conv_P1 on the CUDA cores (conv_p1_fast_kernel) also when a tcgen05 conv family is selected
    bool role_prof = false;                // AYQ_ROLE_PROF=1: per-op warp-role cycle counters (conv_tma only), dumped at destroy
    long long* d_role = nullptr;
    TcState tc;                            // tcgen05 path state
    TmaState tma;                          // TMA-fed tcgen05 path: driver entry point for tensor-map encoding
    std::vector<int> conv_impl_used;       // per op: implementation that ran it in the last pass (ayq_get_conv_impls)
    std::vector<TmaLaunch> tma_cache;      // per op: tensor maps + stage plan for the last pass size
    std::vector<std::vector<TmaSeg>> tma_segs;   // per op: source buffers of the conv input
};

#define AYQ_GUARD_BYTES 4096
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// The engine owns ONE workspace: entries on different streams are ordered through ev_busy (see ayq.h "stream semantics").
static int wait_busy(ayq_engine* e, cudaStream_t st) {
    if (e->busy_recorded) CK(cudaStreamWaitEvent(st, e->ev_busy, 0));
    return 0;
}
static int mark_busy(ayq_engine* e, cudaStream_t st) {
    CK(cudaEventRecord(e->ev_busy, st));
    e->busy_recorded = true;
    e->busy_from_host = false;
    return 0;
}
static inline float f_from_bits(int32_t b) { float f; memcpy(&f, &b, 4); return f; }

extern "C" const char* ayq_last_error(void) { return g_err.c_str(); }
extern "C" int ayq_version(void) { return AYQ_PLAN_VERSION; }

static void drop_graphs(ayq_engine* e);
static void free_workspace(ayq_engine* e) {
    if (e->ws) cudaDeviceSynchronize();                            // asynchronous host calls may still be using it
    drop_graphs(e);
    if (e->ws) cudaFree(e->ws);
    e->ws = nullptr; e->ws_bytes = 0; e->cap = 0;
    for (auto p : e->d_kc) if (p) cudaFree(p);
    e->d_kc.clear();
    for (auto p : e->acc_taps) if (p) cudaFree(p);
    e->acc_taps.clear();
    tc_release(e->tc);
}

static int ensure_workspace_impl(ayq_engine* e, int n);
static void build_conv_args(ayq_engine* e, int opi, int n, ConvArgs& a);
static int prepare_tma_conv(ayq_engine* e, int opi, int n, const ConvArgs& a);
static int tune_conv(ayq_engine* e, int opi, int n, const ConvArgs& a);
// launch-plan variants the load-time tuner chooses from (all compute the same bits; they differ in how the CTA's shared memory,
// TMEM and control warps are split)
static const int AYQ_TUNE_VARIANTS = 3;
static const char* const AYQ_TUNE_NAMES[AYQ_TUNE_VARIANTS] = {"default", "one chain per pipeline", "one accumulator per epilogue group"};
// (measured and dropped from the list: weights resident up to 160 KB / streamed above 32 KB -- never faster, up to 50 % slower;
//  control warps on the highest warp ids / no halo boxes for 16-channel inputs -- within the timing noise)
static void tune_variant(TmaState& s, int v) {
    switch (v) {
    case 1: s.force_nq1 = 1; break;
    case 2: s.nbuf_mul = 1; break;
    default: break;
    }
}
static int ensure_workspace(ayq_engine* e, int n) {
    if (n <= e->cap) return 0;
    const int rc = ensure_workspace_impl(e, n);
    if (rc) free_workspace(e);                                     // never leave a half-built workspace behind (cap stays 0)
    return rc;
}
static int ensure_workspace_impl(ayq_engine* e, int n) {
    free_workspace(e);
    e->guard_off.clear();
    const int cap = n;
    size_t off = 0;
    e->buf_off.resize(e->bufs.size());
    for (size_t i = 0; i < e->bufs.size(); ++i) {
        const BufDesc& b = e->bufs[i];
        e->buf_off[i] = off;
        off = align_up(off + (size_t)b.nplanes * cap * b.H * b.W * 16 * b.elem_bytes, 1024);
        if (e->guard) { e->guard_off.push_back(off); off += AYQ_GUARD_BYTES; }
    }
    const int A = e->hdr.n_anchors;
    e->off_amax = off; off = align_up(off + sizeof(float) * cap + sizeof(unsigned) * (cap + 1), 1024);   // amax[cap] | fused Conv_P1: ticket, band counters[cap]
    e->off_dbox = off; off = align_up(off + sizeof(float4) * (size_t)cap * A, 1024);
    e->off_conf = off; off = align_up(off + sizeof(int) * (size_t)cap * A, 1024);
    e->off_cls = off;  off = align_up(off + sizeof(int) * (size_t)cap * A, 1024);
    CK(cudaMalloc(&e->ws, off));
    CK(cudaMemset(e->ws, 0, off));
    for (size_t g : e->guard_off) CK(cudaMemset(e->ws + g, 0xA5, AYQ_GUARD_BYTES));
    e->ws_bytes = off;
    e->cap = cap;
    // resolve K-chunk tables and accumulator taps
    e->d_kc.assign(e->ops.size(), nullptr);
    e->h_kc.assign(e->ops.size(), std::vector<KChunk>());
    e->tma_cache.assign(e->ops.size(), TmaLaunch());
    e->tma_segs.assign(e->ops.size(), std::vector<TmaSeg>());
    int ntaps = 0;
    for (size_t i = 0; i < e->ops.size(); ++i) {
        const int32_t* f = e->ops[i].f;
        int tap = -1; size_t elems = 0;
        if (f[0] == OP_CONV) {
            const int nkc = f[CF_NKC];
            const int32_t* src = (const int32_t*)(e->host_data.data() + f[CF_KC_OFF]);
            const int pad = f[CF_KSIZE] / 2;
            std::vector<KChunk> kc(nkc);
            std::vector<int> seg_buf;
            for (int k = 0; k < nkc; ++k) {
                const int buf = src[4 * k];
                kc[k].off = (long long)e->buf_off[buf];
                kc[k].plane = src[4 * k + 1];
                kc[k].dy = src[4 * k + 2] - pad;
                kc[k].dx = src[4 * k + 3] - pad;
                int sg = -1;
                for (size_t q = 0; q < seg_buf.size(); ++q) if (seg_buf[q] == buf) sg = (int)q;
                if (sg < 0) {
                    sg = (int)seg_buf.size();
                    seg_buf.push_back(buf);
                    TmaSeg ts; ts.base = e->ws + e->buf_off[buf]; ts.nplanes = e->bufs[buf].nplanes;
                    e->tma_segs[i].push_back(ts);
                }
                kc[k].pad_ = sg;                               // index into tma_segs[op]
            }
            CK(cudaMalloc(&e->d_kc[i], sizeof(KChunk) * nkc));
            CK(cudaMemcpy(e->d_kc[i], kc.data(), sizeof(KChunk) * nkc, cudaMemcpyHostToDevice));
            e->h_kc[i] = kc;
            tap = f[CF_ACC_TAP];
            elems = (size_t)f[CF_COUT] * f[CF_HOUT] * f[CF_WOUT];
        } else if (f[0] == OP_CONV_P1) {
            tap = f[P1_ACC_TAP];
            elems = (size_t)16 * f[P1_HOUT] * f[P1_WOUT];
        }
        if (tap >= 0) {
            if ((int)e->acc_taps.size() <= tap) { e->acc_taps.resize(tap + 1, nullptr); e->acc_tap_elems.resize(tap + 1, 0); }
            CK(cudaMalloc(&e->acc_taps[tap], elems * cap * sizeof(int)));
            e->acc_tap_elems[tap] = elems;
            ++ntaps;
        }
    }
    (void)ntaps;
    // No fallback in the conv path: every convolution of the plan must be covered by the TMA-fed tcgen05 kernel.  Checked here,
    // when the workspace (and with it the tensor maps) is built, so that an uncovered shape is a load-time error.
    if (e->conv_impl == 2) {
        e->cap = cap;
        for (size_t i = 0; i < e->ops.size(); ++i) {
            if (e->ops[i].f[0] != OP_CONV) continue;
            ConvArgs a;
            build_conv_args(e, (int)i, cap, a);
            if (e->conv_nq1.size() != e->ops.size()) e->conv_nq1.assign(e->ops.size(), -1);
            if (!prepare_tma_conv(e, (int)i, cap, a))
                return fail(-38, "conv op %zu (%s): shape not covered by the TMA-fed tcgen05 kernel (cout %d, %dx%d, stride %d, %d K chunks)", i,
                            (const char*)(e->host_data.data() + e->ops[i].f[CF_NAME_OFF]), a.cout, a.Hout, a.Wout, a.stride, a.nkc);
            int rc = tune_conv(e, (int)i, cap, a);
            if (rc) return rc;
        }
    }
    return 0;
}

extern "C" int ayq_create(const void* plan_blob, size_t nbytes, int device, ayq_handle* out) {
    if (!plan_blob || !out || nbytes < sizeof(PlanHeader)) return fail(-22, "ayq_create: bad arguments");
    const unsigned char* p = (const unsigned char*)plan_blob;
    PlanHeader h;
    memcpy(&h, p, sizeof h);
    if (h.magic != AYQ_MAGIC) return fail(-22, "ayq_create: bad magic 0x%x", h.magic);
    if (h.version != AYQ_PLAN_VERSION) return fail(-22, "ayq_create: plan version %u, library %d", h.version, AYQ_PLAN_VERSION);
    if (h.K < 2 || h.K > 8) return fail(-22, "ayq_create: unsupported bit width K=%d (2..8: activations are stored as int8)", h.K);
    if (h.data_off + h.data_bytes > nbytes || h.ops_off + (uint64_t)h.n_ops * sizeof(OpDesc) > nbytes ||
        h.bufs_off + (uint64_t)h.n_bufs * sizeof(BufDesc) > nbytes)
        return fail(-22, "ayq_create: truncated plan (%zu bytes)", nbytes);
    // ---- validate everything the kernels will index with before any of it is used (a blob from another compiler version or
    // a corrupted file must be rejected here, not read out of bounds on the device)
    if (h.img_h != 640 || h.img_w != 640 || h.n_anchors != 8400)
        return fail(-22, "ayq_create: plan is for %dx%d images / %d anchors; the Detect-head kernels are built for 640x640 / 8400", h.img_h, h.img_w, h.n_anchors);
    if (h.n_bufs < 1 || h.n_bufs > 4096 || h.n_ops < 1 || h.n_ops > 4096) return fail(-22, "ayq_create: implausible plan (%d buffers, %d ops)", h.n_bufs, h.n_ops);
    {
        const OpDesc* ops = (const OpDesc*)(p + h.ops_off);
        const BufDesc* bufs = (const BufDesc*)(p + h.bufs_off);
        const unsigned char* data = p + h.data_off;
        const uint64_t DB = h.data_bytes;
        auto in_data = [&](int64_t off, uint64_t bytes) { return off >= 0 && (uint64_t)off <= DB && bytes <= DB - (uint64_t)off; };
        auto buf_ok = [&](int b) { return b >= 0 && b < h.n_bufs; };
        for (int b = 0; b < h.n_bufs; ++b)
            if (bufs[b].nplanes < 1 || bufs[b].H < 1 || bufs[b].W < 1 || bufs[b].H > 4096 || bufs[b].W > 4096 || (bufs[b].elem_bytes != 1 && bufs[b].elem_bytes != 2 && bufs[b].elem_bytes != 4))
                return fail(-22, "ayq_create: buffer %d has a bad descriptor", b);
        for (int i = 0; i < h.n_ops; ++i) {
            const int32_t* f = ops[i].f;
            bool ok = true;
            if (f[0] == OP_CONV) {
                const int nkc = f[CF_NKC], cout = f[CF_COUT], nkp = (nkc + 1) & ~1;
                ok = nkc >= 1 && nkc <= 4096 && cout >= 16 && cout <= 256 && cout % 16 == 0 && (f[CF_KSIZE] == 1 || f[CF_KSIZE] == 3) &&
                     (f[CF_STRIDE] == 1 || f[CF_STRIDE] == 2) && f[CF_NOUT] >= 1 && f[CF_NOUT] <= AYQ_MAX_OUT && f[CF_EPI] >= 0 && f[CF_EPI] <= 2 &&
                     in_data(f[CF_KC_OFF], (uint64_t)nkc * 16) && in_data(f[CF_W_OFF], (uint64_t)nkp * cout * 16) && in_data(f[CF_BIAS_OFF], (uint64_t)cout * 4) &&
                     in_data(f[CF_TAB_OFF], (uint64_t)cout * 16) && in_data(f[CF_NAME_OFF], 1) &&
                     (f[CF_EPI] != EPI_SILU || in_data(f[CF_LUT_OFF], (uint64_t)(2 * f[CF_CLAMP] + 1) * 4)) && (f[CF_ACC_BUF] < 0 || buf_ok(f[CF_ACC_BUF]));
                if (ok) {
                    const int32_t* kc = (const int32_t*)(data + f[CF_KC_OFF]);
                    for (int k = 0; k < nkc && ok; ++k)
                        ok = buf_ok(kc[4 * k]) && kc[4 * k + 1] >= 0 && kc[4 * k + 1] < bufs[kc[4 * k]].nplanes && kc[4 * k + 2] >= 0 && kc[4 * k + 2] < f[CF_KSIZE] &&
                             kc[4 * k + 3] >= 0 && kc[4 * k + 3] < f[CF_KSIZE];
                    for (int o = 0; o < f[CF_NOUT] && ok; ++o) {
                        const int32_t* of = f + CF_OUT0 + CF_OUT_STRIDE * o;
                        ok = buf_ok(of[0]) && of[1] >= 0 && of[5] >= 0 && of[5] <= 2;
                    }
                    if (ok && memchr(data + f[CF_NAME_OFF], 0, (size_t)(DB - (uint64_t)f[CF_NAME_OFF])) == nullptr) ok = false;
                }
            } else if (f[0] == OP_CONV_P1) {
                ok = buf_ok(f[P1_OUT_BUF]) && in_data(f[P1_W_OFF], 16 * 32) && in_data(f[P1_BIAS_OFF], 64) && in_data(f[P1_TAB_OFF], 256) &&
                     in_data(f[P1_LUT_OFF], (uint64_t)(2 * f[P1_CLAMP] + 1) * 4);
            } else if (f[0] == OP_POOL) {
                ok = buf_ok(f[PL_IN_BUF]) && buf_ok(f[PL_OUT_BUF]) && f[PL_NPLANES] >= 1 && f[PL_H] >= 1 && f[PL_W] >= 1;
            } else if (f[0] == OP_HEAD) {
                for (int l = 0; l < 3; ++l) ok = ok && buf_ok(f[HD_BOX_BUF0 + l]) && buf_ok(f[HD_CLS_BUF0 + l]);
                ok = ok && in_data(f[HD_LUT_EXP_OFF], (uint64_t)(1u << h.K) * 4) && in_data(f[HD_LUT16_OFF], 65535 * 2) && in_data(f[HD_LO16_OFF], 65535 * 2) &&
                     in_data(f[HD_DFLW_OFF], 64) && in_data(f[HD_ANCH_OFF], (uint64_t)h.n_anchors * 8);
            } else if (f[0] == OP_HEAD_FLOAT) {
                for (int l = 0; l < 3; ++l) ok = ok && buf_ok(f[HF_BOX_BUF0 + l]) && buf_ok(f[HF_CLS_BUF0 + l]);
                ok = ok && in_data(f[HF_BOX_SCALE_OFF], 3 * 64 * 4) && in_data(f[HF_CLS_SCALE_OFF], 3 * 80 * 4) && in_data(f[HF_DFLW_OFF], 64);
            } else if (f[0] != OP_NMS && f[0] != OP_NMS_FLOAT) ok = false;
            if (!ok) return fail(-22, "ayq_create: plan op %d (kind %d) references data outside the blob or an unknown buffer", i, f[0]);
        }
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev <= 0)
        return fail(-19, "ayq_create: no CUDA device (%s); this engine has no CPU fallback", cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) return fail(-22, "ayq_create: device %d out of range (%d devices)", device, ndev);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(-19, "ayq_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    ayq_engine* e = new ayq_engine();
    struct Guard { ayq_engine* e; ~Guard() { if (e) { if (e->d_data) cudaFree(e->d_data); if (e->d_lutrep) cudaFree(e->d_lutrep); if (e->ev_busy) cudaEventDestroy(e->ev_busy); if (e->d_role) cudaFree(e->d_role); delete e; } } } guard{e};
    e->device = device;
    e->hdr = h;
    e->bufs.resize(h.n_bufs);
    memcpy(e->bufs.data(), p + h.bufs_off, sizeof(BufDesc) * h.n_bufs);
    e->ops.resize(h.n_ops);
    memcpy(e->ops.data(), p + h.ops_off, sizeof(OpDesc) * h.n_ops);
    e->host_data.assign(p + h.data_off, p + h.data_off + h.data_bytes);
    if (cudaMalloc(&e->d_data, h.data_bytes) != cudaSuccess ||
        cudaMemcpy(e->d_data, e->host_data.data(), h.data_bytes, cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(-12, "ayq_create: cannot upload %llu bytes of plan data", (unsigned long long)h.data_bytes);
    CK(cudaEventCreateWithFlags(&e->ev_busy, cudaEventDisableTiming));
    {   // replicated sigmoid tables: entry i of a block = table[clamp(i - 128, -M, M) + M], 32 (convs) or 8 (Conv_P1) copies each
        std::map<std::pair<int, int>, size_t> seen;               // (table offset, M) -> float offset of its two blocks
        std::vector<float> rep;
        const size_t blk = (size_t)AYQ_LUTREP_N * 32 + (size_t)AYQ_LUTREP_N * 8;
        e->op_lutrep.assign(h.n_ops, nullptr);
        std::vector<size_t> op_off(h.n_ops, (size_t)-1);
        for (uint32_t i = 0; i < h.n_ops; ++i) {
            const int32_t* f = e->ops[i].f;
            int off, M;
            if (f[0] == OP_CONV && f[CF_EPI] == 0) { off = f[CF_LUT_OFF]; M = f[CF_CLAMP]; }
            else if (f[0] == OP_CONV_P1) { off = f[P1_LUT_OFF]; M = f[P1_CLAMP]; }
            else continue;
            if (M < 1 || M > 127) continue;
            auto key = std::make_pair(off, M);
            auto it = seen.find(key);
            if (it == seen.end()) {
                const float* t = (const float*)(e->host_data.data() + off);
                const size_t base = rep.size();
                rep.resize(base + blk);
                for (int j = 0; j < AYQ_LUTREP_N; ++j) {
                    const int r = std::max(-M, std::min(M, j - 128));
                    for (int c = 0; c < 32; ++c) rep[base + (size_t)j * 32 + c] = t[r + M];
                    for (int c = 0; c < 8; ++c) rep[base + (size_t)AYQ_LUTREP_N * 32 + (size_t)j * 8 + c] = t[r + M];
                }
                it = seen.emplace(key, base).first;
            }
            op_off[i] = it->second + (f[0] == OP_CONV_P1 ? (size_t)AYQ_LUTREP_N * 32 : 0);
        }
        if (!rep.empty()) {
            if (cudaMalloc(&e->d_lutrep, rep.size() * sizeof(float)) != cudaSuccess ||
                cudaMemcpy(e->d_lutrep, rep.data(), rep.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
                return fail(-12, "ayq_create: cannot upload the replicated sigmoid tables");
            for (uint32_t i = 0; i < h.n_ops; ++i) if (op_off[i] != (size_t)-1) e->op_lutrep[i] = e->d_lutrep + op_off[i];
        }
    }
#ifdef AYQ_TEST_BUILD
    CK(cudaFuncSetAttribute(conv_dp4a_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(conv_dp4a_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
#endif
    CK(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NMS_SMEM));
    CK(cudaFuncSetAttribute(nms_float_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nmsf_smem_bytes(h.n_anchors)));
    e->debug_sync = getenv("AYQ_DEBUG_SYNC") != nullptr;
    e->use_graph = getenv("AYQ_NO_GRAPH") == nullptr;
    e->role_prof = getenv("AYQ_ROLE_PROF") != nullptr;
#ifndef AYQ_ROLE_PROF_BUILD
    if (e->role_prof) { return fail(-22, "AYQ_ROLE_PROF=1 needs the profiling build of the library (libayq_prof.so: python -m alpha_yolo_quant_b200.build --prof)"); }
#endif
    e->p1_dp4a = getenv("AYQ_P1_DP4A") != nullptr;
    e->guard = getenv("AYQ_WS_GUARD") != nullptr;
    e->autotune = getenv("AYQ_AUTOTUNE") && atoi(getenv("AYQ_AUTOTUNE")) != 0;   // opt-in: see tune_conv()
    e->p1_fuse = getenv("AYQ_P1_FUSE") != nullptr;           // off: measured slower than the two kernels (843 vs 184 + 439 us per 256 images)
    if (const char* ev = getenv("AYQ_P1_CHUNK")) e->p1_chunk = atoi(ev);
    if (const char* ev = getenv("AYQ_HOST_PASS")) e->host_pass = atoi(ev);
    if (const char* ev = getenv("AYQ_HOST_RAMP")) e->host_ramp = atoi(ev);
    if (const char* ev = getenv("AYQ_HOST_TAIL")) {
        for (const char* q = ev; *q;) { e->host_tail.push_back(atoi(q)); while (*q && *q != ',') ++q; if (*q == ',') ++q; }
    }
    if (e->role_prof) {
        e->use_graph = false;
        CK(cudaMalloc(&e->d_role, sizeof(long long) * h.n_ops * 148 * AYQ_DBG_SLOTS));
        CK(cudaMemset(e->d_role, 0, sizeof(long long) * h.n_ops * 148 * AYQ_DBG_SLOTS));
    }
    g_pdl = getenv("AYQ_NO_PDL") == nullptr ? 1 : 0;
    tc_init(e->tc);
    tma_init(e->tma);
    {   // exhaustive check of the DFL division shortcut: e in [0, 127], S in [1, 16 * 127]
        int* d_bad = nullptr;
        const int emax = 127, smax = 16 * 127;
        if (cudaMalloc(&d_bad, sizeof(int)) == cudaSuccess) {
            cudaMemset(d_bad, 0, sizeof(int));
            div_selfcheck_kernel<<<((emax + 1) * smax + 255) / 256, 256>>>(emax, smax, d_bad);
            int bad = 1;
            if (cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess) e->fast_div = bad == 0 ? 1 : 0;
            cudaFree(d_bad);
        }
        if (getenv("AYQ_NO_FAST_DIV")) e->fast_div = 0;
    }
    e->op_ms.assign(h.n_ops + 1, 0.f);
    e->op_calls.assign(h.n_ops + 1, 0);
    e->conv_impl_used.assign(h.n_ops, -1);
    guard.e = nullptr;                                             // success: the caller owns the engine now
    *out = e;
    return 0;
}

extern "C" int ayq_destroy(ayq_handle e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    if (e->role_prof && e->d_role) {
        cudaDeviceSynchronize();
        std::vector<long long> h((size_t)e->ops.size() * 148 * AYQ_DBG_SLOTS);
        cudaMemcpy(h.data(), e->d_role, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        fprintf(stderr, "role profile of the last pass (kilo-cycles, mean over CTAs): op name | prod0 total/wait_empty | prod1 | mma0 total/wait_tempty/wait_full | mma1 | epi0 total/wait_tfull | epi1\n");
        for (size_t i = 0; i < e->ops.size(); ++i) {
            if (e->ops[i].f[0] != OP_CONV) continue;
            double m[16] = {0};
            int cnt = 0;
            for (int b = 0; b < 148; ++b) {
                const long long* r = &h[(i * 148 + b) * AYQ_DBG_SLOTS];
                if (r[6] == 0) continue;
                ++cnt;
                for (int k = 0; k < 16; ++k) m[k] += (double)r[k];
            }
            if (!cnt) continue;
            const char* nm = (const char*)(e->host_data.data() + e->ops[i].f[CF_NAME_OFF]);
            fprintf(stderr, "%3zu %-20s | %6.1f/%6.1f | %6.1f/%6.1f | %6.1f/%6.1f/%6.1f | %6.1f/%6.1f/%6.1f | %6.1f/%6.1f | %6.1f/%6.1f\n", i, nm,
                    m[0] / cnt / 1e3, m[1] / cnt / 1e3, m[2] / cnt / 1e3, m[3] / cnt / 1e3,
                    m[6] / cnt / 1e3, m[7] / cnt / 1e3, m[8] / cnt / 1e3, m[9] / cnt / 1e3, m[10] / cnt / 1e3, m[11] / cnt / 1e3,
                    m[12] / cnt / 1e3, m[13] / cnt / 1e3, m[14] / cnt / 1e3, m[15] / cnt / 1e3);
        }
        // kernel-to-kernel timeline from the globaltimer stamps (us): how the fixed cost of a launch splits up
        fprintf(stderr, "timeline of the last pass (us): op | span = last exit - first entry | prologue (mean) | dependency wait (mean) | body = wait done -> exit (mean) | "
                        "exit spread: last - mean, last - first | gap: this op's first 'wait done' - previous conv's last exit\n");
        long long prev_last_exit = 0, prev_first_exit = 0;
        double sum_span = 0, sum_body = 0, sum_tail = 0, sum_gap = 0;
        for (size_t i = 0; i < e->ops.size(); ++i) {
            if (e->ops[i].f[0] != OP_CONV) continue;
            long long first_in = 0, last_out = 0, first_out = 0, first_go = 0;
            double pro = 0, wt = 0, body = 0, mean_out = 0;
            int cnt = 0;
            for (int b = 0; b < 148; ++b) {
                const long long* r = &h[(i * 148 + b) * AYQ_DBG_SLOTS];
                if (r[16] == 0 || r[19] == 0) continue;
                if (!cnt || r[16] < first_in) first_in = r[16];
                if (!cnt || r[19] > last_out) last_out = r[19];
                if (!cnt || r[19] < first_out) first_out = r[19];
                if (!cnt || r[18] < first_go) first_go = r[18];
                pro += (double)(r[17] - r[16]); wt += (double)(r[18] - r[17]); body += (double)(r[19] - r[18]);
                ++cnt;
            }
            if (!cnt) continue;
            for (int b = 0; b < 148; ++b) { const long long* r = &h[(i * 148 + b) * AYQ_DBG_SLOTS]; if (r[16] && r[19]) mean_out += (double)(r[19] - first_in); }
            mean_out /= cnt;
            const char* nm = (const char*)(e->host_data.data() + e->ops[i].f[CF_NAME_OFF]);
            const double gap = prev_last_exit ? (double)(first_go - prev_last_exit) / 1e3 : 0.0;
            fprintf(stderr, "%3zu %-20s | %7.2f | %6.2f | %6.2f | %7.2f | %6.2f %6.2f | %6.2f | entry-prev first exit %6.2f, entry-prev last exit %6.2f\n", i, nm, (double)(last_out - first_in) / 1e3, pro / cnt / 1e3, wt / cnt / 1e3,
                    body / cnt / 1e3, ((double)(last_out - first_in) - mean_out) / 1e3, (double)(last_out - first_out) / 1e3, gap,
                    prev_last_exit ? (double)(first_in - prev_first_exit) / 1e3 : 0.0, prev_last_exit ? (double)(first_in - prev_last_exit) / 1e3 : 0.0);
            sum_span += (double)(last_out - first_in) / 1e3; sum_body += body / cnt / 1e3; sum_tail += ((double)(last_out - first_in) - mean_out) / 1e3;
            if (prev_last_exit && gap < 100.0) sum_gap += gap;
            prev_last_exit = last_out; prev_first_exit = first_out;
        }
        if (const char* path = getenv("AYQ_TIMELINE_RAW")) {          // every CTA's four stamps (ns), one line per (op, CTA)
            if (FILE* fp = fopen(path, "w")) {
                for (size_t i = 0; i < e->ops.size(); ++i)
                    for (int b = 0; b < 148; ++b) {
                        const long long* r = &h[(i * 148 + b) * AYQ_DBG_SLOTS];
                        if (r[16]) fprintf(fp, "%zu %d %lld %lld %lld %lld %lld\n", i, b, r[16], r[17], r[18], r[19], r[20]);
                    }
                fclose(fp);
            }
        }
        fprintf(stderr, "timeline sums (us): span %.1f, mean body %.1f, exit spread (last - mean) %.1f, gaps %.1f\n", sum_span, sum_body, sum_tail, sum_gap);
        cudaFree(e->d_role);
    }
    free_workspace(e);
    if (e->d_data) cudaFree(e->d_data);
    if (e->d_lutrep) cudaFree(e->d_lutrep);
    for (int i = 0; i < 2; ++i) {
        if (e->d_img[i]) cudaFree(e->d_img[i]);
        if (e->d_img_u8[i]) cudaFree(e->d_img_u8[i]);
        if (e->d_dets[i]) cudaFree(e->d_dets[i]);
        if (e->d_counts[i]) cudaFree(e->d_counts[i]);
        if (e->ev_h2d[i]) cudaEventDestroy(e->ev_h2d[i]);
        if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]);
        if (e->ev_d2h[i]) cudaEventDestroy(e->ev_d2h[i]);
        if (e->ev_img[i]) cudaEventDestroy(e->ev_img[i]);
    }
    if (e->s_copy) cudaStreamDestroy(e->s_copy);
    if (e->s_comp) cudaStreamDestroy(e->s_comp);
    if (e->s_d2h) cudaStreamDestroy(e->s_d2h);
    if (e->s_cap) cudaStreamDestroy(e->s_cap);
    for (auto ev : e->prof_ev) cudaEventDestroy(ev);
    if (e->ev_busy) cudaEventDestroy(e->ev_busy);
    delete e;
    return 0;
}

extern "C" int ayq_set_max_batch(ayq_handle e, int max_batch) {
    if (!e || max_batch < 1 || max_batch > 4096) return fail(-22, "ayq_set_max_batch: 1..4096");
    e->max_batch = max_batch;
    return 0;
}
extern "C" size_t ayq_workspace_bytes(ayq_handle e) { return e ? e->ws_bytes : 0; }
extern "C" int ayq_set_conv_impl(ayq_handle e, int impl) {
    if (!e || impl < 0 || impl > 2) return fail(-22, "ayq_set_conv_impl: 0 (dp4a), 1 (tcgen05, cp.async feed) or 2 (tcgen05, TMA feed)");
#ifndef AYQ_TEST_BUILD
    if (impl != 2) return fail(-38, "ayq_set_conv_impl: this is the product library, which contains the TMA-fed tcgen05 convolution only; "
                                    "the dp4a / cp.async cross-check kernels live in the test build (libayq_test.so)");
#endif
    if (e->conv_impl != impl) drop_graphs(e);
    e->conv_impl = impl;
    return 0;
}
extern "C" int ayq_set_profiling(ayq_handle e, int enabled) {
    if (!e) return fail(-22, "null handle");
    e->profiling = enabled != 0;
    if (e->profiling && e->prof_ev.empty()) {
        e->prof_ev.resize(e->ops.size() + 2);
        for (auto& ev : e->prof_ev) if (cudaEventCreate(&ev) != cudaSuccess) return fail(-5, "cudaEventCreate");
    }
    std::fill(e->op_ms.begin(), e->op_ms.end(), 0.f);
    std::fill(e->op_calls.begin(), e->op_calls.end(), 0);
    return 0;
}
extern "C" int ayq_get_op_times(ayq_handle e, float* ms, int32_t* calls, int cap) {
    if (!e) return fail(-22, "null handle");
    const int n = (int)e->op_ms.size();
    for (int i = 0; i < n && i < cap; ++i) { ms[i] = e->op_ms[i]; calls[i] = e->op_calls[i]; }
    return n;
}
extern "C" int ayq_get_conv_impls(ayq_handle e, int32_t* impl, int cap) {
    if (!e || !impl) return fail(-22, "ayq_get_conv_impls: bad arguments");
    const int n = (int)e->ops.size();
    for (int i = 0; i < n && i < cap; ++i)
        impl[i] = e->ops[i].f[0] == OP_CONV ? (i < (int)e->conv_impl_used.size() ? e->conv_impl_used[i] : -1) : -2;
    return n < cap ? n : cap;
}
extern "C" int ayq_get_conv_variants(ayq_handle e, int32_t* variant, int cap) {
    if (!e || !variant) return fail(-22, "ayq_get_conv_variants: bad arguments");
    const int n = (int)e->ops.size();
    for (int i = 0; i < n && i < cap; ++i)
        variant[i] = e->ops[i].f[0] == OP_CONV ? (i < (int)e->conv_nq1.size() ? (int)e->conv_nq1[i] : -1) : -2;
    return n < cap ? n : cap;
}
// Own bounds check (compute-sanitizer is closed on the GPU pool this was developed on): with AYQ_WS_GUARD=1 every activation buffer
// of the workspace is followed by a 4 KB canary zone; returns the number of canary bytes any kernel has overwritten (0 = clean).
extern "C" int ayq_check_guards(ayq_handle e) {
    if (!e) return fail(-22, "ayq_check_guards: null handle");
    if (!e->guard) return fail(-22, "ayq_check_guards: create the engine with AYQ_WS_GUARD=1");
    CK(cudaSetDevice(e->device));
    CK(cudaDeviceSynchronize());
    std::vector<unsigned char> h(AYQ_GUARD_BYTES);
    int bad = 0;
    for (size_t g : e->guard_off) {
        CK(cudaMemcpy(h.data(), e->ws + g, AYQ_GUARD_BYTES, cudaMemcpyDeviceToHost));
        for (unsigned char c : h) bad += c != 0xA5;
    }
    return bad;
}
extern "C" int ayq_launches_per_pass(ayq_handle e) {
    if (!e) return fail(-22, "null handle");
    return (int)e->ops.size() + (e->last_fused ? 0 : 1);   // + the abs-max reduction unless it is fused into Conv_P1; the memset node is not a kernel
}

// ---- one pass over n <= cap images ---------------------------------------------------------------------
static void build_conv_args(ayq_engine* e, int opi, int n, ConvArgs& a) {
    const int32_t* f = e->ops[opi].f;
    memset(&a, 0, sizeof a);
    a.kc = e->d_kc[opi];
    a.nkc = f[CF_NKC];
    a.ws = (const int8_t*)e->ws;
    a.in_plane_bytes = (size_t)n * f[CF_HIN] * f[CF_WIN] * 16;
    a.w = (const int8_t*)(e->d_data + f[CF_W_OFF]);
    a.bias = (const int*)(e->d_data + f[CF_BIAS_OFF]);
    a.tab = (const float*)(e->d_data + f[CF_TAB_OFF]);
    a.lut = (const float*)(e->d_data + f[CF_LUT_OFF]);
    a.lut_rep = e->op_lutrep[opi];
    a.n = n; a.Hin = f[CF_HIN]; a.Win = f[CF_WIN]; a.Hout = f[CF_HOUT]; a.Wout = f[CF_WOUT];
    a.stride = f[CF_STRIDE]; a.cout = f[CF_COUT]; a.epi = f[CF_EPI]; a.M = f[CF_CLAMP];
    a.nout = f[CF_NOUT];
    for (int o = 0; o < a.nout; ++o) {
        const int32_t* of = f + CF_OUT0 + CF_OUT_STRIDE * o;
        const BufDesc& b = e->bufs[of[0]];
        a.out[o].base = e->ws + e->buf_off[of[0]] + (size_t)of[1] * n * b.H * b.W * 16 * b.elem_bytes;
        a.out[o].mode = of[2];
        a.out[o].k = f_from_bits(of[3]);
        a.out[o].inv = f_from_bits(of[4]);
        a.out[o].up = of[5];
    }
    a.acc_tap = f[CF_ACC_TAP] >= 0 ? e->acc_taps[f[CF_ACC_TAP]] : nullptr;
    if (f[CF_ACC_BUF] >= 0) a.acc_tap = (int*)(e->ws + e->buf_off[f[CF_ACC_BUF]]);   // raw accumulators feed the float head
    a.half = 0.5f;
    a.dbg = e->role_prof ? e->d_role + (size_t)opi * 148 * AYQ_DBG_SLOTS : nullptr;
    a.dbg_mode = getenv("AYQ_EPI_SKIP") ? atoi(getenv("AYQ_EPI_SKIP")) : 0;
}
// tensor maps + stage plan of conv op `opi` for passes of n images (cached); returns L.ok
static int prepare_tma_conv(ayq_engine* e, int opi, int n, const ConvArgs& a) {
    const int32_t* f = e->ops[opi].f;
    TmaLaunch& L = e->tma_cache[opi];
    const float* h_tab = (const float*)(e->host_data.data() + f[CF_TAB_OFF]);
    const int* h_bias = (const int*)(e->host_data.data() + f[CF_BIAS_OFF]);
    if (L.n != n) {
        TmaState ts = e->tma;
        if ((size_t)opi < e->conv_nq1.size() && e->conv_nq1[opi] > 0) tune_variant(ts, e->conv_nq1[opi]);
        tma_prepare(ts, L, a, e->h_kc[opi].data(), e->tma_segs[opi].data(), (int)e->tma_segs[opi].size(), h_tab, h_bias,
                    (const float*)(e->host_data.data() + f[CF_LUT_OFF]), (const int8_t*)(e->host_data.data() + f[CF_W_OFF]));
        if (getenv("AYQ_PLAN_DUMP"))
            fprintf(stderr, "plan %-20s n=%d ok=%d cout=%3d %3dx%-3d s%d nkc=%3d | %s fast=%d gen=%d resB=%d NS=%2d slot=%5dB nbuf=%d tiles=%d smem=%zuK grid=%u\n",
                    (const char*)(e->host_data.data() + f[CF_NAME_OFF]), n, L.ok, a.cout, a.Hout, a.Wout, a.stride, a.nkc,
                    L.pl.halo ? "halo " : "boxes", L.fast, L.gen_outs, L.tp.resident_b, L.tp.NS, L.pl.a_slot_bytes, L.tp.nbuf, L.tp.ntiles, L.smem / 1024, L.grid);
    }
    return L.ok;
}
// Load-time tuner (runs when the workspace is built, never inside a pass).  The default launch plan of a layer (two producer ->
// issuer chains per pipeline, two accumulators per epilogue group, weights resident up to 96 KB, ...) is the best on average;
// single layers measure a few us faster with another split of the CTA's resources (e.g. ONE chain with the whole ring on the
// 20x20 layers with streamed weights and on some halo layers: fewer, deeper rings balance the few tiles a CTA gets).  Every
// variant computes the same bits; each sufficiently large layer is timed with all of them on whatever the workspace holds (best
// of three pairs of launches, 2 % hysteresis in favour of the default) and the choice is kept for every pass size of the engine.
// OPT-IN (AYQ_AUTOTUNE=1): measured +0.9 % on short runs (3.707 vs 3.74 ms per 256 images over 10 passes), nothing significant once
// the power cap sets the clocks (50 passes: 3.73-3.77 vs 3.76-3.78 ms), and the picks vary a little from run to run -- not worth a
// non-deterministic launch plan by default.
static int tune_conv(ayq_engine* e, int opi, int n, const ConvArgs& a) {
    TmaLaunch& L = e->tma_cache[opi];
    if (e->conv_nq1[opi] != -1) return 0;
    e->conv_nq1[opi] = 0;
    if (!e->autotune || e->role_prof) return 0;
    if (!L.ok || L.tp.ntiles < 2 * (int)L.grid) return 0;         // too small to matter
    const int32_t* f = e->ops[opi].f;
    std::vector<TmaLaunch> cand(AYQ_TUNE_VARIANTS);
    std::vector<int> live;
    cand[0] = L; live.push_back(0);
    for (int v = 1; v < AYQ_TUNE_VARIANTS; ++v) {
        TmaState ts = e->tma;
        tune_variant(ts, v);
        tma_prepare(ts, cand[v], a, e->h_kc[opi].data(), e->tma_segs[opi].data(), (int)e->tma_segs[opi].size(), (const float*)(e->host_data.data() + f[CF_TAB_OFF]),
                    (const int*)(e->host_data.data() + f[CF_BIAS_OFF]), (const float*)(e->host_data.data() + f[CF_LUT_OFF]), (const int8_t*)(e->host_data.data() + f[CF_W_OFF]));
        if (!cand[v].ok) continue;
        const tc::TcParams &x = cand[v].tp, &y = L.tp;            // identical to the default plan: nothing to measure
        if (x.NS == y.NS && x.KS == y.KS && x.nbuf == y.nbuf && x.nq == y.nq && x.resident_b == y.resident_b && x.role_hi == y.role_hi &&
            cand[v].pl.halo == L.pl.halo && cand[v].smem == L.smem) continue;
        live.push_back(v);
    }
    if (live.size() < 2) return 0;
    cudaEvent_t ev0, ev1;
    CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
    std::vector<float> best(AYQ_TUNE_VARIANTS, 1e30f);
    for (int rep = 0; rep < 4; ++rep)                             // rep 0: warm-up
        for (int v : live) {
            CK(cudaEventRecord(ev0, 0));
            for (int k = 0; k < 2; ++k)
                if (tma_launch(cand[v], a, 0) != 0) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); return fail(-5, "tuner: conv launch failed for op %d (variant %d)", opi, v); }
            CK(cudaEventRecord(ev1, 0));
            CK(cudaEventSynchronize(ev1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ev0, ev1));
            if (rep > 0 && ms < best[v]) best[v] = ms;
        }
    CK(cudaEventDestroy(ev0)); CK(cudaEventDestroy(ev1));
    int pick = 0;
    for (int v : live) if (v && best[v] < 0.98f * best[0] && best[v] < best[pick]) pick = v;
    if (pick) { e->conv_nq1[opi] = (signed char)pick; L = cand[pick]; }
    if (getenv("AYQ_PLAN_DUMP")) {
        fprintf(stderr, "tune %-20s", (const char*)(e->host_data.data() + f[CF_NAME_OFF]));
        for (int v : live) fprintf(stderr, " v%d %.1f", v, best[v] * 500.f);
        fprintf(stderr, " us -> %s\n", AYQ_TUNE_NAMES[pick]);
    }
    return 0;
}
static int launch_conv(ayq_engine* e, int opi, int n, cudaStream_t st) {
    const int32_t* f = e->ops[opi].f;
    ConvArgs a;
    build_conv_args(e, opi, n, a);
    if (e->conv_impl == 2) {
        prepare_tma_conv(e, opi, n, a);
        TmaLaunch& L = e->tma_cache[opi];
        if (L.ok) {
            if (tma_launch(L, a, st) == 0) { e->conv_impl_used[opi] = 2; return 0; }
            return fail(-5, "TMA conv launch failed for op %d: %s", opi, cudaGetErrorString(cudaGetLastError()));
        }
        // no silent fallback: a shape the TMA kernel does not cover is an error (ensure_workspace checks every conv op up front)
        return fail(-38, "conv op %d (%s): shape not covered by the TMA-fed tcgen05 kernel", opi, (const char*)(e->host_data.data() + f[CF_NAME_OFF]));
    }
#ifndef AYQ_TEST_BUILD
    return fail(-38, "conv op %d: no kernel family selected", opi);
#else
    if (e->conv_impl >= 1) {
        int rc = tc_launch_conv(e->tc, a, e->h_kc[opi].data(), (const float*)(e->host_data.data() + f[CF_TAB_OFF]), (const int*)(e->host_data.data() + f[CF_BIAS_OFF]), st);
        if (rc == 0) { e->conv_impl_used[opi] = 1; return 0; }
        if (rc != 1) return fail(-5, "tcgen05 conv launch failed for op %d: %s", opi, cudaGetErrorString(cudaGetLastError()));
        // rc == 1: shape not covered by the tcgen05 kernel -> CUDA-core kernel below
    }
    e->conv_impl_used[opi] = 0;
    const size_t npix = (size_t)n * a.Hout * a.Wout;
    const unsigned gx = (unsigned)((npix + 127) / 128);
    const size_t lut_bytes = a.epi == 0 ? (size_t)AYQ_LUT256 * 4 : 0;       // the sigmoid table is only read by the SiLU epilogue
    if (a.cout % 32 == 0) {
        CK(launch_k(conv_dp4a_kernel<32>, dim3(gx, a.cout / 32), dim3(128), (size_t)a.nkc * 32 * 16 + lut_bytes, st, a));
    } else {
        CK(launch_k(conv_dp4a_kernel<16>, dim3(gx, a.cout / 16), dim3(128), (size_t)a.nkc * 16 * 16 + lut_bytes, st, a));
    }
    return 0;
#endif
}

struct PassArgs { const float* img; const uint8_t* img_u8; int n; float* dbox_cls; float* dets; int32_t* counts; int p1_img0 = 0; int p1_cnt = -1; bool p1_fused = false;
                  cudaEvent_t ev_img = nullptr; };   // recorded right after Conv_P1, the last kernel that reads the input images

// Conv_P1 takes the lean tensor-core kernel (conv_p1_tc_kernel, MAGIC epilogue) when K = 8, no accumulator tap is asked for, the
// geometry is the 640 -> 320 one it is written for and the host can prove the epilogue's range conditions from the plan.
static bool p1_lean(ayq_engine* e, const int32_t* f, const PassArgs& pa) {
    const int H = e->hdr.img_h, W = e->hdr.img_w, n = pa.n;
    const int Hout = f[P1_HOUT], Wout = f[P1_WOUT], M = f[P1_CLAMP];
    const int8_t* hw = (const int8_t*)(e->host_data.data() + f[P1_W_OFF]);
    const float* ht = (const float*)(e->host_data.data() + f[P1_TAB_OFF]);
    const int* hb = (const int*)(e->host_data.data() + f[P1_BIAS_OFF]);
    long long sw[16];
    for (int co = 0; co < 16; ++co) { sw[co] = 0; for (int k = 0; k < 27; ++k) sw[co] += hw[co * 32 + k] < 0 ? -hw[co * 32 + k] : hw[co * 32 + k]; }
    bool range_ok = M <= 127;                                       // K != 8: explicit clamps, only the magic int -> float range has to hold
    for (int co = 0; co < 16 && range_ok; ++co) range_ok = (hb[co] < 0 ? -(long long)hb[co] : (long long)hb[co]) + (long long)M * sw[co] < (1ll << 22) - 1;
    return f[P1_ACC_TAP] < 0 && H == 2 * Hout && W == 2 * Wout && Wout % P1_TW == 0 && Hout % P1_TH == 0 && W % 4 == 0 &&
           (unsigned long long)n * Hout * Wout < (1ull << 28) &&
           (pa.img_u8 ? ((uintptr_t)pa.img_u8 & 15) == 0 || !pa.p1_fused && ((uintptr_t)pa.img_u8 & 3) == 0 : ((uintptr_t)pa.img & 15) == 0) &&
           (M == 127 ? magic_coeffs_ok(16, M, ht, hb, (const float*)(e->host_data.data() + f[P1_LUT_OFF]), sw) : range_ok);
}

// launch plan op i of a pass
static int launch_op(ayq_engine* e, size_t i, const PassArgs& pa, cudaStream_t st) {
    const int H = e->hdr.img_h, W = e->hdr.img_w, A = e->hdr.n_anchors;
    const int n = pa.n;
    const float* img = pa.img;
    float* dbox_cls = pa.dbox_cls; float* dets = pa.dets; int32_t* counts = pa.counts;
    float* amax = (float*)(e->ws + e->off_amax);
    float4* dbox = (float4*)(e->ws + e->off_dbox);
    int* conf = (int*)(e->ws + e->off_conf);
    int* cls = (int*)(e->ws + e->off_cls);
    const int32_t* f = e->ops[i].f;
    switch (f[0]) {
    case OP_CONV_P1: {
        P1Args a;
        a.img = img; a.img_u8 = pa.img_u8; a.amax = amax;
        a.lut = (const float*)(e->d_data + f[P1_LUT_OFF]);
        a.lut_rep8 = e->op_lutrep[i];
        a.n = n; a.H = H; a.W = W; a.Hout = f[P1_HOUT]; a.Wout = f[P1_WOUT]; a.M = f[P1_CLAMP];
        a.out = (int8_t*)(e->ws + e->buf_off[f[P1_OUT_BUF]]);
        a.acc_tap = f[P1_ACC_TAP] >= 0 ? e->acc_taps[f[P1_ACC_TAP]] : nullptr;
        a.half = 0.5f;
        a.ps = f[P1_OUT_PS];
        a.img0 = pa.p1_img0;
        const int nz = pa.p1_cnt >= 0 ? pa.p1_cnt : n;                // images of this launch (chunked abs-max -> Conv_P1 interleave)
        const bool lean = p1_lean(e, f, pa);
#ifdef AYQ_TEST_BUILD
        const bool lean_ok = lean && (f[P1_CLAMP] == 127 || (e->conv_impl >= 1 && !e->p1_dp4a));   // the CUDA-core cross-check kernel is K = 8 only
#else
        const bool lean_ok = lean;
#endif
        const bool fold = a.M == 127 || lean_ok;                   // folded coefficients k * 2^-s (exact): K = 8, and every lean launch
        P1Const pc;
        const int8_t* hw = (const int8_t*)(e->host_data.data() + f[P1_W_OFF]);         // [16][32], k = (ky*3+kx)*3 + c
        const float* ht = (const float*)(e->host_data.data() + f[P1_TAB_OFF]);         // [4][16]
        const int* hb = (const int*)(e->host_data.data() + f[P1_BIAS_OFF]);
        for (int tap = 0; tap < 9; ++tap)
            for (int co = 0; co < 16; ++co) {
                const int8_t* w = hw + co * 32 + tap * 3;
                pc.w4[tap][co] = (unsigned)(uint8_t)w[0] | ((unsigned)(uint8_t)w[1] << 8) | ((unsigned)(uint8_t)w[2] << 16);
            }
        for (int co = 0; co < 16; ++co) {
            pc.k1[co] = fold ? ht[co] * ht[16 + co] : ht[co]; pc.i1[co] = ht[16 + co];
            pc.k2[co] = fold ? ht[32 + co] * ht[48 + co] : ht[32 + co]; pc.i2[co] = ht[48 + co]; pc.bias[co] = hb[co];
        }
        a.amax_rw = amax;
        a.sync = (unsigned*)(amax + e->cap);
        if (lean_ok) {                                             // MAGIC epilogue: accumulators start at bias + 0x4B400000, i1 = -k1p * C
            for (int co = 0; co < 16; ++co) { pc.i1[co] = -(pc.k1[co] * AYQ_MAGIC_F); pc.bias[co] = hb[co] + AYQ_MAGIC_I; }   // (i1: the test build's conv_p1_fast_kernel)
            P1Const pc2 = pc;                                      // tensor-core kernel: MAGIC2 epilogue, first coefficient pre-scaled by 2^-8
            for (int co = 0; co < 16; ++co) pc2.k1[co] = pc.k1[co] * 0.00390625f;
#ifdef AYQ_TEST_BUILD
            const bool p1_tc = e->conv_impl >= 1 && !e->p1_dp4a;
#else
            const bool p1_tc = true;                               // product library: the tensor-core Conv_P1 only
#endif
            if (p1_tc) {                                           // tensor-core Conv_P1 (conv_tc.cuh): per-parity weight matrices with the byte shift
                tc::P1B wb;
                memset(&wb, 0, sizeof wb);
                for (int par = 0; par < 2; ++par)
                    for (int ky = 0; ky < 3; ++ky)
                        for (int co = 0; co < 16; ++co)
                            for (int t = 0; t < 9; ++t)            // t = kx * 3 + c
                                wb.b[par][ky == 2 ? 0 : ky + 2][co][(par ? 3 : 1) + t] = hw[co * 32 + ky * 9 + t];
                if (pa.p1_fused) {                                 // abs-max inside the kernel, one image ahead (conv_tc.cuh): nb * (n + 1) tickets
                    a.fuse_d = getenv("AYQ_P1_FUSE_D") ? atoi(getenv("AYQ_P1_FUSE_D")) : 0;
                    const unsigned nblk = (unsigned)(a.Hout / P1_TH) * (unsigned)(n + a.fuse_d);
                    if (pa.img_u8) CK(launch_k(tc::conv_p1_tc_kernel<true, true>, dim3(nblk), dim3(P1TC_THREADS), 0, st, a, pc2, wb));
                    else CK(launch_k(tc::conv_p1_tc_kernel<false, true>, dim3(nblk), dim3(P1TC_THREADS), 0, st, a, pc2, wb));
                }
                else if (a.M != 127) {                             // K = 6 / 4: explicit clamp to +-M
                    if (pa.img_u8) CK(launch_k(tc::conv_p1_tc_kernel<true, false, true>, dim3(1, a.Hout / P1_TH, nz), dim3(P1TC_THREADS), 0, st, a, pc2, wb));
                    else CK(launch_k(tc::conv_p1_tc_kernel<false, false, true>, dim3(1, a.Hout / P1_TH, nz), dim3(P1TC_THREADS), 0, st, a, pc2, wb));
                }
                else if (pa.img_u8) CK(launch_k(tc::conv_p1_tc_kernel<true, false>, dim3(1, a.Hout / P1_TH, nz), dim3(P1TC_THREADS), 0, st, a, pc2, wb));
                else CK(launch_k(tc::conv_p1_tc_kernel<false, false>, dim3(1, a.Hout / P1_TH, nz), dim3(P1TC_THREADS), 0, st, a, pc2, wb));
            }
#ifdef AYQ_TEST_BUILD
            else if (pa.img_u8) CK(launch_k(conv_p1_fast_kernel<true>, dim3(1, a.Hout / P1_TH, nz), dim3(256), 0, st, a, pc));
            else CK(launch_k(conv_p1_fast_kernel<false>, dim3(1, a.Hout / P1_TH, nz), dim3(256), 0, st, a, pc));
#endif
        } else if (pa.img_u8) CK(launch_k(conv_p1_kernel<true>, dim3((a.Wout + P1_TW - 1) / P1_TW, (a.Hout + P1_TH - 1) / P1_TH, nz), dim3(256), 0, st, a, pc));
        else CK(launch_k(conv_p1_kernel<false>, dim3((a.Wout + P1_TW - 1) / P1_TW, (a.Hout + P1_TH - 1) / P1_TH, nz), dim3(256), 0, st, a, pc));
        break;
    }
    case OP_CONV: {
        int rc = launch_conv(e, (int)i, n, st);
        if (rc) return rc;
        if (f[CF_ACC_BUF] >= 0 && f[CF_ACC_TAP] >= 0)             // parity taps of a float-head plan: same NCHW int32 layout
            CK(cudaMemcpyAsync(e->acc_taps[f[CF_ACC_TAP]], e->ws + e->buf_off[f[CF_ACC_BUF]],
                               (size_t)n * f[CF_COUT] * f[CF_HOUT] * f[CF_WOUT] * sizeof(int), cudaMemcpyDeviceToDevice, st));
        break;
    }
    case OP_POOL: {
        const BufDesc& ib = e->bufs[f[PL_IN_BUF]];
        const BufDesc& ob = e->bufs[f[PL_OUT_BUF]];
        const size_t ppx = (size_t)n * f[PL_H] * f[PL_W] * 16;
        const int8_t* in = (const int8_t*)(e->ws + e->buf_off[f[PL_IN_BUF]]) + (size_t)f[PL_IN_PLANE0] * ppx;
        int8_t* out = (int8_t*)(e->ws + e->buf_off[f[PL_OUT_BUF]]) + (size_t)f[PL_OUT_PLANE0] * ppx;
        (void)ib; (void)ob;
        if (f[PL_H] * f[PL_W] > 1024) return fail(-38, "SPPF pool: %dx%d map (one thread per pixel, at most 1024)", f[PL_H], f[PL_W]);
        CK(launch_k(sppf_pool_kernel, dim3(f[PL_NPLANES], n), dim3((unsigned)((f[PL_H] * f[PL_W] + 31) & ~31)), (size_t)f[PL_H] * f[PL_W] * 16 * 2, st, in, out, n, (int)f[PL_H], (int)f[PL_W], (int)f[PL_NPLANES]));
        break;
    }
    case OP_HEAD: {
        HeadArgs a;
        for (int l = 0; l < 3; ++l) {
            a.box[l] = (const int8_t*)(e->ws + e->buf_off[f[HD_BOX_BUF0 + l]]);
            a.cls[l] = (const int16_t*)(e->ws + e->buf_off[f[HD_CLS_BUF0 + l]]);
        }
        a.lut_exp = (const float*)(e->d_data + f[HD_LUT_EXP_OFF]);
        a.lut16 = (const int16_t*)(e->d_data + f[HD_LUT16_OFF]);
        a.lo16 = (const int16_t*)(e->d_data + f[HD_LO16_OFF]);
        a.mono = f[HD_MONO];
        a.fast_div = e->fast_div;
        a.dflw = (const int*)(e->d_data + f[HD_DFLW_OFF]);
        a.anchors = (const int*)(e->d_data + f[HD_ANCH_OFF]);
        a.kd = f_from_bits(f[HD_KD]); a.id = f_from_bits(f[HD_ID]);
        a.n = n; a.K = e->hdr.K; a.A = A;
        a.dbox = dbox; a.conf = conf; a.cls_id = cls; a.dbox_cls = dbox_cls;
        CK(launch_k(head_kernel, dim3((unsigned)(((size_t)n * A + 127) / 128)), dim3(128), 0, st, a));
        break;
    }
    case OP_NMS: {
        NmsArgs a;
        a.dbox = dbox; a.conf = conf; a.cls_id = cls; a.boxes = nullptr; a.scores = nullptr; a.n = n; a.A = A; a.mode = 0; a.max_keep = NMS_MAXDET; a.dets = dets; a.counts = counts;
        CK(launch_k(nms_kernel, dim3(n), dim3(NMS_THREADS), NMS_SMEM, st, a));
        break;
    }
    case OP_HEAD_FLOAT: {
        HeadFloatArgs a;
        for (int l = 0; l < 3; ++l) {
            a.box[l] = (const int*)(e->ws + e->buf_off[f[HF_BOX_BUF0 + l]]);
            a.cls[l] = (const int*)(e->ws + e->buf_off[f[HF_CLS_BUF0 + l]]);
        }
        a.box_scale = (const float*)(e->d_data + f[HF_BOX_SCALE_OFF]);
        a.cls_scale = (const float*)(e->d_data + f[HF_CLS_SCALE_OFF]);
        a.dflw = (const float*)(e->d_data + f[HF_DFLW_OFF]);
        a.n = n; a.A = A;
        a.dbox = dbox; a.conf = (float*)conf; a.cls_id = cls; a.dbox_cls = dbox_cls;
        CK(launch_k(head_float_kernel, dim3((unsigned)(((size_t)n * A + 127) / 128)), dim3(128), 0, st, a));
        break;
    }
    case OP_NMS_FLOAT: {
        NmsFloatArgs a;
        a.dbox = dbox; a.conf = (const float*)conf; a.cls_id = cls; a.n = n; a.A = A; a.max_keep = NMS_MAXDET; a.dets = dets; a.counts = counts;
        CK(launch_k(nms_float_kernel, dim3(n), dim3(NMS_THREADS), nmsf_smem_bytes(A), st, a));
        break;
    }
    default:
        return fail(-22, "plan op %zu has unknown kind %d", i, f[0]);
    }
    return 0;
}

static void drop_graphs(ayq_engine* e) {
    for (auto& kv : e->graphs) if (kv.second) cudaGraphExecDestroy(kv.second);
    e->graphs.clear();
}

// One pass = memset + abs-max + the plan ops.  The ops between Conv_P1 and q_NMS touch only engine-owned memory, so for
// a given pass size they are captured once into a CUDA graph (with the PDL edges) and replayed; Conv_P1 / q_NMS carry the
// caller's pointers and are launched directly.
// img (fp32) or img_u8 (uint8, ToTensor fused into the abs-max and Conv_P1 kernels): exactly one is non-null
static int run_pass(ayq_engine* e, const float* img, const uint8_t* img_u8, int n, float* dbox_cls, float* dets, int32_t* counts, cudaStream_t st,
                    cudaEvent_t ev_img = nullptr) {
    const int H = e->hdr.img_h, W = e->hdr.img_w;
    float* amax = (float*)(e->ws + e->off_amax);
    const bool prof = e->profiling;
    PassArgs pa{img, img_u8, n, dbox_cls, dets, counts};
    pa.ev_img = ev_img;
    // fused abs-max + Conv_P1 (one read of the image from HBM instead of two): the lean tensor-core Conv_P1 of the product path
    pa.p1_fused = e->p1_fuse && !e->p1_dp4a && e->conv_impl >= 1 && e->ops.size() && e->ops[0].f[0] == OP_CONV_P1;
    if (pa.p1_fused) pa.p1_fused = e->ops[0].f[P1_CLAMP] == 127 && p1_lean(e, e->ops[0].f, pa);
    e->last_fused = pa.p1_fused;
    int pe = 0;
    if (prof) CK(cudaEventRecord(e->prof_ev[pe++], st));
    CK(cudaMemsetAsync(amax, 0, sizeof(float) * e->cap + sizeof(unsigned) * (e->cap + 1), st));   // amax[] and the fused kernel's ticket / band counters
    // The image is read twice (per-image abs-max, then quantise + Conv_P1).  With fp32 images a pass does not fit the L2, so
    // the two kernels are interleaved over chunks of p1_chunk images: the second read of a chunk then comes from the L2.
    const int chunk = (!pa.p1_fused && !prof && !e->debug_sync && !img_u8 && e->p1_chunk > 0 && e->ops.size() && e->ops[0].f[0] == OP_CONV_P1) ? e->p1_chunk : 0;
    if (chunk && n > chunk) {
        const size_t per = (size_t)3 * H * W;
        for (int i0 = 0; i0 < n; i0 += chunk) {
            const int m = n - i0 < chunk ? n - i0 : chunk;
            CK(launch_k(absmax_kernel, dim3(64, m), dim3(256), 0, st, img + (size_t)i0 * per, amax + i0, per));
            pa.p1_img0 = i0; pa.p1_cnt = m;
            int rc = launch_op(e, 0, pa, st);
            if (rc) return rc;
        }
        pa.p1_img0 = 0; pa.p1_cnt = -1;
        if (pa.ev_img) CK(cudaEventRecord(pa.ev_img, st));
    } else if (pa.p1_fused) { /* the abs-max runs inside Conv_P1 */ }
    else if (img_u8) CK(launch_k(absmax_u8_kernel, dim3(32, n), dim3(256), 0, st, img_u8, amax, (size_t)3 * H * W));
    else CK(launch_k(absmax_kernel, dim3(64, n), dim3(256), 0, st, img, amax, (size_t)3 * H * W));
    const size_t op_first = (chunk && n > chunk) ? 1 : 0;            // Conv_P1 already launched chunk by chunk
    if (prof) CK(cudaEventRecord(e->prof_ev[pe++], st));
    if (e->debug_sync) {
        cudaError_t de = cudaStreamSynchronize(st);
        if (de == cudaSuccess) de = cudaGetLastError();
        if (de != cudaSuccess) return fail(-5, "absmax failed: %s", cudaGetErrorString(de));
    }
    const size_t nops = e->ops.size();
    size_t g0 = 0, g1 = nops;                                      // graphable range [g0, g1)
    while (g0 < nops && e->ops[g0].f[0] == OP_CONV_P1) ++g0;
    while (g1 > g0 && (e->ops[g1 - 1].f[0] == OP_NMS || e->ops[g1 - 1].f[0] == OP_NMS_FLOAT)) --g1;
    bool graphable = e->use_graph && !prof && !e->debug_sync && !dbox_cls && e->conv_impl == 2 && g1 > g0;
    for (size_t i = g0; i < g1 && graphable; ++i)
        if (e->ops[i].f[0] == OP_CONV_P1 || e->ops[i].f[0] == OP_NMS || e->ops[i].f[0] == OP_NMS_FLOAT) graphable = false;
    for (size_t i = op_first; i < nops; ++i) {
        if (graphable && i == g0) {
            auto it = e->graphs.find(n);
            if (it == e->graphs.end()) {
                // record on an engine-owned stream (the caller's may be the legacy default stream, which cannot capture)
                if (!e->s_cap) CK(cudaStreamCreateWithFlags(&e->s_cap, cudaStreamNonBlocking));
                cudaGraph_t graph = nullptr;
                CK(cudaStreamBeginCapture(e->s_cap, cudaStreamCaptureModeThreadLocal));
                int rc = 0;
                for (size_t j = g0; j < g1 && rc == 0; ++j) rc = launch_op(e, j, pa, e->s_cap);
                cudaError_t ce = cudaStreamEndCapture(e->s_cap, &graph);
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (ce != cudaSuccess) return fail(-5, "graph capture failed: %s", cudaGetErrorString(ce));
                cudaGraphExec_t exec = nullptr;
                ce = cudaGraphInstantiate(&exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ce != cudaSuccess) return fail(-5, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
                it = e->graphs.emplace(n, exec).first;
            }
            CK(cudaGraphLaunch(it->second, st));
            i = g1 - 1;
            continue;
        }
        int rc = launch_op(e, i, pa, st);
        if (rc) return rc;
        if (pa.ev_img && e->ops[i].f[0] == OP_CONV_P1) CK(cudaEventRecord(pa.ev_img, st));   // the staging buffer of the images is free again
        if (prof) CK(cudaEventRecord(e->prof_ev[pe++], st));
        if (e->debug_sync) {
            cudaError_t de = cudaStreamSynchronize(st);
            if (de == cudaSuccess) de = cudaGetLastError();
            if (de != cudaSuccess) return fail(-5, "op %zu (kind %d) failed: %s", i, e->ops[i].f[0], cudaGetErrorString(de));
        }
    }
    CK(cudaGetLastError());
    if (prof) {
        CK(cudaStreamSynchronize(st));
        for (int i = 0; i + 1 < pe; ++i) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e->prof_ev[i], e->prof_ev[i + 1]));
            e->op_ms[i] += ms;
            e->op_calls[i] += 1;
        }
    }
    e->last_n = n;
    return 0;
}

static int forward_device(ayq_handle e, const float* img, const uint8_t* img_u8, int n, float* dbox_cls, float* dets, int32_t* counts, void* stream);
extern "C" int ayq_forward(ayq_handle e, const float* img, int n, float* dbox_cls, float* dets, int32_t* counts, void* stream) {
    if (!e || !img || !dets || !counts || n < 0) return fail(-22, "ayq_forward: bad arguments");
    return forward_device(e, img, nullptr, n, dbox_cls, dets, counts, stream);
}
extern "C" int ayq_forward_u8(ayq_handle e, const uint8_t* img_u8, int n, float* dbox_cls, float* dets, int32_t* counts, void* stream) {
    if (!e || !img_u8 || !dets || !counts || n < 0) return fail(-22, "ayq_forward_u8: bad arguments");
    return forward_device(e, nullptr, img_u8, n, dbox_cls, dets, counts, stream);
}
static int forward_device(ayq_handle e, const float* img, const uint8_t* img_u8, int n, float* dbox_cls, float* dets, int32_t* counts, void* stream) {
    if (n == 0) return 0;
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int mb = e->max_batch;
    int rc = ensure_workspace(e, n < mb ? n : mb);
    if (rc) return rc;
    rc = wait_busy(e, st);                                         // a pass of another stream / the host pipeline may own the workspace
    if (rc) return rc;
    const size_t img_elems = (size_t)3 * e->hdr.img_h * e->hdr.img_w;
    for (int i0 = 0; i0 < n; i0 += mb) {
        const int m = (n - i0) < mb ? (n - i0) : mb;
        rc = run_pass(e, img ? img + (size_t)i0 * img_elems : nullptr, img_u8 ? img_u8 + (size_t)i0 * img_elems : nullptr, m,
                      dbox_cls ? dbox_cls + (size_t)i0 * 84 * e->hdr.n_anchors : nullptr,
                      dets + (size_t)i0 * AYQ_MAX_DET * AYQ_DET_STRIDE, counts + i0, st);
        if (rc) { mark_busy(e, st); return rc; }
    }
    return mark_busy(e, st);
}

// ---- host-buffer entry: double-buffered H2D / compute / D2H --------------------------------------------
static int ensure_host_pipeline(ayq_engine* e, int m, bool u8) {
    if (!e->s_copy) {
        CK(cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&e->ev_h2d[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&e->ev_d2h[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&e->ev_img[i], cudaEventDisableTiming));
        }
    }
    if (m > e->host_cap || (u8 && !e->d_img_u8[0])) {
        const int cap = m > e->host_cap ? m : e->host_cap;
        const size_t img_elems = (size_t)3 * e->hdr.img_h * e->hdr.img_w;
        for (int i = 0; i < 2; ++i) {
            if (cap > e->host_cap) {
                if (e->d_img[i]) cudaFree(e->d_img[i]);
                if (e->d_dets[i]) cudaFree(e->d_dets[i]);
                if (e->d_counts[i]) cudaFree(e->d_counts[i]);
                if (e->d_img_u8[i]) { cudaFree(e->d_img_u8[i]); e->d_img_u8[i] = nullptr; }
                CK(cudaMalloc(&e->d_img[i], img_elems * cap * sizeof(float)));
                CK(cudaMalloc(&e->d_dets[i], (size_t)cap * AYQ_MAX_DET * AYQ_DET_STRIDE * sizeof(float)));
                CK(cudaMalloc(&e->d_counts[i], (size_t)cap * sizeof(int)));
            }
            if (u8 && !e->d_img_u8[i]) CK(cudaMalloc(&e->d_img_u8[i], img_elems * cap));
        }
        e->host_cap = cap;
    }
    return 0;
}

static int host_sync(ayq_engine* e) {
    cudaError_t r0 = e->s_d2h ? cudaStreamSynchronize(e->s_d2h) : cudaSuccess;
    cudaError_t r1 = e->s_copy ? cudaStreamSynchronize(e->s_copy) : cudaSuccess;
    cudaError_t r2 = e->s_comp ? cudaStreamSynchronize(e->s_comp) : cudaSuccess;
    const cudaError_t r = r0 != cudaSuccess ? r0 : (r1 != cudaSuccess ? r1 : r2);
    if (r != cudaSuccess) return fail(-5, "host pipeline failed: %s", cudaGetErrorString(r));
    return 0;
}

// Enqueues the whole call on the engine's three streams (H2D | kernels | D2H) and returns; `sync` waits for the results.
// Slots, events and the pass counter persist across calls, so a second asynchronous call queued behind the first overlaps its
// first H2D with the first call's last pass (nothing but the very first copy and the very last pass of a SEQUENCE is exposed).
static int forward_host_impl(ayq_engine* e, const void* img_host, bool u8, int n, float* dets_host, int32_t* counts_host, bool sync) {
    if (!e || !img_host || !dets_host || !counts_host || n < 0) return fail(-22, "ayq_forward_host: bad arguments");
    if (n == 0) return 0;
    CK(cudaSetDevice(e->device));
    // pass size of the host pipeline: H2D of pass i+1 overlaps the kernels of pass i, so the first copy and the last pass are
    // exposed -- smaller passes shorten both (64 images keep the kernels within ~10 % of their large-batch throughput)
    const int host_pass = e->host_pass > 0 ? e->host_pass : 64;
    const int mb = e->max_batch < host_pass ? e->max_batch : host_pass;
    const int m_max = n < mb ? n : mb;
    // AYQ_HOST_RAMP=1 makes the passes at both ends smaller still (16, 32, ..., 32, 16).  Measured on B200 / PCIe 5 (256 uint8
    // images per call): uniform 64-image passes 35.7 k images/s, ramp 31.8 k, uniform 32 28.9 k, uniform 128 28.4 k -- a pass
    // has a fixed cost of ~0.5 ms (67 launches), so small passes fall behind the copy engine; the ramp stays off by default.
    // Splitting only the last pass (AYQ_HOST_TAIL) measured 33.5 k (32,32) / 33.3 k (48,16) / 31.4 k (32,16,16): also off.
    std::vector<int> sizes;
    if (n <= mb || e->host_ramp == 0) {
        // uniform passes; the LAST full pass may be split (AYQ_HOST_TAIL="32,32"): only the final pass's kernels are exposed
        // after the last byte has arrived, so a smaller final pass ends the call sooner
        int rem = n;
        int tail_sum = 0;
        for (int t : e->host_tail) tail_sum += t;
        const bool split = !e->host_tail.empty() && n > mb && tail_sum <= mb && n % mb == 0 && tail_sum == mb;
        for (; rem > (split ? mb : 0); rem -= mb) sizes.push_back(rem < mb ? rem : mb);
        if (split) for (int t : e->host_tail) sizes.push_back(t);
    } else {
        int rem = n;
        std::vector<int> tail;
        for (int t : {16, 32}) if (t < mb && rem - t >= mb) { tail.push_back(t); rem -= t; }
        for (int h : {16, 32}) if (h < mb && rem - h >= mb / 2) { sizes.push_back(h); rem -= h; }
        for (; rem > 0; rem -= mb) sizes.push_back(rem < mb ? rem : mb);
        for (auto it = tail.rbegin(); it != tail.rend(); ++it) sizes.push_back(*it);
    }
    if (m_max > e->cap || m_max > e->host_cap || (u8 && !e->d_img_u8[0])) {   // (re)allocation: nothing may be in flight
        int rc = host_sync(e);
        if (rc) return rc;
    }
    int rc = ensure_workspace(e, m_max);                           // the passes of this entry are at most m_max images
    if (rc) return rc;
    rc = ensure_host_pipeline(e, m_max, u8);
    if (rc) return rc;
    if (e->busy_recorded && !e->busy_from_host) {                  // a device-entry pass on a caller's stream may own the workspace
        CK(cudaStreamWaitEvent(e->s_comp, e->ev_busy, 0));         // (after another host call the pipeline's own stream order and slot events
        CK(cudaStreamWaitEvent(e->s_copy, e->ev_busy, 0));         //  suffice: making the copy stream wait for the previous call's last pass would
    }                                                              //  serialise the upload of this call behind it)
    const size_t img_elems = (size_t)3 * e->hdr.img_h * e->hdr.img_w;
    int i0 = 0;
    auto enqueue = [&]() -> int {
        for (size_t pi = 0; pi < sizes.size(); i0 += sizes[pi], ++pi, ++e->host_passes) {
            const int m = sizes[pi];
            const int slot = (int)(e->host_passes & 1);
            const bool reuse = e->host_passes >= 2;
            // three streams: H2D of pass i+1 and D2H of pass i-1 overlap the kernels of pass i
            if (reuse) CK(cudaStreamWaitEvent(e->s_copy, e->ev_img[slot], 0));    // d_img[slot] is read by abs-max and Conv_P1 only: free long before the pass ends
            if (u8) CK(cudaMemcpyAsync(e->d_img_u8[slot], (const uint8_t*)img_host + (size_t)i0 * img_elems, img_elems * m, cudaMemcpyHostToDevice, e->s_copy));
            else CK(cudaMemcpyAsync(e->d_img[slot], (const float*)img_host + (size_t)i0 * img_elems, img_elems * m * sizeof(float), cudaMemcpyHostToDevice, e->s_copy));
            CK(cudaEventRecord(e->ev_h2d[slot], e->s_copy));
            CK(cudaStreamWaitEvent(e->s_comp, e->ev_h2d[slot], 0));
            if (reuse) CK(cudaStreamWaitEvent(e->s_comp, e->ev_d2h[slot], 0));    // d_dets[slot] still draining
            int prc = run_pass(e, u8 ? nullptr : e->d_img[slot], u8 ? e->d_img_u8[slot] : nullptr, m, nullptr, e->d_dets[slot], e->d_counts[slot], e->s_comp, e->ev_img[slot]);
            if (prc) return prc;
            CK(cudaEventRecord(e->ev_done[slot], e->s_comp));
            CK(cudaStreamWaitEvent(e->s_d2h, e->ev_done[slot], 0));
            CK(cudaMemcpyAsync(dets_host + (size_t)i0 * AYQ_MAX_DET * AYQ_DET_STRIDE, e->d_dets[slot],
                               (size_t)m * AYQ_MAX_DET * AYQ_DET_STRIDE * sizeof(float), cudaMemcpyDeviceToHost, e->s_d2h));
            CK(cudaMemcpyAsync(counts_host + i0, e->d_counts[slot], (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, e->s_d2h));
            CK(cudaEventRecord(e->ev_d2h[slot], e->s_d2h));
        }
        return 0;
    };
    rc = enqueue();
    if (rc) {                                                      // DMA into the caller's buffers may be in flight: drain before reporting
        const std::string msg = g_err;
        host_sync(e);
        g_err = msg;
        return rc;
    }
    CK(cudaEventRecord(e->ev_busy, e->s_comp));                    // later device-entry calls wait for the pipeline's last pass
    e->busy_recorded = true;
    e->busy_from_host = true;
    return sync ? host_sync(e) : 0;
}

extern "C" int ayq_forward_host(ayq_handle e, const float* img_host, int n, float* dets_host, int32_t* counts_host) {
    return forward_host_impl(e, img_host, false, n, dets_host, counts_host, true);
}
extern "C" int ayq_forward_host_u8(ayq_handle e, const uint8_t* img_host, int n, float* dets_host, int32_t* counts_host) {
    return forward_host_impl(e, img_host, true, n, dets_host, counts_host, true);
}
extern "C" int ayq_forward_host_async(ayq_handle e, const void* img_host, int is_u8, int n, float* dets_host, int32_t* counts_host) {
    return forward_host_impl(e, img_host, is_u8 != 0, n, dets_host, counts_host, false);
}
extern "C" int ayq_wait(ayq_handle e) {
    if (!e) return fail(-22, "ayq_wait: null handle");
    CK(cudaSetDevice(e->device));
    return host_sync(e);
}

// ---- taps ---------------------------------------------------------------------------------------------
extern "C" int ayq_buffer_shape(ayq_handle e, int buf, int* channels, int* height, int* width) {
    if (!e || buf < 0 || buf >= (int)e->bufs.size()) return fail(-22, "ayq_buffer_shape: bad buffer %d", buf);
    if (channels) *channels = e->bufs[buf].nplanes * 16;
    if (height) *height = e->bufs[buf].H;
    if (width) *width = e->bufs[buf].W;
    return 0;
}
extern "C" int ayq_export_buffer(ayq_handle e, int buf, int n, int32_t* dst, void* stream) {
    if (!e || buf < 0 || buf >= (int)e->bufs.size() || !dst) return fail(-22, "ayq_export_buffer: bad arguments");
    if (n != e->last_n) return fail(-22, "ayq_export_buffer: n=%d but the last pass had %d images", n, e->last_n);
    const BufDesc& b = e->bufs[buf];
    int rc = wait_busy(e, (cudaStream_t)stream);
    if (rc) return rc;
    export_planes_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(e->ws + e->buf_off[buf], b.elem_bytes, b.nplanes, n, b.H, b.W, dst);
    CK(cudaGetLastError());
    return mark_busy(e, (cudaStream_t)stream);
}
extern "C" int ayq_export_acc_tap(ayq_handle e, int tap, int n, int32_t* dst, void* stream) {
    if (!e || tap < 0 || tap >= (int)e->acc_taps.size() || !dst) return fail(-22, "ayq_export_acc_tap: bad tap %d (plan compiled without taps?)", tap);
    if (n != e->last_n) return fail(-22, "ayq_export_acc_tap: n=%d but the last pass had %d images", n, e->last_n);
    int rc = wait_busy(e, (cudaStream_t)stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(dst, e->acc_taps[tap], e->acc_tap_elems[tap] * n * sizeof(int), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return mark_busy(e, (cudaStream_t)stream);
}

// ---- unit-level layer library -------------------------------------------------------------------------
static inline unsigned grid_for(size_t total) {
    size_t g = (total + 255) / 256;
    return (unsigned)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}
extern "C" int ayq_requantize_f32(const float* x, float* y, const float* k, const float* inv2s, int per_channel,
                                  int n, int c, int hw, int bits, void* stream) {
    if (!x || !y || !k || !inv2s || bits < 2 || bits > 24) return fail(-22, "ayq_requantize_f32: bad arguments");
    const size_t total = (size_t)n * c * hw;
    if (!total) return 0;
    requantize_f32_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, y, k, inv2s, per_channel, c, hw, total, (1 << (bits - 1)) - 1);
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ayq_silu_f32(const float* acc, float* y, const float* tab, const float* lut, int n, int c, int hw, int bits, void* stream) {
    if (!acc || !y || !tab || !lut || bits < 2 || bits > 16) return fail(-22, "ayq_silu_f32: bad arguments");
    const size_t total = (size_t)n * c * hw;
    if (!total) return 0;
    silu_f32_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(acc, y, tab, lut, c, hw, total, (1 << (bits - 1)) - 1);
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ayq_lut_f32(const float* x, float* y, const float* lut, int key_min, int key_max, size_t count, void* stream) {
    if (!x || !y || !lut || key_max < key_min) return fail(-22, "ayq_lut_f32: bad arguments");
    if (!count) return 0;
    lut_f32_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(x, y, lut, key_min, key_max, count);
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ayq_absmax_f32(const float* x, float* out, int n, size_t per_image, void* stream) {
    if (!x || !out || n < 0) return fail(-22, "ayq_absmax_f32: bad arguments");
    if (!n) return 0;
    CK(cudaMemsetAsync(out, 0, sizeof(float) * n, (cudaStream_t)stream));
    if (!per_image) return 0;
    size_t blocks = (per_image / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 64) blocks = 64;
    absmax_kernel<<<dim3((unsigned)blocks, n), 256, 0, (cudaStream_t)stream>>>(x, out, per_image);
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ayq_quant_input_f32(const float* x, float* y, float* amax, float* scales, int n, size_t per_image, int bits, void* stream) {
    if (!x || !y || !amax || !scales || bits < 2 || bits > 16) return fail(-22, "ayq_quant_input_f32: bad arguments");
    int rc = ayq_absmax_f32(x, amax, n, per_image, stream);
    if (rc) return rc;
    const size_t total = per_image * n;
    if (!total) return 0;
    quant_input_f32_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, y, amax, scales, per_image, n, (1 << (bits - 1)) - 1);
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ayq_calib_conv_f32(const float* x, const float* w, const float* b, float* y, float* amax, int n, int cin, int H, int W,
                                  int cout, int ks, int stride, void* stream) {
    if (!x || !w || !b || !y || n < 0 || cin < 1 || cout < 1 || H < 1 || W < 1 || (ks != 1 && ks != 3) || (stride != 1 && stride != 2))
        return fail(-22, "ayq_calib_conv_f32: bad arguments (kernel 1 or 3, stride 1 or 2)");
    if (!n) return 0;
    const int pad = ks / 2, Hout = (H + 2 * pad - ks) / stride + 1, Wout = (W + 2 * pad - ks) / stride + 1;
    if (n > 65535 || (cout + CALIB_CO - 1) / CALIB_CO > 65535) return fail(-22, "ayq_calib_conv_f32: batch / channel count too large");
    calib_conv_f32_kernel<<<dim3((unsigned)((Hout * Wout + 127) / 128), (unsigned)((cout + CALIB_CO - 1) / CALIB_CO), (unsigned)n), 128, 0, (cudaStream_t)stream>>>(
        x, w, b, y, amax, cin, H, W, cout, Hout, Wout, ks, stride);
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ayq_calib_silu_f32(float* x, size_t count, void* stream) {
    if (!x) return fail(-22, "ayq_calib_silu_f32: null pointer");
    if (!count) return 0;
    calib_silu_f32_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(x, count);
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ayq_calib_maxpool5_f32(const float* x, float* y, int planes, int H, int W, void* stream) {
    if (!x || !y || planes < 0 || H < 1 || W < 1) return fail(-22, "ayq_calib_maxpool5_f32: bad arguments");
    if (!planes) return 0;
    calib_maxpool5_f32_kernel<<<grid_for((size_t)planes * H * W), 256, 0, (cudaStream_t)stream>>>(x, y, planes, H, W);
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ayq_calib_upsample2_f32(const float* x, float* y, int planes, int H, int W, void* stream) {
    if (!x || !y || planes < 0 || H < 1 || W < 1) return fail(-22, "ayq_calib_upsample2_f32: bad arguments");
    if (!planes) return 0;
    calib_upsample2_f32_kernel<<<grid_for((size_t)planes * H * W * 4), 256, 0, (cudaStream_t)stream>>>(x, y, planes, H, W);
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ayq_quant_weights_f32(const float* w, const float* bias, int cout, size_t per_channel, int bits, double scale_input,
                                     int8_t* qw, int64_t* qb, double* scale_res, void* stream) {
    if (!w || !bias || !qw || !qb || !scale_res || cout < 0 || bits < 2 || bits > 8) return fail(-22, "ayq_quant_weights_f32: bad arguments (bits 2..8)");
    if (!cout || !per_channel) return 0;
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    quant_weights_kernel<<<cout, 256, 0, (cudaStream_t)stream>>>(w, bias, per_channel, (1 << (bits - 1)) - 1, scale_input, qw, (long long*)qb, scale_res);
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ayq_nms(ayq_handle e, const float* dbox_cls, int n, float* dets, int32_t* counts, void* stream) {
    if (!e || !dbox_cls || !dets || !counts || n < 0) return fail(-22, "ayq_nms: bad arguments");
    if (!n) return 0;
    CK(cudaSetDevice(e->device));
    const int A = e->hdr.n_anchors;
    const int mb = e->max_batch;
    int rc = ensure_workspace(e, n < mb ? n : mb);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = wait_busy(e, st);
    if (rc) return rc;
    for (int i0 = 0; i0 < n; i0 += e->cap) {
        const int m = (n - i0) < e->cap ? (n - i0) : e->cap;
        float4* dbox = (float4*)(e->ws + e->off_dbox);
        int* conf = (int*)(e->ws + e->off_conf);
        int* cls = (int*)(e->ws + e->off_cls);
        pred_to_cand_kernel<<<(unsigned)(((size_t)m * A + 127) / 128), 128, 0, st>>>(dbox_cls + (size_t)i0 * 84 * A, m, A, dbox, conf, cls);
        NmsArgs a;
        a.dbox = dbox; a.conf = conf; a.cls_id = cls; a.boxes = nullptr; a.scores = nullptr; a.n = m; a.A = A; a.mode = 0; a.max_keep = NMS_MAXDET;
        a.dets = dets + (size_t)i0 * AYQ_MAX_DET * AYQ_DET_STRIDE; a.counts = counts + i0;
        nms_kernel<<<m, NMS_THREADS, NMS_SMEM, st>>>(a);
    }
    CK(cudaGetLastError());
    return mark_busy(e, st);
}

extern "C" int ayq_coord_float(ayq_handle e, const float* dbox_cls, int n, float* dets, int32_t* counts, void* stream) {
    if (!e || !dbox_cls || !dets || !counts || n < 0) return fail(-22, "ayq_coord_float: bad arguments");
    if (!n) return 0;
    CK(cudaSetDevice(e->device));
    const int A = e->hdr.n_anchors;
    const int mb = e->max_batch;
    int rc = ensure_workspace(e, n < mb ? n : mb);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    rc = wait_busy(e, st);
    if (rc) return rc;
    for (int i0 = 0; i0 < n; i0 += e->cap) {
        const int m = (n - i0) < e->cap ? (n - i0) : e->cap;
        float4* dbox = (float4*)(e->ws + e->off_dbox);
        float* conf = (float*)(e->ws + e->off_conf);
        int* cls = (int*)(e->ws + e->off_cls);
        pred_to_cand_float_kernel<<<(unsigned)(((size_t)m * A + 127) / 128), 128, 0, st>>>(dbox_cls + (size_t)i0 * 84 * A, m, A, dbox, conf, cls);
        NmsFloatArgs a;
        a.dbox = dbox; a.conf = conf; a.cls_id = cls; a.n = m; a.A = A; a.max_keep = NMS_MAXDET;
        a.dets = dets + (size_t)i0 * AYQ_MAX_DET * AYQ_DET_STRIDE; a.counts = counts + i0;
        nms_float_kernel<<<m, NMS_THREADS, nmsf_smem_bytes(A), st>>>(a);
    }
    CK(cudaGetLastError());
    return mark_busy(e, st);
}

extern "C" int ayq_nms_boxes(const float* boxes, const float* scores, int nb, float* keep, int32_t* count, void* stream) {
    if (!boxes || !scores || !keep || !count || nb < 0 || nb > NMS_SORT_N) return fail(-22, "ayq_nms_boxes: 0 <= nb <= %d", NMS_SORT_N);
    if (nb == 0) { CK(cudaMemsetAsync(count, 0, sizeof(int), (cudaStream_t)stream)); return 0; }
    CK(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NMS_SMEM));   // per device: set on the current one
    NmsArgs a;
    a.dbox = nullptr; a.conf = nullptr; a.cls_id = nullptr; a.boxes = boxes; a.scores = scores;
    a.n = 1; a.A = nb; a.mode = 1; a.max_keep = NMS_TOPK; a.dets = keep; a.counts = count;
    nms_kernel<<<1, NMS_THREADS, NMS_SMEM, (cudaStream_t)stream>>>(a);
    CK(cudaGetLastError());
    return 0;
}
