// conv_tma.cuh -- TMA-fed variant of the tcgen05 / TMEM implicit-GEMM quantised convolution (sm_100a).
//
// Same GEMM view, tiles, B layout, TMEM double buffering and fixed-point epilogue as conv_tc.cuh; what changes is the
// A-operand feed.  Activation buffers are 16-channel planes [plane][n][H][W][16 B], i.e. a rank-5 uint8 tensor
// {16, W, H, n, planes}.  For one K-chunk group (source buffer, filter tap, np consecutive planes) ONE tensor-map TMA
// (cp.async.bulk.tensor.5d) with box {16, bw*s, bh*s, bn, np} and element strides {1, s, s, 1, 1} lands exactly np K-chunks
// of the A tile in shared memory, each [128 pixels][16 B] = the canonical K-major no-swizzle core-matrix layout:
//   * the filter tap is the box origin (x0*s + kx - pad, y0*s + ky - pad): out-of-range coordinates are zero filled
//     by the hardware, which IS the convolution padding (and the image overhang of the last tile);
//   * the convolution stride is the TMA element stride;
//   * concat / residual-sum inputs are further groups from other tensor maps (plan.py's segment lists).
// One elected thread issues the loads (no producer warps, no per-pixel address arithmetic, no proxy fences); stages are
// packed on the host from whole boxes.  Warp roles: warps 2, 3 = TMA producers (stages round-robin;
// A, and B when the weights are streamed; warp 2 also loads resident weights), warps 0, 1 = MMA issuers (even / odd tiles;
// warp 1 also owns the TMEM allocation),
// warps 4-7 / 8-11 = epilogue of even / odd tiles.
#pragma once
#include <cuda.h>
#include "conv_tc.cuh"

namespace ayq {

namespace tc {

// Role-level cycle counters (AYQ_ROLE_PROF=1) exist only in the profiling build of the library (libayq_prof.so, built with
// -DAYQ_ROLE_PROF_BUILD): a predicated-off clock read still costs an issue slot, and these sit in every per-tile loop.
#define AYQ_DBG_SLOTS 24          // per CTA: 16 role counters + 4 globaltimer stamps (entry, prologue done, dependency wait done, exit)
#ifdef AYQ_ROLE_PROF_BUILD
#define AYQ_DBG(a) ((a).dbg != nullptr)
#define AYQ_STAMP(a, k) do { if ((a).dbg && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); (a).dbg[blockIdx.x * AYQ_DBG_SLOTS + 16 + (k)] = (long long)t_; } } while (0)
#define AYQ_CLK(a) ((a).dbg ? clock64() : 0)
#define AYQ_XMODE(a, bit) (((a).dbg_mode & (bit)) != 0)     // experiments (AYQ_EPI_SKIP bit mask): 2 no epilogue work, 4 no loads, 8 no MMAs
#else
#define AYQ_DBG(a) false
#define AYQ_STAMP(a, k) do { } while (0)
#define AYQ_CLK(a) 0ll
#define AYQ_XMODE(a, bit) false
#endif

constexpr int TMA_MAX_MAPS = 8;
#define AYQ_MAX_OUT_ 3
constexpr int TMA_MAX_OPS = 56;
constexpr int TMA_MAX_STAGES = 40;
// block = 256 + 256 * EG threads: warps 0-1 and 6-7 MMA issuers (two per pipeline), 2-5 TMA producers, then EG epilogue groups (4 warps) per pipeline

struct TmaOp { int map; int dx, dy, p0; uint32_t dst_off; };          // box origin offsets (tap - pad), first plane, byte offset in the slot
struct TmaStage { int op0, nops, nchunks, chunk0; };
constexpr int TMA_MAX_HMMA = 80;
// one K = 32 MMA of a halo tile = two K chunks: A start (16-byte units, low 16 bits) | LBO (16-byte units) << 16, and the same for
// B relative to the resident weights (first chunk * N | chunk distance * N << 16)
struct HaloMma { uint32_t a_off_lbo; uint32_t b_off_lbo; };
struct TmaPlan {
    int nstages, nops, stride, slot_chunks;                            // slot_chunks = K chunks one ring slot can hold
    int merged_cx;                                                     // 1: stride-1 conv, rank-4 maps with (channel, x) merged into one 16*W byte row
    int a_slot_bytes;                                                  // bytes of one ring slot of A
    // halo mode (3x3 stride-1 convs on 8 x 16 pixel tiles): ONE box per input segment per tile = the (8+2) x (16+2) pixel halo
    // of np planes, [plane][18][10][16 B]; the nine taps are shifted windows of it, addressed by the MMA descriptors
    // (start += (ky*10 + kx) * 16 B, SBO = 160 B = one halo row, LBO = distance to the chunk that forms the K = 32 pair)
    int halo, n_hmma, halo_tx_bytes;
    HaloMma hm[TMA_MAX_HMMA];
    TmaStage st[TMA_MAX_STAGES];
    TmaOp op[TMA_MAX_OPS];
};
struct TmaMaps { CUtensorMap m[TMA_MAX_MAPS]; };

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

__device__ __forceinline__ uint32_t elect_one() {          // one lane of the (fully converged) warp
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(pred));
    return pred;
}

// dynamic smem: [A ring NS*slot_chunks*2048][B: resident nkc_pad*N*16 | ring NS*slot_chunks*N*16][tab 4N f32][bias N i32][lut 256 f32]
// EG = epilogue groups per pipeline: 2 (one group per TMEM buffer) for cout <= 32, where the epilogue is the issue-bound
// stage and the kernels are small enough in registers for 768 threads; 1 otherwise (the group alternates buffers).
// EG = epilogue groups (of 4 warps) per pipeline = TMEM accumulator buffers per pipeline when EG >= 2 (group k owns buffer k and
// every EG-th tile of its pipeline).  3 for cout <= 32 (1024 threads at <= 64 registers: these layers are bound by instruction
// issue in the epilogue, more resident epilogue warps hide its latencies), 2 for the other compile-time-cout kernels, 1 otherwise.
// NBC == 0 (cout 128 / 256, coefficients from shared memory): two groups per pipeline as well, but they SPLIT THE COLUMNS of the
// same tile (group k drains channels [k * N / 2, (k + 1) * N / 2)): sixteen epilogue warps instead of eight on the layers
// whose tiles are the longest to drain, without a second accumulator per group (cout 256 fills the TMEM with one per pipeline).
#define TMA_EG(NBC) ((NBC) == 1 || (NBC) == 2 ? 3 : 2)
#define TMA_CSPLIT(NBC) ((NBC) == 0)
constexpr int TMA_NB = 6;                                          // barrier slots per pipeline for the accumulator buffers (ring of up to 6)
template <int NBC, int EPI, int FAST, int EG = TMA_EG(NBC)>
__global__ void __launch_bounds__(256 + 256 * EG, 1) conv_tma_kernel(const __grid_constant__ ConvArgs a, const __grid_constant__ TcParams tp,
                                                                  const __grid_constant__ EpiTab et, const __grid_constant__ TmaPlan pl,
                                                                  const __grid_constant__ TmaMaps maps) {
    constexpr int TMA_THREADS = 256 + 256 * EG;
#ifdef AYQ_ROLE_PROF_BUILD
    unsigned long long t_entry_;                                   // first instruction of the CTA, before any parameter is touched
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_entry_));
#endif
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2 * TC_MAX_NS + 2 * 2 * TMA_NB + 2];   // full[NS], empty[NS], tfull[2][NB], tempty[2][NB], wfull, lfull
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31;
    // Role of a warp.  The SMSP arbiter favours the highest warp ids, so with role_hi the eight control warps (MMA issuers, TMA
    // producers: few instructions, but every one of them sits on the tile-to-tile critical path) take the LAST eight warp ids and
    // the epilogue groups the first ones; `warp` below is the role index (0-1 MMA, 2-5 producers, 8.. epilogue).  The shift is a
    // multiple of four, so role index and hardware warp id agree modulo 4 (the TMEM lane quarter a warp may read).
    constexpr int TMA_WARPS = TMA_THREADS / 32;
    const int hw_warp = tid >> 5;
    const int warp = tp.role_hi ? (hw_warp + 8 >= TMA_WARPS ? hw_warp + 8 - TMA_WARPS : hw_warp + 8) : hw_warp;
    const int N = a.cout, NS = tp.NS;
    const uint32_t a_slot_bytes = (uint32_t)pl.a_slot_bytes, b_slot_bytes = (uint32_t)pl.slot_chunks * N * 16u;
    unsigned char* sA = smem;
    unsigned char* sB = smem + (size_t)NS * a_slot_bytes;
    const size_t b_bytes = tp.resident_b ? (size_t)tp.nkc_pad * N * 16 : (size_t)NS * b_slot_bytes;
    float* tab_s = (float*)(sB + b_bytes);
    int* bias_s = (int*)(tab_s + 4 * N);
    float* lut_s = (float*)(bias_s + N);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[TC_MAX_NS]);
    const uint32_t tfull0 = smem_u32(&bars[2 * TC_MAX_NS]), tempty0 = smem_u32(&bars[2 * TC_MAX_NS + 2 * TMA_NB]), wfull = smem_u32(&bars[2 * TC_MAX_NS + 4 * TMA_NB]), lfull = wfull + 8;
    // TMEM accumulator ring of a pipeline: tile j of the pipeline (j = 0, 1, ...) lands in buffer j % nbuf and is drained by
    // epilogue group j % EG.  nbuf = 2 * EG where the 512 columns allow it: a group then never waits for the refill of the buffer
    // it has just handed back (barrier hand-off + MMA issue + MMA execution, several hundred cycles) -- its next tile is already
    // complete in the other buffer.
    const int nbuf = tp.nbuf;

    pdl_trigger();
    AYQ_STAMP(a, 0);
#ifdef AYQ_ROLE_PROF_BUILD
    if (a.dbg && threadIdx.x == 0) a.dbg[blockIdx.x * AYQ_DBG_SLOTS + 20] = (long long)t_entry_;
#endif
    // ---- prologue: touches only engine constants (tables, weights), so it may overlap the previous kernel's tail ----
    if (NBC == 0) {
        if (FAST) {                                                 // folded coefficients: rows 0 / 2 hold k * 2^-s (exact)
            for (int i = tid; i < N; i += TMA_THREADS) {
                const float k1p = __fmul_rn(a.tab[i], a.tab[N + i]);
                tab_s[i] = (EPI == 0 && FAST >= 2) ? __fmul_rn(k1p, 0.00390625f) : k1p;        // MAGIC2: coefficient of the first requant pre-scaled by 2^-8 (exact)
                tab_s[2 * N + i] = __fmul_rn(a.tab[2 * N + i], a.tab[3 * N + i]);
                tab_s[N + i] = 0.f; tab_s[3 * N + i] = 0.f;
            }
        } else {
            for (int i = tid; i < 4 * N; i += TMA_THREADS) tab_s[i] = a.tab[i];
        }
        for (int i = tid; i < N; i += TMA_THREADS) bias_s[i] = a.bias[i] + (FAST == 2 ? AYQ_MAGIC_I : 0);
    }
    if (EPI == 0 && FAST >= 2) {
        // per-lane replicated sigmoid table (fixedpoint.cuh: silu_magic2): the engine built it once in global memory, two bulk copies
        // bring its 32.1 KB in while the rest of the prologue runs (filling it from the 255-entry table with dependent loads took
        // ~2 us per CTA, all of it between the previous kernel's last store and this kernel's first load)
        if (warp == 3 && lane == 0) {
            mbar_init(lfull, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_arrive_expect_tx(lfull, (uint32_t)AYQ_LUTREP_BYTES);
            bulk_g2s(smem_u32(lut_s), a.lut_rep, AYQ_LUTREP_BYTES / 2, lfull);
            bulk_g2s(smem_u32(lut_s) + AYQ_LUTREP_BYTES / 2, (const unsigned char*)a.lut_rep + AYQ_LUTREP_BYTES / 2, AYQ_LUTREP_BYTES / 2, lfull);
        }
        if (a.gen_outs) {                                          // 256-byte table per requantised output: index = SiLU result + 128
            unsigned char* rq = (unsigned char*)lut_s + AYQ_LUTREP_BYTES;
            for (int i = tid; i < 256 * a.nout; i += TMA_THREADS) {
                const OutSpec& os = a.out[i >> 8];
                rq[i] = (unsigned char)(os.mode == 1 ? requant8((float)((i & 255) - 128), os.k, os.inv, a.M) : ((i & 255) - 128));
            }
        }
    }
    else if (EPI == 0) fill_lut256(lut_s, a.lut, a.M, tid, TMA_THREADS);
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2 * TMA_NB; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, TMA_CSPLIT(NBC) ? 128 * EG : 128); }
        if (!tp.resident_b) mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)tp.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 2 && tp.resident_b) {                               // resident weights: one bulk-TMA burst, still in the prologue
        if (lane == 0) {
            mbar_init(wfull, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint32_t total = (uint32_t)tp.nkc_pad * N * 16u;
            mbar_arrive_expect_tx(wfull, total);
            for (uint32_t o = 0; o < total; o += 32768u) {
                const uint32_t bytes = total - o < 32768u ? total - o : 32768u;
                bulk_g2s(smem_u32(sB) + o, a.w + o, bytes, wfull);
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    AYQ_STAMP(a, 1);
    pdl_wait();                                                     // activations of the previous layer are complete from here on
    AYQ_STAMP(a, 2);

    // Two pipelines share the CTA: pipeline m (m = 0, 1) owns the tiles i with i % 2 == m, its accumulator ring in TMEM and its EG
    // epilogue groups.  Each pipeline is fed by TWO independent chains (q = 0, 1: the pipeline's even / odd tiles), each chain =
    // one TMA producer warp -> a private ring of NSQ = NS / 4 smem slots -> one MMA issuer warp.  The per-tile issue path of an
    // issuer (barrier probes, fences, elect, descriptor moves to uniform registers, MMAs, commits) is a serial chain of ~200
    // instructions that shares a scheduler with six epilogue warps (measured ~1100 busy cycles per tile); four chains overlap
    // them.  Private rings keep every mbarrier strictly in order for its one producer and one consumer: a parity wait may never
    // skip a phase (a consumer that has not seen phase 0 of a barrier finds "parity 1" already complete on a fresh barrier).
    const int nq = tp.nq;                                         // chains per pipeline: 2, or 1 when the pipeline has a single accumulator (cout 256)
    const int NSQ = NS / (2 * nq);
    if (warp >= 2 && warp < 6) {
        // ===== TMA producer of chain (m, q): warp-uniform control flow, one elected lane issues (all operands of the tensor loads
        // live in uniform registers: no per-lane waterfall loops) =====
        const int m = (warp - 2) & 1, q = (warp - 2) >> 1;
        const int nstages = pl.nstages, stride = pl.stride;
        const int slot0 = (nq * m + q) * NSQ;
        int slot = 0;
        uint32_t ephase = 1;                                       // fresh barrier: parity 1 passes immediately
        long long d_wait = 0, d_t0 = AYQ_CLK(a);
        for (int t = q < nq ? blockIdx.x + (m + 2 * q) * gridDim.x : tp.ntiles; t < tp.ntiles; t += 2 * nq * gridDim.x) {
            const TileCoord tc0 = tile_coord(t, tp);
            const int xs = tc0.x0 * stride, ys = tc0.y0 * stride;
            for (int s = 0; s < nstages; ++s) {
                const TmaStage sg = pl.st[s];
                const int gs = slot0 + slot;
                const long long w0 = AYQ_CLK(a);
                mbar_wait_relaxed<768>(empty0 + 8 * gs, ephase);
                if (AYQ_DBG(a)) d_wait += clock64() - w0;
                if (AYQ_XMODE(a, 4)) { if (elect_one()) mbar_arrive(full0 + 8 * gs); }
                else if (elect_one()) {
                    const uint32_t bar = full0 + 8 * gs;
                    const uint32_t nch_b = (uint32_t)((sg.nchunks + 1) & ~1);
                    mbar_arrive_expect_tx(bar, (pl.halo ? (uint32_t)pl.halo_tx_bytes : (uint32_t)sg.nchunks * 2048u) + (tp.resident_b ? 0u : nch_b * N * 16u));
                    const uint32_t dst0 = smem_u32(sA) + gs * a_slot_bytes;
                    for (int o = sg.op0; o < sg.op0 + sg.nops; ++o) {
                        const TmaOp op = pl.op[o];
                        if (pl.merged_cx) tma_load_4d(dst0 + op.dst_off, &maps.m[op.map], (xs + op.dx) * 16, ys + op.dy, tc0.img0, op.p0, bar);
                        else tma_load_5d(dst0 + op.dst_off, &maps.m[op.map], 0, xs + op.dx, ys + op.dy, tc0.img0, op.p0, bar);
                    }
                    if (!tp.resident_b)
                        bulk_g2s(smem_u32(sB) + gs * b_slot_bytes, a.w + (size_t)sg.chunk0 * N * 16, nch_b * N * 16u, bar);
                }
                __syncwarp();
                if (++slot == NSQ) { slot = 0; ephase ^= 1; }
            }
        }
        if (AYQ_DBG(a) && lane == 0 && q == 0) { a.dbg[blockIdx.x * AYQ_DBG_SLOTS + m * 2] = clock64() - d_t0; a.dbg[blockIdx.x * AYQ_DBG_SLOTS + m * 2 + 1] = d_wait; }
    } else if (warp < 2 || warp == 6 || warp == 7) {
        // ===== MMA issuer of chain (m, par): warp-uniform control flow, one elected lane issues; the descriptor low words (address |
        // LBO) are stepped with 32-bit adds.  Warps m and 6 + m issue the even / odd tiles of pipeline m from their private rings; the
        // two share the pipeline's accumulator ring (tile j -> buffer j % nbuf), whose barriers each of them always finds in the
        // phase it expects: a buffer cannot be released twice before the issuer that waits for it has refilled it. =====
        const int m = warp & 1, par = warp >= 6 ? 1 : 0;
        const uint32_t idesc = make_idesc_i8(N);
        const int nstages = pl.nstages;
        if (tp.resident_b) mbar_wait(wfull, 0);
        const uint32_t d_hi = (128u >> 4) | (1u << 14);            // SBO = 128 B, descriptor version 1 (bits 32.. of make_desc)
        const uint32_t h_hi = (160u >> 4) | (1u << 14);            // halo mode: 8-pixel row groups are one halo row (10 pixels) apart
        const uint32_t a_lo0 = ((smem_u32(sA) & 0x3ffffu) >> 4) | ((2048u >> 4) << 16);
        const uint32_t b_base = (smem_u32(sB) & 0x3ffffu) >> 4;    // halo mode: the plan's words carry the chunk offset and the LBO
        const uint32_t b_lo0 = b_base | ((((uint32_t)N * 16u) >> 4) << 16);
        const uint32_t a_step = a_slot_bytes >> 4, b_step = tp.resident_b ? 0u : (b_slot_bytes >> 4);
        const uint32_t b_pair = 2u * (uint32_t)N;                  // two K chunks of B, in 16-byte units
        const TmaStage sg0 = pl.st[0];
        const int slot0 = (nq * m + par) * NSQ;
        int slot = 0, buf = par;                                   // chain `par` owns the accumulators j % nbuf of its tiles j = par, par + nq, ...: with
                                                                   // nq = 2 and nbuf even, the even / odd buffers -- private, like its ring slots
        uint32_t fphase = 0, ephase = 1;                           // parity to wait for on tempty[m][buf]: fresh barriers pass parity 1; flips per ring turn
        long long d_we = 0, d_wf = 0, d_t0 = AYQ_CLK(a);
        for (int t = par < nq ? blockIdx.x + (m + 2 * par) * gridDim.x : tp.ntiles; t < tp.ntiles; t += 2 * nq * gridDim.x) {
            const long long w0 = AYQ_CLK(a);
            // When the epilogue is the bottleneck the issuers spend most of their time here, a whole accumulator ring ahead: a tight
            // probe loop was 6 % of all issued instructions (ncu source view); look again every ~0.25 us instead.
            mbar_wait_relaxed<256>(tempty0 + 8 * (TMA_NB * m + buf), ephase);
            if (AYQ_DBG(a)) d_we += clock64() - w0;
            tc_fence_after();
            const uint32_t dcol = tmem_base + (uint32_t)((m * nbuf + buf) * N);
            uint32_t accum = 0;
            for (int s = 0; s < nstages; ++s) {
                TmaStage sg = sg0;
                if (s > 0) sg = pl.st[s];
                const int gs = slot0 + slot;
                const long long w1 = AYQ_CLK(a);
                mbar_wait(full0 + 8 * gs, fphase);
                if (AYQ_DBG(a)) d_wf += clock64() - w1;
                tc_fence_after();
                if (elect_one()) {
                    uint32_t acc = accum;
                    if (AYQ_XMODE(a, 8)) { }
                    else if (pl.halo) {
                        const uint32_t abase = ((smem_u32(sA) & 0x3ffffu) >> 4) + (uint32_t)gs * a_step;
                        const int n_hmma = pl.n_hmma;
                        for (int j = 0; j < n_hmma; ++j) {
                            const HaloMma hmj = pl.hm[j];
                            mma_i8_lh(dcol, abase + hmj.a_off_lbo, h_hi, b_base + hmj.b_off_lbo, d_hi, idesc, acc);
                            acc = 1;
                        }
                    } else {
                        uint32_t alo = a_lo0 + (uint32_t)gs * a_step;
                        uint32_t blo = tp.resident_b ? b_lo0 + (uint32_t)sg.chunk0 * (uint32_t)N : b_lo0 + (uint32_t)gs * b_step;
                        const int pairs = (sg.nchunks + 1) >> 1;  // an odd tail pairs with stale smem x zero weights
                        for (int j = 0; j < pairs; ++j) {
                            mma_i8_lh(dcol, alo, d_hi, blo, d_hi, idesc, acc);
                            acc = 1; alo += 256u; blo += b_pair;
                        }
                    }
                    mma_commit(empty0 + 8 * gs);                  // frees the smem slot when these MMAs retire
                    if (s == nstages - 1) mma_commit(tfull0 + 8 * (TMA_NB * m + buf));   // accumulator complete -> epilogue group m
                }
                __syncwarp();
                accum = 1;
                if (++slot == NSQ) { slot = 0; fphase ^= 1; }
            }
            buf += nq;
            if (buf >= nbuf) { buf -= nbuf; ephase ^= 1u; }
        }
        if (AYQ_DBG(a) && lane == 0 && par == 0) { a.dbg[blockIdx.x * AYQ_DBG_SLOTS + 6 + 3 * m] = clock64() - d_t0; a.dbg[blockIdx.x * AYQ_DBG_SLOTS + 7 + 3 * m] = d_we; a.dbg[blockIdx.x * AYQ_DBG_SLOTS + 8 + 3 * m] = d_wf; }
    } else if (warp >= 8) {
        // ===== epilogue: TMEM -> registers -> fixed-point SiLU / requant -> 16-byte plane rows =====
        const int gq = (warp - 8) >> 2;
        const int grp = gq / EG, gk = gq - grp * EG;              // pipeline (tile parity) this group drains; its index inside the pipeline
        constexpr bool CS = TMA_CSPLIT(NBC);                     // column split: every group of the pipeline drains every tile (its share of the channels)
        const int tstep = CS ? 2 : 2 * EG;                       // otherwise group gk takes every EG-th tile of the pipeline
        const int row = ((warp & 3) << 5) | lane;                // TMEM lane == GEMM row; warp w may touch lanes 32*(w%4)..
        const int dx = row & ((1 << tp.bw_log) - 1), dy = (row >> tp.bw_log) & ((1 << tp.bh_log) - 1), dn = row >> (tp.bw_log + tp.bh_log);
        const uint32_t lane_quad = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t tphase = 0;                                     // parity to wait for on tfull[grp][buf] (flips per ring turn)
        int buf = CS ? 0 : gk;                                   // tile j = gk, gk + EG, ... of the pipeline -> buffer j % nbuf
        while (buf >= nbuf) { buf -= nbuf; tphase ^= 1u; }
        long long d_wt = 0, d_t0 = AYQ_CLK(a);
        if (EPI == 0 && FAST >= 2) mbar_wait(lfull, 0);          // the sigmoid table has landed (prologue bulk copy)
        const EpiPairs cp = epi_pairs();                         // packed-epilogue constants, once per thread
        const StoreOff so_inv = FAST ? store_off(a, 0, 0, 0) : StoreOff{0u, 0u, 0u, 0u, 0u};   // its loop-invariant fields
        const StoreOffThread so_th = store_off_thread(a, dn, dy, dx);
        for (int t = blockIdx.x + (grp + (CS ? 0 : 2 * gk)) * gridDim.x; t < tp.ntiles; t += tstep * gridDim.x) {
            const uint32_t lane_base = lane_quad + (uint32_t)((grp * nbuf + buf) * N);
            const uint32_t tfull_b = tfull0 + 8 * (TMA_NB * grp + buf), tempty_b = tempty0 + 8 * (TMA_NB * grp + buf);
            const TileCoord tc0 = tile_coord(t, tp);
            const int img = tc0.img0 + dn, oy = tc0.y0 + dy, ox = tc0.x0 + dx;
            const bool valid = img < a.n && oy < a.Hout && !AYQ_XMODE(a, 2);          // overhanging tiles: images past the batch, rows past the map (halo mode)
            const StoreOff so = !FAST ? StoreOff{0u, 0u, 0u, 0u, 0u}                                // per-tile part of the store addresses
                                : (((tc0.x0 | tc0.y0) & 1) == 0 ? store_off_tile(a, so_inv, so_th, tc0.img0, tc0.y0, tc0.x0) : store_off(a, img, oy, ox));
            const long long w0 = AYQ_CLK(a);
            // ONE warp of the group polls the accumulator barrier; the other three park in a hardware barrier that costs no issue
            // slots (every polling warp re-executes its probe loop whenever any barrier of the CTA changes: measured 8 % of all
            // issued instructions with four pollers per group).
            if ((warp & 3) == 0) mbar_wait(tfull_b, tphase);
            group_bar_sync(1 + gq, 128);
            if (AYQ_DBG(a)) d_wt += clock64() - w0;
            tc_fence_after();
            int accA[16], accB[16];
            const int nbh = CS ? (N / 16) / EG : N / 16;          // 16-channel groups this epilogue group drains (even, checked by the host)
            const int g0 = CS ? gk * nbh : 0;
            tmem_ld16(lane_base + (uint32_t)(g0 * 16), accA);
            if (NBC > 0) {
#pragma unroll
                for (int gch = 0; gch < NBC; ++gch) {
                    int* cur = (gch & 1) ? accB : accA;
                    int* nxt = (gch & 1) ? accA : accB;
                    tmem_ld_wait16(cur);
                    if (gch + 1 < NBC) tmem_ld16(lane_base + (uint32_t)((gch + 1) * 16), nxt);
                    else { tc_fence_before(); mbar_arrive(tempty_b); }   // accumulator fully read: hand it back to the MMA warp
                    if (valid) epilogue16_t<EPI, true, FAST>(a, et, cur, gch * 16, img, oy, ox, tab_s, bias_s, lut_s, so, cp);
                }
            } else {
                const int nb = g0 + nbh;
                for (int gch = g0; gch < nb; gch += 2) {
                    tmem_ld_wait16(accA);
                    tmem_ld16(lane_base + (uint32_t)((gch + 1) * 16), accB);
                    if (valid) epilogue16_t<EPI, false, FAST>(a, et, accA, gch * 16, img, oy, ox, tab_s, bias_s, lut_s, so, cp);
                    tmem_ld_wait16(accB);
                    if (gch + 2 < nb) tmem_ld16(lane_base + (uint32_t)((gch + 2) * 16), accA);
                    else { tc_fence_before(); mbar_arrive(tempty_b); }
                    if (valid) epilogue16_t<EPI, false, FAST>(a, et, accB, (gch + 1) * 16, img, oy, ox, tab_s, bias_s, lut_s, so, cp);
                }
            }
            buf += CS ? 1 : EG;
            while (buf >= nbuf) { buf -= nbuf; tphase ^= 1u; }
        }
        if (AYQ_DBG(a) && (warp & 3) == 0 && lane == 0 && gk == 0) { a.dbg[blockIdx.x * AYQ_DBG_SLOTS + 12 + 2 * grp] = clock64() - d_t0; a.dbg[blockIdx.x * AYQ_DBG_SLOTS + 13 + 2 * grp] = d_wt; }
    }
    tc_fence_before();
#ifdef AYQ_ROLE_PROF_BUILD
    // exit stamp = arrival of the CTA's LAST warp at the final barrier.  (A stamp read by one thread after the barrier is not that:
    // BAR.SYNC.DEFER_BLOCKING lets a warp run ahead to the next instruction that needs the barrier, and a timer read does not.)
    if (a.dbg && lane == 0) {
        unsigned long long t_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_) :: "memory");
        atomicMax((unsigned long long*)&a.dbg[blockIdx.x * AYQ_DBG_SLOTS + 19], t_);
    }
#endif
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tp.tmem_cols) : "memory");
    }
}

}  // namespace tc

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmaLaunch {            // everything one launch needs, cached per (op, images in the pass)
    int n = -1;               // images the cache was built for (-1 = empty), 0 = shape not covered
    int ok = 0;
    tc::TcParams tp;
    tc::EpiTab et;
    tc::TmaPlan pl;
    tc::TmaMaps maps;
    size_t smem = 0;
    unsigned grid = 0;
    int fast = 0;             // epilogue variant: 0 generic, 1 FAST (folded coefficients), 2 MAGIC (silu_magic)
    int gen_outs = 0;         // MAGIC with a general output list
};

struct TmaState { PFN_tmapEncodeTiled encode = nullptr; int num_sms = 148; int ready = 0; int halo_min_np = 1; int budget_kb = 208; int resident_kb = 96; int role_hi = 0; int nbuf_mul = 2;
                  int force_nq1 = 0; };   // force_nq1: one producer -> issuer chain per pipeline (AYQ_ONE_ISSUER=1, or the per-layer choice of the load-time tuner)

typedef void (*TmaKernel)(const ConvArgs, const tc::TcParams, const tc::EpiTab, const tc::TmaPlan, const tc::TmaMaps);
template <int FAST>
static inline TmaKernel tma_pick_t(int N, int epi) {
    using namespace tc;
    if (epi == 0) {
        switch (N) {
        case 16: return conv_tma_kernel<1, 0, FAST>;
        case 32: return conv_tma_kernel<2, 0, FAST>;
        case 64: return conv_tma_kernel<4, 0, FAST>;
        case 80: return conv_tma_kernel<5, 0, FAST>;
        default: return (N % 64 == 0) ? conv_tma_kernel<0, 0, FAST> : nullptr;   // column split: two groups x an even number of 16-channel groups
        }
    }
    constexpr int F = FAST == 2 ? 2 : (FAST ? 1 : 0);             // requant epilogues: generic, folded, folded + magic int -> float
    if (epi == 1) return N == 64 ? conv_tma_kernel<4, 1, F> : (N % 64 == 0 ? conv_tma_kernel<0, 1, F> : nullptr);
    if (epi == 2) return N == 80 ? conv_tma_kernel<5, 2, F> : (N % 64 == 0 ? conv_tma_kernel<0, 2, F> : nullptr);
    return nullptr;
}
static inline TmaKernel tma_pick(int N, int epi, int fast) {
    return fast == 3 ? tma_pick_t<3>(N, epi) : fast == 2 ? tma_pick_t<2>(N, epi) : fast ? tma_pick_t<1>(N, epi) : tma_pick_t<0>(N, epi);
}
// FAST epilogue: clamp 127 (16-bit logits: 32767), a single identity output, no accumulator tap
// epilogue groups per pipeline of the kernel tma_pick() returns for this cout (must mirror TMA_EG / the switch in tma_pick_t)
static inline int tma_eg(int N, int epi) {
    if (epi == 0) return (N == 16 || N == 32) ? 3 : 2;
    return 2;
}
// column split (TMA_CSPLIT): the generic-cout kernels; their groups share every tile, so the accumulator ring is per pipeline
static inline bool tma_csplit(int N, int epi) {
    if (epi == 0) return !(N == 16 || N == 32 || N == 64 || N == 80);
    return !((epi == 1 && N == 64) || (epi == 2 && N == 80));
}
static inline bool tma_fast(const ConvArgs& a) {
    if (a.acc_tap) return false;
    if ((unsigned long long)a.n * a.cout * a.Hout * a.Wout * (a.epi == 2 ? 2 : 1) >= (1ull << 32)) return false;   // 32-bit store offsets
    if (a.epi == 2) return true;
    if (a.M != 127) return false;
    if (a.epi == 1) return true;
    // one identity output (plain or phase-split), or the plain tensor plus its phase-split copy
    if (a.nout == 1) return a.out[0].mode == 0 && a.out[0].up != 1;
    return a.nout == 2 && a.out[0].mode == 0 && a.out[0].up == 0 && a.out[1].mode == 0 && a.out[1].up == 2;
}

// MAGIC epilogue (fixedpoint.cuh: silu_magic) is exact iff, for every output channel, (1) |bias| + 127 * sum|w| < 2^22, so
// that accumulator + bias + 0x4B400000 reinterprets as 1.5 * 2^23 + acc, and (2) the SiLU result can never reach -128, so
// that the saturating conversion alone implements the clamp to [-127, 127]: for a negative accumulator with first requant
// r1 (> -128) we have acc >= (r1 - 1/2) / k1p, hence k2p * lut[r1] * acc >= k2p * lut[r1] * (r1 - 1/2) / k1p; r1 = -128 is the
// saturated end where the table must be 0.  Both are checked here on the host from the plan's weights and coefficients.
static inline bool magic_coeffs_ok(int N, int M, const float* h_tab, const int* h_bias, const float* h_lut, const long long* sum_abs_w) {
    if (M != 127 || !h_lut || getenv("AYQ_NO_MAGIC")) return false;
    if (h_lut[0] != 0.f) return false;                            // table[-M]: value used by every saturated negative r1
    for (int c = 0; c < N; ++c) {
        const long long bound = (h_bias[c] < 0 ? -(long long)h_bias[c] : h_bias[c]) + 127ll * sum_abs_w[c];
        if (bound >= (1ll << 22) - 1) return false;
        const double k1p = (double)h_tab[c] * h_tab[N + c], k2p = (double)h_tab[2 * N + c] * h_tab[3 * N + c];
        if (!(k1p > 0.0) || !(k2p > 0.0)) return false;
        for (int r1 = -M; r1 < 0; ++r1) {
            const double l = h_lut[r1 + M];
            if (l < 0.0) return false;
            if (k2p * l * (r1 - 0.5) / k1p * 1.001 < -126.5) return false;
        }
    }
    return true;
}
static inline bool magic_epilogue_ok(const ConvArgs& a, const float* h_tab, const int* h_bias, const float* h_lut, const int8_t* h_w, int nkc_pad) {
    if (!h_w || a.cout > 256 || getenv("AYQ_NO_MAGIC")) return false;
    const int N = a.cout;
    long long sw[256];
    for (int c = 0; c < N; ++c) {
        long long t = 0;
        for (int k = 0; k < nkc_pad; ++k) {
            const int8_t* w = h_w + ((size_t)k * N + c) * 16;
            for (int j = 0; j < 16; ++j) t += w[j] < 0 ? -w[j] : w[j];
        }
        sw[c] = t;
    }
    if (a.epi != 0) {   // requantize-only epilogues (requant_last_layers / exponent_requant): only the magic int -> float range has to hold; inputs are K-bit activations (|x| <= 127)
        for (int c = 0; c < N; ++c)
            if ((h_bias[c] < 0 ? -(long long)h_bias[c] : (long long)h_bias[c]) + 127ll * sw[c] >= (1ll << 22) - 1) return false;
        return a.epi == 2 || a.M == 127;
    }
    return magic_coeffs_ok(N, a.M, h_tab, h_bias, h_lut, sw);
}

static inline void tma_init(TmaState& s) {
    const int ns[] = {16, 32, 64, 80, 128};
    const bool carve = getenv("AYQ_CARVEOUT") != nullptr && atoi(getenv("AYQ_CARVEOUT")) != 0;
    for (int epi = 0; epi < 3; ++epi)
        for (int N : ns)
            for (int fast = 0; fast < 4; ++fast) {
                TmaKernel k = tma_pick(N, epi, fast);
                if (k) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
                // one L1 / shared-memory split for every conv launch: consecutive kernels with different dynamic sizes otherwise
                // get different carve-outs, and an SM re-partitions only when idle (measured: see DESIGN.md, launch-to-launch gap)
                if (k && carve) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&s.num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (const char* ev = getenv("AYQ_HALO_MIN_NP")) s.halo_min_np = atoi(ev);
    if (const char* ev = getenv("AYQ_ROLE_HI")) s.role_hi = atoi(ev);
    if (const char* ev = getenv("AYQ_ONE_ISSUER")) s.force_nq1 = atoi(ev) != 0;
    if (const char* ev = getenv("AYQ_NBUF_MUL")) s.nbuf_mul = atoi(ev);   // 1: one accumulator buffer per epilogue group (round-1 behaviour)
    if (const char* ev = getenv("AYQ_SMEM_KB")) s.budget_kb = atoi(ev);          // experiments: smaller CTAs let consecutive kernels co-reside
    if (const char* ev = getenv("AYQ_RESIDENT_KB")) s.resident_kb = atoi(ev);   // 16-channel inputs (np = 1) pair taps 16 B apart: slower than plain boxes
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn) {
        s.encode = (PFN_tmapEncodeTiled)fn;
        s.ready = 1;
    }
    cudaGetLastError();
}

// One buffer segment of the conv input as the TMA sees it.
struct TmaSeg { const void* base; int nplanes; };   // base = first byte of the buffer (plane 0), planes in the buffer

// Build the launch record.  h_kc: K-chunk list (pad_ = index into segs).  Returns L.ok.
static inline int tma_prepare(TmaState& s, TmaLaunch& L, const ConvArgs& a, const KChunk* h_kc, const TmaSeg* segs, int nsegs,
                              const float* h_tab, const int* h_bias, const float* h_lut = nullptr, const int8_t* h_w = nullptr) {
    L.n = a.n; L.ok = 0;
    if (!s.ready) return 0;
    const int N = a.cout;
    // MAGIC epilogue: any output list (32-bit store offsets: the largest output, 4x for an upsampled copy, stays below 4 GB)
    bool any_up = false;
    for (int o = 0; o < a.nout; ++o) any_up |= a.out[o].up == 1;
    const bool magic = !a.acc_tap && (unsigned long long)a.n * a.cout * a.Hout * a.Wout * (any_up ? 4 : 1) < (1ull << 32) &&
                       magic_epilogue_ok(a, h_tab, h_bias, h_lut, h_w, (a.nkc + 1) & ~1);
    // WIDE form of the same epilogue (fixedpoint.cuh): any clamp (K = 8 / 6 / 4), any accumulator range, explicit result clamps
    // (used where the plain FAST epilogue does not apply -- another clamp, or a general output list; where it does, FAST keeps its
    // 1 KB table, which matters for the ring depth of exactly these large-K layers: measured)
    const bool wide = !magic && a.epi == 0 && !a.acc_tap && h_lut && a.M <= 127 && (a.M != 127 || !tma_fast(a)) && !getenv("AYQ_NO_WIDE") &&
                      (unsigned long long)a.n * a.cout * a.Hout * a.Wout * (any_up ? 4 : 1) < (1ull << 32);
    const bool fast = magic || wide || tma_fast(a);               // FAST epilogue: folded coefficients k * 2^-s (exact), see fixedpoint.cuh
    L.fast = magic ? 2 : wide ? 3 : fast ? 1 : 0;
    L.gen_outs = a.epi == 0 && (magic || wide) && !tma_fast(a) ? 1 : 0;         // beyond "one identity output (+ phase-split copy)"
    if (N % 16 != 0 || N < 16 || N > 256 || !tma_pick(N, a.epi, 0)) return 0;
    tc::TcParams& tp = L.tp;
    tp.role_hi = s.role_hi;
    int bw_log = 4;
    while (bw_log > 0 && (a.Wout % (1 << bw_log))) --bw_log;
    int bh_log = 7 - bw_log;
    while (bh_log > 0 && (a.Hout % (1 << bh_log))) --bh_log;
    tp.bw_log = bw_log; tp.bh_log = bh_log;
    const int bn = 128 >> (bw_log + bh_log);
    tp.tiles_x = a.Wout >> bw_log;
    tp.tiles_y = a.Hout >> bh_log;
    tp.ntiles = tp.tiles_x * tp.tiles_y * ((a.n + bn - 1) / bn);
    tp.mul_x = tc_magic(tp.tiles_x); tp.mul_y = tc_magic(tp.tiles_y);
    if ((unsigned long long)tp.ntiles * (unsigned)tp.tiles_x >= (1ull << 32)) return 0;
    if ((1 << bw_log) * a.stride > 256 || (1 << bh_log) * a.stride > 256 || bn > 256) return 0;
    tp.nkc_pad = (a.nkc + 1) & ~1;
    // ring slot = 32 KB / 16 KB of A, or the whole (small) K extent: small slots leave room for a deep ring
    const int slot_cap = N <= 64 ? 16 : 8;
    int slot_chunks = tp.nkc_pad < slot_cap ? tp.nkc_pad : slot_cap;
    {   // the ring needs at least four slots (two per pipeline): shrink the slot until they fit next to the tables / resident weights
        const size_t lut_b = a.epi == 0 ? (((magic && a.epi == 0) || wide) ? (size_t)AYQ_LUTREP_BYTES : (size_t)AYQ_LUT256 * 8) + 256 * AYQ_MAX_OUT_ : 0;
        const size_t fixed_b = (size_t)N * 20 + lut_b + 64, w_b = (size_t)tp.nkc_pad * N * 16;
        const bool res = w_b <= (size_t)s.resident_kb * 1024;
        // ... and prefer eight (two per chain, so that a chain can load its next stage while the current one is multiplied) as long as
        // a slot keeps at least four K chunks
        while (slot_chunks > 2) {
            const size_t per = (size_t)slot_chunks * 2048 + (res ? 0 : (size_t)slot_chunks * N * 16);
            const size_t fit = ((size_t)s.budget_kb * 1024 - fixed_b - (res ? w_b : 0)) / per;
            if (fit >= 8 || (fit >= 4 && slot_chunks <= 4)) break;
            if (fit >= 4 && (tp.nkc_pad + slot_chunks / 2 - 1) / (slot_chunks / 2) > tc::TMA_MAX_STAGES - 2) break;   // stage table size
            slot_chunks = (slot_chunks / 2 + 1) & ~1;
            if (slot_chunks < 2) slot_chunks = 2;
        }
    }
    tp.KS = slot_chunks; tp.nst = 0; tp.lag = 0;
    int cols = 32;
    const int eg = tma_eg(N, a.epi);
    int nbuf = tma_csplit(N, a.epi) ? 2 : s.nbuf_mul * eg;        // accumulator ring per pipeline: two buffers per epilogue group ...
    while (nbuf > 1 && 2 * nbuf * N > 512) --nbuf;                // ... as far as the 512 TMEM columns go (cout 80: three, 256: one)
    if (nbuf > tc::TMA_NB) nbuf = tc::TMA_NB;
    // two producer -> ring -> issuer chains per pipeline when its accumulators split evenly between them (private even / odd
    // buffers); an odd count (cout 80: three) keeps them all with a single chain -- measured faster than two chains with one each
    // ... and layers whose weights are streamed through the ring (K x cout x 16 B > 96 KB: the 1152- / 2304-K convs at 20x20) keep one
    // chain per pipeline as well: half as many, twice as deep rings (measured 2-5 us per layer, consistently: DESIGN.md, tuner)
    const bool streamed_b = (size_t)tp.nkc_pad * N * 16 > (size_t)s.resident_kb * 1024;
    tp.nq = nbuf >= 2 && (nbuf & 1) == 0 && !s.force_nq1 && !streamed_b ? 2 : 1;
    tp.nbuf = nbuf;
    while (cols < 2 * nbuf * N) cols <<= 1;                       // two pipelines x nbuf accumulators
    tp.tmem_cols = cols;

    tc::TmaPlan& pl = L.pl;
    pl.halo = 0; pl.n_hmma = 0; pl.halo_tx_bytes = 0;
    L.smem = 0;
    // ---- halo mode: 3x3 stride-1 convs whose map tiles into 8 x 16 pixel boxes ----
    // Hout need not be a multiple of 16: the last row of tiles overhangs (TMA zero-fills the reads, the epilogue masks the stores)
    // as long as at least 80 % of the GEMM rows stay useful (40 x 40 maps: 83 %).
    const int halo_ty = (a.Hout + 15) / 16;
    if (a.stride == 1 && a.Wout % 8 == 0 && a.Hout * 5 >= halo_ty * 16 * 4 && a.Win == a.Wout && a.Hin == a.Hout && a.nkc >= 9 &&
        tp.nkc_pad * N * 16 <= s.resident_kb * 1024) {
        const int HW = 10, HH = 18, PLANE16 = HW * HH;           // halo pixels per plane (= 16-byte units)
        struct Blk { int seg, p0, np, i0; uint32_t reg16; };
        Blk blk[8];
        int nblk = 0, i = 0;
        uint32_t reg16 = 0;                                       // running region start, 16-byte units
        bool ok = true;
        while (ok && i < a.nkc) {                                 // a block = 9 taps (raster order) x np planes of one buffer segment
            int np = 1;
            while (i + np < a.nkc && h_kc[i + np].pad_ == h_kc[i].pad_ && h_kc[i + np].dy == h_kc[i].dy && h_kc[i + np].dx == h_kc[i].dx &&
                   h_kc[i + np].plane == h_kc[i].plane + np) ++np;
            if (nblk == 8 || i + 9 * np > a.nkc || a.nkc > 2 * tc::TMA_MAX_HMMA || np < s.halo_min_np) { ok = false; break; }
            for (int t = 0; t < 9 && ok; ++t)
                for (int q = 0; q < np && ok; ++q) {
                    const KChunk& k = h_kc[i + t * np + q];
                    if (k.pad_ != h_kc[i].pad_ || k.plane != h_kc[i].plane + q || k.dy != t / 3 - 1 || k.dx != t % 3 - 1) ok = false;
                }
            blk[nblk].seg = h_kc[i].pad_; blk[nblk].p0 = h_kc[i].plane; blk[nblk].np = np; blk[nblk].reg16 = reg16; blk[nblk].i0 = i;
            ++nblk;
            reg16 += (uint32_t)((np * PLANE16 + 7) & ~7);         // keep every box start 128-byte aligned
            i += 9 * np;
        }
        // Pair the K chunks into K = 32 MMAs.  Chunk (block b, tap t, plane q) sits at A address reg16_b + q * PLANE16 + tap offset and
        // at B index i0_b + t * np_b + q.  Both descriptors need a POSITIVE distance from the first to the second chunk of a
        // pair: planes (q, q + 1) of one tap pair up (A: one plane apart, B: adjacent); what is left over when np is odd -- one
        // chunk per tap, all of them when the input has 16 channels -- is paired in (block, tap) order, where A addresses and
        // B indices both increase; a last odd chunk meets the zero chunk that pads the weights to an even count.
        int npairs = 0;
        struct Single { int a16, bidx; };
        Single singles[2 * tc::TMA_MAX_HMMA];
        int nsingle = 0;
        auto add_pair = [&](int a0, int a1, int b0, int b1) {
            const int lbo = a1 - a0, dch = b1 - b0;
            if (npairs >= tc::TMA_MAX_HMMA || lbo <= 0 || lbo >= (1 << 14) || a0 >= (1 << 14) || dch <= 0 || dch * N >= (1 << 14) || b0 * N >= (1 << 14)) { ok = false; return; }
            pl.hm[npairs].a_off_lbo = (uint32_t)a0 | ((uint32_t)lbo << 16);
            pl.hm[npairs].b_off_lbo = (uint32_t)(b0 * N) | ((uint32_t)(dch * N) << 16);
            ++npairs;
        };
        for (int b = 0; b < nblk && ok; ++b)
            for (int t = 0; t < 9 && ok; ++t) {
                const int a_t = (int)blk[b].reg16 + (t / 3) * HW + (t % 3), b_t = blk[b].i0 + t * blk[b].np;
                int q = 0;
                for (; q + 1 < blk[b].np && ok; q += 2) add_pair(a_t + q * PLANE16, a_t + (q + 1) * PLANE16, b_t + q, b_t + q + 1);
                if (q < blk[b].np) { singles[nsingle].a16 = a_t + q * PLANE16; singles[nsingle].bidx = b_t + q; ++nsingle; }
            }
        for (int j = 0; j + 1 < nsingle && ok; j += 2) add_pair(singles[j].a16, singles[j + 1].a16, singles[j].bidx, singles[j + 1].bidx);
        if (ok && (nsingle & 1)) {                                // odd chunk count: second half = any valid smem x the zero chunk at index nkc
            if (!(a.nkc & 1)) ok = false;
            else add_pair(singles[nsingle - 1].a16, singles[nsingle - 1].a16 + 1, singles[nsingle - 1].bidx, a.nkc);
        }
        if (nblk > tc::TMA_MAX_MAPS) ok = false;
        if (ok) {
            tp.bw_log = 3; tp.bh_log = 4;
            tp.tiles_x = a.Wout >> 3; tp.tiles_y = halo_ty;
            tp.ntiles = tp.tiles_x * tp.tiles_y * a.n;
            tp.mul_x = tc_magic(tp.tiles_x); tp.mul_y = tc_magic(tp.tiles_y);
            if ((unsigned long long)tp.ntiles * (unsigned)tp.tiles_x >= (1ull << 32)) ok = false;
        }
        if (ok) {
            pl.halo = 1; pl.n_hmma = npairs; pl.halo_tx_bytes = 0;
            pl.nstages = 1; pl.nops = nblk; pl.stride = 1; pl.slot_chunks = 2; pl.merged_cx = 1;
            pl.st[0].op0 = 0; pl.st[0].nops = nblk; pl.st[0].nchunks = a.nkc; pl.st[0].chunk0 = 0;
            for (int b = 0; b < nblk && ok; ++b) {
                tc::TmaOp& op = pl.op[b];
                op.map = b; op.dx = -1; op.dy = -1; op.p0 = blk[b].p0; op.dst_off = blk[b].reg16 * 16u;
                pl.halo_tx_bytes += blk[b].np * PLANE16 * 16;
                if (blk[b].seg < 0 || blk[b].seg >= nsegs) { ok = false; break; }
                const TmaSeg& sg = segs[blk[b].seg];
                const cuuint64_t gdim[4] = {(cuuint64_t)a.Win * 16, (cuuint64_t)a.Hin, (cuuint64_t)a.n, (cuuint64_t)sg.nplanes};
                const cuuint64_t gstr[3] = {(cuuint64_t)a.Win * 16, (cuuint64_t)a.Hin * a.Win * 16, (cuuint64_t)a.n * a.Hin * a.Win * 16};
                const cuuint32_t box[4] = {(cuuint32_t)(HW * 16), (cuuint32_t)HH, 1, (cuuint32_t)blk[b].np};
                const cuuint32_t estr[4] = {1, 1, 1, 1};
                if (s.encode(&L.maps.m[b], CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(sg.base), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) ok = false;
            }
            for (int q = nblk; q < tc::TMA_MAX_MAPS; ++q) L.maps.m[q] = L.maps.m[0];
        }
        if (ok) {
            pl.a_slot_bytes = (int)(((size_t)reg16 * 16 + 1023) & ~(size_t)1023);
            tp.KS = 2; tp.nst = 0; tp.lag = 0;
            tp.resident_b = 1;
            const size_t lut_bytes = a.epi == 0 ? (((magic && a.epi == 0) || wide) ? (size_t)AYQ_LUTREP_BYTES : (size_t)AYQ_LUT256 * 8) + 256 * AYQ_MAX_OUT_ : 0;   // sigmoid table (MAGIC2: replicated per lane) + requant byte tables
            const size_t fixed = (size_t)N * 20 + lut_bytes + 64;
            const size_t w_bytes = (size_t)tp.nkc_pad * N * 16;
            const size_t avail = (size_t)s.budget_kb * 1024 - fixed - w_bytes;
            int ns = (int)(avail / (size_t)pl.a_slot_bytes);
            if (ns > tc::TC_MAX_NS) ns = tc::TC_MAX_NS;
            ns &= ~3;
            if (ns >= 4) {
                tp.NS = ns;
                L.smem = fixed + w_bytes + (size_t)ns * pl.a_slot_bytes;
                L.grid = (unsigned)tp.ntiles < (unsigned)s.num_sms ? (unsigned)tp.ntiles : (unsigned)s.num_sms;
                if (N <= TC_CT_MAXN) {
                    for (int c = 0; c < N; ++c) {
                        L.et.k1[c] = fast ? h_tab[c] * h_tab[N + c] * (((magic && a.epi == 0) || wide) ? 0.00390625f : 1.f) : h_tab[c]; L.et.i1[c] = h_tab[N + c];   // MAGIC2: k1 * 2^-s1 * 2^-8
                        L.et.k2[c] = fast ? h_tab[2 * N + c] * h_tab[3 * N + c] : h_tab[2 * N + c]; L.et.i2[c] = h_tab[3 * N + c];
                        L.et.bias[c] = h_bias[c] + (magic ? AYQ_MAGIC_I : 0);
                    }
                    for (int c = 0; c + 1 < N; c += 2) { memcpy(&L.et.k1x2[c >> 1], &L.et.k1[c], 8); memcpy(&L.et.k2x2[c >> 1], &L.et.k2[c], 8); }
                }
                L.ok = 1;
                return 1;
            }
        }
        pl.halo = 0; pl.n_hmma = 0;                               // not eligible after all: generic plan below
        tp.bw_log = bw_log; tp.bh_log = bh_log;
        tp.tiles_x = a.Wout >> bw_log; tp.tiles_y = a.Hout >> bh_log;
        tp.ntiles = tp.tiles_x * tp.tiles_y * ((a.n + bn - 1) / bn);
        tp.mul_x = tc_magic(tp.tiles_x); tp.mul_y = tc_magic(tp.tiles_y);
    }
    // groups -> boxes (power-of-two plane counts, so that a stage never ends on an odd chunk before the tile's last stage)
    pl.nstages = 0; pl.nops = 0; pl.stride = a.stride; pl.slot_chunks = slot_chunks; pl.a_slot_bytes = slot_chunks * 2048;
    // stride 1: a tile row of bw pixels is 16*bw contiguous bytes -> describe (channel, x) as ONE dimension so the TMA moves
    // 256-byte rows instead of 16-byte ones (the x tap shift becomes a +-16 byte coordinate; dimension 0 cannot be strided,
    // so stride-2 convs keep the rank-5 form)
    pl.merged_cx = (a.stride == 1 && (16 << bw_log) <= 256) ? 1 : 0;
    struct MapKey { int seg, boxp; };
    MapKey keys[tc::TMA_MAX_MAPS];
    int nmaps = 0;
    int stage_chunks = 0, chunk = 0;
    auto open_stage = [&]() -> bool {
        if (pl.nstages == tc::TMA_MAX_STAGES) return false;
        tc::TmaStage& st = pl.st[pl.nstages++];
        st.op0 = pl.nops; st.nops = 0; st.nchunks = 0; st.chunk0 = chunk;
        stage_chunks = 0;
        return true;
    };
    if (!open_stage()) return 0;
    int i = 0;
    while (i < a.nkc) {
        int np = 1;
        while (i + np < a.nkc && h_kc[i + np].pad_ == h_kc[i].pad_ && h_kc[i + np].dy == h_kc[i].dy && h_kc[i + np].dx == h_kc[i].dx &&
               h_kc[i + np].plane == h_kc[i].plane + np) ++np;
        int done = 0;
        while (done < np) {                                       // split the group into power-of-two boxes; fill every slot to capacity
            if (stage_chunks == slot_chunks && !open_stage()) return 0;   // (so only the tile's last stage can hold an odd chunk count)
            int bp = 1;
            while (bp * 2 <= np - done && bp * 2 <= slot_chunks - stage_chunks) bp *= 2;
            int m = -1;
            for (int q = 0; q < nmaps; ++q) if (keys[q].seg == h_kc[i].pad_ && keys[q].boxp == bp) m = q;
            if (m < 0) {
                if (nmaps == tc::TMA_MAX_MAPS) return 0;
                m = nmaps++;
                keys[m].seg = h_kc[i].pad_; keys[m].boxp = bp;
            }
            if (pl.nops == tc::TMA_MAX_OPS) return 0;
            tc::TmaOp& op = pl.op[pl.nops++];
            op.map = m; op.dx = h_kc[i].dx; op.dy = h_kc[i].dy; op.p0 = h_kc[i].plane + done; op.dst_off = (uint32_t)stage_chunks * 2048u;
            tc::TmaStage& st = pl.st[pl.nstages - 1];
            ++st.nops; st.nchunks += bp;
            stage_chunks += bp; chunk += bp; done += bp;
        }
        i += np;
    }
    // tensor maps: rank 5 uint8 {16, W, H, n, planes}
    for (int q = 0; q < nmaps; ++q) {
        if (keys[q].seg < 0 || keys[q].seg >= nsegs) return 0;
        const TmaSeg& sg = segs[keys[q].seg];
        CUresult r;
        if (pl.merged_cx) {
            const cuuint64_t gdim[4] = {(cuuint64_t)a.Win * 16, (cuuint64_t)a.Hin, (cuuint64_t)a.n, (cuuint64_t)sg.nplanes};
            const cuuint64_t gstr[3] = {(cuuint64_t)a.Win * 16, (cuuint64_t)a.Hin * a.Win * 16, (cuuint64_t)a.n * a.Hin * a.Win * 16};
            const cuuint32_t box[4] = {(cuuint32_t)(16 << bw_log), (cuuint32_t)(1 << bh_log), (cuuint32_t)bn, (cuuint32_t)keys[q].boxp};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            r = s.encode(&L.maps.m[q], CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(sg.base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            const cuuint64_t gdim[5] = {16, (cuuint64_t)a.Win, (cuuint64_t)a.Hin, (cuuint64_t)a.n, (cuuint64_t)sg.nplanes};
            const cuuint64_t gstr[4] = {16, (cuuint64_t)a.Win * 16, (cuuint64_t)a.Hin * a.Win * 16, (cuuint64_t)a.n * a.Hin * a.Win * 16};
            const cuuint32_t box[5] = {16, (cuuint32_t)((1 << bw_log) * a.stride), (cuuint32_t)((1 << bh_log) * a.stride), (cuuint32_t)bn, (cuuint32_t)keys[q].boxp};
            const cuuint32_t estr[5] = {1, (cuuint32_t)a.stride, (cuuint32_t)a.stride, 1, 1};
            r = s.encode(&L.maps.m[q], CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, const_cast<void*>(sg.base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) return 0;
    }
    for (int q = nmaps; q < tc::TMA_MAX_MAPS; ++q) L.maps.m[q] = L.maps.m[0];
    // shared memory budget
    const size_t lut_bytes = a.epi == 0 ? (((magic && a.epi == 0) || wide) ? (size_t)AYQ_LUTREP_BYTES : (size_t)AYQ_LUT256 * 8) + 256 * AYQ_MAX_OUT_ : 0;   // sigmoid table (MAGIC2: replicated per lane) + requant byte tables
    const size_t fixed = (size_t)N * 20 + lut_bytes + 64;
    const size_t w_bytes = (size_t)tp.nkc_pad * N * 16;
    const size_t budget = (size_t)s.budget_kb * 1024;
    tp.resident_b = w_bytes <= (size_t)s.resident_kb * 1024 ? 1 : 0;
    const size_t per_slot = (size_t)slot_chunks * 2048 + (tp.resident_b ? 0 : (size_t)slot_chunks * N * 16);
    const size_t avail = budget - fixed - (tp.resident_b ? w_bytes : 0);
    int ns = (int)(avail / per_slot);
    if (ns > tc::TC_MAX_NS) ns = tc::TC_MAX_NS;
    ns &= ~3;                                                     // two pipelines, each with a private ring of ns / 2 slots (even: two producers)
    if (ns < 4) return 0;
    tp.NS = ns;
    L.smem = fixed + (tp.resident_b ? w_bytes : 0) + (size_t)ns * per_slot;
    L.grid = (unsigned)tp.ntiles < (unsigned)s.num_sms ? (unsigned)tp.ntiles : (unsigned)s.num_sms;
    if (N <= TC_CT_MAXN) {
        for (int c = 0; c < N; ++c) {
            L.et.k1[c] = fast ? h_tab[c] * h_tab[N + c] * (((magic && a.epi == 0) || wide) ? 0.00390625f : 1.f) : h_tab[c]; L.et.i1[c] = h_tab[N + c];   // MAGIC2: k1 * 2^-s1 * 2^-8
            L.et.k2[c] = fast ? h_tab[2 * N + c] * h_tab[3 * N + c] : h_tab[2 * N + c]; L.et.i2[c] = h_tab[3 * N + c];
            L.et.bias[c] = h_bias[c] + (magic ? AYQ_MAGIC_I : 0);
        }
        for (int c = 0; c + 1 < N; c += 2) { memcpy(&L.et.k1x2[c >> 1], &L.et.k1[c], 8); memcpy(&L.et.k2x2[c >> 1], &L.et.k2[c], 8); }
    }
    L.ok = 1;
    return 1;
}

static inline int tma_launch(const TmaLaunch& L, const ConvArgs& a, cudaStream_t st) {
    TmaKernel kern = tma_pick(a.cout, a.epi, L.fast);
    ConvArgs a2 = a;
    a2.gen_outs = L.gen_outs;
#ifdef AYQ_ROLE_PROF_BUILD
    static const int skip_tiles = getenv("AYQ_SKIP_TILES") ? atoi(getenv("AYQ_SKIP_TILES")) : 0;   // experiment: prologue + exit only (results are garbage)
    if (skip_tiles) {
        tc::TcParams tp0 = L.tp;
        { const long long cap = (long long)(skip_tiles - 1) * (long long)L.grid; if (cap < tp0.ntiles) tp0.ntiles = (int)cap; }   // k - 1 tiles per CTA
        return launch_k(kern, dim3(L.grid), dim3(256 + 256 * tma_eg(a.cout, a.epi)), L.smem, st, a2, tp0, L.et, L.pl, L.maps) == cudaSuccess ? 0 : -1;
    }
#endif
    return launch_k(kern, dim3(L.grid), dim3(256 + 256 * tma_eg(a.cout, a.epi)), L.smem, st, a2, L.tp, L.et, L.pl, L.maps) == cudaSuccess ? 0 : -1;
}

}  // namespace ayq
