// libbagpu: C ABI (include/bagpu.h) + host driver of the Levenberg-Marquardt loop.
//
// The LM control flow restates OptimizationAlgorithmLevenberg::solve and SparseOptimizer::optimize
// (Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-169, sparse_optimizer.cpp:354-419):
// decisions (rho, lambda, nu, the _nBad rule, the stop flag) are taken on the host in FP64 from sums the device
// reduced in a fixed order; all per-edge / per-landmark / per-camera work is in the kernels.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <atomic>
#include <chrono>

#include "../../include/bagpu.h"
#include <cub/cub.cuh>

#include "ba_kernels.cuh"
#include "schur_pairs.cuh"
#include "schur_tiles.cuh"
#include "chol.cuh"
#include "chol_parts.cuh"
#include "chol_small.cuh"
#include "pcg.cuh"
#include "pose_opt.cuh"

namespace {

struct DevBuf {                       // grow-only device buffer (the context's arena is a set of these)
    void *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {                       // grow-only pinned host buffer
    void *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (h) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        CommAbort = (decltype(CommAbort))dlsym(h, "ncclCommAbort");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
    }
};
NcclApi g_nccl;

enum { EV_BUILD = 0, EV_LINSOLVE = 1, EV_UPDATE = 2, EV_KINDS = 3 };

}  // namespace

namespace {
// Host-side planning runs over millions of observations per upload: split the landmark / observation ranges over a few
// threads (the calling SLAM thread blocks in bagpu_upload anyway). fn(thread, begin, end).
constexpr int kPlanThreads = 8;
template <class F>
void parallel_ranges(int64_t n, int64_t min_chunk, F fn) {
    static const int hw = std::max(1u, std::min((unsigned)kPlanThreads, std::thread::hardware_concurrency()));
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(hw, n / std::max<int64_t>(1, min_chunk)));
    if (nt <= 1) { fn(0, (int64_t)0, n); return; }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int t = 1; t < nt; t++) th.emplace_back([=] { fn(t, n * t / nt, n * (t + 1) / nt); });
    fn(0, (int64_t)0, n / nt);
    for (auto &x : th) x.join();
}

}  // namespace

// Development switches, read ONCE per context in bagpu_init (INTEGRATION.md lists them); none is needed in production.
struct BagpuOptions {
    bool debug = false;          // BAGPU_DEBUG: progress and timing lines on stderr
    bool no_twoway = false;      // BAGPU_NO_TWOWAY: one factorisation front instead of two
    bool compare = false;        // BAGPU_COMPARE: check the reduced system element-wise against the v1 global-atomic build
    bool no_tiles = false;       // BAGPU_NO_TILES: v1 build_kernel / update_kernel instead of stage + pair
    bool no_overlap = false;     // BAGPU_NO_OVERLAP (or a tool that serialises kernels is attached): solve after pair_kernel
    int  stage_first = 2;        // BAGPU_STAGE_FIRST: launch order of stage_kernel and the factorisation clusters
    bool update_relin = false;   // BAGPU_UPDATE_RELIN: re-linearising update kernel instead of update_z_kernel
    bool no_band = false;        // BAGPU_NO_BAND: tiled chol_solve_kernel even for narrow envelopes
    bool no_cluster = false;     // BAGPU_NO_CLUSTER: cooperative launch instead of one cluster for chol_solve_kernel
    bool no_back3 = false;       // BAGPU_NO_BACK3: two-buffer backward substitution
    int  parts = 0;              // BAGPU_PARTS: number of partitions of the partitioned band solver (0 = automatic, 1 = off)
    bool sep_tiled = false;      // BAGPU_SEP_TILED: separator system through the tiled band solver instead of block cyclic reduction
    bool small_chol = false;     // BAGPU_SMALL_CHOL: chained loop solves systems of n <= 224 through chol_small_kernel (one CTA, one barrier per column;
                                 // measured slower than the cluster kernel: 130 against 62 us for n = 120, 268 against 90 us for n = 180 -- off)
    bool no_chain = false;       // BAGPU_NO_CHAIN: host-stepped LM loop (one status read per trial) even for small maps
    long long chain_obs = 600000; // BAGPU_CHAIN_OBS: largest map (observations) whose LM loop runs chained on the device
    bool tile_fma = false;       // BAGPU_TILE_FMA: the tile contraction with FMAs in registers (pair_tile_kernel) instead of the FP64 tensor pipe (pair_tile_mma_kernel)
    bool chunks = false;         // BAGPU_CHUNKS: multi-GPU trial with the rows of the reduced system accumulated, all-reduced and factored in three chunks of fronts
                                 // (measured on 2 GPUs, config 5: 81.6 ms per step against 78.8 with one all-reduce: every front is as long as the others, so a full
                                 // front still follows the last collective, and the early fronts take SMs from pair_tile_mma_kernel -- off by default)
    bool spike_v1 = false;       // BAGPU_SPIKE_V1: spike_forward_kernel without the cp.async prefetch of the panel operands
    bool pair_list = false;      // BAGPU_PAIR_LIST: pair_kernel over the per-pair entry list instead of pair_tile_kernel over (tile, landmark) records
    void read() {
        auto on = [](const char *k) { return getenv(k) != nullptr; };
        debug = on("BAGPU_DEBUG"); no_twoway = on("BAGPU_NO_TWOWAY"); compare = on("BAGPU_COMPARE"); no_tiles = on("BAGPU_NO_TILES");
        // tools that serialise kernel launches (ncu, compute-sanitizer) would leave the Cholesky spinning on pair_kernel
        no_overlap = on("BAGPU_NO_OVERLAP") || compare || no_tiles || on("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || on("CUDA_INJECTION64_PATH") ||
                     on("NV_NSIGHT_INJECTION_PORT_BASE") || on("CUDA_LAUNCH_BLOCKING");
        if (getenv("BAGPU_STAGE_FIRST")) stage_first = atoi(getenv("BAGPU_STAGE_FIRST"));
        update_relin = on("BAGPU_UPDATE_RELIN"); no_band = on("BAGPU_NO_BAND"); no_cluster = on("BAGPU_NO_CLUSTER"); no_back3 = on("BAGPU_NO_BACK3");
        if (getenv("BAGPU_PARTS")) parts = atoi(getenv("BAGPU_PARTS"));
        sep_tiled = on("BAGPU_SEP_TILED"); pair_list = on("BAGPU_PAIR_LIST"); spike_v1 = on("BAGPU_SPIKE_V1"); chunks = on("BAGPU_CHUNKS"); tile_fma = on("BAGPU_TILE_FMA"); no_chain = on("BAGPU_NO_CHAIN"); small_chol = on("BAGPU_SMALL_CHOL");
        if (getenv("BAGPU_CHAIN_OBS")) chain_obs = atoll(getenv("BAGPU_CHAIN_OBS"));
    }
};

// Partitioned band solver (chol_parts.cuh): host-side plan and buffers of one reduced camera system.
struct PartPlan {
    bool on = false;
    int P = 0, w = 0, n = 0, ld = 0;
    PartTable T;
    int nc = 1, maxr = 0, n_max = 0;                     // cluster size, reach and largest sub-system of the factorisation launch
    int nS = 0, ldM = 1, gridM = 1, maxrM = 0;           // separator system
    int spike_ctas = 0, inv_panels = 0, gram_ctas = 0, apply_rows = 0;
    std::vector<int> h_subce, h_seprow, h_ceM;           // kept alive: copied without a sync
    std::vector<CholArgs> h_tab;                         // [0, P): factorisation, [P, 2P): backward substitution
    DevBuf d_tab, d_subce, d_seprow, d_V, d_linv, d_Dp, d_Ep, d_gp, d_SM, d_rhsM, d_zeroM, d_yM, d_xM, d_ceM, d_dinvM;
    // separator system by block cyclic reduction: dense blocks, per-level launch lists
    bool cr = false; int cr_nc = 1;
    CrPlan C;
    std::vector<int> cr_stride, cr_cnt, cr_off, h_cedense;     // per level: stride s, eliminated blocks, offset into the tables
    std::vector<CholArgs> h_crtab;                             // [0, K): factorisation in level order, [K, 2K): backward substitution
    DevBuf d_crtab, d_Dd, d_Cc, d_gg, d_yy, d_Fa, d_Fb, d_linvB, d_cedense;
    void release() {
        DevBuf *b[] = {&d_tab, &d_subce, &d_seprow, &d_V, &d_linv, &d_Dp, &d_Ep, &d_gp, &d_SM, &d_rhsM, &d_zeroM, &d_yM, &d_xM, &d_ceM, &d_dinvM,
                       &d_crtab, &d_Dd, &d_Cc, &d_gg, &d_yy, &d_Fa, &d_Fb, &d_linvB, &d_cedense};
        for (DevBuf *x : b) x->release();
    }
};

struct bagpu_ctx {
    int device = 0;
    int n_sm = 148;
    BagpuOptions opt;
    size_t band_smem_cap = 200 * 1024;     // dynamic shared memory chol_band_kernel may use on this device
    bool overlap_off = false;              // set when the overlapped solve starved once (watchdog): the solve then runs after pair_kernel
    cudaStream_t stream = nullptr;
    cudaStream_t stream_chol = nullptr;    // the band Cholesky runs beside pair_kernel on its own (high-priority) stream
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t stream_chol2 = nullptr;   // second half of the two-way factorisation
    cudaEvent_t ev_tw[4] = {nullptr, nullptr, nullptr, nullptr};
    // multi-GPU trial of the partitioned solver: the rows of the reduced system are accumulated, all-reduced and factored chunk by chunk
    // (chunk = a group of fronts): the collective of chunk k and its factorisation fronts run while pair_tile_mma_kernel is still on k + 1
    static constexpr int kChunks = 3;
    cudaStream_t stream_comm = nullptr, stream_front[kChunks] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_chunk[3 * kChunks + 2] = {};
    struct ChunkPlan { bool on = false; int n = 0; int part0[kChunks + 1]; int item0[kChunks + 1]; size_t elem0[kChunks + 1]; } chunks;
    struct TwoWay {                        // band factored from both ends towards a separator block M
        bool on = false;
        int k = 0, rT = 0, n1 = 0, n2 = 0, nM = 0, ldM = 0;
        int grid1 = 1, maxr1 = 0, grid2 = 1, maxr2 = 0, gridM = 1, maxrM = 0;
        size_t s2_elems = 0, sM_elems = 0;
    } tw;
    PartPlan parts;                        // more than two fronts (long keyframe chains)
    DevBuf d_pcg_vec, d_pcg_minv, d_pcg_part, d_pcg_scal;   // BAGPU_SOLVER_PCG work space
    DevBuf d_colend1, d_colend2, d_colendM, d_y2, d_SM, d_rhsM, d_zeroM, d_yM, d_xM, d_rowpos, d_rowofpos;
    char err[512] = {0};
    // ---- communicator (multi-GPU global BA)
    ncclComm_t comm = nullptr; int world = 1, rank = 0;
    // ---- BA problem state
    bool have_problem = false;
    int n_poses = 0, n_points = 0, n_free = 0, n_cams = 0, n_rigs = 0;
    int64_t n_obs = 0;
    bool identity_perm = true;
    DevBuf d_lm_ptr, d_o_pose, d_o_point, d_o_meta, d_o_u, d_o_v, d_o_ur, d_o_w, d_cams, d_rigs, d_hidx, d_perm;
    DevBuf d_raw8a, d_raw8b, d_raw16a, d_raw16b, d_rawd;     // raw upload staging on the device
    DevBuf d_pose_a, d_pose_b, d_pose_init, d_pt_a, d_pt_b, d_pt_init, d_meta_init;
    DevBuf d_sys;                      // [S (n*ld) | bp (n) | bs (n) | S2 (two-way: mirrored lower half) | hpp_diag (n)]; S .. S2 contiguous (one all-reduce)
    DevBuf d_y, d_colend, d_dinv, d_widelist, d_tasks;
    DevBuf d_Z, d_Dr, d_entries, d_items, d_pk_keys, d_pk_keys2, d_pk_vals, d_npairs, d_pairoff, d_blkcnt, d_blkoff, d_itemcnt, d_itemoff, d_cubtmp, d_rowdone, d_part, d_blkdone, d_Lm;
    int n_wide = 0, n_tasks = 0, stage_grid = 1, stage_wide_grid = 1, upd_grid = 1, updz_grid = 1, parts_stride = 1;
    int n_items = 0, pair_grid = 1, pair_occ = 1, stage_occ = 1; long long n_entries = 0;
    // tile-major Schur pass (schur_tiles.cuh): records sorted by tile, tile order, what the factorisation clusters wait on
    bool tiles = false; int ntile = 0, tbw1 = 1, diag_grid = 1;
    DevBuf d_tp_raw, d_tp_recs, d_tp_idx, d_tp_idx2, d_tile_pos, d_tile_of_pos, d_cam_tpos, d_tp_work;
    int wait_bw1 = 1; const int *wait_rowpos = nullptr;    // item_off is indexed [wait_rowpos[camera] * wait_bw1 ..] by chol_band_kernel
    size_t s_elems = 0, scratch_elems = 0; int chol_grid = 1; int chol_maxr = 0; int band_blocks = 0;
    DevBuf d_xp, d_parts, d_status, d_chi2, d_depth, d_out_chi2, d_out_u8a, d_out_u8b, d_fail, d_count;
    PinBuf h_status, h_stage;
    DevBuf d_lm, d_lm_trace; PinBuf h_lm;
    DevBuf d_lastrow; PinBuf h_lastrow, h_lmptr; cudaEvent_t ev_lastrow = nullptr;   // upload: envelope computed on the device while the host checks the order      // chained mode: LM state and per-iteration records on the device
    std::vector<int> h_hidx;
    std::vector<int> h_colend, h_ce1, h_ce2, h_ceM;      // envelope arrays: kept alive, copied on the main stream without a sync
    double *pose_cur = nullptr, *pose_trial = nullptr, *pt_cur = nullptr, *pt_trial = nullptr;
    int n_sys = 0, ld = 0;
    int build_grid = 0;
    // ---- pose batch state
    bool have_pose = false;
    int pb_frames = 0; int64_t pb_obs = 0;
    DevBuf p_pose0, p_ptr, p_cams, p_rigs, p_xw, p_meta, p_u, p_v, p_ur, p_w, p_chi2, p_out, p_pose_out, p_ninl, p_fchi;
    double pb_delta_mono = 0, pb_delta_stereo = 0; float pb_gate_mono = 0, pb_gate_stereo = 0;
    // ---- timing
    bagpu_timing tm;
    std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
    struct Pending { cudaEvent_t a, b; int kind; };
    std::vector<Pending> pending;
    cudaEvent_t ev_phase[4] = {nullptr, nullptr, nullptr, nullptr};
};

namespace {

// where the pieces of d_sys live: [S | bp | bs | S2 | y (n) | y2 (n2) | yM (nM) | row_done (nf x u32) | fail (4 x i32) | hpp_diag (n)]
struct SysLayout { double *S, *bp, *bs, *S2, *y1, *y2, *yM, *hpp; unsigned *row_done; int *fail; size_t sys_count; };
SysLayout sys_layout(const bagpu_ctx *ctx) {
    SysLayout L;
    const size_t n1 = (size_t)std::max(1, ctx->n_sys);
    L.S = ctx->d_sys.as<double>();
    L.bp = L.S + ctx->s_elems; L.bs = L.bp + n1; L.S2 = L.bs + n1;
    L.sys_count = ctx->s_elems + 2 * n1 + ctx->tw.s2_elems;                    // what a trial all-reduces
    L.y1 = L.S2 + ctx->tw.s2_elems; L.y2 = L.y1 + n1; L.yM = L.y2 + ctx->tw.n2;
    L.row_done = reinterpret_cast<unsigned *>(L.yM + ctx->tw.nM);
    L.fail = reinterpret_cast<int *>(L.yM + ctx->tw.nM + ((size_t)std::max(1, ctx->n_free) + 1) / 2);
    L.hpp = L.S + L.sys_count + ctx->scratch_elems;
    return L;
}

int fail(bagpu_ctx *c, int code, const char *fmt, ...) {
    if (c) {
        va_list ap; va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) return fail(ctx, BAGPU_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)
#define CKN(call)                                                                                         \
    do {                                                                                                  \
        ncclResult_t r__ = (call);                                                                        \
        if (r__ != ncclSuccess) return fail(ctx, BAGPU_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error"); \
    } while (0)

void normalize_qt(const double *in, double *out) {      // SE3Quat(q,t) ctor: flip w<0, normalise (se3quat.h:62-64,280-285)
    double x = in[3], y = in[4], z = in[5], w = in[6];
    if (w < 0) { x *= -1; y *= -1; z *= -1; w *= -1; }
    const double n = std::sqrt(x * x + y * y + z * z + w * w);
    out[0] = in[0]; out[1] = in[1]; out[2] = in[2];
    out[3] = x / n; out[4] = y / n; out[5] = z / n; out[6] = w / n;
}

cudaEvent_t get_event(bagpu_ctx *ctx) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e; cudaEventCreate(&e); ctx->ev_pool.push_back(e);
    }
    return ctx->ev_pool[ctx->ev_used++];
}
struct ScopedEv {
    bagpu_ctx *c; cudaEvent_t a, b; int kind; cudaStream_t s;
    ScopedEv(bagpu_ctx *ctx, int k, cudaStream_t stream = nullptr) : c(ctx), kind(k), s(stream ? stream : ctx->stream) { a = get_event(ctx); b = get_event(ctx); cudaEventRecord(a, s); }
    ~ScopedEv() { cudaEventRecord(b, s); c->pending.push_back({a, b, kind}); }
};
void resolve_events(bagpu_ctx *ctx) {      // call after a stream sync
    for (auto &p : ctx->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            if (p.kind == EV_BUILD) { ctx->tm.build_ms += ms; ctx->tm.build_launches++; }
            else if (p.kind == EV_LINSOLVE) { ctx->tm.linsolve_ms += ms; ctx->tm.linsolve_launches++; }
            else { ctx->tm.update_ms += ms; ctx->tm.update_launches++; }
        }
    }
    ctx->pending.clear();
    ctx->ev_used = 0;
}

// compose the observation meta word on the device from the caller's raw arrays
__global__ void compose_meta_kernel(int64_t n, const uint8_t *kind, const int16_t *cam, const int16_t *rig, const uint8_t *flags, uint32_t *meta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = (rig[i] < 0) ? 255u : (uint32_t)rig[i];
    meta[i] = (uint32_t)kind[i] | ((uint32_t)cam[i] << 2) | (r << 10) | ((flags[i] & BAGPU_FLAG_ROBUST) ? META_ROBUST : 0u);
}
template <typename T>
__global__ void gather_perm_kernel(int64_t n, const int *__restrict__ perm, const T *__restrict__ in, T *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}
// ---- two-way band factorisation: glue between the two half-systems and the separator block M (rows rT .. rT + nM - 1)
// S_M = S1[M,M] + mirror(S2[M,M]) (partial Schur complements, no lambda), rhs_M = b_p + b_s + forward updates of both halves
__global__ void tw_merge_kernel(int nM, int rT, int n, int ld, const double *__restrict__ S1, const double *__restrict__ S2,
                                const double *__restrict__ bp, const double *__restrict__ bs, const double *__restrict__ y1,
                                const double *__restrict__ y2, double *SM, int ldM, double *rhsM) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nM * nM) return;
    const int R = idx / nM, C = idx - R * nM;
    if (C < R) return;
    double v = 0.0;
    if (C - R <= ld) {
        const int Ro = rT + R, Co = rT + C;
        v = S1[(size_t)Ro * ld + Co] + S2[(size_t)(n - 1 - Co) * ld + (n - 1 - Ro)];
    }
    SM[(size_t)R * ldM + C] = v;
    if (R == C) { const int Ro = rT + R; rhsM[R] = bp[Ro] + bs[Ro] + y1[Ro] + y2[n - 1 - Ro]; }
}
// x_M into the final x and into the right-hand sides of the two backward substitutions
__global__ void tw_scatter_kernel(int nM, int rT, int n, const double *__restrict__ xM, double *y1, double *y2, double *x) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nM) return;
    const int Ro = rT + k;
    const double v = xM[k];
    y1[Ro] = v; y2[n - 1 - Ro] = v; x[Ro] = v;
}
// pos -> row and row -> pos: identity, or 0, nf-1, 1, nf-2, ... (two-way factorisation)
__global__ void row_order_kernel(int nf, int zigzag, int *row_pos, int *row_of_pos) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nf) return;
    int pos = a;
    if (zigzag) { const int half = (nf + 1) / 2; pos = (a < half) ? 2 * a : 2 * (nf - 1 - a) + 1; }
    row_pos[a] = pos; row_of_pos[pos] = a;
}
// FP64 issue-rate probes for the roofline of the Schur pass (K3 is FP64-bound): 8 independent DFMA chains per
// thread, and 8 independent m8n8k4 FP64 MMA accumulators per warp, whole device, timed with CUDA events by bagpu_test_fp64_peak.
__global__ void __launch_bounds__(1024) fp64_dfma_probe_kernel(double *o, int iters, double seed) {
    const double a = seed + threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-12;
    double c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = k * 1e-3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = fma(a, b, c[k]);
    }
    double sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) sum += c[k];
    if (sum == 12345.678) o[blockIdx.x * blockDim.x + threadIdx.x] = sum;      // keeps the chains alive, never true
}
__global__ void __launch_bounds__(1024) fp64_dmma_probe_kernel(double *o, int iters, double seed) {
    const double a = seed + threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-12;
    double c[16];
#pragma unroll
    for (int k = 0; k < 16; k++) c[k] = 0.0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[2 * k]), "+d"(c[2 * k + 1]) : "d"(a), "d"(b));
    }
    double sum = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) sum += c[k];
    if (sum == 12345.678) o[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}
__global__ void atan2f_test_kernel(int64_t n, const float *y, const float *x, float *o) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = baf_atan2f(y[i], x[i]);
}

// envelope of the reduced system from the landmark structure: lastrow[h] = largest free index among the poses that share a landmark with
// free pose h (observations landmark-major, pose-ascending: the partner with the largest index is the last free observation of the run)
__global__ void lastrow_kernel(int64_t n_obs, int n_poses, const int *__restrict__ o_pose, const int *__restrict__ o_point, const int *__restrict__ hidx, int *lastrow) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_obs) return;
    // (runs on the arrays as uploaded, BEFORE the host has validated them: an index out of range is skipped here and reported there)
    const int ip = o_pose[e];
    if ((unsigned)ip >= (unsigned)n_poses) return;
    const int h = hidx[ip];
    if (h < 0) return;
    const int j = o_point[e];
    int mx = h;
    for (int64_t e2 = e + 1; e2 < n_obs && o_point[e2] == j; e2++) {
        const int i2 = o_pose[e2];
        if ((unsigned)i2 < (unsigned)n_poses) mx = max(mx, hidx[i2]);
    }
    atomicMax(lastrow + h, mx);
}

inline int grid_for(int64_t n, int threads) { return (int)std::max<int64_t>(1, (n + threads - 1) / threads); }

BaDev make_dev(bagpu_ctx *ctx, double delta_mono, double delta_stereo) {
    BaDev D;
    D.n_points = ctx->n_points; D.n_poses = ctx->n_poses; D.n_free = ctx->n_free; D.n_obs = ctx->n_obs;
    D.lm_ptr = ctx->d_lm_ptr.as<int>(); D.o_pose = ctx->d_o_pose.as<int>(); D.o_point = ctx->d_o_point.as<int>();
    D.o_meta = ctx->d_o_meta.as<uint32_t>();
    D.o_u = ctx->d_o_u.as<double>(); D.o_v = ctx->d_o_v.as<double>(); D.o_ur = ctx->d_o_ur.as<double>(); D.o_w = ctx->d_o_w.as<double>();
    D.cams = ctx->d_cams.as<bagpu_camera>(); D.rigs = ctx->d_rigs.as<double>(); D.pose_hidx = ctx->d_hidx.as<int>();
    D.delta_mono = delta_mono; D.delta_stereo = delta_stereo;
    return D;
}

int validate_problem(bagpu_ctx *ctx, const bagpu_problem *p) {
    if (!p) return fail(ctx, BAGPU_ERR_ARG, "problem is NULL");
    if (p->n_poses <= 0 || p->n_points <= 0 || p->n_obs <= 0 || p->n_cameras <= 0) return fail(ctx, BAGPU_ERR_ARG, "empty problem (poses %d points %d obs %lld cameras %d)", p->n_poses, p->n_points, (long long)p->n_obs, p->n_cameras);
    if (p->n_obs >= (int64_t)INT32_MAX) return fail(ctx, BAGPU_ERR_ARG, "n_obs too large for one device shard");
    if (!p->pose_qt || !p->pose_fixed || !p->points || !p->cameras || !p->obs_pose || !p->obs_point || !p->obs_cam || !p->obs_rig ||
        !p->obs_kind || !p->obs_flags || !p->obs_u || !p->obs_v || !p->obs_inv_sigma2)
        return fail(ctx, BAGPU_ERR_ARG, "NULL array in problem");
    if (p->n_cameras > 255 || p->n_rigs > 254) return fail(ctx, BAGPU_ERR_ARG, "too many cameras/rigs");
    return BAGPU_OK;
}
// the per-observation checks: one pass, split over a few host threads (2 M observations: 0.5 ms instead of 3); bagpu_upload runs it
// while the observation arrays are already crossing PCIe (the bytes of a rejected problem are never used)
// (the same pass checks the order -- landmark-major, pose-ascending inside a landmark -- and notes where each landmark's run starts)
int validate_observations(bagpu_ctx *ctx, const bagpu_problem *p, std::vector<int> &lm_ptr, bool &sorted_out) {
    std::atomic<int> any_stereo_a{0};
    std::atomic<bool> sorted_a{true};
    std::atomic<long long> bad_e{-1};
    std::atomic<int> bad_why{0};
    const int n_poses = p->n_poses, n_points = p->n_points, n_cameras = p->n_cameras, n_rigs = p->n_rigs;
    const int *op = p->obs_pose, *opt = p->obs_point; const int16_t *oc = p->obs_cam, *org = p->obs_rig; const uint8_t *ok = p->obs_kind;
    parallel_ranges(p->n_obs, 1 << 16, [&](int, int64_t e0, int64_t e1) {
        // branch-free first (the compiler vectorises these reductions); the exact culprit is looked up only when something is wrong
        unsigned bad = 0u, st = 0u, uns = 0u;
        for (int64_t e = e0; e < e1; e++) {
            bad |= (unsigned)((unsigned)op[e] >= (unsigned)n_poses) | (unsigned)((unsigned)opt[e] >= (unsigned)n_points) |
                   (unsigned)((unsigned)(int)oc[e] >= (unsigned)n_cameras) | (unsigned)(ok[e] > 2) |
                   (unsigned)((ok[e] == BAGPU_EDGE_BODY) & ((unsigned)(int)org[e] >= (unsigned)n_rigs));
            st |= (unsigned)(ok[e] == BAGPU_EDGE_STEREO);
        }
        if (bad) {
            for (int64_t e = e0; e < e1; e++) {
                int why = 0;
                if (op[e] < 0 || op[e] >= n_poses || opt[e] < 0 || opt[e] >= n_points) why = 1;
                else if (oc[e] < 0 || oc[e] >= n_cameras) why = 2;
                else if (ok[e] > 2) why = 3;
                else if (ok[e] == BAGPU_EDGE_BODY && (org[e] < 0 || org[e] >= n_rigs)) why = 4;
                if (why) { long long exp = -1; if (bad_e.compare_exchange_strong(exp, (long long)e)) bad_why.store(why); return; }
            }
        }
        for (int64_t e = std::max<int64_t>(e0, 1); e < e1; e++)
            uns |= (unsigned)(opt[e - 1] > opt[e]) | (unsigned)((opt[e - 1] == opt[e]) & (op[e - 1] > op[e]));
        if (e0 == 0) lm_ptr[(size_t)opt[0]] = 0;
        for (int64_t e = std::max<int64_t>(e0, 1); e < e1; e++)
            if (opt[e - 1] != opt[e]) lm_ptr[(size_t)opt[e]] = (int)e;           // only meaningful (and only used) when the order holds
        if (uns) sorted_a.store(false, std::memory_order_relaxed);
        if (st) any_stereo_a.store(1, std::memory_order_relaxed);
    });
    sorted_out = sorted_a.load();
    if (bad_e.load() >= 0) {
        static const char *msg[] = {"", "vertex index out of range", "camera index out of range", "bad kind", "rig index out of range"};
        return fail(ctx, BAGPU_ERR_ARG, "observation %lld: %s", bad_e.load(), msg[bad_why.load()]);
    }
    const bool any_stereo = any_stereo_a.load() != 0;
    if (any_stereo && !p->obs_ur) return fail(ctx, BAGPU_ERR_ARG, "stereo edges need obs_ur");
    return BAGPU_OK;
}

}  // namespace

// ======================================================================================= API
extern "C" {

const char *bagpu_strerror(int code) {
    switch (code) {
        case BAGPU_OK: return "ok";
        case BAGPU_TERMINATE_TRIALS: return "LM terminated: 10 failed trials or rho == 0";
        case BAGPU_TERMINATE_NBAD: return "LM terminated: relative chi2 decrease < 1e-3 for 3 iterations";
        case BAGPU_STOPPED: return "stopped by the caller's flag";
        case BAGPU_ERR_CUDA: return "CUDA error";
        case BAGPU_ERR_ARG: return "bad argument";
        case BAGPU_ERR_NCCL: return "NCCL error";
        case BAGPU_ERR_NO_DEVICE: return "no CUDA device";
        case BAGPU_ERR_ALLOC: return "allocation failed";
        default: return "unknown status";
    }
}
const char *bagpu_last_error(const bagpu_ctx *ctx) { return ctx ? ctx->err : "no context"; }

int bagpu_init(int device_id, bagpu_ctx **out) {
    if (!out) return BAGPU_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return BAGPU_ERR_NO_DEVICE;   // no CPU fallback, by design
    bagpu_ctx *ctx = new (std::nothrow) bagpu_ctx();
    if (!ctx) return BAGPU_ERR_ALLOC;
    if (device_id < 0) { if (cudaGetDevice(&device_id) != cudaSuccess) device_id = 0; }
    if (device_id >= ndev) { delete ctx; return BAGPU_ERR_NO_DEVICE; }
    ctx->device = device_id;
    ctx->opt.read();
    if (cudaSetDevice(device_id) != cudaSuccess) { delete ctx; return BAGPU_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) == cudaSuccess) ctx->n_sm = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return BAGPU_ERR_CUDA; }
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&ctx->stream_chol, cudaStreamNonBlocking, hi) != cudaSuccess) { cudaStreamDestroy(ctx->stream); delete ctx; return BAGPU_ERR_CUDA; }
    }
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi); cudaStreamCreateWithPriority(&ctx->stream_chol2, cudaStreamNonBlocking, hi); }
    for (int i = 0; i < 4; i++) cudaEventCreateWithFlags(&ctx->ev_tw[i], cudaEventDisableTiming);
    {
        int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamCreateWithPriority(&ctx->stream_comm, cudaStreamNonBlocking, hi);
        for (int i = 0; i < bagpu_ctx::kChunks; i++) cudaStreamCreateWithPriority(&ctx->stream_front[i], cudaStreamNonBlocking, hi);
        for (auto &e : ctx->ev_chunk) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    }
    for (int i = 0; i < 4; i++) cudaEventCreate(&ctx->ev_phase[i]);
    cudaEventCreateWithFlags(&ctx->ev_lastrow, cudaEventDisableTiming);
    {
        // Load every kernel now. With CUDA's lazy module loading the FIRST launch of a kernel may synchronise the context; a
        // launch that does so while the Cholesky clusters spin on pair_kernel's counters (pair_kernel not yet enqueued) would
        // block the host until the watchdog fires.
        cudaFuncAttributes fa;
        const void *fns[] = {(const void *)compose_meta_kernel, (const void *)gather_perm_kernel<int>, (const void *)gather_perm_kernel<double>,
                             (const void *)gather_perm_kernel<uint32_t>, (const void *)lastrow_kernel, (const void *)tw_merge_kernel, (const void *)tw_scatter_kernel, (const void *)row_order_kernel,
                             (const void *)atan2f_test_kernel, (const void *)fp64_dfma_probe_kernel, (const void *)fp64_dmma_probe_kernel, (const void *)build_kernel, (const void *)update_kernel, (const void *)update_packed_kernel, (const void *)update_z_kernel,
                             (const void *)gate_kernel, (const void *)count_active_kernel, (const void *)pose_update_kernel, (const void *)reduce_partials_kernel, (const void *)finish_trial_kernel,
                             (const void *)scatter_perm_kernel<double>, (const void *)scatter_perm_kernel<uint8_t>, (const void *)level_from_meta_kernel,
                             (const void *)pair_count_kernel, (const void *)pair_gen_kernel, (const void *)pair_item_count_kernel, (const void *)pair_item_fill_kernel,
                             (const void *)stage_kernel, (const void *)stage_wide_kernel, (const void *)pair_kernel, (const void *)chol_band_kernel,
                             (const void *)chol_solve_kernel<true>, (const void *)chol_solve_kernel<false>, (const void *)pose_opt_kernel,
                             (const void *)panel_inverse_kernel, (const void *)spike_forward_kernel, (const void *)spike_forward2_kernel, (const void *)spike_gram_kernel, (const void *)sep_assemble_kernel,
                             (const void *)sep_scatter_kernel, (const void *)spike_apply_kernel, (const void *)row_order_parts_kernel,
                             (const void *)pcg_prec_kernel, (const void *)pcg_init_kernel, (const void *)pcg_init_finish_kernel, (const void *)pcg_spmv_kernel, (const void *)pcg_update_kernel, (const void *)pcg_dir_kernel,
                             (const void *)chol_small_kernel, (const void *)lm_init_kernel, (const void *)lm_decide_kernel, (const void *)pair_tile_kernel, (const void *)pair_tile_mma_kernel, (const void *)tile_diag_kernel, (const void *)tile_order_kernel, (const void *)tile_plan_kernel<false>, (const void *)tile_plan_kernel<true>, (const void *)tile_gather_kernel,
                             (const void *)tile_item_count_kernel, (const void *)tile_item_fill_kernel, (const void *)cr_assemble_kernel, (const void *)block_inverse_kernel, (const void *)block_spike_kernel, (const void *)block_gram_kernel, (const void *)block_apply_kernel};
        for (const void *f : fns) if (cudaFuncGetAttributes(&fa, f) != cudaSuccess) { cudaGetLastError(); }
        // Function attributes are PER DEVICE: every context sets them for its own device (idempotent, no process-wide flag),
        // so a second context on another GPU of the same process gets its large dynamic shared memory and cluster sizes too.
        if (cudaFuncGetAttributes(&fa, chol_band_kernel) == cudaSuccess) ctx->band_smem_cap = 232448 - fa.sharedSizeBytes - 1024;
        bool ok_attr = true;
        ok_attr &= cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM_BYTES) == cudaSuccess;
        ok_attr &= cudaFuncSetAttribute(pair_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM_BYTES) == cudaSuccess;
        ok_attr &= cudaFuncSetAttribute(pair_tile_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM_BYTES) == cudaSuccess;
        ok_attr &= cudaFuncSetAttribute(chol_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * CH_MAX_SMEM_N)) == cudaSuccess;
        ok_attr &= cudaFuncSetAttribute(chol_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * CH_MAX_SMEM_N)) == cudaSuccess;
        ok_attr &= cudaFuncSetAttribute(chol_solve_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        ok_attr &= cudaFuncSetAttribute(chol_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->band_smem_cap) == cudaSuccess;
        ok_attr &= cudaFuncSetAttribute(chol_band_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        ok_attr &= cudaFuncSetAttribute(chol_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_small_smem(CS_MAX_N)) == cudaSuccess;
        ok_attr &= cudaFuncSetAttribute(spike_forward2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->band_smem_cap) == cudaSuccess;
        if (!ok_attr) { cudaGetLastError(); bagpu_destroy(ctx); return BAGPU_ERR_CUDA; }
    }
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    *out = ctx;
    return BAGPU_OK;
}

void bagpu_destroy(bagpu_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    // Teardown must not depend on the peers: a rank that leaves early (an exception in the host code, a failed assertion in a test)
    // would otherwise block in ncclCommDestroy while the others wait for it elsewhere. Every operation of this context has completed
    // (stream synchronised above), so aborting the communicator only frees its resources.
    if (ctx->comm) { if (g_nccl.CommAbort) g_nccl.CommAbort(ctx->comm); else if (g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm); }
    DevBuf *bufs[] = {&ctx->d_lm_ptr, &ctx->d_o_pose, &ctx->d_o_point, &ctx->d_o_meta, &ctx->d_o_u, &ctx->d_o_v, &ctx->d_o_ur, &ctx->d_o_w,
                      &ctx->d_cams, &ctx->d_rigs, &ctx->d_hidx, &ctx->d_perm, &ctx->d_raw8a, &ctx->d_raw8b, &ctx->d_raw16a, &ctx->d_raw16b,
                      &ctx->d_rawd, &ctx->d_pose_a, &ctx->d_pose_b, &ctx->d_pose_init, &ctx->d_pt_a, &ctx->d_pt_b, &ctx->d_pt_init, &ctx->d_meta_init, &ctx->d_sys, &ctx->d_xp, &ctx->d_y, &ctx->d_colend, &ctx->d_dinv, &ctx->d_widelist, &ctx->d_tasks, &ctx->d_Z, &ctx->d_Dr, &ctx->d_entries, &ctx->d_items, &ctx->d_pk_keys, &ctx->d_pk_keys2, &ctx->d_pk_vals,
                      &ctx->d_npairs, &ctx->d_pairoff, &ctx->d_blkcnt, &ctx->d_blkoff, &ctx->d_itemcnt, &ctx->d_itemoff, &ctx->d_cubtmp, &ctx->d_rowdone, &ctx->d_part, &ctx->d_blkdone, &ctx->d_Lm,
                      &ctx->d_parts, &ctx->d_status, &ctx->d_chi2, &ctx->d_depth, &ctx->d_out_chi2, &ctx->d_out_u8a, &ctx->d_out_u8b,
                      &ctx->d_fail, &ctx->d_count, &ctx->p_pose0, &ctx->p_ptr, &ctx->p_cams, &ctx->p_rigs, &ctx->p_xw, &ctx->p_meta,
                      &ctx->p_u, &ctx->p_v, &ctx->p_ur, &ctx->p_w, &ctx->p_chi2, &ctx->p_out, &ctx->p_pose_out, &ctx->p_ninl, &ctx->p_fchi};
    for (DevBuf *b : bufs) b->release();
    ctx->h_status.release(); ctx->h_stage.release(); ctx->d_lm.release(); ctx->d_lm_trace.release(); ctx->h_lm.release();
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < 4; i++) if (ctx->ev_phase[i]) cudaEventDestroy(ctx->ev_phase[i]);
    for (int i = 0; i < 4; i++) if (ctx->ev_tw[i]) cudaEventDestroy(ctx->ev_tw[i]);
    if (ctx->stream_chol2) { cudaStreamSynchronize(ctx->stream_chol2); cudaStreamDestroy(ctx->stream_chol2); }
    if (ctx->stream_comm) { cudaStreamSynchronize(ctx->stream_comm); cudaStreamDestroy(ctx->stream_comm); }
    for (auto &f : ctx->stream_front) if (f) { cudaStreamSynchronize(f); cudaStreamDestroy(f); }
    for (auto &e : ctx->ev_chunk) if (e) cudaEventDestroy(e);
    { DevBuf *tb[] = {&ctx->d_colend1, &ctx->d_colend2, &ctx->d_colendM, &ctx->d_y2, &ctx->d_SM, &ctx->d_rhsM, &ctx->d_zeroM, &ctx->d_yM, &ctx->d_xM, &ctx->d_rowpos, &ctx->d_rowofpos};
      for (DevBuf *x : tb) x->release(); }
    ctx->parts.release();
    { DevBuf *tb2[] = {&ctx->d_tp_raw, &ctx->d_tp_recs, &ctx->d_tp_idx, &ctx->d_tp_idx2, &ctx->d_tile_pos, &ctx->d_tile_of_pos, &ctx->d_cam_tpos, &ctx->d_tp_work}; for (DevBuf *x : tb2) x->release(); }
    ctx->d_pcg_vec.release(); ctx->d_pcg_minv.release(); ctx->d_pcg_part.release(); ctx->d_pcg_scal.release();
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    cudaStreamSynchronize(ctx->stream_chol);
    cudaStreamDestroy(ctx->stream_chol);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int bagpu_comm_unique_id(uint8_t id_out[128]) {
    if (!g_nccl.load()) return BAGPU_ERR_NCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return BAGPU_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(id_out, &id, 128);
    return BAGPU_OK;
}

int bagpu_comm_init(bagpu_ctx *ctx, int world_size, int rank, const uint8_t id[128]) {
    if (!ctx || world_size < 1 || rank < 0 || rank >= world_size) return fail(ctx, BAGPU_ERR_ARG, "bad communicator arguments");
    if (world_size == 1) { ctx->world = 1; ctx->rank = 0; return BAGPU_OK; }
    if (!g_nccl.load()) return fail(ctx, BAGPU_ERR_NCCL, "cannot dlopen libnccl.so.2");
    CK(cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    if (ctx->opt.debug) fprintf(stderr, "[bagpu r%d] ncclCommInitRank world %d dev %d ...\n", rank, world_size, ctx->device);
    CKN(g_nccl.CommInitRank(&ctx->comm, world_size, uid, rank));
    if (ctx->opt.debug) fprintf(stderr, "[bagpu r%d] ncclCommInitRank done\n", rank);
    ctx->world = world_size; ctx->rank = rank;
    return BAGPU_OK;
}

int bagpu_pin_host(void *p, size_t bytes) { return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? BAGPU_OK : BAGPU_ERR_CUDA; }
int bagpu_unpin_host(void *p) { return cudaHostUnregister(p) == cudaSuccess ? BAGPU_OK : BAGPU_ERR_CUDA; }

// ------------------------------------------------------------------------------- upload
namespace {
int chol_plan_grid(bagpu_ctx *ctx, int n, int max_below, int *grid_out, int *maxr_out);
int parts_plan(bagpu_ctx *ctx, PartPlan &pp, int n, int ld, const std::vector<int> &col_end, int want_parts, bool overlap_possible, cudaStream_t st);
int parts_bind(bagpu_ctx *ctx, PartPlan &pp, double *S, double *bp, double *bs, double *y, double *x, double *dinv, int *failp,
               const unsigned *row_done, const unsigned *item_off, int bw1, const int *row_pos, cudaStream_t st);
}

int bagpu_upload(bagpu_ctx *ctx, const bagpu_problem *p) {
    if (!ctx) return BAGPU_ERR_ARG;
    int rc = validate_problem(ctx, p);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->ev_phase[0], st));
    ctx->have_problem = false;
    struct DrainOnError { cudaStream_t s; bool armed = true; ~DrainOnError() { if (armed) cudaStreamSynchronize(s); } } drain{st};   // an early return must not leave copies of the caller's arrays in flight
    const int Nt = p->n_poses, Np = p->n_points;
    const int64_t Ne = p->n_obs;
    ctx->n_poses = Nt; ctx->n_points = Np; ctx->n_obs = Ne; ctx->n_cams = p->n_cameras; ctx->n_rigs = p->n_rigs;
    int64_t h2d = 0;

    const bool dbg_t = ctx->opt.debug;
    auto wall = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tw0 = wall();
    double lap_t = tw0; std::string laps;
    auto lap = [&](const char *name) { if (dbg_t) { const double t = wall(); char buf[64]; snprintf(buf, sizeof(buf), " %s %.2f", name, t - lap_t); laps += buf; lap_t = t; } };
    // --- the observation arrays go up FIRST, straight into their final places, as if they were already in the order the kernels want
    // (Optimizer.cc's per-MapPoint loops produce it): PCIe is the long pole of an upload, and the host's own passes (order check, CSR
    // pointers, envelope) run while the bytes cross. If the check below finds another order, the arrays are uploaded again through the
    // staging buffer and gathered (the speculative copies are overwritten: same stream).
    const size_t ne = (size_t)Ne;
    CK(ctx->d_o_pose.ensure(4 * ne)); CK(ctx->d_o_point.ensure(4 * ne)); CK(ctx->d_o_meta.ensure(4 * ne));
    CK(ctx->d_o_u.ensure(8 * ne)); CK(ctx->d_o_v.ensure(8 * ne)); CK(ctx->d_o_w.ensure(8 * ne)); CK(ctx->d_o_ur.ensure(8 * ne));
    CK(ctx->d_raw8a.ensure(ne)); CK(ctx->d_raw8b.ensure(ne)); CK(ctx->d_raw16a.ensure(2 * ne)); CK(ctx->d_raw16b.ensure(2 * ne));
    CK(ctx->d_chi2.ensure(8 * ne)); CK(ctx->d_depth.ensure(ne));
    cudaStream_t sp = ctx->stream_chol;
    ctx->h_hidx.assign(Nt, -1);
    int nf = 0;
    for (int i = 0; i < Nt; i++) if (!p->pose_fixed[i]) ctx->h_hidx[i] = nf++;
    ctx->n_free = nf;
    CK(ctx->d_hidx.ensure(sizeof(int) * (size_t)Nt));
    CK(cudaMemcpyAsync(ctx->d_hidx.p, ctx->h_hidx.data(), sizeof(int) * (size_t)Nt, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_o_pose.p, p->obs_pose, 4 * ne, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_o_point.p, p->obs_point, 4 * ne, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev_fork, st));                 // (sorted case) the plan stream needs only the index arrays
    {   // ... and so does the envelope: computed on the plan stream now, read by the host after its own passes
        CK(ctx->d_lastrow.ensure(sizeof(int) * (size_t)std::max(1, nf))); CK(ctx->h_lastrow.ensure(sizeof(int) * (size_t)std::max(1, nf)));
        CK(cudaStreamWaitEvent(sp, ctx->ev_fork, 0));
        CK(cudaMemsetAsync(ctx->d_lastrow.p, 0xff, sizeof(int) * (size_t)std::max(1, nf), sp));
        lastrow_kernel<<<grid_for(Ne, 256), 256, 0, sp>>>(Ne, Nt, ctx->d_o_pose.as<int>(), ctx->d_o_point.as<int>(), ctx->d_hidx.as<int>(), ctx->d_lastrow.as<int>());
        // every rank must lay the reduced camera system out identically: the envelope is the union over the shards
        if (ctx->world > 1 && nf > 0) CKN(g_nccl.AllReduce(ctx->d_lastrow.p, ctx->d_lastrow.p, (size_t)nf, ncclInt32, ncclMax, ctx->comm, sp));
        CK(cudaMemcpyAsync(ctx->h_lastrow.p, ctx->d_lastrow.p, sizeof(int) * (size_t)std::max(1, nf), cudaMemcpyDeviceToHost, sp));
        CK(cudaEventRecord(ctx->ev_lastrow, sp));
    }
    CK(cudaMemcpyAsync(ctx->d_raw8a.p, p->obs_kind, ne, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_raw8b.p, p->obs_flags, ne, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_raw16a.p, p->obs_cam, 2 * ne, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_raw16b.p, p->obs_rig, 2 * ne, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_o_u.p, p->obs_u, 8 * ne, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_o_v.p, p->obs_v, 8 * ne, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_o_w.p, p->obs_inv_sigma2, 8 * ne, cudaMemcpyHostToDevice, st));
    if (p->obs_ur) CK(cudaMemcpyAsync(ctx->d_o_ur.p, p->obs_ur, 8 * ne, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->d_chi2.p, 0, 8 * ne, st));
    lap("obs-enqueue");
    // --- validation + order: landmark-major, pose-ascending inside a landmark, insertion order preserved among equals. One pass checks
    // every observation AND the order AND (for the sorted case) notes the run boundaries lm_ptr[j] = first observation of landmark j
    std::vector<int> lm_ptr((size_t)Np + 1, -1);
    bool sorted = true;
    rc = validate_observations(ctx, p, lm_ptr, sorted);
    if (rc) return rc;
    lap("validate+check");
    if (sorted) {
        // landmarks without observations inherit the next one
        lm_ptr[(size_t)Np] = (int)Ne;
        for (int j = Np - 1; j >= 0; j--) if (lm_ptr[j] < 0) lm_ptr[j] = lm_ptr[j + 1];
    } else {
        std::fill(lm_ptr.begin(), lm_ptr.end(), 0);
        for (int64_t e = 0; e < Ne; e++) lm_ptr[(size_t)p->obs_point[e] + 1]++;
        for (int j = 0; j < Np; j++) lm_ptr[j + 1] += lm_ptr[j];
    }
    std::vector<int> perm;
    if (!sorted) {
        perm.resize(Ne);
        std::vector<int> cursor(lm_ptr.begin(), lm_ptr.end() - 1);
        for (int64_t e = 0; e < Ne; e++) perm[cursor[p->obs_point[e]]++] = (int)e;       // stable counting sort by point
        for (int j = 0; j < Np; j++)                                                        // stable insertion sort by pose
            for (int a = lm_ptr[j] + 1; a < lm_ptr[j + 1]; a++) {
                const int v = perm[a]; int b = a - 1;
                while (b >= lm_ptr[j] && p->obs_pose[perm[b]] > p->obs_pose[v]) { perm[b + 1] = perm[b]; b--; }
                perm[b + 1] = v;
            }
    }
    ctx->identity_perm = sorted;
    lap("lm_ptr");
    // packed tasks / wide list (a serial greedy pass over the landmarks) on a helper thread, beside the envelope computation:
    // tasks = runs of whole landmarks with <= 32 observations in total (stage_kernel, update kernels); a landmark with more
    // than 32 observations is "wide" (stage_wide_kernel, update_kernel: warp = landmark)
    std::vector<int2> tasks;
    std::vector<int> wide_list;
    struct Joiner { std::thread t; ~Joiner() { if (t.joinable()) t.join(); } } tasks_thread;
    // (a fixed number of segments, each packed greedily on its own thread: the partition -- and with it every summation order -- depends
    // on the map only, not on the host's core count)
    constexpr int kTaskSegs = 16;
    tasks_thread.t = std::thread([&tasks, &wide_list, &lm_ptr, Np] {
        std::vector<int2> seg_tasks[kTaskSegs];
        std::vector<int> seg_wide[kTaskSegs];
        auto pack = [&](int sgm) {
            const int j0 = (int)((long long)Np * sgm / kTaskSegs), j1 = (int)((long long)Np * (sgm + 1) / kTaskSegs);
            std::vector<int2> &T = seg_tasks[sgm];
            T.reserve((size_t)(j1 - j0) / 2 + 16);
            int tb = -1, tobs = 0;
            for (int j = j0; j <= j1; j++) {
                const int k = (j < j1) ? lm_ptr[j + 1] - lm_ptr[j] : 0;
                const bool brk = j == j1 || k > 32;
                if (tb >= 0 && (brk || tobs + k > 32)) { T.push_back(make_int2(tb, j)); tb = -1; tobs = 0; }
                if (!brk) { if (tb < 0) tb = j; tobs += k; }
                else if (j < j1) seg_wide[sgm].push_back(j);
            }
        };
        if (Np >= (1 << 16)) {
            std::vector<std::thread> th;
            for (int sgm = 1; sgm < kTaskSegs; sgm++) th.emplace_back(pack, sgm);
            pack(0);
            for (auto &x : th) x.join();
        } else for (int sgm = 0; sgm < kTaskSegs; sgm++) pack(sgm);
        size_t nt = 0, nw = 0;
        for (int sgm = 0; sgm < kTaskSegs; sgm++) { nt += seg_tasks[sgm].size(); nw += seg_wide[sgm].size(); }
        tasks.reserve(nt); wide_list.reserve(nw);
        for (int sgm = 0; sgm < kTaskSegs; sgm++) {
            tasks.insert(tasks.end(), seg_tasks[sgm].begin(), seg_tasks[sgm].end());
            wide_list.insert(wide_list.end(), seg_wide[sgm].begin(), seg_wide[sgm].end());
        }
    });

    // --- poses / free index
    CK(ctx->h_stage.ensure(sizeof(double) * 7 * (size_t)Nt + sizeof(double) * 7 * (size_t)std::max(1, p->n_rigs)));
    double *hp = ctx->h_stage.as<double>();
    for (int i = 0; i < Nt; i++) normalize_qt(p->pose_qt + 7 * (size_t)i, hp + 7 * (size_t)i);
    double *hr = hp + 7 * (size_t)Nt;
    for (int i = 0; i < p->n_rigs; i++) normalize_qt(p->rigs[i].qt, hr + 7 * (size_t)i);

    CK(ctx->d_pose_a.ensure(sizeof(double) * 7 * (size_t)Nt)); CK(ctx->d_pose_b.ensure(sizeof(double) * 7 * (size_t)Nt));
    CK(ctx->d_pose_init.ensure(sizeof(double) * 7 * (size_t)Nt));
    CK(ctx->d_pt_a.ensure(sizeof(double) * 3 * (size_t)Np)); CK(ctx->d_pt_b.ensure(sizeof(double) * 3 * (size_t)Np));
    CK(ctx->d_cams.ensure(sizeof(bagpu_camera) * (size_t)p->n_cameras));
    CK(ctx->d_rigs.ensure(sizeof(double) * 7 * (size_t)std::max(1, p->n_rigs)));
    CK(ctx->d_lm_ptr.ensure(sizeof(int) * ((size_t)Np + 1)));
    CK(cudaMemcpyAsync(ctx->d_pose_a.p, hp, sizeof(double) * 7 * (size_t)Nt, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_pose_init.p, ctx->d_pose_a.p, sizeof(double) * 7 * (size_t)Nt, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_pose_b.p, ctx->d_pose_a.p, sizeof(double) * 7 * (size_t)Nt, cudaMemcpyDeviceToDevice, st));
    if (p->n_rigs) CK(cudaMemcpyAsync(ctx->d_rigs.p, hr, sizeof(double) * 7 * (size_t)p->n_rigs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_pt_a.p, p->points, sizeof(double) * 3 * (size_t)Np, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_pt_b.p, ctx->d_pt_a.p, sizeof(double) * 3 * (size_t)Np, cudaMemcpyDeviceToDevice, st));   // landmarks without observations are never rewritten
    CK(cudaMemcpyAsync(ctx->d_cams.p, p->cameras, sizeof(bagpu_camera) * (size_t)p->n_cameras, cudaMemcpyHostToDevice, st));
    CK(ctx->h_lmptr.ensure(sizeof(int) * ((size_t)Np + 1)));                     // page-locked copy: a pageable source would make this call wait for the stream
    memcpy(ctx->h_lmptr.p, lm_ptr.data(), sizeof(int) * ((size_t)Np + 1));
    // on the plan stream: the tile plan needs it, and on the main stream it would queue behind the bulk of the observation data
    CK(cudaMemcpyAsync(ctx->d_lm_ptr.p, ctx->h_lmptr.p, sizeof(int) * ((size_t)Np + 1), cudaMemcpyHostToDevice, sp));
    h2d += sizeof(double) * (7 * (int64_t)Nt + 3 * (int64_t)Np) + sizeof(int) * ((int64_t)Nt + Np + 1);

    lap("poses");
    // --- observations (uploaded above): meta packing on the device; another order than landmark-major: permutation + gathers
    const int g = grid_for(Ne, 256);
    h2d += 6 * Ne;
    if (sorted) {
        CK(cudaStreamWaitEvent(sp, ctx->ev_fork, 0));
        compose_meta_kernel<<<g, 256, 0, st>>>(Ne, ctx->d_raw8a.as<uint8_t>(), ctx->d_raw16a.as<int16_t>(), ctx->d_raw16b.as<int16_t>(),
                                               ctx->d_raw8b.as<uint8_t>(), ctx->d_o_meta.as<uint32_t>());
    } else {
        CK(ctx->d_perm.ensure(4 * ne)); CK(ctx->d_rawd.ensure(8 * ne));
        CK(cudaMemcpyAsync(ctx->d_perm.p, perm.data(), 4 * ne, cudaMemcpyHostToDevice, st));
        h2d += 4 * Ne;
        // meta in caller order first (into d_rawd as scratch), then gather everything through perm
        uint32_t *meta_raw = ctx->d_rawd.as<uint32_t>();
        compose_meta_kernel<<<g, 256, 0, st>>>(Ne, ctx->d_raw8a.as<uint8_t>(), ctx->d_raw16a.as<int16_t>(), ctx->d_raw16b.as<int16_t>(),
                                               ctx->d_raw8b.as<uint8_t>(), meta_raw);
        gather_perm_kernel<uint32_t><<<g, 256, 0, st>>>(Ne, ctx->d_perm.as<int>(), meta_raw, ctx->d_o_meta.as<uint32_t>());
        struct { const void *src; void *dst; int w; } cols[] = {
            {p->obs_pose, ctx->d_o_pose.p, 4}, {p->obs_point, ctx->d_o_point.p, 4}, {p->obs_u, ctx->d_o_u.p, 8},
            {p->obs_v, ctx->d_o_v.p, 8}, {p->obs_inv_sigma2, ctx->d_o_w.p, 8}, {p->obs_ur, ctx->d_o_ur.p, 8}};
        for (auto &c : cols) {
            if (!c.src) continue;
            CK(cudaMemcpyAsync(ctx->d_rawd.p, c.src, (size_t)c.w * ne, cudaMemcpyHostToDevice, st));
            if (c.w == 4) gather_perm_kernel<int><<<g, 256, 0, st>>>(Ne, ctx->d_perm.as<int>(), ctx->d_rawd.as<int>(), (int *)c.dst);
            else gather_perm_kernel<double><<<g, 256, 0, st>>>(Ne, ctx->d_perm.as<int>(), ctx->d_rawd.as<double>(), (double *)c.dst);
        }
        h2d += (4 + 4 + 8 + 8 + 8 + (p->obs_ur ? 8 : 0)) * Ne;      // uploaded twice
        CK(cudaEventRecord(ctx->ev_fork, st));
        CK(cudaStreamWaitEvent(sp, ctx->ev_fork, 0));
    }
    h2d += (4 + 4 + 8 + 8 + 8 + (p->obs_ur ? 8 : 0)) * Ne;
    CK(cudaGetLastError());
    // snapshot of the initial estimates / edge flags for bagpu_reset_resident
    CK(ctx->d_pt_init.ensure(sizeof(double) * 3 * (size_t)Np)); CK(ctx->d_meta_init.ensure(4 * ne));
    CK(cudaMemcpyAsync(ctx->d_pt_init.p, ctx->d_pt_a.p, sizeof(double) * 3 * (size_t)Np, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_meta_init.p, ctx->d_o_meta.p, 4 * ne, cudaMemcpyDeviceToDevice, st));

    lap("meta");
    const double tw1 = wall();
    // --- reduced camera system: envelope of Hschur from the landmark structure, band or dense storage
    const int n = 6 * nf;
    ctx->n_sys = n;
    {
        std::vector<int> lastrow(std::max(1, nf));
        for (int h = 0; h < nf; h++) lastrow[h] = h;
        CK(cudaEventSynchronize(ctx->ev_lastrow));            // lastrow_kernel on the plan stream (all-reduced over the ranks there)
        bool any_unsorted = !sorted;
        if (ctx->world > 1) {
            // the ranks must take the same path below (it holds a collective): agree on whether any shard came in another order
            CK(ctx->d_count.ensure(sizeof(unsigned long long) * 2));
            int *flag = reinterpret_cast<int *>(ctx->d_count.as<unsigned long long>() + 1);
            int hf = sorted ? 0 : 1;
            CK(cudaMemcpyAsync(flag, &hf, sizeof(int), cudaMemcpyHostToDevice, sp));
            CKN(g_nccl.AllReduce(flag, flag, 1, ncclInt32, ncclMax, ctx->comm, sp));
            CK(cudaMemcpyAsync(&hf, flag, sizeof(int), cudaMemcpyDeviceToHost, sp));
            CK(cudaStreamSynchronize(sp));
            any_unsorted = hf != 0;
        }
        if (sorted) {                                          // computed on the device while the host ran its passes
            const int *dl = ctx->h_lastrow.as<int>();
            for (int h = 0; h < nf; h++) lastrow[h] = std::max(lastrow[h], dl[h]);
        }
        if (any_unsorted) {
            // (a shard in another order: its speculative device result is meaningless -- at worst it widens the envelope, which stays valid;
            // its own rows come from the host pass below, and one more collective gives every rank the union)
            if (!sorted) {
            std::vector<std::vector<int>> part(kPlanThreads);
            parallel_ranges(Np, 1 << 12, [&](int t, int64_t j0, int64_t j1) {
                std::vector<int> &lr = part[t];
                lr.assign(std::max(1, nf), -1);
                for (int64_t j = j0; j < j1; j++) {
                    int mx = -1;
                    for (int a = lm_ptr[j]; a < lm_ptr[j + 1]; a++) mx = std::max(mx, ctx->h_hidx[p->obs_pose[perm[a]]]);
                    if (mx < 0) continue;
                    for (int a = lm_ptr[j]; a < lm_ptr[j + 1]; a++) {
                        const int h = ctx->h_hidx[p->obs_pose[perm[a]]];
                        if (h >= 0) lr[h] = std::max(lr[h], mx);
                    }
                }
            });
            for (auto &lr : part) for (size_t h = 0; h < lr.size() && (int)h < nf; h++) lastrow[h] = std::max(lastrow[h], lr[h]);
            }
            if (ctx->world > 1 && nf > 0) {
                CK(ctx->d_colend.ensure(sizeof(int) * (size_t)nf));
                CK(cudaMemcpyAsync(ctx->d_colend.p, lastrow.data(), sizeof(int) * (size_t)nf, cudaMemcpyHostToDevice, st));
                CKN(g_nccl.AllReduce(ctx->d_colend.p, ctx->d_colend.p, (size_t)nf, ncclInt32, ncclMax, ctx->comm, st));
                CK(cudaMemcpyAsync(lastrow.data(), ctx->d_colend.p, sizeof(int) * (size_t)nf, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
            }
        }
        lap("lastrow");
        int bwb = 0;
        for (int h = 1; h < nf; h++) lastrow[h] = std::max(lastrow[h], lastrow[h - 1]);     // monotone: bounds the fill too
        for (int h = 0; h < nf; h++) bwb = std::max(bwb, lastrow[h] - h);
        ctx->band_blocks = bwb;
        std::vector<int> &col_end = ctx->h_colend; col_end.assign(std::max(1, n), 0);
        int max_below = 0;
        int band = 6 * (bwb + 1);                            // max (i - j) + 1 inside the envelope ...
        for (int h = 0; h < nf; h++) for (int r = 0; r < 6; r++) col_end[6 * h + r] = 6 * lastrow[h] + 5;
        for (int p0 = 0; p0 < n; p0 += CH_NB) {
            const int nb = std::min(CH_NB, n - p0);
            const int rend = std::min(n - 1, col_end[p0 + nb - 1]);
            max_below = std::max(max_below, rend - (p0 + nb) + 1);
            band = std::max(band, rend - p0 + 1);            // ... and every column of a panel stores the panel's rows
        }
        ctx->ld = std::max(1, std::min(band - 1, n));        // band storage when it is narrower than the matrix
        { int rc2 = chol_plan_grid(ctx, n, max_below, &ctx->chol_grid, &ctx->chol_maxr); if (rc2) return rc2; }
        // --- two-way factorisation: rows [0, rT) from the top, rows [n - rT, n) from the bottom (mirrored), separator M between
        ctx->tw = bagpu_ctx::TwoWay();
        ctx->parts.on = false;
        std::vector<int> &ce1 = ctx->h_ce1, &ce2 = ctx->h_ce2, &ceM = ctx->h_ceM;
        if (ctx->chol_maxr > 0 && !ctx->opt.no_twoway && !ctx->opt.compare && !ctx->opt.no_tiles) {
            // long keyframe chains: more than two factorisation fronts (chol_parts.cuh)
            int rcp = parts_plan(ctx, ctx->parts, n, ctx->ld, col_end, ctx->opt.parts, ctx->world == 1 && !ctx->opt.no_overlap && !ctx->overlap_off, st);
            if (rcp) return rcp;
        }
        if (!ctx->parts.on && ctx->chol_maxr > 0 && !ctx->opt.no_twoway && !ctx->opt.compare && !ctx->opt.no_tiles) {
            int band_rows = 1;
            for (int j = 0; j < n; j++) band_rows = std::max(band_rows, std::min(n - 1, col_end[j]) - j + 1);
            const int k = (n - band_rows - CH_NB) / (2 * CH_NB);
            if (k >= 8) {
                bagpu_ctx::TwoWay &T = ctx->tw;
                T.k = k; T.rT = CH_NB * k; T.n1 = n - CH_NB * k; T.n2 = n - T.rT; T.nM = T.n1 - T.rT;
                ce1.resize(T.n1); ce2.resize(T.n2); ceM.assign(T.nM, T.nM - 1);
                for (int j = 0; j < T.n1; j++) ce1[j] = std::min(col_end[j], T.n1 - 1);
                std::vector<int> first_row(n, 0);           // first row with a possibly-nonzero entry in column C of the upper triangle
                { int c0 = 0; for (int R = 0; R < n; R++) { const int ce = std::min(n - 1, col_end[R]); for (; c0 <= ce; c0++) first_row[c0] = R; } }
                for (int j = 0; j < T.n2; j++) ce2[j] = std::min(T.n2 - 1, n - 1 - first_row[n - 1 - j]);
                auto metrics = [&](const std::vector<int> &ce, int nn, int &bnd, int &mb) {
                    bnd = 1; mb = 0;
                    for (int j = 0; j < nn; j++) bnd = std::max(bnd, ce[j] - j + 1);
                    for (int p0 = 0; p0 < nn; p0 += CH_NB) {
                        const int nb = std::min(CH_NB, nn - p0);
                        const int rend = std::min(nn - 1, ce[p0 + nb - 1]);
                        mb = std::max(mb, rend - (p0 + nb) + 1);
                        bnd = std::max(bnd, rend - p0 + 1);
                    }
                };
                int b1, mb1, b2, mb2, bM, mbM;
                metrics(ce1, T.n1, b1, mb1); metrics(ce2, T.n2, b2, mb2); metrics(ceM, T.nM, bM, mbM);
                ctx->ld = std::max(ctx->ld, std::max(1, std::min(std::max(b1, b2) - 1, n)));   // one row stride for S, S1 (= S) and S2
                T.ldM = std::max(1, std::min(bM - 1, T.nM));
                T.s2_elems = (size_t)T.n2 * (ctx->ld + 1) + 8; T.sM_elems = (size_t)T.nM * (T.ldM + 1) + 8;
                int rc2 = chol_plan_grid(ctx, T.n1, mb1, &T.grid1, &T.maxr1); if (rc2) return rc2;
                rc2 = chol_plan_grid(ctx, T.n2, mb2, &T.grid2, &T.maxr2); if (rc2) return rc2;
                rc2 = chol_plan_grid(ctx, T.nM, mbM, &T.gridM, &T.maxrM); if (rc2) return rc2;
                T.on = T.maxr1 > 0 && T.maxr2 > 0 && T.maxrM > 0 && T.grid1 + T.grid2 + 8 <= ctx->n_sm;
            }
        }
        ctx->s_elems = (size_t)std::max(1, n) * (ctx->ld + 1) + 8;   // Lm(i,j) = S[j*ld + i], i in [j, j+band): last index (n-1)*(ld+1)
        if (ctx->tw.on) {
            const bagpu_ctx::TwoWay &T = ctx->tw;
            CK(ctx->d_colend1.ensure(4 * (size_t)T.n1)); CK(ctx->d_colend2.ensure(4 * (size_t)T.n2)); CK(ctx->d_colendM.ensure(4 * (size_t)T.nM));
            CK(ctx->d_y2.ensure(8 * (size_t)T.n2)); CK(ctx->d_SM.ensure(8 * T.sM_elems)); CK(ctx->d_rhsM.ensure(8 * (size_t)T.nM));
            CK(ctx->d_zeroM.ensure(8 * (size_t)T.nM)); CK(ctx->d_yM.ensure(8 * (size_t)T.nM)); CK(ctx->d_xM.ensure(8 * (size_t)T.nM));
            CK(cudaMemcpyAsync(ctx->d_colend1.p, ce1.data(), 4 * (size_t)T.n1, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(ctx->d_colend2.p, ce2.data(), 4 * (size_t)T.n2, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(ctx->d_colendM.p, ceM.data(), 4 * (size_t)T.nM, cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(ctx->d_zeroM.p, 0, 8 * (size_t)T.nM, st));
            if (ctx->opt.debug) fprintf(stderr, "[bagpu] two-way: k=%d rT=%d n1=%d n2=%d nM=%d ld=%d ldM=%d grids %d/%d/%d maxr %d/%d/%d\n",
                                               T.k, T.rT, T.n1, T.n2, T.nM, ctx->ld, T.ldM, T.grid1, T.grid2, T.gridM, T.maxr1, T.maxr2, T.maxrM);
        }
        // order in which pair_kernel takes the camera rows: from both ends towards the separator when the factorisation is two-way
        // (built on the device: the plan stream must not queue H2D copies behind the bulk of the observation data)
        CK(ctx->d_rowpos.ensure(4 * (size_t)std::max(1, nf))); CK(ctx->d_rowofpos.ensure(4 * (size_t)std::max(1, nf)));
        // (the interleaved orders only serve fronts that run BESIDE the accumulation; otherwise the natural order keeps the records of
        // neighbouring rows together in L2)
        const bool fronts_beside = ctx->world == 1 && !ctx->opt.no_overlap && !ctx->overlap_off;
        if (ctx->parts.on && fronts_beside && ctx->parts.P * ctx->parts.nc <= ctx->n_sm / 4)
            row_order_parts_kernel<<<grid_for(std::max(1, nf), 256), 256, 0, sp>>>(nf, ctx->parts.T, ctx->d_rowpos.as<int>(), ctx->d_rowofpos.as<int>());
        else row_order_kernel<<<grid_for(std::max(1, nf), 256), 256, 0, sp>>>(nf, (ctx->tw.on && fronts_beside) ? 1 : 0, ctx->d_rowpos.as<int>(), ctx->d_rowofpos.as<int>());
        if (ctx->opt.debug) fprintf(stderr, "[bagpu] n=%d band_blocks=%d band=%d ld=%d s_elems=%zu max_below=%d chol_grid=%d\n", n, bwb, band, ctx->ld, ctx->s_elems, max_below, ctx->chol_grid);
        CK(ctx->d_colend.ensure(sizeof(int) * (size_t)std::max(1, n)));
        CK(cudaMemcpyAsync(ctx->d_colend.p, col_end.data(), sizeof(int) * (size_t)std::max(1, n), cudaMemcpyHostToDevice, st));
    }
    lap("envelope-tail");
    const double tw2 = wall();
    // --- plan of the linearise + Schur pass: packed tasks / wide list on the host (O(Np)), block-sorted pair lists on the device
    {
        tasks_thread.t.join();
        lap("tasks");
        const int nw = (int)wide_list.size();
        ctx->n_wide = nw; ctx->n_tasks = (int)tasks.size();
        CK(ctx->d_widelist.ensure(sizeof(int) * std::max<size_t>(1, wide_list.size())));
        if (!wide_list.empty()) CK(cudaMemcpyAsync(ctx->d_widelist.p, wide_list.data(), sizeof(int) * wide_list.size(), cudaMemcpyHostToDevice, st));
        CK(ctx->d_tasks.ensure(sizeof(int2) * std::max<size_t>(1, tasks.size())));
        if (!tasks.empty()) CK(cudaMemcpyAsync(ctx->d_tasks.p, tasks.data(), sizeof(int2) * tasks.size(), cudaMemcpyHostToDevice, st));
        {
            int occ_u = 0, occ_st = 0, occ_sw = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_u, update_packed_kernel, ST_THREADS, 0));
            ctx->upd_grid = std::max(1, std::min(ctx->n_sm * std::max(1, occ_u), (ctx->n_tasks + ST_WARPS - 1) / ST_WARPS));
            int occ_z = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_z, update_z_kernel, ST_THREADS, 0));
            ctx->updz_grid = std::max(1, std::min(ctx->n_sm * std::max(1, occ_z), (ctx->n_tasks + ST_WARPS - 1) / ST_WARPS));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_st, stage_kernel, ST_THREADS, 0));
            ctx->stage_occ = std::max(1, occ_st);
            ctx->stage_grid = std::max(1, std::min(ctx->n_sm * ctx->stage_occ, (ctx->n_tasks + ST_WARPS - 1) / ST_WARPS));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sw, stage_wide_kernel, ST_THREADS, 0));
            ctx->stage_wide_grid = std::max(1, std::min(ctx->n_sm * std::max(1, occ_sw), (nw + ST_WARPS - 1) / ST_WARPS));
        }
        ctx->n_items = 0; ctx->n_entries = 0; ctx->tiles = false;
        if (n > 0) {
            CK(ctx->d_npairs.ensure(4 * (ne + 1))); CK(ctx->d_pairoff.ensure(4 * (ne + 1)));
            unsigned *npairs = ctx->d_npairs.as<unsigned>(), *pairoff = ctx->d_pairoff.as<unsigned>();
            int nblk = 0;
            bool tiles = !ctx->opt.pair_list;
            if (tiles) {
                // (tile, landmark) records for pair_tile_kernel (schur_tiles.cuh), sorted by tile (stable radix sort of the records
                // generated in observation order, so every tile's list is in landmark order).
                const int ntile = (nf + TP_T - 1) / TP_T;
                const int tbw1 = (ctx->band_blocks + TP_T - 1) / TP_T + 1;          // tB - tA <= (band_blocks + 3) / 4
                const long long nblk_ll = (long long)ntile * tbw1;
                if (nblk_ll >= (1ll << 31)) return fail(ctx, BAGPU_ERR_ARG, "reduced system has too many tiles (%lld)", nblk_ll);
                nblk = (int)nblk_ll;
                ctx->ntile = ntile; ctx->tbw1 = tbw1;
                CK(ctx->d_tile_pos.ensure(4 * (size_t)ntile)); CK(ctx->d_tile_of_pos.ensure(4 * (size_t)ntile)); CK(ctx->d_cam_tpos.ensure(4 * (size_t)std::max(1, nf)));
                CK(ctx->d_blkcnt.ensure(4 * ((size_t)nblk + 1))); CK(ctx->d_blkoff.ensure(4 * ((size_t)nblk + 1)));
                CK(ctx->d_itemcnt.ensure(4 * ((size_t)nblk + 1))); CK(ctx->d_itemoff.ensure(4 * ((size_t)nblk + 1)));
                CK(ctx->d_count.ensure(sizeof(unsigned long long) * 2));
                unsigned *blkcnt = ctx->d_blkcnt.as<unsigned>(), *blkoff = ctx->d_blkoff.as<unsigned>();
                unsigned *itemcnt = ctx->d_itemcnt.as<unsigned>(), *itemoff = ctx->d_itemoff.as<unsigned>();
                int *ovf = reinterpret_cast<int *>(ctx->d_count.as<unsigned long long>() + 1);
                tile_order_kernel<<<grid_for(ntile, 128), 128, 0, sp>>>(ntile, nf, ctx->d_rowpos.as<int>(), ctx->d_tile_pos.as<int>(), ctx->d_tile_of_pos.as<int>(), ctx->d_cam_tpos.as<int>());
                CK(cudaMemsetAsync(blkcnt, 0, 4 * ((size_t)nblk + 1), sp));
                CK(cudaMemsetAsync(npairs + ne, 0, 4, sp));
                CK(cudaMemsetAsync(ovf, 0, sizeof(int), sp));
                tile_plan_kernel<false><<<g, 256, 0, sp>>>(Ne, ctx->d_lm_ptr.as<int>(), ctx->d_o_pose.as<int>(), ctx->d_o_point.as<int>(), ctx->d_hidx.as<int>(),
                                                           ctx->d_tile_pos.as<int>(), tbw1, npairs, blkcnt, nullptr, nullptr, nullptr, nullptr, ovf);
                // records per work item: long items on big maps (fewer partial sums through memory), shorter ones where the items would not fill the device
                const unsigned chunk = (Ne >= 1000000) ? TP_CHUNK : TP_CHUNK / 2;
                tile_item_count_kernel<<<grid_for(nblk + 1, 256), 256, 0, sp>>>(nblk + 1, blkcnt, itemcnt, chunk);
                size_t tmp_a = 0, tmp_b = 0, tmp_c = 0;
                CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_a, npairs, pairoff, (int)(ne + 1), sp));
                CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_b, blkcnt, blkoff, nblk + 1, sp));
                CK(ctx->d_cubtmp.ensure(std::max(tmp_a, tmp_b)));
                size_t tmp = ctx->d_cubtmp.cap;
                CK(cub::DeviceScan::ExclusiveSum(ctx->d_cubtmp.p, tmp, npairs, pairoff, (int)(ne + 1), sp));
                tmp = ctx->d_cubtmp.cap;
                CK(cub::DeviceScan::ExclusiveSum(ctx->d_cubtmp.p, tmp, blkcnt, blkoff, nblk + 1, sp));
                tmp = ctx->d_cubtmp.cap;
                CK(cub::DeviceScan::ExclusiveSum(ctx->d_cubtmp.p, tmp, itemcnt, itemoff, nblk + 1, sp));
                unsigned totals[2] = {0, 0}; int h_ovf = 0;
                CK(cudaMemcpyAsync(&totals[0], pairoff + ne, 4, cudaMemcpyDeviceToHost, sp));
                CK(cudaMemcpyAsync(&totals[1], itemoff + nblk, 4, cudaMemcpyDeviceToHost, sp));
                CK(cudaMemcpyAsync(&h_ovf, ovf, 4, cudaMemcpyDeviceToHost, sp));
                CK(cudaStreamSynchronize(sp));
                lap("tile count+scan+sync");
                const size_t nrec = totals[0];
                if (h_ovf || nrec >= (1ull << 31)) tiles = false;      // more than TP_MAX_LAYERS edges on one (pose, point) pair: the pair-list path takes any number
                else {
                    ctx->n_entries = (long long)nrec; ctx->n_items = (int)totals[1];
                    if (nrec > 0) {
                        CK(ctx->d_pk_keys.ensure(4 * nrec)); CK(ctx->d_pk_keys2.ensure(4 * nrec)); CK(ctx->d_tp_idx.ensure(4 * nrec)); CK(ctx->d_tp_idx2.ensure(4 * nrec));
                        CK(ctx->d_tp_raw.ensure(sizeof(TileRec) * nrec)); CK(ctx->d_tp_recs.ensure(sizeof(TileRec) * nrec));
                        CK(ctx->d_items.ensure(sizeof(TileItem) * (size_t)std::max(1, ctx->n_items)));
                        tile_plan_kernel<true><<<g, 256, 0, sp>>>(Ne, ctx->d_lm_ptr.as<int>(), ctx->d_o_pose.as<int>(), ctx->d_o_point.as<int>(), ctx->d_hidx.as<int>(),
                                                                  ctx->d_tile_pos.as<int>(), tbw1, nullptr, nullptr, pairoff, ctx->d_pk_keys.as<unsigned>(), ctx->d_tp_idx.as<unsigned>(),
                                                                  ctx->d_tp_raw.as<TileRec>(), ovf);
                        int end_bit = 1;
                        while (end_bit < 32 && (1ll << end_bit) < nblk_ll) end_bit++;
                        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_c, ctx->d_pk_keys.as<unsigned>(), ctx->d_pk_keys2.as<unsigned>(),
                                                           ctx->d_tp_idx.as<unsigned>(), ctx->d_tp_idx2.as<unsigned>(), (int)nrec, 0, end_bit, sp));
                        CK(ctx->d_cubtmp.ensure(tmp_c));
                        tmp = ctx->d_cubtmp.cap;
                        CK(cub::DeviceRadixSort::SortPairs(ctx->d_cubtmp.p, tmp, ctx->d_pk_keys.as<unsigned>(), ctx->d_pk_keys2.as<unsigned>(),
                                                           ctx->d_tp_idx.as<unsigned>(), ctx->d_tp_idx2.as<unsigned>(), (int)nrec, 0, end_bit, sp));
                        tile_gather_kernel<<<grid_for((int64_t)nrec, 256), 256, 0, sp>>>((unsigned)nrec, ctx->d_tp_idx2.as<unsigned>(), ctx->d_tp_raw.as<TileRec>(), ctx->d_tp_recs.as<TileRec>());
                        tile_item_fill_kernel<<<grid_for(nblk, 256), 256, 0, sp>>>(nblk, tbw1, ctx->d_tile_of_pos.as<int>(), blkoff, blkcnt, itemoff, ctx->d_items.as<TileItem>(), chunk);
                        CK(cudaGetLastError());
                    }
                    CK(ctx->d_part.ensure(sizeof(double) * TP_PART * (size_t)std::max(1, ctx->n_items)));
                    CK(ctx->d_tp_work.ensure(16)); CK(cudaMemsetAsync(ctx->d_tp_work.p, 0, 16, sp));
                    int occ_p = 0;
                    if (ctx->opt.tile_fma) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_p, pair_tile_kernel, TP_THREADS, TP_SMEM_BYTES));
                    else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_p, pair_tile_mma_kernel, TP_THREADS, TM_SMEM_BYTES));
                    ctx->pair_occ = std::max(1, occ_p);
                    ctx->pair_grid = std::max(1, std::min(ctx->n_sm * ctx->pair_occ, (ctx->n_items + TP_WARPS - 1) / TP_WARPS));
                    { int occ_d = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_d, tile_diag_kernel, TP_THREADS, 0)); ctx->diag_grid = std::max(1, ctx->n_sm * std::max(1, occ_d)); }
                    ctx->tiles = true; ctx->wait_bw1 = tbw1; ctx->wait_rowpos = ctx->d_cam_tpos.as<int>();
                }
            }
            if (!tiles) {
            // pair lists: for every upper block (a, b) of the reduced system the observation pairs (e_a, e_b) of the landmarks
            // both cameras see, sorted by block (stable radix sort of the pairs generated in observation order).
            const int bw1 = ctx->band_blocks + 1;
            const long long nblk_ll = (long long)nf * bw1;
            if (nblk_ll >= (1ll << 31)) return fail(ctx, BAGPU_ERR_ARG, "reduced system has too many blocks (%lld)", nblk_ll);
            nblk = (int)nblk_ll;
            ctx->wait_bw1 = bw1; ctx->wait_rowpos = ctx->d_rowpos.as<int>();
            CK(ctx->d_blkcnt.ensure(4 * ((size_t)nblk + 1))); CK(ctx->d_blkoff.ensure(4 * ((size_t)nblk + 1)));
            CK(ctx->d_itemcnt.ensure(4 * ((size_t)nblk + 1))); CK(ctx->d_itemoff.ensure(4 * ((size_t)nblk + 1)));
            unsigned *blkcnt = ctx->d_blkcnt.as<unsigned>(), *blkoff = ctx->d_blkoff.as<unsigned>();
            unsigned *itemcnt = ctx->d_itemcnt.as<unsigned>(), *itemoff = ctx->d_itemoff.as<unsigned>();
            CK(cudaMemsetAsync(blkcnt, 0, 4 * ((size_t)nblk + 1), sp));
            CK(cudaMemsetAsync(npairs + ne, 0, 4, sp));
            pair_count_kernel<<<g, 256, 0, sp>>>(Ne, ctx->d_lm_ptr.as<int>(), ctx->d_o_pose.as<int>(), ctx->d_o_point.as<int>(), ctx->d_hidx.as<int>(),
                                                 ctx->d_rowpos.as<int>(), bw1, npairs, blkcnt);
            pair_item_count_kernel<<<grid_for(nblk + 1, 256), 256, 0, sp>>>(nblk + 1, blkcnt, itemcnt);
            size_t tmp_a = 0, tmp_b = 0, tmp_c = 0;
            CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_a, npairs, pairoff, (int)(ne + 1), sp));
            CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_b, blkcnt, blkoff, nblk + 1, sp));
            CK(ctx->d_cubtmp.ensure(std::max(tmp_a, tmp_b)));
            size_t tmp = ctx->d_cubtmp.cap;
            CK(cub::DeviceScan::ExclusiveSum(ctx->d_cubtmp.p, tmp, npairs, pairoff, (int)(ne + 1), sp));
            tmp = ctx->d_cubtmp.cap;
            CK(cub::DeviceScan::ExclusiveSum(ctx->d_cubtmp.p, tmp, blkcnt, blkoff, nblk + 1, sp));
            tmp = ctx->d_cubtmp.cap;
            CK(cub::DeviceScan::ExclusiveSum(ctx->d_cubtmp.p, tmp, itemcnt, itemoff, nblk + 1, sp));
            unsigned totals[2] = {0, 0};
            CK(cudaMemcpyAsync(&totals[0], pairoff + ne, 4, cudaMemcpyDeviceToHost, sp));
            CK(cudaMemcpyAsync(&totals[1], itemoff + nblk, 4, cudaMemcpyDeviceToHost, sp));
            CK(cudaStreamSynchronize(sp));
            lap("count+scan+sync");
            const size_t npr = totals[0];
            if (npr >= (1ull << 31)) return fail(ctx, BAGPU_ERR_ARG, "too many observation pairs for one device shard (%zu)", npr);
            ctx->n_entries = (long long)npr; ctx->n_items = (int)totals[1];
            if (npr > 0) {
                CK(ctx->d_pk_keys.ensure(4 * npr)); CK(ctx->d_pk_keys2.ensure(4 * npr));
                CK(ctx->d_pk_vals.ensure(8 * npr)); CK(ctx->d_entries.ensure(8 * npr));
                CK(ctx->d_items.ensure(sizeof(PairItem) * (size_t)std::max(1, ctx->n_items)));
                pair_gen_kernel<<<g, 256, 0, sp>>>(Ne, ctx->d_lm_ptr.as<int>(), ctx->d_o_pose.as<int>(), ctx->d_o_point.as<int>(), ctx->d_hidx.as<int>(),
                                                   ctx->d_rowpos.as<int>(), bw1, pairoff, ctx->d_pk_keys.as<unsigned>(), ctx->d_pk_vals.as<int2>());
                int end_bit = 1;
                while (end_bit < 32 && (1ll << end_bit) < nblk_ll) end_bit++;
                static_assert(sizeof(unsigned long long) == sizeof(int2), "pair entry size");
                CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_c, ctx->d_pk_keys.as<unsigned>(), ctx->d_pk_keys2.as<unsigned>(),
                                                   ctx->d_pk_vals.as<unsigned long long>(), ctx->d_entries.as<unsigned long long>(), (int)npr, 0, end_bit, sp));
                CK(ctx->d_cubtmp.ensure(tmp_c));
                tmp = ctx->d_cubtmp.cap;
                CK(cub::DeviceRadixSort::SortPairs(ctx->d_cubtmp.p, tmp, ctx->d_pk_keys.as<unsigned>(), ctx->d_pk_keys2.as<unsigned>(),
                                                   ctx->d_pk_vals.as<unsigned long long>(), ctx->d_entries.as<unsigned long long>(), (int)npr, 0, end_bit, sp));
                pair_item_fill_kernel<<<grid_for(nblk, 256), 256, 0, sp>>>(nblk, bw1, ctx->d_rowofpos.as<int>(), blkoff, blkcnt, itemoff, ctx->d_items.as<PairItem>());
                CK(cudaGetLastError());
            }
            }
            CK(ctx->d_Z.ensure(sizeof(double) * ZR_STRIDE * ne)); CK(ctx->d_Dr.ensure(sizeof(double) * DR_STRIDE * ne));
            CK(ctx->d_Lm.ensure(sizeof(double) * LM_STRIDE * (size_t)Np));
            CK(ctx->d_blkdone.ensure(sizeof(unsigned) * ((size_t)nblk + 1)));
            CK(cudaMemsetAsync(ctx->d_blkdone.p, 0, sizeof(unsigned) * ((size_t)nblk + 1), sp));
            if (!ctx->tiles) {
                CK(ctx->d_part.ensure(sizeof(double) * PK_PART * (size_t)std::max(1, ctx->n_items)));
                int occ_p = 0;
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_p, pair_kernel, PK_THREADS, PK_SMEM_BYTES));
                ctx->pair_occ = std::max(1, occ_p);
                ctx->pair_grid = std::max(1, std::min(ctx->n_sm * ctx->pair_occ, (ctx->n_items + PK_WARPS - 1) / PK_WARPS));
            }
            CK(ctx->d_rowdone.ensure(sizeof(unsigned) * (size_t)std::max(1, nf)));
        }
        if (ctx->opt.debug) fprintf(stderr, "[bagpu] pair plan (%s): tasks=%d wide=%d entries=%lld items=%d stage_grid=%d pair_grid=%d\n",
                                           ctx->tiles ? "tile records" : "pair list", ctx->n_tasks, nw, ctx->n_entries, ctx->n_items, ctx->stage_grid, ctx->pair_grid);
    }
    // [S | bp | bs | S2 | trial scratch: y (n), y2 (n2), yM (nM), row_done (nf x u32), fail | hpp_diag]: one memset per trial covers S .. fail
    ctx->scratch_elems = (size_t)std::max(1, n) + (size_t)ctx->tw.n2 + (size_t)ctx->tw.nM + ((size_t)std::max(1, nf) + 1) / 2 + 2;
    CK(ctx->d_sys.ensure(sizeof(double) * (ctx->s_elems + 3 * (size_t)std::max(1, n) + ctx->tw.s2_elems + ctx->scratch_elems)));
    CK(ctx->d_xp.ensure(sizeof(double) * (size_t)std::max(1, n)));
    CK(ctx->d_y.ensure(sizeof(double) * (size_t)std::max(1, n)));
    CK(ctx->d_dinv.ensure(sizeof(double) * (size_t)std::max(1, n)));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, build_kernel, BUILD_THREADS, 0));
    ctx->build_grid = std::max(1, std::min(ctx->n_sm * std::max(1, occ), (Np + BUILD_WARPS - 1) / BUILD_WARPS));
    ctx->parts_stride = std::max(std::max(ctx->build_grid, ctx->stage_grid + ctx->stage_wide_grid), std::max(ctx->upd_grid, ctx->updz_grid));
    CK(ctx->d_parts.ensure(sizeof(double) * 7 * (size_t)ctx->parts_stride));
    CK(ctx->d_status.ensure(sizeof(double) * 32));
    CK(ctx->d_fail.ensure(sizeof(int) * 4));
    CK(ctx->d_count.ensure(sizeof(unsigned long long) * 2));
    CK(ctx->h_status.ensure(sizeof(double) * 32));
    ctx->pose_cur = ctx->d_pose_a.as<double>(); ctx->pose_trial = ctx->d_pose_b.as<double>();
    ctx->pt_cur = ctx->d_pt_a.as<double>(); ctx->pt_trial = ctx->d_pt_b.as<double>();
    if (ctx->parts.on) {
        const SysLayout L = sys_layout(ctx);
        int rcb = parts_bind(ctx, ctx->parts, L.S, L.bp, L.bs, L.y1, ctx->d_xp.as<double>(), ctx->d_dinv.as<double>(), L.fail,
                             L.row_done, ctx->d_itemoff.as<unsigned>(), ctx->wait_bw1, ctx->wait_rowpos, st);
        if (rcb) return rcb;
    }
    // multi-GPU trial in chunks of fronts: row / item / element boundaries (the items are in natural row order when the fronts do not
    // run beside the accumulation)
    ctx->chunks.on = false;
    if (ctx->parts.on && ctx->world > 1 && ctx->tiles && !ctx->opt.tile_fma && ctx->opt.chunks && ctx->parts.P >= 2 * bagpu_ctx::kChunks && ctx->n_items > 0) {
        bagpu_ctx::ChunkPlan &C = ctx->chunks;
        C.n = bagpu_ctx::kChunks;
        unsigned h_item[bagpu_ctx::kChunks + 1];
        for (int k = 0; k <= C.n; k++) {
            C.part0[k] = (int)((long long)ctx->parts.P * k / C.n);
            const int row = (k == C.n) ? n : ctx->parts.T.d[C.part0[k]].r0;          // sub-system q owns rows [r0_q, r0_{q+1}): interior + bottom separator
            C.elem0[k] = (size_t)row * (ctx->ld + 1);
            const int tile = std::min(ctx->ntile, (row / 6) / TP_T);
            CK(cudaMemcpyAsync(&h_item[k], ctx->d_itemoff.as<unsigned>() + (size_t)tile * ctx->tbw1, 4, cudaMemcpyDeviceToHost, sp));
        }
        CK(cudaStreamSynchronize(sp));
        for (int k = 0; k <= C.n; k++) C.item0[k] = (int)h_item[k];
        C.item0[C.n] = ctx->n_items;
        C.on = true;
        if (ctx->opt.debug) fprintf(stderr, "[bagpu r%d] chunked trial: fronts %d/%d/%d/%d items %d/%d/%d/%d\n", ctx->rank, C.part0[0], C.part0[1], C.part0[2], C.part0[3],
                                    C.item0[0], C.item0[1], C.item0[2], C.item0[3]);
    }
    lap("plan-rest");
    const double tw3 = wall();
    CK(cudaEventRecord(ctx->ev_join, sp));
    CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));          // the plan stream joins the upload
    CK(cudaEventRecord(ctx->ev_phase[1], st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_phase[0], ctx->ev_phase[1]);
    ctx->tm.h2d_ms = ms; ctx->tm.h2d_bytes = h2d;
    if (dbg_t) fprintf(stderr, "[bagpu] upload laps:%s\n", laps.c_str());
    if (dbg_t) fprintf(stderr, "[bagpu] upload host ms: order+copies %.2f envelope %.2f plan %.2f tail+sync %.2f | stream %.2f\n", tw1 - tw0, tw2 - tw1, tw3 - tw2, wall() - tw3, ms);
    ctx->have_problem = true;
    drain.armed = false;                                   // the stream was synchronised above
    return BAGPU_OK;
}

// ------------------------------------------------------------------------------- solve
namespace {

int all_reduce_sum(bagpu_ctx *ctx, double *buf, size_t count, cudaStream_t stream = nullptr) {
    if (ctx->world <= 1) return BAGPU_OK;
    if (ctx->opt.debug) fprintf(stderr, "[bagpu r%d] allreduce sum %zu\n", ctx->rank, count);
    CKN(g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclSum, ctx->comm, stream ? stream : ctx->stream));
    return BAGPU_OK;
}
int all_reduce_max(bagpu_ctx *ctx, double *buf, size_t count) {
    if (ctx->world <= 1) return BAGPU_OK;
    CKN(g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclMax, ctx->comm, ctx->stream));
    return BAGPU_OK;
}

// read `count` doubles of the device status record (one stream sync)
int read_status(bagpu_ctx *ctx, double *out, int count) {
    CK(cudaMemcpyAsync(ctx->h_status.p, ctx->d_status.p, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->h_status.p, sizeof(double) * count);
    resolve_events(ctx);
    return BAGPU_OK;
}

size_t chol_dyn_smem(int n) { return n <= CH_MAX_SMEM_N ? sizeof(double) * (size_t)std::max(1, n) : 0; }

// Launch plan for a system whose widest panel has `max_below` rows under its diagonal block.
//  * max_below + 32 <= CB_MAXR: chol_band_kernel, ONE cluster of NC = 2^k >= ceil(maxr / 32) CTAs (<= 16), the window of
//    the factorisation in shared memory;
//  * otherwise chol_solve_kernel: one CTA per trailing tile (or TRSM row pass) plus the CTA that owns the right-hand side;
//    up to CH_CLUSTER_MAX CTAs as one cluster (hardware barrier), beyond that a cooperative launch.
#define CH_CLUSTER_MAX 16
size_t chol_band_smem_min(int maxr) { return sizeof(double) * 2 * (size_t)(maxr + CB_PAD) * CB_LD; }
// dynamic shared memory of the band kernel: the two block-column buffers of the factorisation, or -- when it still fits --
// the right-hand side plus two prefetched panels for the backward substitution, whichever is larger
size_t chol_band_smem(const bagpu_ctx *ctx, int n, int maxr) {
    const size_t cap = ctx->band_smem_cap;
    const size_t need = chol_band_smem_min(maxr);
    const size_t bs = (size_t)((maxr + 1) | 1);
    const size_t back = sizeof(double) * ((size_t)n + 2 * CH_NB * bs);
    const size_t back3 = sizeof(double) * ((size_t)n + 3 * CH_NB * bs);      // pipelined backward substitution: three panel buffers
    const size_t ysm = sizeof(double) * (size_t)n;
    if (back3 <= cap && !ctx->opt.no_back3) return std::max(need, back3);
    if (back <= cap) return std::max(need, back);
    if (ysm <= cap) return std::max(need, ysm);
    return need;
}

int chol_plan_grid(bagpu_ctx *ctx, int n, int max_below, int *grid_out, int *maxr_out) {
    const bool no_band = ctx->opt.no_band;
    const int maxr = max_below + CH_NB;
    if (maxr <= CB_MAXR && !no_band) {
        int nc = 2;
        while (nc * CH_NB < maxr) nc *= 2;
        *grid_out = nc; *maxr_out = maxr;
        return BAGPU_OK;
    }
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, chol_solve_kernel<false>, CH_THREADS, chol_dyn_smem(n)));
    if (occ < 1) return fail(ctx, BAGPU_ERR_CUDA, "cholesky kernel does not fit");
    const int nt = (max_below + CH_TB - 1) / CH_TB;
    const int want = std::max(nt * (nt + 1) / 2, (max_below + CH_TR - 1) / CH_TR);
    *grid_out = std::max(1, std::min(ctx->n_sm * occ, want + (want > 0 ? 1 : 0)));
    *maxr_out = 0;
    return BAGPU_OK;
}

int launch_chol(bagpu_ctx *ctx, CholArgs &a, int grid, int maxr, cudaStream_t stream = nullptr) {
    if (!stream) stream = ctx->stream;
    const bool no_cluster = ctx->opt.no_cluster;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = grid; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(grid); cfg.stream = stream; cfg.attrs = at; cfg.numAttrs = 1;
    if (maxr > 0) {
        cfg.blockDim = dim3(CB_THREADS);
        cfg.dynamicSmemBytes = chol_band_smem(ctx, a.n, maxr);
        CK(cudaLaunchKernelEx(&cfg, chol_band_kernel, a, maxr, (int)(cfg.dynamicSmemBytes / sizeof(double)), (const CholArgs *)nullptr));
        return BAGPU_OK;
    }
    if (grid <= CH_CLUSTER_MAX && !no_cluster) {
        cfg.blockDim = dim3(CH_THREADS);
        cfg.dynamicSmemBytes = chol_dyn_smem(a.n);
        CK(cudaLaunchKernelEx(&cfg, chol_solve_kernel<true>, a));
        return BAGPU_OK;
    }
    void *args[] = {&a};
    CK(cudaLaunchCooperativeKernel((void *)chol_solve_kernel<false>, dim3(grid), dim3(CH_THREADS), args, chol_dyn_smem(a.n), stream));
    return BAGPU_OK;
}

// ---- partitioned band solver: plan (which fronts, which buffers), bind (argument tables), enqueue
void envelope_metrics(const std::vector<int> &ce, int off, int nn, int &band, int &max_below) {
    band = 1; max_below = 0;
    for (int j = 0; j < nn; j++) band = std::max(band, ce[off + j] - j + 1);
    for (int p0 = 0; p0 < nn; p0 += CH_NB) {
        const int nb = std::min(CH_NB, nn - p0);
        const int rend = std::min(nn - 1, ce[off + p0 + nb - 1]);
        max_below = std::max(max_below, rend - (p0 + nb) + 1);
        band = std::max(band, rend - p0 + 1);
    }
}

// Decide whether (and into how many fronts) to cut the band; on success pp.on, the partition table, the static device arrays and all
// buffers are in place. want_parts: 0 = automatic (cost model below), 1 = never, >= 3 = that many (when the system is long enough).
int parts_plan(bagpu_ctx *ctx, PartPlan &pp, int n, int ld, const std::vector<int> &col_end, int want_parts, bool overlap_possible, cudaStream_t st) {
    pp.on = false; pp.P = 0;
    if (want_parts == 1 || want_parts == 2 || n < 32 * 24) return BAGPU_OK;
    int band_rows = 1;
    for (int j = 0; j < n; j++) band_rows = std::max(band_rows, std::min(n - 1, col_end[j]) - j + 1);
    const int w = ((std::max(band_rows - 1, 32) + 95) / 96) * 96;       // separator rows: >= the reach of a column, boundaries on 16-camera / 3-panel marks
    auto interior_of = [&](int P) { return ((n - (P - 1) * w) / P / 96) * 96; };
    int P = 0;
    if (want_parts >= 3) {
        P = std::min(want_parts, PS_MAX_PARTS);
        while (P >= 3 && interior_of(P) < std::max(w, 192)) P--;
        if (P < 3) return BAGPU_OK;
    } else {
        // Beside pair_kernel (one GPU) the two-front solver wins: its chain nearly keeps pace with the accumulation, while P spinning
        // clusters take SMs away from pair_kernel and leave a longer tail (spikes + separator system) after it -- measured on config 5:
        // 13.5 ms per trial with two fronts, 16.0 / 17.7 ms with 4 / 8. The fronts pay off where the solve is exposed: after the
        // all-reduce of a multi-GPU trial, or when the overlap is off.
        // One GPU: the two-front chain runs beside the accumulation, so P fronts only win when that chain is much longer than the pass it
        // hides behind (config 5: two-front chain 7.7 ms -- 11 ms beside pair_tile_mma_kernel -- against a 5.2 ms pass + 2.7 ms of
        // twelve fronts after it; config 4: 1.0 ms chain beside a 0.6 ms pass, two fronts stay).
        const double pass_us = 0.28e-3 * (double)ctx->n_obs;
        // cost model (us; measured on B200, config 5: ~19 us per 32-column panel of a front, ~160 us per level of the separator
        // system's cyclic reduction, ~0.5 ms for the spikes, their products and the extra launches; the two-front solver runs at
        // ~16 us per panel). More than ~100 SMs of spinning clusters slow each other down through L2 (16 fronts: 3.7 ms, 12: 2.8 ms).
        const double panels = n / 32.0, two_way = panels / 2 * 16.0 + 250.0;
        double best = 0.8 * two_way;
        // Fronts: measured on config 5 (B200, clusters of 8 CTAs), 10 / 12 / 13 / 14 / 15 / 16 fronts: 2.98 / 2.54 / 2.42 / 2.78 / 3.14 / 3.48 ms
        // per solve. Beyond 13 clusters (104 of 148 SMs) the fronts slow each other down although cudaOccupancyMaxActiveClusters reports
        // 15 or more co-resident (an 8-GPU run with 15 fronts: 3.12 ms): the bound is the measured 70 % of the SMs, not the calculator's.
        int nc_guess = 2;
        while (nc_guess * CH_NB < std::min(CB_MAXR, band_rows + CH_NB)) nc_guess *= 2;
        const int max_fronts = std::max(3, (45 * ctx->n_sm) / (64 * nc_guess));
        for (int q = 3; q <= PS_MAX_PARTS && q <= max_fronts; q++) {
            const int per = interior_of(q);
            if (per < std::max(2 * w, 576)) break;
            int levels = 0;
            while ((1 << levels) < q) levels++;
            const double c = (per / 32.0 + w / 32.0) * 19.0 + levels * 160.0 + 500.0;
            if (c < best) { best = c; P = q; }
        }
        if (P < 3) return BAGPU_OK;
        if (overlap_possible && pass_us + best > 0.9 * std::max(pass_us, 1.4 * two_way)) return BAGPU_OK;
    }
    const int per = interior_of(P);
    PartTable &T = pp.T;
    memset(&T, 0, sizeof(T));
    T.P = P; T.w = w;
    pp.h_subce.clear(); pp.h_seprow.clear();
    long long v_off = 0, linv_off = 0;
    int cta0 = 0, pan0 = 0, tile0 = 0, r = 0, maxr = 0, n_max = 0, apply_rows = 0;
    const int nt = w / 32;
    for (int i = 0; i < P; i++) {
        PartDesc &D = T.d[i];
        D.r0 = r;
        D.m = (i + 1 < P) ? per : n - r;
        D.k = (D.m + 31) / 32;
        D.wT = (i > 0) ? w : 0; D.wB = (i + 1 < P) ? w : 0;
        D.sep = i - 1;
        D.v_off = v_off; D.linv_off = linv_off; D.ce_off = (int)pp.h_subce.size();
        D.cta0 = cta0; D.pan0 = pan0; D.tile0 = tile0;
        const int ni = D.m + D.wB;
        for (int j = 0; j < ni; j++) pp.h_subce.push_back(std::min(std::min(n - 1, col_end[r + j]) - r, ni - 1));
        int bnd, mb;
        envelope_metrics(pp.h_subce, D.ce_off, ni, bnd, mb);
        if (bnd - 1 > ld) return fail(ctx, BAGPU_ERR_ARG, "partition band %d exceeds the row stride %d", bnd, ld);
        maxr = std::max(maxr, mb + CH_NB); n_max = std::max(n_max, ni);
        if (i > 0) {
            v_off += (long long)D.m * w; linv_off += (long long)D.k * 1024;
            cta0 += (w + PS_NCOL - 1) / PS_NCOL; pan0 += D.k; tile0 += nt * (nt + 1) / 2 + nt * nt + 1; apply_rows += D.m;
            pp.h_seprow.push_back(r - w);
        }
        r += D.m + D.wB;
    }
    if (maxr > CB_MAXR || ctx->opt.no_band) return BAGPU_OK;             // envelope too wide for the shared-memory window: keep the one-front solvers
    int nc = 2;
    while (nc * CH_NB < maxr) nc *= 2;
    if (P * nc > ctx->n_sm) return BAGPU_OK;
    pp.P = P; pp.w = w; pp.n = n; pp.ld = ld; pp.nc = nc; pp.maxr = maxr; pp.n_max = n_max;
    pp.spike_ctas = cta0; pp.inv_panels = pan0; pp.gram_ctas = tile0; pp.apply_rows = apply_rows;
    // separator system: block tridiagonal, (P - 1) blocks of w rows
    pp.nS = (P - 1) * w;
    pp.h_ceM.resize(pp.nS);
    for (int j = 0; j < pp.nS; j++) pp.h_ceM[j] = std::min(pp.nS - 1, (j / w + 2) * w - 1);
    int bM, mbM;
    envelope_metrics(pp.h_ceM, 0, pp.nS, bM, mbM);
    pp.ldM = std::max(1, std::min(bM - 1, pp.nS));
    { int rc = chol_plan_grid(ctx, pp.nS, mbM, &pp.gridM, &pp.maxrM); if (rc) return rc; }
    const size_t nS = (size_t)pp.nS, ww = (size_t)(P - 1) * w * w;
    CK(pp.d_subce.ensure(4 * pp.h_subce.size())); CK(pp.d_seprow.ensure(4 * pp.h_seprow.size())); CK(pp.d_ceM.ensure(4 * nS));
    CK(pp.d_V.ensure(8 * (size_t)std::max<long long>(1, v_off))); CK(pp.d_linv.ensure(8 * (size_t)std::max<long long>(1, linv_off)));
    CK(pp.d_Dp.ensure(8 * ww)); CK(pp.d_Ep.ensure(8 * ww)); CK(pp.d_gp.ensure(8 * nS));
    CK(pp.d_SM.ensure(8 * (nS * (pp.ldM + 1) + 8))); CK(pp.d_rhsM.ensure(8 * nS)); CK(pp.d_zeroM.ensure(8 * nS)); CK(pp.d_yM.ensure(8 * nS));
    CK(pp.d_xM.ensure(8 * nS)); CK(pp.d_dinvM.ensure(8 * nS)); CK(pp.d_tab.ensure(sizeof(CholArgs) * 2 * (size_t)P));
    CK(cudaMemcpyAsync(pp.d_subce.p, pp.h_subce.data(), 4 * pp.h_subce.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(pp.d_seprow.p, pp.h_seprow.data(), 4 * pp.h_seprow.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(pp.d_ceM.p, pp.h_ceM.data(), 4 * nS, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(pp.d_zeroM.p, 0, 8 * nS, st));
    CK(cudaMemsetAsync(pp.d_SM.p, 0, 8 * (nS * (pp.ldM + 1) + 8), st));
    CK(cudaMemsetAsync(pp.d_Ep.p, 0, 8 * ww, st));
    // separator system: block cyclic reduction over the K = P - 1 dense blocks (levels of independent eliminations)
    pp.cr = !ctx->opt.sep_tiled && w <= CR_MAXW && w <= CB_MAXR;
    if (pp.cr) {
        const int K = P - 1;
        pp.cr_stride.clear(); pp.cr_cnt.clear(); pp.cr_off.clear();
        int off = 0;
        for (int sft = 1; sft - 1 < K; sft *= 2) {
            const int cnt = (K - (sft - 1) + 2 * sft - 1) / (2 * sft);
            if (cnt <= 0) break;
            pp.cr_stride.push_back(sft); pp.cr_cnt.push_back(cnt); pp.cr_off.push_back(off);
            off += cnt;
        }
        pp.cr_nc = 2;
        while (pp.cr_nc * CH_NB < w) pp.cr_nc *= 2;
        pp.h_cedense.assign(w, w - 1);
        const size_t kw = (size_t)K * w;
        CK(pp.d_Dd.ensure(8 * ww)); CK(pp.d_Cc.ensure(8 * ww)); CK(pp.d_Fa.ensure(8 * ww)); CK(pp.d_Fb.ensure(8 * ww));
        CK(pp.d_gg.ensure(8 * kw)); CK(pp.d_yy.ensure(8 * kw)); CK(pp.d_linvB.ensure(8 * (size_t)K * (w / 32) * 1024));
        CK(pp.d_cedense.ensure(4 * (size_t)w)); CK(pp.d_crtab.ensure(sizeof(CholArgs) * 2 * (size_t)K));
        CK(cudaMemcpyAsync(pp.d_cedense.p, pp.h_cedense.data(), 4 * (size_t)w, cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(pp.d_Dd.p, 0, 8 * ww, st)); CK(cudaMemsetAsync(pp.d_Cc.p, 0, 8 * ww, st));
        pp.C.K = K; pp.C.w = w; pp.C.Dd = pp.d_Dd.as<double>(); pp.C.Cc = pp.d_Cc.as<double>(); pp.C.gg = pp.d_gg.as<double>(); pp.C.yy = pp.d_yy.as<double>();
        pp.C.xx = pp.d_xM.as<double>(); pp.C.Fa = pp.d_Fa.as<double>(); pp.C.Fb = pp.d_Fb.as<double>(); pp.C.linv = pp.d_linvB.as<double>();
    }
    pp.on = true;
    if (ctx->opt.debug) fprintf(stderr, "[bagpu] partitioned solver: P=%d w=%d interior=%d (last %d) nc=%d maxr=%d | separator system n=%d ld=%d grid=%d maxr=%d | V %.1f MB\n",
                                P, w, per, T.d[P - 1].m, nc, maxr, pp.nS, pp.ldM, pp.gridM, pp.maxrM, 8.0 * v_off / 1e6);
    if (ctx->opt.debug && pp.cr) fprintf(stderr, "[bagpu] separator system by block cyclic reduction: %d blocks of %d rows, %zu levels\n", P - 1, w, pp.cr_stride.size());
    return BAGPU_OK;
}

// Argument tables of the two table launches (factorisation: one cluster per sub-system; backward substitution: one CTA each).
int parts_bind(bagpu_ctx *ctx, PartPlan &pp, double *S, double *bp, double *bs, double *y, double *x, double *dinv, int *failp,
               const unsigned *row_done, const unsigned *item_off, int bw1, const int *row_pos, cudaStream_t st) {
    const int P = pp.P;
    pp.h_tab.assign(2 * (size_t)P, CholArgs());
    for (int i = 0; i < P; i++) {
        const PartDesc &D = pp.T.d[i];
        CholArgs a;
        a.S = S + (size_t)D.r0 * pp.ld + D.r0; a.n = D.m + D.wB; a.ld = pp.ld; a.lambda = 0.0;
        a.bp = bp + D.r0; a.bs = bs + D.r0; a.col_end = pp.d_subce.as<int>() + D.ce_off; a.y = y + D.r0; a.dinv = dinv + D.r0; a.x = x + D.r0;
        a.fail = failp; a.prof = nullptr;
        a.row_done = row_done; a.item_off = item_off; a.bw1 = bw1; a.row_pos = row_pos; a.row_base = D.r0;
        a.p_stop = (i + 1 < P) ? D.k : 0; a.no_back = (i + 1 < P) ? 0 : 1;
        pp.h_tab[i] = a;
        CholArgs b = a;
        b.row_done = nullptr; b.p_stop = 0; b.no_back = 0; b.back_from = D.k;
        pp.h_tab[P + i] = b;
    }
    CK(cudaMemcpyAsync(pp.d_tab.p, pp.h_tab.data(), sizeof(CholArgs) * 2 * (size_t)P, cudaMemcpyHostToDevice, st));
    if (pp.cr) {
        const int K = pp.C.K, w = pp.C.w;
        pp.h_crtab.assign(2 * (size_t)K, CholArgs());
        int pos = 0;
        for (size_t lv = 0; lv < pp.cr_stride.size(); lv++)
            for (int t = 0; t < pp.cr_cnt[lv]; t++, pos++) {
                const int j = pp.cr_stride[lv] - 1 + 2 * pp.cr_stride[lv] * t;
                CholArgs a;
                a.S = pp.C.Dd + (size_t)j * w * w; a.n = w; a.ld = w; a.lambda = 0.0; a.bp = pp.C.gg + (size_t)j * w; a.bs = pp.d_zeroM.as<double>();
                a.col_end = pp.d_cedense.as<int>(); a.y = pp.C.yy + (size_t)j * w; a.dinv = pp.d_dinvM.as<double>() + (size_t)j * w; a.x = pp.C.xx + (size_t)j * w;
                a.fail = failp; a.prof = nullptr; a.no_back = 1;
                pp.h_crtab[pos] = a;
                CholArgs b = a; b.no_back = 0; b.back_from = w / CH_NB;
                pp.h_crtab[K + pos] = b;
            }
        CK(cudaMemcpyAsync(pp.d_crtab.p, pp.h_crtab.data(), sizeof(CholArgs) * 2 * (size_t)K, cudaMemcpyHostToDevice, st));
    }
    return BAGPU_OK;
}

int parts_launch_table(bagpu_ctx *ctx, const CholArgs *table, int count, int cluster, int maxr, int n_max, double lambda, const unsigned *wait, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(count * cluster); cfg.blockDim = dim3(CB_THREADS); cfg.stream = stream; cfg.attrs = at; cfg.numAttrs = 1;
    cfg.dynamicSmemBytes = chol_band_smem(ctx, n_max, maxr);
    CholArgs a; a.lambda = lambda; a.row_done = wait; a.S = nullptr; a.n = 0; a.ld = 0; a.bp = a.bs = nullptr; a.col_end = nullptr;
    a.y = a.dinv = a.x = nullptr; a.fail = nullptr; a.prof = nullptr;
    CK(cudaLaunchKernelEx(&cfg, chol_band_kernel, a, maxr, (int)(cfg.dynamicSmemBytes / sizeof(double)), table));
    return BAGPU_OK;
}
// the P factorisation fronts (one launch); wait != nullptr: the clusters consume camera rows as pair_kernel completes them
int parts_enqueue_factor(bagpu_ctx *ctx, const PartPlan &pp, double lambda, const unsigned *wait, cudaStream_t stream) {
    return parts_launch_table(ctx, pp.d_tab.as<CholArgs>(), pp.P, pp.nc, pp.maxr, pp.n_max, lambda, wait, stream);
}
// spikes, separator system, backward substitutions
int parts_enqueue_rest(bagpu_ctx *ctx, PartPlan &pp, const double *S, const double *bp, const double *bs, double *y, double *x, double lambda, int *failp, cudaStream_t st) {
    const PartTable &T = pp.T;
    const int nsep = pp.P - 1, w = pp.w;
    panel_inverse_kernel<<<(pp.inv_panels + PS_INV_WARPS - 1) / PS_INV_WARPS, 32 * PS_INV_WARPS, 0, st>>>(T, S, pp.ld, pp.d_linv.as<double>(), pp.inv_panels);
    {
        const int ldr = (std::max(32, pp.maxr - CH_NB) + 1) & ~1;                 // rows of L below a panel's diagonal block
        const size_t sm2 = spike_forward2_smem(ldr);
        if (!ctx->opt.spike_v1 && sm2 <= ctx->band_smem_cap) {
            const int per = (w + PS2_NCOL - 1) / PS2_NCOL;
            spike_forward2_kernel<<<nsep * per, PS2_THREADS, sm2, st>>>(T, S, pp.ld, pp.d_subce.as<int>(), pp.d_linv.as<double>(), pp.d_V.as<double>(), ldr, per);
        } else
            spike_forward_kernel<<<pp.spike_ctas, PS_THREADS, 0, st>>>(T, S, pp.ld, pp.d_subce.as<int>(), pp.d_linv.as<double>(), pp.d_V.as<double>());
    }
    spike_gram_kernel<<<pp.gram_ctas, 256, 0, st>>>(T, S, pp.ld, pp.d_V.as<double>(), y, pp.d_Dp.as<double>(), pp.d_Ep.as<double>(), pp.d_gp.as<double>());
    const long long na = (long long)nsep * w * 2 * w;
    if (pp.cr) {
        const CrPlan &C = pp.C;
        const int K = C.K, np = w / 32, nt = w / 32;
        cr_assemble_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(C, pp.d_seprow.as<int>(), S, pp.ld, bp, bs, y, pp.d_Dp.as<double>(), pp.d_Ep.as<double>(), pp.d_gp.as<double>());
        CK(cudaMemsetAsync(pp.d_yy.p, 0, 8 * (size_t)K * w, st));
        const size_t nl = pp.cr_stride.size();
        for (size_t lv = 0; lv < nl; lv++) {
            const int sft = pp.cr_stride[lv], cnt = pp.cr_cnt[lv];
            { int rc = parts_launch_table(ctx, pp.d_crtab.as<CholArgs>() + pp.cr_off[lv], cnt, pp.cr_nc, w, w, lambda, nullptr, st); if (rc) return rc; }
            const int n_keep = (K >= 2 * sft) ? (K - (2 * sft - 1) + 2 * sft - 1) / (2 * sft) : 0;
            if (n_keep > 0) {
                block_inverse_kernel<<<(cnt * np + PS_INV_WARPS - 1) / PS_INV_WARPS, 32 * PS_INV_WARPS, 0, st>>>(C, sft, cnt);
                block_spike_kernel<<<cnt * 2 * (w / PS_NCOL), PS_THREADS, 0, st>>>(C, sft, cnt);
                block_gram_kernel<<<n_keep * (2 * nt * nt + 1), 256, 0, st>>>(C, sft, n_keep);
            }
        }
        for (size_t lv = nl; lv-- > 0;) {
            const int sft = pp.cr_stride[lv], cnt = pp.cr_cnt[lv];
            if (lv + 1 < nl) block_apply_kernel<<<(cnt * w + 7) / 8, 256, 0, st>>>(C, sft, cnt);
            { int rc = parts_launch_table(ctx, pp.d_crtab.as<CholArgs>() + K + pp.cr_off[lv], cnt, 1, w, w, lambda, nullptr, st); if (rc) return rc; }
        }
    } else {
    sep_assemble_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(nsep, w, pp.d_seprow.as<int>(), S, pp.ld, bp, bs, y, pp.d_Dp.as<double>(), pp.d_Ep.as<double>(),
                                                                    pp.d_gp.as<double>(), pp.d_SM.as<double>(), pp.ldM, pp.d_rhsM.as<double>());
    CK(cudaMemsetAsync(pp.d_yM.p, 0, 8 * (size_t)pp.nS, st));
    CholArgs cM; cM.S = pp.d_SM.as<double>(); cM.n = pp.nS; cM.ld = pp.ldM; cM.lambda = lambda; cM.bp = pp.d_rhsM.as<double>(); cM.bs = pp.d_zeroM.as<double>();
    cM.col_end = pp.d_ceM.as<int>(); cM.y = pp.d_yM.as<double>(); cM.dinv = pp.d_dinvM.as<double>(); cM.x = pp.d_xM.as<double>(); cM.fail = failp; cM.prof = nullptr;
    { int rc = launch_chol(ctx, cM, pp.gridM, pp.maxrM, st); if (rc) return rc; }
    }
    sep_scatter_kernel<<<(nsep * w + 255) / 256, 256, 0, st>>>(nsep, w, pp.d_seprow.as<int>(), pp.d_xM.as<double>(), y, x);
    spike_apply_kernel<<<(pp.apply_rows + 7) / 8, 256, 0, st>>>(T, pp.d_V.as<double>(), pp.d_xM.as<double>(), y, pp.apply_rows);
    CK(cudaGetLastError());
    return parts_launch_table(ctx, pp.d_tab.as<CholArgs>() + pp.P, pp.P, 1, pp.maxr, pp.n_max, lambda, nullptr, st);
}

// Block-Jacobi PCG on the band-stored reduced system (pcg.cuh). The host reads the residual every 25 iterations (one small D2H and
// a sync); no convergence within max_iter, or a non-positive curvature, raises the failure flag = a rejected LM trial.
int pcg_solve(bagpu_ctx *ctx, const double *S, int n, int ld, double lambda, const double *bp, const double *bs, double *x, int *failp,
              double tol, int max_iter, int *iters_out) {
    cudaStream_t st = ctx->stream;
    const size_t nn = (size_t)n;
    CK(ctx->d_pcg_vec.ensure(8 * 4 * nn)); CK(ctx->d_pcg_minv.ensure(8 * 36 * (nn / 6 + 1))); CK(ctx->d_pcg_part.ensure(8 * 3 * PCG_MAXB)); CK(ctx->d_pcg_scal.ensure(64));
    PcgArgs A;
    A.S = S; A.n = n; A.ld = ld; A.band = std::min(ld, n - 1); A.lambda = lambda; A.bp = bp; A.bs = bs; A.x = x;
    A.r = ctx->d_pcg_vec.as<double>(); A.z = A.r + nn; A.p = A.z + nn; A.q = A.p + nn;
    A.Minv = ctx->d_pcg_minv.as<double>(); A.part = ctx->d_pcg_part.as<double>(); A.scal = ctx->d_pcg_scal.as<double>();
    const int ncam = n / 6;
    A.nblk = std::max(1, std::min(PCG_MAXB, (ncam + PCG_THREADS - 1) / PCG_THREADS));
    const int nblk_spmv = std::max(1, std::min(PCG_MAXB, (n + PCG_THREADS / 32 - 1) / (PCG_THREADS / 32)));
    if (tol <= 0) tol = 1e-10;
    if (max_iter <= 0) max_iter = 20000;
    CK(cudaMemsetAsync(A.scal, 0, 64, st));
    pcg_prec_kernel<<<(ncam + 127) / 128, 128, 0, st>>>(A, failp);
    pcg_init_kernel<<<A.nblk, PCG_THREADS, 0, st>>>(A);
    pcg_init_finish_kernel<<<1, PCG_THREADS, 0, st>>>(A);
    int it = 0;
    bool converged = false;
    double h[8];
    while (it < max_iter && !converged) {
        const int chunk = std::min(25, max_iter - it);
        for (int k = 0; k < chunk; k++) {
            pcg_spmv_kernel<<<nblk_spmv, PCG_THREADS, 0, st>>>(A);
            pcg_update_kernel<<<A.nblk, PCG_THREADS, 0, st>>>(A, nblk_spmv);
            pcg_dir_kernel<<<A.nblk, PCG_THREADS, 0, st>>>(A);
        }
        it += chunk;
        ctx->tm.total_launches += 3 * chunk;
        CK(cudaMemcpyAsync(ctx->h_status.as<double>() + 16, A.scal, 64, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        memcpy(h, ctx->h_status.as<double>() + 16, 64);
        const double rr = h[3], bb = h[4];
        if (!(rr == rr) || !(bb == bb)) break;                               // NaN: give up
        if (bb == 0.0 || rr <= tol * tol * bb) converged = true;
        if (!(h[1] > 0.0) && bb > 0.0) break;                                // p.q <= 0: not positive definite
    }
    if (!converged) { static const int one = 1; CK(cudaMemcpyAsync(failp, &one, sizeof(int), cudaMemcpyHostToDevice, st)); }
    if (ctx->opt.debug) fprintf(stderr, "[bagpu] pcg: %d iterations, |r|/|b| = %.3e%s\n", it, (h[4] > 0) ? std::sqrt(h[3] / h[4]) : 0.0, converged ? "" : " (NOT converged)");
    ctx->tm.pcg_iterations += it;
    if (iters_out) *iters_out = it;
    CK(cudaGetLastError());
    return BAGPU_OK;
}

// the camera half of the Schur complement: pair_tile_kernel over (tile, landmark) records, or pair_kernel over the pair list
// parts_of_pass: 1 = the Dr sums of the diagonal tiles, 2 = the tiles themselves (items [item_begin, item_end), -1 = to the end), 3 = both
void launch_pair(bagpu_ctx *ctx, const SysLayout &L, int grid, unsigned *row_done, double *hpp_diag, double *S2, int n1, cudaStream_t st, const LmDev *lm = nullptr,
                 int parts_of_pass = 3, int item_begin = 0, int item_end = -1) {
    if (ctx->tiles) {
        TileArgs A; A.items = ctx->d_items.as<TileItem>(); A.n_items = ctx->n_items; A.recs = ctx->d_tp_recs.as<TileRec>();
        A.Z = ctx->d_Z.as<double>(); A.Dr = ctx->d_Dr.as<double>(); A.S = L.S; A.ld = ctx->ld; A.bp = L.bp; A.bs = L.bs;
        A.part = ctx->d_part.as<double>(); A.blk_done = ctx->d_blkdone.as<unsigned>(); A.tbw1 = ctx->tbw1; A.tile_pos = ctx->d_tile_pos.as<int>();
        A.row_done = row_done; A.n_free = ctx->n_free; A.S2 = S2; A.n_tot = ctx->n_sys; A.n1 = n1; A.hpp_diag = hpp_diag; A.work = ctx->d_tp_work.as<unsigned>(); A.lm = lm;
        if (ctx->opt.tile_fma) { pair_tile_kernel<<<grid, TP_THREADS, TP_SMEM_BYTES, st>>>(A); return; }
        // the Dr sums of the diagonal tiles first (they only need the stage's records), then the tiles on the FP64 tensor pipe
        if (parts_of_pass & 1) tile_diag_kernel<<<std::min(grid, ctx->diag_grid), TP_THREADS, 0, st>>>(A);
        A.item_begin = item_begin; A.item_end = item_end;
        if (!hpp_diag && (parts_of_pass & 2)) pair_tile_mma_kernel<<<grid, TP_THREADS, TM_SMEM_BYTES, st>>>(A);
        return;
    }
    PairArgs PA; PA.items = ctx->d_items.as<PairItem>(); PA.n_items = ctx->n_items; PA.entries = ctx->d_entries.as<int2>();
    PA.Z = ctx->d_Z.as<double>(); PA.Dr = ctx->d_Dr.as<double>(); PA.S = L.S; PA.ld = ctx->ld; PA.bp = L.bp; PA.bs = L.bs;
    PA.part = ctx->d_part.as<double>(); PA.blk_done = ctx->d_blkdone.as<unsigned>(); PA.bw1 = ctx->band_blocks + 1;
    PA.row_done = row_done; PA.hpp_diag = hpp_diag; PA.S2 = S2; PA.n_tot = ctx->n_sys; PA.n1 = n1;
    pair_kernel<<<grid, PK_THREADS, PK_SMEM_BYTES, st>>>(PA);
}

// Chained mode: a whole optimize() of a small map without a host round trip per trial. lambda, the accept / reject decision, the
// buffer swap and the stop rules live on the device (LmDev, lm_decide_kernel); the host queues trials a few at a time, reads the state
// once per chunk (where it also polls the caller's stop flag) and stops queueing when the device reports the loop finished -- trials
// queued beyond that point return at their first instruction. The solve runs after the accumulation on the same stream (no second
// stream, no spinning clusters); maps with the two-way / partitioned solvers, PCG or more than one rank use the host-stepped loop.
bool chain_eligible(const bagpu_ctx *ctx, const bagpu_schedule *s) {
    return !ctx->opt.no_chain && ctx->world == 1 && ctx->tiles && !ctx->opt.tile_fma && !ctx->opt.no_tiles && !ctx->opt.compare && !ctx->opt.update_relin &&
           s->linear_solver != BAGPU_SOLVER_PCG && !ctx->tw.on && !ctx->parts.on && ctx->n_sys > 0 && ctx->n_items > 0 && ctx->n_tasks > 0 &&
           ctx->n_obs <= ctx->opt.chain_obs;
}

int optimize_chained(bagpu_ctx *ctx, const bagpu_schedule *s, int round, int iterations, int64_t n_active, bagpu_result *res, int *status_out) {
    cudaStream_t st = ctx->stream;
    BaDev D = make_dev(ctx, s->delta_mono, s->delta_stereo);
    const int n = ctx->n_sys, G = ctx->build_grid, PS = ctx->parts_stride;
    const SysLayout L = sys_layout(ctx);
    double *parts = ctx->d_parts.as<double>();
    double *part_chi_b = parts, *part_max = parts + PS, *part_chi_u = parts + 2 * PS, *part_scale = parts + 3 * PS, *part_chi_w = parts + 4 * PS,
           *part_chi_uw = parts + 5 * PS, *part_scale_w = parts + 6 * PS;
    double *dstat = ctx->d_status.as<double>();
    auto stop = [&]() { return s->stop_flag && *s->stop_flag; };
    *status_out = BAGPU_OK;
    if (iterations <= 0) return BAGPU_OK;
    if (stop()) { *status_out = BAGPU_STOPPED; return BAGPU_OK; }
    CK(ctx->d_lm.ensure(sizeof(LmDev))); CK(ctx->h_lm.ensure(sizeof(LmDev) + sizeof(LmTraceDev) * (size_t)iterations));
    CK(ctx->d_lm_trace.ensure(sizeof(LmTraceDev) * (size_t)iterations));
    LmDev *lm = ctx->d_lm.as<LmDev>();
    LmTraceDev *dtrace = ctx->d_lm_trace.as<LmTraceDev>();
    double *pose_a = ctx->pose_cur, *pose_b = ctx->pose_trial, *pt_a = ctx->pt_cur, *pt_b = ctx->pt_trial;
    StageArgs SA; SA.tasks = ctx->d_tasks.as<int2>(); SA.n_tasks = ctx->n_tasks; SA.lm_list = ctx->d_widelist.as<int>(); SA.n_list = ctx->n_wide;
    SA.Z = ctx->d_Z.as<double>(); SA.Dr = ctx->d_Dr.as<double>(); SA.Lm = ctx->d_Lm.as<double>(); SA.fail = L.fail;
    {   // computeLambdaInit: the same stage / diagonal kernels as the trials, lambda set on the device
        CK(cudaMemsetAsync(L.hpp, 0, sizeof(double) * n, st));
        cudaEvent_t e0 = get_event(ctx), e1 = get_event(ctx);
        cudaEventRecord(e0, st);
        StageArgs S0 = SA; S0.lambda = 1.0; S0.part_chi2 = part_chi_b; S0.part_maxdiag = part_max;
        stage_kernel<<<ctx->stage_grid, ST_THREADS, 0, st>>>(D, pose_a, pt_a, S0);
        int n_part0 = ctx->stage_grid;
        if (ctx->n_wide > 0) {
            StageArgs SW = S0; SW.part_chi2 = part_chi_b + ctx->stage_grid; SW.part_maxdiag = part_max + ctx->stage_grid;
            stage_wide_kernel<<<ctx->stage_wide_grid, ST_THREADS, 0, st>>>(D, pose_a, pt_a, SW);
            n_part0 += ctx->stage_wide_grid; ctx->tm.total_launches++;
        }
        launch_pair(ctx, L, ctx->pair_grid, nullptr, L.hpp, nullptr, n, st);
        cudaEventRecord(e1, st);
        ctx->pending.push_back({e0, e1, EV_BUILD});
        reduce_partials_kernel<<<1, 256, 0, st>>>(n_part0, part_chi_b, nullptr, part_max, nullptr, L.hpp, n, dstat);
        lm_init_kernel<<<1, 32, 0, st>>>(lm, dstat, s->lambda_init, iterations, round, iterations);
        ctx->tm.total_launches += 4;
        ctx->tm.edge_linearisations += n_active;
    }
    struct TrialEv { cudaEvent_t e[4]; };
    std::vector<TrialEv> tev;
    auto enqueue_trial = [&]() -> int {
        TrialEv T;
        for (int k = 0; k < 4; k++) T.e[k] = get_event(ctx);
        CK(cudaMemsetAsync(L.S, 0, sizeof(double) * (L.sys_count + ctx->scratch_elems), st));
        cudaEventRecord(T.e[0], st);
        StageArgs S1 = SA; S1.lambda = 0.0; S1.part_chi2 = part_chi_b; S1.part_maxdiag = nullptr; S1.lm = lm; S1.pose_b = pose_b; S1.pt_b = pt_b;
        stage_kernel<<<ctx->stage_grid, ST_THREADS, 0, st>>>(D, pose_a, pt_a, S1);
        if (ctx->n_wide > 0) {
            StageArgs SW = S1; SW.part_chi2 = part_chi_w;
            stage_wide_kernel<<<ctx->stage_wide_grid, ST_THREADS, 0, st>>>(D, pose_a, pt_a, SW);
            ctx->tm.total_launches++;
        }
        launch_pair(ctx, L, ctx->pair_grid, nullptr, nullptr, nullptr, n, st, lm);
        cudaEventRecord(T.e[1], st);
        CholArgs ca; ca.S = L.S; ca.n = n; ca.ld = ctx->ld; ca.lambda = 0.0; ca.bp = L.bp; ca.bs = L.bs; ca.prof = nullptr; ca.x = ctx->d_xp.as<double>(); ca.y = L.y1;
        ca.dinv = ctx->d_dinv.as<double>(); ca.col_end = ctx->d_colend.as<int>(); ca.fail = L.fail; ca.lm = lm;
        if (n <= CS_MAX_N && ctx->opt.small_chol)               // opt-in: one CTA, packed in shared memory, one barrier per column (measured slower)
            chol_small_kernel<<<1, CS_THREADS, chol_small_smem(n), st>>>(L.S, n, ctx->ld, 0.0, L.bp, L.bs, ctx->d_xp.as<double>(), L.fail, lm);
        else { int rc = launch_chol(ctx, ca, ctx->chol_grid, ctx->chol_maxr, st); if (rc) return rc; }
        cudaEventRecord(T.e[2], st);
        pose_update_kernel<<<1, 256, 0, st>>>(ctx->n_poses, ctx->d_hidx.as<int>(), pose_a, pose_b, ctx->d_xp.as<double>(), L.bp, 0.0, dstat + 4, lm);
        UpdateOut U; U.lambda = 0.0; U.xp = ctx->d_xp.as<double>(); U.pose_trial = pose_b; U.pt_trial = pt_b;
        U.edge_chi2 = ctx->d_chi2.as<double>(); U.part_chi2 = part_chi_u; U.part_scale = part_scale; U.lm_list = nullptr; U.n_list = 0;
        U.lm = lm; U.pose_a = pose_a; U.pose_b = pose_b; U.pt_a = pt_a; U.pt_b = pt_b;
        UpdateTasks K; K.tasks = ctx->d_tasks.as<int2>(); K.n_tasks = ctx->n_tasks;
        update_z_kernel<<<ctx->updz_grid, ST_THREADS, 0, st>>>(D, pt_a, U, K, ctx->d_Z.as<double>(), ctx->d_Lm.as<double>());
        const bool upd_wide = ctx->n_wide > 0;
        if (upd_wide) {
            UpdateOut UW = U; UW.part_chi2 = part_chi_uw; UW.part_scale = part_scale_w; UW.lm_list = ctx->d_widelist.as<int>(); UW.n_list = ctx->n_wide;
            update_kernel<<<G, BUILD_THREADS, 0, st>>>(D, pose_a, pt_a, UW);
            ctx->tm.total_launches++;
        }
        cudaEventRecord(T.e[3], st);
        TrialSums TS;
        TS.a[0] = part_chi_b; TS.na[0] = ctx->stage_grid; TS.b[0] = ctx->n_wide > 0 ? part_chi_w : nullptr; TS.nb[0] = ctx->stage_wide_grid;
        TS.a[1] = part_chi_u; TS.na[1] = ctx->updz_grid; TS.b[1] = upd_wide ? part_chi_uw : nullptr; TS.nb[1] = G;
        TS.a[2] = part_scale; TS.na[2] = ctx->updz_grid; TS.b[2] = upd_wide ? part_scale_w : nullptr; TS.nb[2] = G;
        TS.fail = L.fail; TS.out = dstat;
        lm_decide_kernel<<<1, 256, 0, st>>>(TS, dstat + 4, lm, dtrace);
        ctx->tm.total_launches += 8;
        tev.push_back(T);
        CK(cudaGetLastError());
        return BAGPU_OK;
    };
    LmDev h; memset(&h, 0, sizeof(h));
    h.iterations = iterations;
    bool stopped = false;
    for (;;) {
        const int chunk = std::max(1, std::min(4, iterations - h.it));
        for (int k = 0; k < chunk; k++) { int rc = enqueue_trial(); if (rc) return rc; }
        CK(cudaMemcpyAsync(ctx->h_lm.p, lm, sizeof(LmDev), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        memcpy(&h, ctx->h_lm.p, sizeof(LmDev));
        if (h.done) break;
        if (stop()) { stopped = true; break; }
    }
    if (h.starved) return fail(ctx, BAGPU_ERR_CUDA, "reduced-system solve reported a wait time-out in chained mode");
    // per-trial phase times (events on the library stream) -> call totals and the per-iteration records
    std::vector<float> tb(h.trials, 0.f), ts(h.trials, 0.f), tu(h.trials, 0.f), tt(h.trials, 0.f);
    for (int k = 0; k < h.trials && k < (int)tev.size(); k++) {
        cudaEventElapsedTime(&tb[k], tev[k].e[0], tev[k].e[1]); cudaEventElapsedTime(&ts[k], tev[k].e[1], tev[k].e[2]);
        cudaEventElapsedTime(&tu[k], tev[k].e[2], tev[k].e[3]); cudaEventElapsedTime(&tt[k], tev[k].e[0], tev[k].e[3]);
        ctx->tm.build_ms += tb[k]; ctx->tm.build_launches++; ctx->tm.linsolve_ms += ts[k]; ctx->tm.linsolve_launches++;
        ctx->tm.update_ms += tu[k]; ctx->tm.update_launches++;
    }
    resolve_events(ctx);
    if (h.n_trace > 0) {
        LmTraceDev *ht = reinterpret_cast<LmTraceDev *>(ctx->h_lm.as<char>() + sizeof(LmDev));
        CK(cudaMemcpyAsync(ht, dtrace, sizeof(LmTraceDev) * (size_t)h.n_trace, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        int k0 = 0;
        for (int i = 0; i < h.n_trace; i++) {
            const LmTraceDev &d = ht[i];
            if (res && res->trace && res->n_trace < s->max_trace) {
                bagpu_trace &t = res->trace[res->n_trace++];
                t.round = round; t.iteration = d.iteration; t.chi2_before = d.chi2_before; t.chi2_after = d.chi2_after; t.lambda = d.lambda; t.trials = d.trials; t.status = d.status;
                t.active_edges = n_active;
                double b = 0, sv = 0, u = 0, w = 0;
                for (int k = k0; k < k0 + d.trials && k < h.trials; k++) { b += tb[k]; sv += ts[k]; u += tu[k]; w += tt[k]; }
                t.linearise_schur_us = 1e3 * b; t.linear_solve_us = 1e3 * sv; t.update_us = 1e3 * u; t.iteration_us = 1e3 * w;
            }
            k0 += d.trials;
        }
    }
    ctx->tm.lm_iterations += h.it; ctx->tm.lm_trials += h.trials;
    ctx->tm.edge_linearisations += n_active * (int64_t)h.trials; ctx->tm.edge_evaluations += n_active * (int64_t)h.trials;
    if (h.flip) { std::swap(ctx->pose_cur, ctx->pose_trial); std::swap(ctx->pt_cur, ctx->pt_trial); }
    int status = h.status;
    if ((stopped || stop()) && status == BAGPU_OK) status = BAGPU_STOPPED;
    *status_out = status;
    return BAGPU_OK;
}

int optimize(bagpu_ctx *ctx, const bagpu_schedule *s, int round, int iterations, int64_t n_active, bagpu_result *res, int *status_out) {
    cudaStream_t st = ctx->stream;
    BaDev D = make_dev(ctx, s->delta_mono, s->delta_stereo);
    const int n = ctx->n_sys, ld = ctx->ld, G = ctx->build_grid, PS = ctx->parts_stride;
    const SysLayout L = sys_layout(ctx);
    double *S = L.S, *bp = L.bp, *bs = L.bs, *S2 = L.S2;
    const size_t sys_count = L.sys_count;                                                       // what a trial all-reduces
    double *y1p = L.y1, *y2p = L.y2, *yMp = L.yM;                                               // trial scratch, zeroed with the system
    unsigned *rowdone_p = L.row_done;
    int *fail_p = L.fail;
    double *hpp = L.hpp;
    double *parts = ctx->d_parts.as<double>();
    double *part_chi_b = parts, *part_max = parts + PS, *part_chi_u = parts + 2 * PS, *part_scale = parts + 3 * PS, *part_chi_w = parts + 4 * PS,
           *part_chi_uw = parts + 5 * PS, *part_scale_w = parts + 6 * PS;
    double *dstat = ctx->d_status.as<double>();
    auto stop = [&]() { return s->stop_flag && *s->stop_flag; };

    int status = BAGPU_OK;
    double lambda = -1, ni = 2; int nBad = 0;
    bool ok = true;
    for (int it = 0; it < iterations && !stop() && ok; it++) {
        double currentChi = 0, iniChi = 0;
        const auto it_t0 = std::chrono::steady_clock::now();
        const double ph0[3] = {ctx->tm.build_ms, ctx->tm.linsolve_ms, ctx->tm.update_ms};
        if (it == 0) {
            // computeLambdaInit: tau * max diagonal of Hpp and Hll (optimization_algorithm_levenberg.cpp:171-185)
            CK(cudaMemsetAsync(hpp, 0, sizeof(double) * std::max(1, n), st));
            int n_part0 = G;
            if (n > 0 && ctx->n_items > 0 && !ctx->opt.no_tiles) {
                // the same stage / pair kernels as the trials (fixed summation order): Dr records, then the diagonal blocks only
                ScopedEv ev(ctx, EV_BUILD);
                StageArgs SA; SA.tasks = ctx->d_tasks.as<int2>(); SA.n_tasks = ctx->n_tasks; SA.lm_list = ctx->d_widelist.as<int>(); SA.n_list = ctx->n_wide;
                SA.Z = ctx->d_Z.as<double>(); SA.Dr = ctx->d_Dr.as<double>(); SA.Lm = ctx->d_Lm.as<double>(); SA.lambda = 1.0; SA.part_chi2 = part_chi_b; SA.part_maxdiag = part_max;
                SA.fail = fail_p;
                stage_kernel<<<ctx->stage_grid, ST_THREADS, 0, st>>>(D, ctx->pose_cur, ctx->pt_cur, SA);
                n_part0 = ctx->stage_grid;
                if (ctx->n_wide > 0) {
                    StageArgs SW = SA; SW.part_chi2 = part_chi_b + ctx->stage_grid; SW.part_maxdiag = part_max + ctx->stage_grid;
                    stage_wide_kernel<<<ctx->stage_wide_grid, ST_THREADS, 0, st>>>(D, ctx->pose_cur, ctx->pt_cur, SW);
                    n_part0 += ctx->stage_wide_grid; ctx->tm.total_launches++;
                }
                launch_pair(ctx, L, ctx->pair_grid, nullptr, hpp, nullptr, n, st);
                ctx->tm.total_launches++;
            } else {
                BuildOut O; O.lambda = 0; O.mode = 0; O.S = S; O.ld = ld; O.bp = bp; O.bs = bs; O.hpp_diag = hpp;
                O.part_chi2 = part_chi_b; O.part_maxdiag = part_max; O.lm_list = nullptr; O.n_list = 0;
                ScopedEv ev(ctx, EV_BUILD); build_kernel<<<G, BUILD_THREADS, 0, st>>>(D, ctx->pose_cur, ctx->pt_cur, O);
            }
            ctx->tm.total_launches++;
            int rc = all_reduce_sum(ctx, hpp, std::max(1, n)); if (rc) return rc;
            reduce_partials_kernel<<<1, 256, 0, st>>>(n_part0, part_chi_b, nullptr, part_max, nullptr, hpp, n, dstat);
            ctx->tm.total_launches++;
            if (ctx->world > 1) { rc = all_reduce_sum(ctx, dstat, 1); if (rc) return rc; rc = all_reduce_max(ctx, dstat + 2, 1); if (rc) return rc; }
            double h[3];
            rc = read_status(ctx, h, 3); if (rc) return rc;
            lambda = (s->lambda_init > 0) ? s->lambda_init : 1e-5 * h[2];
            ni = 2; nBad = 0;
            ctx->tm.edge_linearisations += n_active;
        }
        double rho = 0; int qmax = 0;
        bool first = true;
        bool redo = false;
        do {
            redo = false;
            // buildSystem + setLambda + Schur complement, scattered straight into the reduced system
            CK(cudaMemsetAsync(S, 0, sizeof(double) * (sys_count + ctx->scratch_elems), st));      // system + y, y2, yM, row_done, fail
            // Single GPU, band solver: the Cholesky cluster starts beside pair_kernel and consumes block columns as their camera
            // rows complete (row_done counters), so the accumulation of the reduced system hides behind the factorisation chain.
            const bool use_pcg = s->linear_solver == BAGPU_SOLVER_PCG && n > 0;
            const bool overlap = !use_pcg && !ctx->opt.no_overlap && !ctx->overlap_off && n > 0 && ctx->world == 1 && ctx->chol_maxr > 0 && ctx->n_items > 0 &&
                                 (!ctx->parts.on || ctx->parts.P * ctx->parts.nc <= ctx->n_sm / 4);     // many spinning clusters would starve pair_kernel
            CholArgs ca; ca.S = S; ca.n = n; ca.ld = ld; ca.lambda = lambda; ca.bp = bp; ca.bs = bs;
            ca.prof = nullptr; ca.x = ctx->d_xp.as<double>(); ca.y = y1p; ca.dinv = ctx->d_dinv.as<double>(); ca.col_end = ctx->d_colend.as<int>(); ca.fail = fail_p;
            // the linear solve on stream sc: one cluster, or (two-way) two clusters from both ends of the band + the separator.
            // Two parts, so that the host can put stage_kernel / pair_kernel in the queue between them: the factorisation launches
            // (which must be queued before the stage to own their SMs) and everything that runs after pair_kernel has ended anyway.
            CholArgs c1, c2;
            auto enqueue_solver_head = [&](cudaStream_t sc, bool waits) -> int {
                if (waits) { ca.row_done = rowdone_p; ca.item_off = ctx->d_itemoff.as<unsigned>(); ca.bw1 = ctx->wait_bw1; ca.row_pos = ctx->wait_rowpos; }
                if (ctx->parts.on) { ctx->tm.total_launches++; return parts_enqueue_factor(ctx, ctx->parts, lambda, waits ? rowdone_p : nullptr, sc); }
                if (!ctx->tw.on) { ctx->tm.total_launches++; return launch_chol(ctx, ca, ctx->chol_grid, ctx->chol_maxr, sc); }
                const bagpu_ctx::TwoWay &T = ctx->tw;
                cudaStream_t s2 = ctx->stream_chol2;
                c1 = ca; c1.n = T.n1; c1.col_end = ctx->d_colend1.as<int>(); c1.p_stop = T.k;
                c2 = ca; c2.S = S2; c2.n = T.n2; c2.col_end = ctx->d_colend2.as<int>(); c2.y = y2p; c2.p_stop = T.k;
                c2.mirror_n = n; c2.wait_band = ld + 1;
                CK(cudaEventRecord(ctx->ev_tw[0], sc)); CK(cudaStreamWaitEvent(s2, ctx->ev_tw[0], 0));
                int r = launch_chol(ctx, c1, T.grid1, T.maxr1, sc); if (r) return r;
                r = launch_chol(ctx, c2, T.grid2, T.maxr2, s2); if (r) return r;
                ctx->tm.total_launches += 2;
                return BAGPU_OK;
            };
            auto enqueue_solver_tail = [&](cudaStream_t sc) -> int {
                if (ctx->parts.on) { ctx->tm.total_launches += 8; return parts_enqueue_rest(ctx, ctx->parts, S, bp, bs, y1p, ctx->d_xp.as<double>(), lambda, fail_p, sc); }
                if (!ctx->tw.on) return BAGPU_OK;
                const bagpu_ctx::TwoWay &T = ctx->tw;
                cudaStream_t s2 = ctx->stream_chol2;
                CK(cudaEventRecord(ctx->ev_tw[1], s2)); CK(cudaStreamWaitEvent(sc, ctx->ev_tw[1], 0));
                tw_merge_kernel<<<grid_for((int64_t)T.nM * T.nM, 256), 256, 0, sc>>>(T.nM, T.rT, n, ld, S, S2, bp, bs, y1p, y2p,
                                                                                  ctx->d_SM.as<double>(), T.ldM, ctx->d_rhsM.as<double>());
                CholArgs cM; cM.S = ctx->d_SM.as<double>(); cM.n = T.nM; cM.ld = T.ldM; cM.lambda = lambda; cM.bp = ctx->d_rhsM.as<double>(); cM.bs = ctx->d_zeroM.as<double>();
                cM.col_end = ctx->d_colendM.as<int>(); cM.y = yMp; cM.dinv = ctx->d_dinv.as<double>(); cM.x = ctx->d_xM.as<double>(); cM.fail = fail_p; cM.prof = nullptr;
                int r = launch_chol(ctx, cM, T.gridM, T.maxrM, sc); if (r) return r;
                tw_scatter_kernel<<<grid_for(T.nM, 128), 128, 0, sc>>>(T.nM, T.rT, n, ctx->d_xM.as<double>(), y1p, y2p, ctx->d_xp.as<double>());
                CK(cudaEventRecord(ctx->ev_tw[2], sc)); CK(cudaStreamWaitEvent(s2, ctx->ev_tw[2], 0));
                CholArgs b1 = c1; b1.p_stop = 0; b1.back_from = T.k; b1.row_done = nullptr;
                CholArgs b2 = c2; b2.p_stop = 0; b2.back_from = T.k; b2.row_done = nullptr;
                r = launch_chol(ctx, b1, 1, T.maxr1, sc); if (r) return r;
                r = launch_chol(ctx, b2, 1, T.maxr2, s2); if (r) return r;
                CK(cudaEventRecord(ctx->ev_tw[3], s2)); CK(cudaStreamWaitEvent(sc, ctx->ev_tw[3], 0));
                ctx->tm.total_launches += 5;
                return BAGPU_OK;
            };
            const int chol_sms = ctx->parts.on ? ctx->parts.P * ctx->parts.nc : (ctx->tw.on ? ctx->tw.grid1 + ctx->tw.grid2 : ctx->chol_grid);
            BuildOut O; O.lambda = lambda; O.mode = 1; O.S = S; O.ld = ld; O.bp = bp; O.bs = bs; O.hpp_diag = hpp;
            O.part_chi2 = part_chi_b; O.part_maxdiag = part_max; O.lm_list = nullptr; O.n_list = 0;
            const bool tiled = n > 0 && !ctx->opt.no_tiles;
            const bool chunked = tiled && ctx->chunks.on && ctx->parts.on && ctx->world > 1 && !use_pcg;
            ScopedEv *ev_solve = nullptr;                     // overlap: the solve's timing bracket spans stage and pair on the other stream
            bool have_wide_part = false;
            int n_part_b = G, n_part_w = G;
            {
                ScopedEv ev(ctx, EV_BUILD);
                if (tiled) {
                    // Host order: stage_kernel goes into the queue first and on all SMs; the factorisation clusters are queued right behind it
                    // (their stream only waits for the memsets) and take their SMs as the stage drains, ahead of pair_kernel, which
                    // cannot start before the stage has ended. BAGPU_STAGE_FIRST=0 queues the clusters before the stage instead.
                    const int stage_first = ctx->opt.stage_first;
                    auto start_solver = [&]() -> int {
                        CK(cudaStreamWaitEvent(ctx->stream_chol, ctx->ev_fork, 0));
                        ev_solve = new ScopedEv(ctx, EV_LINSOLVE, ctx->stream_chol);
                        int rcc = enqueue_solver_head(ctx->stream_chol, true); if (rcc) { delete ev_solve; ev_solve = nullptr; }
                        return rcc;
                    };
                    if (overlap) {
                        CK(cudaEventRecord(ctx->ev_fork, st));
                        if (!stage_first) { int rcc = start_solver(); if (rcc) return rcc; }
                    }
                    const int sm_avail = overlap ? std::max(1, ctx->n_sm - chol_sms) : ctx->n_sm;
                    const int sgrid = std::max(1, std::min(ctx->stage_grid, (stage_first == 2 ? ctx->n_sm : sm_avail) * ctx->stage_occ));
                    StageArgs SA; SA.tasks = ctx->d_tasks.as<int2>(); SA.n_tasks = ctx->n_tasks; SA.lm_list = ctx->d_widelist.as<int>(); SA.n_list = ctx->n_wide;
                    SA.Z = ctx->d_Z.as<double>(); SA.Dr = ctx->d_Dr.as<double>(); SA.Lm = ctx->d_Lm.as<double>(); SA.lambda = lambda; SA.part_chi2 = part_chi_b; SA.part_maxdiag = nullptr; SA.fail = fail_p;
                    stage_kernel<<<sgrid, ST_THREADS, 0, st>>>(D, ctx->pose_cur, ctx->pt_cur, SA);
                    n_part_b = sgrid;
                    if (overlap && stage_first) { int rcc = start_solver(); if (rcc) return rcc; }
                    if (ctx->n_wide > 0) {             // landmarks with more than 32 observations: warp = landmark
                        StageArgs SW = SA; SW.part_chi2 = part_chi_w;
                        stage_wide_kernel<<<ctx->stage_wide_grid, ST_THREADS, 0, st>>>(D, ctx->pose_cur, ctx->pt_cur, SW);
                        have_wide_part = true; n_part_w = ctx->stage_wide_grid; ctx->tm.total_launches++;
                    }
                    if (ctx->n_items > 0 && chunked) {
                        // multi-GPU, partitioned solver: Dr sums first (b_p, b_s and the Hpp blocks are complete after them), then the tiles chunk by
                        // chunk of fronts; chunk k's rows are all-reduced on the communication stream and factored on their own stream while
                        // pair_tile_mma_kernel works on chunk k + 1
                        const bagpu_ctx::ChunkPlan &C = ctx->chunks;
                        cudaEvent_t *ev = ctx->ev_chunk;
                        launch_pair(ctx, L, ctx->pair_grid, nullptr, nullptr, nullptr, n, st, nullptr, 1);
                        CK(cudaEventRecord(ev[0], st)); CK(cudaStreamWaitEvent(ctx->stream_comm, ev[0], 0));
                        { int rcc = all_reduce_sum(ctx, bp, 2 * (size_t)n, ctx->stream_comm); if (rcc) return rcc; }
                        for (int k = 0; k < C.n; k++) {
                            launch_pair(ctx, L, ctx->pair_grid, nullptr, nullptr, nullptr, n, st, nullptr, 2, C.item0[k], C.item0[k + 1]);
                            CK(cudaEventRecord(ev[1 + k], st)); CK(cudaStreamWaitEvent(ctx->stream_comm, ev[1 + k], 0));
                            { int rcc = all_reduce_sum(ctx, S + C.elem0[k], C.elem0[k + 1] - C.elem0[k], ctx->stream_comm); if (rcc) return rcc; }
                            CK(cudaEventRecord(ev[1 + C.n + k], ctx->stream_comm)); CK(cudaStreamWaitEvent(ctx->stream_front[k], ev[1 + C.n + k], 0));
                            { int rcc = parts_launch_table(ctx, ctx->parts.d_tab.as<CholArgs>() + C.part0[k], C.part0[k + 1] - C.part0[k], ctx->parts.nc, ctx->parts.maxr,
                                                           ctx->parts.n_max, lambda, nullptr, ctx->stream_front[k]); if (rcc) return rcc; }
                            CK(cudaEventRecord(ev[1 + 2 * C.n + k], ctx->stream_front[k]));
                            ctx->tm.total_launches += 2;
                        }
                        ctx->tm.total_launches++;
                    } else if (ctx->n_items > 0) {
                        const int pgrid = std::max(1, std::min(ctx->pair_grid, sm_avail * ctx->pair_occ));
                        launch_pair(ctx, L, pgrid, overlap ? rowdone_p : nullptr, nullptr, (ctx->tw.on && !use_pcg) ? S2 : nullptr, ctx->tw.n1, st);      // PCG reads the whole system from S
                        ctx->tm.total_launches++;
                    }
                    if (overlap) {                     // the rest of the solve goes into the queue behind the factorisations
                        int rcc = enqueue_solver_tail(ctx->stream_chol);
                        delete ev_solve; ev_solve = nullptr;
                        if (rcc) return rcc;
                        CK(cudaEventRecord(ctx->ev_join, ctx->stream_chol));
                    }
                } else {
                    build_kernel<<<G, BUILD_THREADS, 0, st>>>(D, ctx->pose_cur, ctx->pt_cur, O);
                }
            }
            if (tiled && ctx->opt.compare) {      // debug: the same system through the global-atomic kernel, compared element-wise
                const size_t cnt = ctx->s_elems + 2 * (size_t)n;
                std::vector<double> a(cnt), b2(cnt);
                cudaStreamSynchronize(st);
                cudaMemcpy(a.data(), S, 8 * cnt, cudaMemcpyDeviceToHost);
                cudaMemset(S, 0, 8 * cnt);
                BuildOut OC = O; OC.lm_list = nullptr; OC.n_list = 0; OC.part_chi2 = part_chi_w;
                build_kernel<<<G, BUILD_THREADS, 0, st>>>(D, ctx->pose_cur, ctx->pt_cur, OC);
                cudaStreamSynchronize(st);
                cudaMemcpy(b2.data(), S, 8 * cnt, cudaMemcpyDeviceToHost);
                cudaMemcpy(S, a.data(), 8 * cnt, cudaMemcpyHostToDevice);
                double worst = 0; size_t wi = 0; int nbad = 0;
                for (size_t i = 0; i < cnt; i++) { const double d = fabs(a[i] - b2[i]) / (1e-9 + fabs(b2[i])); if (d > 1e-9) nbad++; if (d > worst || a[i] != a[i]) { worst = d; wi = i; } }
                const size_t R = wi / ld, Cc = wi % ld;
                fprintf(stderr, "[bagpu] compare tiled vs atomic: worst rel %.3e at %zu (row %zu col %zu; S elems %zu) tiled %.6e atomic %.6e, %d elements differ\n",
                        worst, wi, R, Cc, ctx->s_elems, a[wi], b2[wi], nbad);
            }
            ctx->tm.total_launches++;
            ctx->tm.edge_linearisations += n_active;
            int rc = chunked ? BAGPU_OK : all_reduce_sum(ctx, S, sys_count); if (rc) return rc;
            if (overlap) CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
            else if (n > 0 && chunked) {
                ScopedEv ev(ctx, EV_LINSOLVE);                  // from the end of the accumulation: what is left of the fronts, then spikes, separators, backward passes
                for (int k = 0; k < ctx->chunks.n; k++) CK(cudaStreamWaitEvent(st, ctx->ev_chunk[1 + 2 * ctx->chunks.n + k], 0));
                ctx->tm.total_launches += 8;
                rc = parts_enqueue_rest(ctx, ctx->parts, S, bp, bs, y1p, ctx->d_xp.as<double>(), lambda, fail_p, st); if (rc) return rc;
            } else if (n > 0) {
                ScopedEv ev(ctx, EV_LINSOLVE);
                if (use_pcg) { rc = pcg_solve(ctx, S, n, ld, lambda, bp, bs, ctx->d_xp.as<double>(), fail_p, s->pcg_tolerance, s->pcg_max_iterations, nullptr); if (rc) return rc; }
                else { rc = enqueue_solver_head(st, false); if (rc) return rc; rc = enqueue_solver_tail(st); if (rc) return rc; }
            }
            pose_update_kernel<<<1, 256, 0, st>>>(ctx->n_poses, ctx->d_hidx.as<int>(), ctx->pose_cur, ctx->pose_trial,
                                                  ctx->d_xp.as<double>(), bp, lambda, dstat + 4);
            UpdateOut U; U.lambda = lambda; U.xp = ctx->d_xp.as<double>(); U.pose_trial = ctx->pose_trial; U.pt_trial = ctx->pt_trial;
            U.edge_chi2 = ctx->d_chi2.as<double>(); U.part_chi2 = part_chi_u; U.part_scale = part_scale; U.lm_list = nullptr; U.n_list = 0;
            const bool packed = ctx->n_tasks > 0 && !ctx->opt.no_tiles;
            int n_part_u = G; bool upd_wide = false;
            {
                ScopedEv ev(ctx, EV_UPDATE);
                if (packed) {
                    UpdateTasks K; K.tasks = ctx->d_tasks.as<int2>(); K.n_tasks = ctx->n_tasks;
                    if (tiled && ctx->n_items > 0 && !ctx->opt.update_relin) {
                        // back-substitution from the Z records of this trial's stage (no second linearisation)
                        update_z_kernel<<<ctx->updz_grid, ST_THREADS, 0, st>>>(D, ctx->pt_cur, U, K, ctx->d_Z.as<double>(), ctx->d_Lm.as<double>());
                        n_part_u = ctx->updz_grid;
                    } else {
                        update_packed_kernel<<<ctx->upd_grid, ST_THREADS, 0, st>>>(D, ctx->pose_cur, ctx->pt_cur, U, K);
                        n_part_u = ctx->upd_grid;
                    }
                    if (ctx->n_wide > 0) {
                        UpdateOut UW = U; UW.part_chi2 = part_chi_uw; UW.part_scale = part_scale_w; UW.lm_list = ctx->d_widelist.as<int>(); UW.n_list = ctx->n_wide;
                        update_kernel<<<G, BUILD_THREADS, 0, st>>>(D, ctx->pose_cur, ctx->pt_cur, UW);
                        upd_wide = true; ctx->tm.total_launches++;
                    }
                } else {
                    update_kernel<<<G, BUILD_THREADS, 0, st>>>(D, ctx->pose_cur, ctx->pt_cur, U);
                }
            }
            ctx->tm.total_launches += 2;
            ctx->tm.edge_linearisations += 0; ctx->tm.edge_evaluations += n_active;
            // dstat[0] = chi2 at the linearisation point, [1] = trial chi2, [2] = landmark part of scale, [4] = pose part of scale
            {
                TrialSums TS;
                TS.a[0] = part_chi_b; TS.na[0] = n_part_b; TS.b[0] = have_wide_part ? part_chi_w : nullptr; TS.nb[0] = n_part_w;
                TS.a[1] = part_chi_u; TS.na[1] = n_part_u; TS.b[1] = upd_wide ? part_chi_uw : nullptr; TS.nb[1] = G;
                TS.a[2] = part_scale; TS.na[2] = n_part_u; TS.b[2] = upd_wide ? part_scale_w : nullptr; TS.nb[2] = G;
                TS.fail = fail_p; TS.out = dstat;
                finish_trial_kernel<<<3, 256, 0, st>>>(TS);
                ctx->tm.total_launches++;
            }
            rc = all_reduce_sum(ctx, dstat, 3); if (rc) return rc;
            rc = all_reduce_max(ctx, dstat + 5, 1); if (rc) return rc;       // a landmark factor can fail on one rank only
            CK(cudaGetLastError());
            double h[6];
            rc = read_status(ctx, h, 6); if (rc) return rc;
            const int failflag = (int)h[5];
            if (failflag >= 2) {
                // The factorisation clusters waited 2 s for pair_kernel's row counters: the device was starved (another context's
                // burst, MPS, a tool serialising kernels), NOT a numerical failure. Never feed that into rho / lambda: re-run this
                // trial with the solve queued after pair_kernel, and keep that order for the rest of this context's life.
                if (overlap) {
                    ctx->overlap_off = true; ctx->tm.solve_retries++;
                    ctx->tm.edge_linearisations -= n_active; ctx->tm.edge_evaluations -= n_active;
                    if (ctx->opt.debug) fprintf(stderr, "[bagpu] overlapped solve starved (flag %d): trial re-run with the solve after pair_kernel\n", failflag);
                    redo = true;
                    continue;
                }
                return fail(ctx, BAGPU_ERR_CUDA, "reduced-system solve reported a wait time-out (flag %d) without overlap", failflag);
            }
            const bool ok2 = (failflag == 0);
            if (ctx->opt.debug) fprintf(stderr, "[bagpu] trial it=%d q=%d lambda=%.6e chi2 %.9e -> %.9e scale %.3e fail=%d\n", it, qmax, lambda, h[0], h[1], h[2] + h[4], failflag);
            if (first) { currentChi = h[0]; iniChi = currentChi; first = false; }
            double tempChi = ok2 ? h[1] : DBL_MAX;
            rho = currentChi - tempChi;
            double scale = h[2] + h[4];
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - std::pow((2 * rho - 1), 3);
                alpha = std::min(alpha, 2. / 3.);
                const double scaleFactor = std::max(1. / 3., alpha);
                lambda *= scaleFactor;
                ni = 2;
                currentChi = tempChi;
                std::swap(ctx->pose_cur, ctx->pose_trial);          // discardTop(): the trial state becomes current
                std::swap(ctx->pt_cur, ctx->pt_trial);
            } else {
                lambda *= ni;
                ni *= 2;                                            // pop(): keep the current state; edge chi2 stay as evaluated
            }
            qmax++; ctx->tm.lm_trials++;
        } while (redo || (rho < 0 && qmax < 10 && !stop()));
        ctx->tm.lm_iterations++;
        int stt = BAGPU_OK;
        if (qmax == 10 || rho == 0) stt = BAGPU_TERMINATE_TRIALS;
        else {
            if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
            if (nBad >= 3) stt = BAGPU_TERMINATE_NBAD;
        }
        if (res && res->trace && res->n_trace < s->max_trace) {
            bagpu_trace &t = res->trace[res->n_trace++];
            t.round = round; t.iteration = it; t.chi2_before = iniChi; t.chi2_after = currentChi; t.lambda = lambda; t.trials = qmax; t.status = stt;
            t.active_edges = n_active;
            t.linearise_schur_us = 1e3 * (ctx->tm.build_ms - ph0[0]); t.linear_solve_us = 1e3 * (ctx->tm.linsolve_ms - ph0[1]);
            t.update_us = 1e3 * (ctx->tm.update_ms - ph0[2]);
            t.iteration_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - it_t0).count();
        }
        status = stt;
        ok = (stt == BAGPU_OK);
    }
    if (stop() && status == BAGPU_OK) status = BAGPU_STOPPED;
    *status_out = status;
    return BAGPU_OK;
}

}  // namespace

int bagpu_solve_resident(bagpu_ctx *ctx, const bagpu_schedule *s, bagpu_result *r) {
    if (!ctx || !s || !s->rounds || s->n_rounds < 1) return fail(ctx, BAGPU_ERR_ARG, "bad schedule");
    if (!ctx->have_problem) return fail(ctx, BAGPU_ERR_ARG, "no problem uploaded");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->ev_used = 0; ctx->pending.clear();
    ctx->tm.build_ms = ctx->tm.linsolve_ms = ctx->tm.update_ms = 0;
    ctx->tm.build_launches = ctx->tm.update_launches = ctx->tm.linsolve_launches = ctx->tm.total_launches = 0;
    ctx->tm.lm_iterations = ctx->tm.lm_trials = ctx->tm.edge_linearisations = ctx->tm.edge_evaluations = 0;
    ctx->tm.pcg_iterations = 0; ctx->tm.solve_retries = 0; ctx->tm.solver_parts = ctx->parts.on ? ctx->parts.P : (ctx->tw.on ? 2 : 1); { const long long nf = ctx->n_free, bw1 = std::min<long long>(nf, ctx->band_blocks + 1); ctx->tm.schur_blocks = (int)(nf * bw1 - bw1 * (bw1 - 1) / 2); }   // blocks (a, b), a <= b <= a + band_blocks
    CK(cudaEventRecord(ctx->ev_phase[0], st));
    if (r) { r->n_trace = 0; r->status = BAGPU_OK; }
    BaDev D = make_dev(ctx, s->delta_mono, s->delta_stereo);
    const int g = grid_for(ctx->n_obs, 256);
    int status = BAGPU_OK;
    auto stop = [&]() { return s->stop_flag && *s->stop_flag; };
    for (int k = 0; k < s->n_rounds; k++) {
        const bagpu_round &rd = s->rounds[k];
        if (rd.reset_pose) CK(cudaMemcpyAsync(ctx->pose_cur, ctx->d_pose_init.p, sizeof(double) * 7 * (size_t)ctx->n_poses, cudaMemcpyDeviceToDevice, st));
        if (stop()) { status = BAGPU_STOPPED; break; }
        // initializeOptimization(0): active edges = level 0 (sparse_optimizer.cpp:199-267)
        CK(cudaMemsetAsync(ctx->d_count.p, 0, sizeof(unsigned long long), st));
        count_active_kernel<<<g, 256, 0, st>>>(D.o_meta, ctx->n_obs, ctx->d_count.as<unsigned long long>());
        unsigned long long n_active = 0;
        CK(cudaMemcpyAsync(ctx->h_status.p, ctx->d_count.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        memcpy(&n_active, ctx->h_status.p, sizeof(n_active));
        bool any_active = n_active > 0;
        if (ctx->world > 1) {       // every rank must take the same branch
            double *ds = ctx->d_status.as<double>();
            double v = (double)n_active;
            CK(cudaMemcpyAsync(ds + 12, &v, sizeof(double), cudaMemcpyHostToDevice, st));
            int rc = all_reduce_sum(ctx, ds + 12, 1); if (rc) return rc;
            CK(cudaMemcpyAsync(&v, ds + 12, sizeof(double), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            any_active = v > 0.5;
        }
        if (any_active) {
            int rc = chain_eligible(ctx, s) ? optimize_chained(ctx, s, k, rd.iterations, (int64_t)n_active, r, &status)
                                            : optimize(ctx, s, k, rd.iterations, (int64_t)n_active, r, &status);
            if (rc) return rc;
        }
        if (rd.gate_after != BAGPU_GATE_NONE || rd.drop_kernel_after) {
            if (rd.gate_after == BAGPU_GATE_LBA && stop()) { status = BAGPU_STOPPED; break; }
            gate_kernel<<<g, 256, 0, st>>>(D, ctx->pose_cur, ctx->pt_cur, ctx->d_chi2.as<double>(), nullptr, rd.gate_after,
                                           rd.gate_mono, rd.gate_stereo, rd.drop_kernel_after);
            ctx->tm.total_launches++;
        }
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev_phase[1], st));
    CK(cudaStreamSynchronize(st));
    resolve_events(ctx);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_phase[0], ctx->ev_phase[1]);
    ctx->tm.solve_ms = ms;
    if (r) {
        r->status = status;
        int rc = bagpu_download(ctx, r);
        if (rc) return rc;
    }
    return status;
}

int bagpu_reset_resident(bagpu_ctx *ctx) {
    if (!ctx) return BAGPU_ERR_ARG;
    if (!ctx->have_problem) return fail(ctx, BAGPU_ERR_ARG, "no problem uploaded");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->pose_cur = ctx->d_pose_a.as<double>(); ctx->pose_trial = ctx->d_pose_b.as<double>();
    ctx->pt_cur = ctx->d_pt_a.as<double>(); ctx->pt_trial = ctx->d_pt_b.as<double>();
    CK(cudaMemcpyAsync(ctx->pose_cur, ctx->d_pose_init.p, sizeof(double) * 7 * (size_t)ctx->n_poses, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(ctx->pose_trial, ctx->d_pose_init.p, sizeof(double) * 7 * (size_t)ctx->n_poses, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(ctx->pt_cur, ctx->d_pt_init.p, sizeof(double) * 3 * (size_t)ctx->n_points, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(ctx->pt_trial, ctx->d_pt_init.p, sizeof(double) * 3 * (size_t)ctx->n_points, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_o_meta.p, ctx->d_meta_init.p, 4 * (size_t)ctx->n_obs, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemsetAsync(ctx->d_chi2.p, 0, 8 * (size_t)ctx->n_obs, st));
    CK(cudaStreamSynchronize(st));
    return BAGPU_OK;
}

int bagpu_update_estimates(bagpu_ctx *ctx, const double *pose_qt, const double *points) {
    if (!ctx) return BAGPU_ERR_ARG;
    if (!ctx->have_problem) return fail(ctx, BAGPU_ERR_ARG, "no problem uploaded");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t pb = sizeof(double) * 7 * (size_t)ctx->n_poses, qb = sizeof(double) * 3 * (size_t)ctx->n_points;
    if (pose_qt) {
        CK(ctx->h_stage.ensure(pb));
        double *hp = ctx->h_stage.as<double>();
        for (int i = 0; i < ctx->n_poses; i++) normalize_qt(pose_qt + 7 * (size_t)i, hp + 7 * (size_t)i);
        CK(cudaMemcpyAsync(ctx->d_pose_init.p, hp, pb, cudaMemcpyHostToDevice, st));
    }
    if (points) CK(cudaMemcpyAsync(ctx->d_pt_init.p, points, qb, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                           // the caller's arrays (and h_stage) may be reused on return
    return bagpu_reset_resident(ctx);                        // both state buffers, edge levels and kernels from the (new) uploaded state
}

int bagpu_download(bagpu_ctx *ctx, bagpu_result *r) {
    if (!ctx || !r) return BAGPU_ERR_ARG;
    if (!ctx->have_problem) return fail(ctx, BAGPU_ERR_ARG, "no problem uploaded");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->ev_phase[2], st));
    const size_t ne = (size_t)ctx->n_obs;
    const int g = grid_for(ctx->n_obs, 256);
    BaDev D = make_dev(ctx, 0, 0);
    int64_t d2h = 0;
    // isDepthPositive on the final estimates, in sorted order
    gate_kernel<<<g, 256, 0, st>>>(D, ctx->pose_cur, ctx->pt_cur, ctx->d_chi2.as<double>(), ctx->d_depth.as<uint8_t>(), BAGPU_GATE_NONE, 0, 0, 0);
    const double *chi_src = ctx->d_chi2.as<double>();
    const uint8_t *dp_src = ctx->d_depth.as<uint8_t>();
    CK(ctx->d_out_u8b.ensure(ne));
    uint8_t *lvl_src = ctx->d_out_u8b.as<uint8_t>();
    if (!ctx->identity_perm) {
        CK(ctx->d_out_chi2.ensure(8 * ne)); CK(ctx->d_out_u8a.ensure(ne));
        scatter_perm_kernel<double><<<g, 256, 0, st>>>(ctx->n_obs, ctx->d_perm.as<int>(), ctx->d_chi2.as<double>(), ctx->d_out_chi2.as<double>());
        scatter_perm_kernel<uint8_t><<<g, 256, 0, st>>>(ctx->n_obs, ctx->d_perm.as<int>(), ctx->d_depth.as<uint8_t>(), ctx->d_out_u8a.as<uint8_t>());
        level_from_meta_kernel<<<g, 256, 0, st>>>(ctx->n_obs, ctx->d_perm.as<int>(), D.o_meta, lvl_src);
        chi_src = ctx->d_out_chi2.as<double>(); dp_src = ctx->d_out_u8a.as<uint8_t>();
    } else {
        level_from_meta_kernel<<<g, 256, 0, st>>>(ctx->n_obs, nullptr, D.o_meta, lvl_src);
    }
    if (r->pose_qt) { CK(cudaMemcpyAsync(r->pose_qt, ctx->pose_cur, sizeof(double) * 7 * (size_t)ctx->n_poses, cudaMemcpyDeviceToHost, st)); d2h += 56 * (int64_t)ctx->n_poses; }
    if (r->points) { CK(cudaMemcpyAsync(r->points, ctx->pt_cur, sizeof(double) * 3 * (size_t)ctx->n_points, cudaMemcpyDeviceToHost, st)); d2h += 24 * (int64_t)ctx->n_points; }
    if (r->edge_chi2) { CK(cudaMemcpyAsync(r->edge_chi2, chi_src, 8 * ne, cudaMemcpyDeviceToHost, st)); d2h += 8 * (int64_t)ne; }
    if (r->edge_depth_pos) { CK(cudaMemcpyAsync(r->edge_depth_pos, dp_src, ne, cudaMemcpyDeviceToHost, st)); d2h += (int64_t)ne; }
    if (r->edge_level) { CK(cudaMemcpyAsync(r->edge_level, lvl_src, ne, cudaMemcpyDeviceToHost, st)); d2h += (int64_t)ne; }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev_phase[3], st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_phase[2], ctx->ev_phase[3]);
    ctx->tm.d2h_ms = ms; ctx->tm.d2h_bytes = d2h;
    return BAGPU_OK;
}

int bagpu_solve_ba(bagpu_ctx *ctx, const bagpu_problem *p, const bagpu_schedule *s, bagpu_result *r) {
    int rc = bagpu_upload(ctx, p);
    if (rc) return rc;
    return bagpu_solve_resident(ctx, s, r);
}

// ------------------------------------------------------------------------------- pose batch
int bagpu_pose_upload(bagpu_ctx *ctx, const bagpu_pose_batch *b) {
    if (!ctx || !b) return BAGPU_ERR_ARG;
    if (b->n_frames <= 0 || !b->pose_qt || !b->frame_ptr || !b->cameras || b->n_cameras <= 0 || b->n_cameras > 255 || b->n_rigs > 254)
        return fail(ctx, BAGPU_ERR_ARG, "bad pose batch");
    const int F = b->n_frames; const int64_t Ne = b->n_obs;
    if (b->frame_ptr[0] != 0 || b->frame_ptr[F] != Ne) return fail(ctx, BAGPU_ERR_ARG, "frame_ptr does not cover the edges");
    if (Ne > 0 && (!b->xw || !b->obs_cam || !b->obs_rig || !b->obs_kind || !b->obs_u || !b->obs_v || !b->obs_inv_sigma2)) return fail(ctx, BAGPU_ERR_ARG, "NULL array in pose batch");
    bool any_stereo = false;
    for (int64_t e = 0; e < Ne; e++) {
        if (b->obs_cam[e] < 0 || b->obs_cam[e] >= b->n_cameras || b->obs_kind[e] > 2) return fail(ctx, BAGPU_ERR_ARG, "edge %lld: bad camera/kind", (long long)e);
        if (b->obs_kind[e] == BAGPU_EDGE_BODY && (b->obs_rig[e] < 0 || b->obs_rig[e] >= b->n_rigs)) return fail(ctx, BAGPU_ERR_ARG, "edge %lld: bad rig", (long long)e);
        if (b->obs_kind[e] == BAGPU_EDGE_STEREO) any_stereo = true;
    }
    if (any_stereo && !b->obs_ur) return fail(ctx, BAGPU_ERR_ARG, "stereo edges need obs_ur");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->ev_phase[0], st));
    ctx->have_pose = false;
    const size_t ne = (size_t)std::max<int64_t>(1, Ne);
    CK(ctx->h_stage.ensure(sizeof(double) * 7 * ((size_t)F + (size_t)std::max(1, b->n_rigs))));
    double *hp = ctx->h_stage.as<double>();
    for (int i = 0; i < F; i++) normalize_qt(b->pose_qt + 7 * (size_t)i, hp + 7 * (size_t)i);
    double *hr = hp + 7 * (size_t)F;
    for (int i = 0; i < b->n_rigs; i++) normalize_qt(b->rigs[i].qt, hr + 7 * (size_t)i);
    CK(ctx->p_pose0.ensure(56 * (size_t)F)); CK(ctx->p_ptr.ensure(8 * ((size_t)F + 1))); CK(ctx->p_cams.ensure(sizeof(bagpu_camera) * (size_t)b->n_cameras));
    CK(ctx->p_rigs.ensure(56 * (size_t)std::max(1, b->n_rigs))); CK(ctx->p_xw.ensure(24 * ne)); CK(ctx->p_meta.ensure(4 * ne));
    CK(ctx->p_u.ensure(8 * ne)); CK(ctx->p_v.ensure(8 * ne)); CK(ctx->p_ur.ensure(8 * ne)); CK(ctx->p_w.ensure(8 * ne));
    CK(ctx->p_chi2.ensure(8 * ne)); CK(ctx->p_out.ensure(ne)); CK(ctx->p_pose_out.ensure(56 * (size_t)F)); CK(ctx->p_ninl.ensure(4 * (size_t)F));
    CK(ctx->p_fchi.ensure(8 * (size_t)F));
    CK(ctx->d_raw8a.ensure(ne)); CK(ctx->d_raw8b.ensure(ne)); CK(ctx->d_raw16a.ensure(2 * ne)); CK(ctx->d_raw16b.ensure(2 * ne));
    CK(cudaMemcpyAsync(ctx->p_pose0.p, hp, 56 * (size_t)F, cudaMemcpyHostToDevice, st));
    if (b->n_rigs) CK(cudaMemcpyAsync(ctx->p_rigs.p, hr, 56 * (size_t)b->n_rigs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->p_ptr.p, b->frame_ptr, 8 * ((size_t)F + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->p_cams.p, b->cameras, sizeof(bagpu_camera) * (size_t)b->n_cameras, cudaMemcpyHostToDevice, st));
    int64_t h2d = 56 * (int64_t)F + 8 * ((int64_t)F + 1);
    if (Ne > 0) {
        CK(cudaMemcpyAsync(ctx->p_xw.p, b->xw, 24 * (size_t)Ne, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->p_u.p, b->obs_u, 8 * (size_t)Ne, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->p_v.p, b->obs_v, 8 * (size_t)Ne, cudaMemcpyHostToDevice, st));
        if (b->obs_ur) CK(cudaMemcpyAsync(ctx->p_ur.p, b->obs_ur, 8 * (size_t)Ne, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->p_w.p, b->obs_inv_sigma2, 8 * (size_t)Ne, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->d_raw8a.p, b->obs_kind, (size_t)Ne, cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(ctx->d_raw8b.p, 0, (size_t)Ne, st));
        CK(cudaMemcpyAsync(ctx->d_raw16a.p, b->obs_cam, 2 * (size_t)Ne, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->d_raw16b.p, b->obs_rig, 2 * (size_t)Ne, cudaMemcpyHostToDevice, st));
        compose_meta_kernel<<<grid_for(Ne, 256), 256, 0, st>>>(Ne, ctx->d_raw8a.as<uint8_t>(), ctx->d_raw16a.as<int16_t>(), ctx->d_raw16b.as<int16_t>(),
                                                             ctx->d_raw8b.as<uint8_t>(), ctx->p_meta.as<uint32_t>());
        h2d += (24 + 8 * (b->obs_ur ? 4 : 3) + 5) * Ne;
    }
    CK(cudaGetLastError());
    ctx->pb_frames = F; ctx->pb_obs = Ne;
    ctx->pb_delta_mono = b->delta_mono; ctx->pb_delta_stereo = b->delta_stereo; ctx->pb_gate_mono = b->gate_mono; ctx->pb_gate_stereo = b->gate_stereo;
    CK(cudaEventRecord(ctx->ev_phase[1], st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev_phase[0], ctx->ev_phase[1]);
    ctx->tm.h2d_ms = ms; ctx->tm.h2d_bytes = h2d;
    ctx->have_pose = true;
    return BAGPU_OK;
}

int bagpu_pose_solve_resident(bagpu_ctx *ctx, bagpu_pose_result *r) {
    if (!ctx) return BAGPU_ERR_ARG;
    if (!ctx->have_pose) return fail(ctx, BAGPU_ERR_ARG, "no pose batch uploaded");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    PoseDev D;
    D.n_frames = ctx->pb_frames; D.pose0 = ctx->p_pose0.as<double>(); D.frame_ptr = ctx->p_ptr.as<int64_t>();
    D.cams = ctx->p_cams.as<bagpu_camera>(); D.rigs = ctx->p_rigs.as<double>(); D.xw = ctx->p_xw.as<double>();
    D.o_meta = ctx->p_meta.as<uint32_t>(); D.o_u = ctx->p_u.as<double>(); D.o_v = ctx->p_v.as<double>(); D.o_ur = ctx->p_ur.as<double>();
    D.o_w = ctx->p_w.as<double>(); D.delta_mono = ctx->pb_delta_mono; D.delta_stereo = ctx->pb_delta_stereo;
    D.gate_mono = ctx->pb_gate_mono; D.gate_stereo = ctx->pb_gate_stereo; D.chi2 = ctx->p_chi2.as<double>();
    D.outlier = ctx->p_out.as<uint8_t>(); D.pose_out = ctx->p_pose_out.as<double>(); D.n_inliers = ctx->p_ninl.as<int>();
    D.final_chi2 = ctx->p_fchi.as<double>();
    CK(cudaEventRecord(ctx->ev_phase[0], st));
    pose_opt_kernel<<<ctx->pb_frames, PO_THREADS, 0, st>>>(D);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev_phase[1], st));
    int64_t d2h = 0;
    if (r) {
        CK(cudaEventRecord(ctx->ev_phase[2], st));
        const size_t F = (size_t)ctx->pb_frames;
        if (r->pose_qt) { CK(cudaMemcpyAsync(r->pose_qt, D.pose_out, 56 * F, cudaMemcpyDeviceToHost, st)); d2h += 56 * (int64_t)F; }
        if (r->outlier && ctx->pb_obs) { CK(cudaMemcpyAsync(r->outlier, D.outlier, (size_t)ctx->pb_obs, cudaMemcpyDeviceToHost, st)); d2h += ctx->pb_obs; }
        if (r->n_inliers) { CK(cudaMemcpyAsync(r->n_inliers, D.n_inliers, 4 * F, cudaMemcpyDeviceToHost, st)); d2h += 4 * (int64_t)F; }
        if (r->final_chi2) { CK(cudaMemcpyAsync(r->final_chi2, D.final_chi2, 8 * F, cudaMemcpyDeviceToHost, st)); d2h += 8 * (int64_t)F; }
        CK(cudaEventRecord(ctx->ev_phase[3], st));
    }
    CK(cudaStreamSynchronize(st));
    float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev_phase[0], ctx->ev_phase[1]);
    ctx->tm.solve_ms = ms; ctx->tm.total_launches = 1;
    if (r) { cudaEventElapsedTime(&ms, ctx->ev_phase[2], ctx->ev_phase[3]); ctx->tm.d2h_ms = ms; ctx->tm.d2h_bytes = d2h; }
    return BAGPU_OK;
}

int bagpu_pose_opt_batch(bagpu_ctx *ctx, const bagpu_pose_batch *b, bagpu_pose_result *r) {
    int rc = bagpu_pose_upload(ctx, b);
    if (rc) return rc;
    return bagpu_pose_solve_resident(ctx, r);
}

int bagpu_get_timing(const bagpu_ctx *ctx, bagpu_timing *out) {
    if (!ctx || !out) return BAGPU_ERR_ARG;
    *out = ctx->tm;
    return BAGPU_OK;
}

// Unit-test hook for the reduced-system solver: dense symmetric A (row-major), envelope col_end, (A + lambda I) x = b.
int bagpu_test_solve(bagpu_ctx *ctx, int n, const int *col_end, const double *A, const double *b, double lambda, double *x, int *fail_out) {
    return bagpu_test_solve_parts(ctx, n, col_end, A, b, lambda, 1, x, fail_out);
}
// parts >= 3: the partitioned solver with that many fronts (BAGPU_ERR_ARG when the system is too short for them); otherwise one front.
int bagpu_test_solve_parts(bagpu_ctx *ctx, int n, const int *col_end, const double *A, const double *b, double lambda, int parts, double *x, int *fail_out) {
    if (!ctx || n <= 0 || !col_end || !A || !b || !x) return BAGPU_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int band = 1, max_below = 0;
    for (int j = 0; j < n; j++) {
        if (col_end[j] < j || col_end[j] >= n || (j && col_end[j] < col_end[j - 1])) return fail(ctx, BAGPU_ERR_ARG, "col_end must be monotone and >= j");
        band = std::max(band, col_end[j] - j + 1);
    }
    for (int p0 = 0; p0 < n; p0 += CH_NB) {
        const int nb = std::min(CH_NB, n - p0);
        const int rend = std::min(n - 1, col_end[p0 + nb - 1]);
        max_below = std::max(max_below, rend - (p0 + nb) + 1);
        band = std::max(band, rend - p0 + 1);
    }
    const int ld = std::max(1, std::min(band - 1, n));
    const size_t s_elems = (size_t)n * (ld + 1) + 8;
    std::vector<double> hS(s_elems, 0.0);
    for (int R = 0; R < n; R++) for (int C = R; C <= col_end[R]; C++) hS[(size_t)R * ld + C] = A[(size_t)R * n + C];
    DevBuf dS, db, dz, dx, dy, dc, df, dd;
    CK(dd.ensure(8 * (size_t)n));
    CK(dS.ensure(8 * s_elems)); CK(db.ensure(8 * (size_t)n)); CK(dz.ensure(8 * (size_t)n)); CK(dx.ensure(8 * (size_t)n));
    CK(dy.ensure(8 * (size_t)n)); CK(dc.ensure(4 * (size_t)n)); CK(df.ensure(16 + 256));
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(dS.p, hS.data(), 8 * s_elems, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(db.p, b, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(dz.p, 0, 8 * (size_t)n, st));
    CK(cudaMemsetAsync(dy.p, 0, 8 * (size_t)n, st));
    CK(cudaMemsetAsync(df.p, 0, 16, st));
    CK(cudaMemcpyAsync(dc.p, col_end, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    int tgrid = 1, tmaxr = 0;
    { int rc2 = chol_plan_grid(ctx, n, max_below, &tgrid, &tmaxr); if (rc2) return rc2; }
    if (parts >= 3) {
        PartPlan pp;
        std::vector<int> ce(col_end, col_end + n);
        int rc2 = parts_plan(ctx, pp, n, ld, ce, parts, false, st);
        if (rc2 == BAGPU_OK && !pp.on) rc2 = fail(ctx, BAGPU_ERR_ARG, "system of %d rows is too short (or its envelope too wide) for %d partitions", n, parts);
        if (rc2 == BAGPU_OK) rc2 = parts_bind(ctx, pp, dS.as<double>(), db.as<double>(), dz.as<double>(), dy.as<double>(), dx.as<double>(), dd.as<double>(), df.as<int>(),
                                             nullptr, nullptr, 0, nullptr, st);
        if (rc2 == BAGPU_OK) rc2 = parts_enqueue_factor(ctx, pp, lambda, nullptr, st);
        if (rc2 == BAGPU_OK) rc2 = parts_enqueue_rest(ctx, pp, dS.as<double>(), db.as<double>(), dz.as<double>(), dy.as<double>(), dx.as<double>(), lambda, df.as<int>(), st);
        int hf = 0;
        if (rc2 == BAGPU_OK) {
            if (cudaMemcpyAsync(x, dx.p, 8 * (size_t)n, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaMemcpyAsync(&hf, df.p, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaStreamSynchronize(st) != cudaSuccess) rc2 = fail(ctx, BAGPU_ERR_CUDA, "partitioned test solve: %s", cudaGetErrorString(cudaGetLastError()));
        } else cudaStreamSynchronize(st);
        if (fail_out) *fail_out = hf;
        pp.release();
        dS.release(); db.release(); dz.release(); dx.release(); dy.release(); dc.release(); df.release(); dd.release();
        return rc2;
    }
    CholArgs ca; ca.S = dS.as<double>(); ca.n = n; ca.ld = ld; ca.lambda = lambda; ca.bp = db.as<double>(); ca.bs = dz.as<double>();
    ca.col_end = dc.as<int>(); ca.y = dy.as<double>(); ca.dinv = dd.as<double>(); ca.prof = ctx->opt.debug ? (long long *)((char *)df.p + 16) : nullptr; ca.x = dx.as<double>(); ca.fail = df.as<int>();
    int rc = launch_chol(ctx, ca, tgrid, tmaxr);
    if (rc) return rc;
    int hf = 0;
    CK(cudaMemcpyAsync(x, dx.p, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&hf, df.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (fail_out) *fail_out = hf;
    if (ctx->opt.debug) {
        long long hp[24];
        cudaMemcpy(hp, (char *)df.p + 16, 192, cudaMemcpyDeviceToHost);
        if (tmaxr > 0 && hf) fprintf(stderr, "[bagpu] band chol FAILED at panel %lld (n=%d maxr=%d grid=%d)\n", hp[23] - 1, n, tmaxr, tgrid);
        else if (tmaxr > 0) {
            fprintf(stderr, "[bagpu] band chol n=%d ld=%d grid=%d maxr=%d (CTA 0 cycles)\n  owner: trsm %lld wb+y %lld barrier %lld take %lld endsync %lld\n"
                    "  next : flagwait %lld diagupd %lld ldbuild %lld potrf %lld clwait %lld endsync %lld\n  other: barrier %lld stage %lld update %lld endsync %lld\n"
                    "  final sync %lld | backward: topsync %lld prefetch+wait %lld dots %lld chain %lld | owner top %lld trsm-loop %lld | potrf chains %lld\n",
                    n, ld, tgrid, tmaxr, hp[0], hp[1], hp[2], hp[3], hp[13], hp[4], hp[5], hp[6], hp[7], hp[8], hp[9], hp[10], hp[11], hp[12], hp[14],
                    hp[15], hp[16], hp[17], hp[18], hp[19], hp[20], hp[21], hp[22]);
        } else {
            fprintf(stderr, "[bagpu] chol cycles n=%d ld=%d grid=%d: potrf %lld trsm %lld sync1 %lld wb+y %lld update %lld sync2 %lld backward %lld\n",
                    n, ld, tgrid, hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], hp[6]);
        }
    }
    dS.release(); db.release(); dz.release(); dx.release(); dy.release(); dc.release(); df.release(); dd.release();
    return BAGPU_OK;
}

// Measured FP64 peaks of this device (TFLOP/s, 2 flops per multiply-add): plain DFMA and the m8n8k4 FP64 MMA, each the best
// of three launches that fill every SM with 1024 threads. bench.py reports the Schur pass against them.
int bagpu_test_fp64_peak(bagpu_ctx *ctx, double *dfma_tflops, double *dmma_tflops) {
    if (!ctx || !dfma_tflops || !dmma_tflops) return BAGPU_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    DevBuf o;
    const int blocks = 2 * ctx->n_sm, threads = 1024, iters = 4096;
    CK(o.ensure(sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0 = get_event(ctx), e1 = get_event(ctx);
    double best[2] = {0, 0};
    for (int which = 0; which < 2; which++)
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(e0, ctx->stream));
            if (which == 0) fp64_dfma_probe_kernel<<<blocks, threads, 0, ctx->stream>>>(o.as<double>(), iters, 0.5);
            else fp64_dmma_probe_kernel<<<blocks, threads, 0, ctx->stream>>>(o.as<double>(), iters, 0.5);
            CK(cudaEventRecord(e1, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            float ms = 0.f; CK(cudaEventElapsedTime(&ms, e0, e1));
            // DFMA: 8 fma per thread per iteration; MMA m8n8k4: 8 x 8 x 4 = 256 multiply-adds per warp instruction, 8 per iteration
            const double mac = which == 0 ? (double)blocks * threads * iters * 8.0 : (double)blocks * (threads / 32) * iters * 8.0 * 256.0;
            if (rep > 0) best[which] = std::max(best[which], 2.0 * mac / (ms * 1e-3) / 1e12);
        }
    ctx->ev_used = 0;
    *dfma_tflops = best[0]; *dmma_tflops = best[1];
    o.release();
    return BAGPU_OK;
}

int bagpu_test_atan2f(bagpu_ctx *ctx, const float *y, const float *x, float *out, int64_t n) {
    if (!ctx || n <= 0) return BAGPU_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    DevBuf a, b, c;
    CK(a.ensure(4 * (size_t)n)); CK(b.ensure(4 * (size_t)n)); CK(c.ensure(4 * (size_t)n));
    CK(cudaMemcpyAsync(a.p, y, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b.p, x, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    atan2f_test_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(n, a.as<float>(), b.as<float>(), c.as<float>());
    CK(cudaMemcpyAsync(out, c.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    a.release(); b.release(); c.release();
    return BAGPU_OK;
}

}  // extern "C"
