// BA hot-path kernels (sm_100a). Landmark-major: observations are sorted by (point, pose) on upload and
// one warp owns one landmark at a time, lane = observation.
//
//  build_kernel   K2+K3 of SURVEY.md section 2: linearise every active edge, reduce the landmark's 3x3 Hll / bl in
//                 the warp, invert (Hll + lambda I), and scatter Hpp, b and the Schur contribution
//                 -W Dinv W^T straight into the reduced camera system. W (Hpl) is never written to memory.
//                 Replaces BlockSolver::buildSystem + the first half of BlockSolver::solve
//                 (Thirdparty/g2o/g2o/core/block_solver.hpp:501-560, 373-439).
//  update_kernel  K5+K6+K1: re-linearise at the old state, back-substitute the landmark step
//                 (block_solver.hpp:459-483), apply it (types_sba.h:50-54), evaluate every edge at the trial
//                 state (SparseOptimizer::computeActiveErrors + activeRobustChi2, sparse_optimizer.cpp:61-114)
//                 and accumulate the landmark part of computeScale (optimization_algorithm_levenberg.cpp:187-194).
#pragma once
#include "ba_math.cuh"

struct BaDev {
    // problem (sorted by landmark)
    int n_points, n_poses, n_free;
    int64_t n_obs;
    const int *lm_ptr;            // [n_points+1]
    const int *o_pose;            // [n_obs]
    const int *o_point;           // [n_obs]
    uint32_t *o_meta;             // [n_obs] kind | cam<<2 | rig<<10 | robust<<18 | level1<<19
    const double *o_u, *o_v, *o_ur, *o_w;
    const bagpu_camera *cams;
    const double *rigs;           // [n_rigs][7] normalised
    const int *pose_hidx;         // [n_poses] free index or -1
    double delta_mono, delta_stereo;
};

// Levenberg-Marquardt state on the device (chained mode: small maps run a whole optimize() without a host round trip per trial).
// Every kernel of a trial reads lambda and which of the two state buffers is current from here; lm_decide_kernel takes the accept /
// reject decision of OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:99-166) at the end of the trial.
struct LmTraceDev { int round, iteration, trials, status; double chi2_before, chi2_after, lambda; };
struct LmDev {
    double lambda, ni, currentChi, iniChi, rho;
    int flip;                     // 0: buffers (a, b) = (current, trial); 1: swapped (an accepted trial flips)
    int qmax, it, nBad, first, done, status;
    int iterations, round, n_trace, max_trace;
    int trials, starved;
};
BA_DEV bool lm_done(const LmDev *lm) { return lm != nullptr && lm->done != 0; }
template <typename T> BA_DEV T *lm_cur(const LmDev *lm, T *a, T *b) { return (lm != nullptr && lm->flip) ? b : a; }
template <typename T> BA_DEV T *lm_trial(const LmDev *lm, T *a, T *b) { return (lm != nullptr && lm->flip) ? a : b; }

#define BUILD_THREADS 256
#define BUILD_WARPS (BUILD_THREADS / 32)

struct LaneEdge {
    EdgeLin L;
    double g0, g1, g2;            // -rho1 * omega * r
    double wgt;                   // rho1 * omega
    double rho0;
    int hidx;                     // free pose index, -1 if fixed / inactive lane
    bool valid;
};

// Linearise the observation `e` (or produce an inert lane).
BA_DEV void lane_linearize(const BaDev &D, const double *__restrict__ pose, double X, double Y, double Z, int e, bool in_range,
                           LaneEdge &E) {
    E.valid = false; E.hidx = -1; E.rho0 = 0.0; E.wgt = 0.0; E.g0 = E.g1 = E.g2 = 0.0;
    if (!in_range) return;
    const uint32_t m = D.o_meta[e];
    if (m & META_LEVEL1) return;
    const int ip = __ldg(D.o_pose + e);
    const int kind = META_KIND(m);
    const Cam c = load_cam(D.cams + META_CAM(m));
    const Pose T = load_pose(pose + 7 * (size_t)ip);
    Pose Trl;
    if (kind == BAGPU_EDGE_BODY) Trl = load_pose(D.rigs + 7 * META_RIG(m));
    const double ou = __ldg(D.o_u + e), ov = __ldg(D.o_v + e);
    const double our = (kind == BAGPU_EDGE_STEREO) ? __ldg(D.o_ur + e) : 0.0;
    const double om = __ldg(D.o_w + e);
    double r0, r1, r2;
    edge_residual(kind, c, T, &Trl, X, Y, Z, ou, ov, our, false, r0, r1, r2);
    const double chi2 = r0 * (om * r0) + r1 * (om * r1) + r2 * (om * r2);
    double rho0 = chi2, rho1 = 1.0;
    if (m & META_ROBUST) huber(chi2, kind == BAGPU_EDGE_STEREO ? D.delta_stereo : D.delta_mono, rho0, rho1);
    edge_linearize(kind, c, T, &Trl, X, Y, Z, E.L);
    E.rho0 = rho0;
    E.wgt = (m & META_ROBUST) ? rho1 * om : om;
    const double s = (m & META_ROBUST) ? rho1 : 1.0;
    E.g0 = -(om * r0) * s; E.g1 = -(om * r1) * s; E.g2 = -(om * r2) * s;
    E.hidx = __ldg(D.pose_hidx + ip);
    E.valid = true;
}

struct BuildOut {
    double lambda;
    int mode;                     // 0: diagonal pass for computeLambdaInit; 1: full system
    double *S; int ld;            // dense reduced camera system; element (R,C), R<=C at S[R*ld + C]
    double *bp;                   // [6 n_free]  pose part of b (kept for computeScale)
    double *bs;                   // [6 n_free]  minus the Schur coefficients  (bschur = bp + bs)
    double *hpp_diag;             // [6 n_free]  mode 0 only
    double *part_chi2;            // [gridDim.x]
    double *part_maxdiag;         // [gridDim.x] mode 0 only (Hll part)
    const int *lm_list; int n_list;   // optional: build_kernel then handles only these landmarks (the "wide" ones)
};

__global__ void __launch_bounds__(BUILD_THREADS) build_kernel(BaDev D, const double *__restrict__ pose,
                                                             const double *__restrict__ pt, BuildOut O) {
    __shared__ double s_chi[BUILD_WARPS], s_max[BUILD_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * BUILD_WARPS + warp, nw = gridDim.x * BUILD_WARPS;
    double chi_acc = 0.0, max_acc = 0.0;

    const int n_lm = O.lm_list ? O.n_list : D.n_points;
    for (int jj = gw; jj < n_lm; jj += nw) {
        const int j = O.lm_list ? __ldg(O.lm_list + jj) : jj;
        const int e0 = __ldg(D.lm_ptr + j), k = __ldg(D.lm_ptr + j + 1) - e0;
        if (k == 0) continue;
        const double X = __ldg(pt + 3 * (size_t)j), Y = __ldg(pt + 3 * (size_t)j + 1), Z = __ldg(pt + 3 * (size_t)j + 2);
        const int nchunk = (k + 31) >> 5;
        // ---- pass 1: Hll (6), bl (3), sum rho0
        double h[6] = {0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0}, chi = 0.0;
        LaneEdge E;
        for (int c = 0; c < nchunk; c++) {
            const int e = e0 + 32 * c + lane;
            lane_linearize(D, pose, X, Y, Z, e, 32 * c + lane < k, E);
            if (E.valid) {
                const double *A = E.L.A;
                h[0] += E.wgt * (A[0] * A[0] + A[3] * A[3] + A[6] * A[6]);
                h[1] += E.wgt * (A[0] * A[1] + A[3] * A[4] + A[6] * A[7]);
                h[2] += E.wgt * (A[0] * A[2] + A[3] * A[5] + A[6] * A[8]);
                h[3] += E.wgt * (A[1] * A[1] + A[4] * A[4] + A[7] * A[7]);
                h[4] += E.wgt * (A[1] * A[2] + A[4] * A[5] + A[7] * A[8]);
                h[5] += E.wgt * (A[2] * A[2] + A[5] * A[5] + A[8] * A[8]);
                bl[0] += A[0] * E.g0 + A[3] * E.g1 + A[6] * E.g2;
                bl[1] += A[1] * E.g0 + A[4] * E.g1 + A[7] * E.g2;
                bl[2] += A[2] * E.g0 + A[5] * E.g1 + A[8] * E.g2;
                chi += E.rho0;
                if (O.mode == 0 && E.hidx >= 0) {
                    const double *B = E.L.B;
#pragma unroll
                    for (int a = 0; a < 6; a++)
                        atomicAdd(O.hpp_diag + 6 * E.hidx + a, E.wgt * (B[a] * B[a] + B[6 + a] * B[6 + a] + B[12 + a] * B[12 + a]));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) h[i] = warp_allsum(h[i]);
#pragma unroll
        for (int i = 0; i < 3; i++) bl[i] = warp_allsum(bl[i]);
        chi = warp_allsum(chi);
        chi_acc += chi;
        if (O.mode == 0) {
            max_acc = fmax(max_acc, fmax(fabs(h[0]), fmax(fabs(h[3]), fabs(h[5]))));
            continue;
        }
        h[0] += O.lambda; h[3] += O.lambda; h[5] += O.lambda;
        double di[6];
        sym3_inverse(h, di);
        const double db0 = di[0] * bl[0] + di[1] * bl[1] + di[2] * bl[2];
        const double db1 = di[1] * bl[0] + di[3] * bl[1] + di[4] * bl[2];
        const double db2 = di[2] * bl[0] + di[4] * bl[1] + di[5] * bl[2];

        // ---- pass 2: per-observation W = wgt B^T A (6x3), diagonal terms, then all pairs of the landmark
        for (int ca = 0; ca < nchunk; ca++) {
            if (nchunk > 1) lane_linearize(D, pose, X, Y, Z, e0 + 32 * ca + lane, 32 * ca + lane < k, E);
            const int ka = min(32, k - 32 * ca);
            const bool va = E.valid && E.hidx >= 0;
            const int ha = E.hidx;
            double Wa[18], Ya[18];
            {
                const double *A = E.L.A, *B = E.L.B;
#pragma unroll
                for (int a = 0; a < 6; a++)
#pragma unroll
                    for (int c = 0; c < 3; c++) Wa[3 * a + c] = E.wgt * (B[a] * A[c] + B[6 + a] * A[3 + c] + B[12 + a] * A[6 + c]);
#pragma unroll
                for (int a = 0; a < 6; a++) {
                    Ya[3 * a + 0] = Wa[3 * a] * di[0] + Wa[3 * a + 1] * di[1] + Wa[3 * a + 2] * di[2];
                    Ya[3 * a + 1] = Wa[3 * a] * di[1] + Wa[3 * a + 1] * di[3] + Wa[3 * a + 2] * di[4];
                    Ya[3 * a + 2] = Wa[3 * a] * di[2] + Wa[3 * a + 1] * di[4] + Wa[3 * a + 2] * di[5];
                }
                if (va) {
                    double *Sd = O.S + (size_t)(6 * ha) * O.ld + 6 * ha;
#pragma unroll
                    for (int a = 0; a < 6; a++) {
#pragma unroll
                        for (int c = a; c < 6; c++)
                            atomicAdd(Sd + (size_t)a * O.ld + c, E.wgt * (B[a] * B[c] + B[6 + a] * B[6 + c] + B[12 + a] * B[12 + c]));
                        atomicAdd(O.bp + 6 * ha + a, B[a] * E.g0 + B[6 + a] * E.g1 + B[12 + a] * E.g2);
                        atomicAdd(O.bs + 6 * ha + a, -(Wa[3 * a] * db0 + Wa[3 * a + 1] * db1 + Wa[3 * a + 2] * db2));
                    }
                }
            }
            for (int cb = ca; cb < nchunk; cb++) {
                LaneEdge Eb;
                double Wb[18];
                int hb;
                bool vb;
                int kb;
                if (cb == ca) {
#pragma unroll
                    for (int i = 0; i < 18; i++) Wb[i] = Wa[i];
                    hb = ha; vb = va; kb = ka;
                } else {
                    lane_linearize(D, pose, X, Y, Z, e0 + 32 * cb + lane, 32 * cb + lane < k, Eb);
                    const double *A = Eb.L.A, *B = Eb.L.B;
#pragma unroll
                    for (int a = 0; a < 6; a++)
#pragma unroll
                        for (int c = 0; c < 3; c++) Wb[3 * a + c] = Eb.wgt * (B[a] * A[c] + B[6 + a] * A[3 + c] + B[12 + a] * A[6 + c]);
                    hb = Eb.hidx; vb = Eb.valid && Eb.hidx >= 0; kb = min(32, k - 32 * cb);
                }
                // rotation: lane a meets lane (a + d). Same chunk: unordered pairs once (d <= ka/2);
                // different chunks: every (a in ca, b in cb) pair, d = 0..31 over the full warp.
                const int nd = (cb == ca) ? (ka / 2 + 1) : 32;
                const int modk = (cb == ca) ? ka : 32;
                for (int d = 0; d < nd; d++) {
                    int src = lane + d;
                    if (src >= modk) src -= modk;
                    if (src < 0 || src > 31) src = 0;
                    double Wp[18];
#pragma unroll
                    for (int i = 0; i < 18; i++) Wp[i] = __shfl_sync(0xffffffffu, Wb[i], src);
                    const int hp = __shfl_sync(0xffffffffu, hb, src);
                    const bool vp = __shfl_sync(0xffffffffu, (int)vb, src) != 0;
                    bool ok = va && vp && lane < ((cb == ca) ? ka : 32) && src < kb;
                    if (cb == ca) {
                        if (lane >= ka) ok = false;
                        // when ka is even, distance ka/2 pairs appear twice: keep the lower lane's copy
                        if (d > 0 && 2 * d == ka && lane >= ka / 2) ok = false;
                    }
                    if (!ok) continue;
                    // M = Ya * Wp^T  (6x6) = W_a Dinv W_p^T, belongs to block (ha, hp)
                    double M[36];
#pragma unroll
                    for (int r = 0; r < 6; r++)
#pragma unroll
                        for (int c = 0; c < 6; c++)
                            M[6 * r + c] = Ya[3 * r] * Wp[3 * c] + Ya[3 * r + 1] * Wp[3 * c + 1] + Ya[3 * r + 2] * Wp[3 * c + 2];
                    if (ha < hp) {
                        double *Sb = O.S + (size_t)(6 * ha) * O.ld + 6 * hp;
#pragma unroll
                        for (int r = 0; r < 6; r++)
#pragma unroll
                            for (int c = 0; c < 6; c++) atomicAdd(Sb + (size_t)r * O.ld + c, -M[6 * r + c]);
                    } else if (ha > hp) {
                        double *Sb = O.S + (size_t)(6 * hp) * O.ld + 6 * ha;
#pragma unroll
                        for (int r = 0; r < 6; r++)
#pragma unroll
                            for (int c = 0; c < 6; c++) atomicAdd(Sb + (size_t)c * O.ld + r, -M[6 * r + c]);
                    } else {
                        double *Sb = O.S + (size_t)(6 * ha) * O.ld + 6 * ha;
                        const bool self = (cb == ca) && d == 0;       // same observation: M is symmetric
#pragma unroll
                        for (int r = 0; r < 6; r++)
#pragma unroll
                            for (int c = r; c < 6; c++)
                                atomicAdd(Sb + (size_t)r * O.ld + c, self ? -M[6 * r + c] : -(M[6 * r + c] + M[6 * c + r]));
                    }
                }
            }
        }
    }
    // deterministic per-block partials (fixed warp order)
    if (lane == 0) { s_chi[warp] = chi_acc; s_max[warp] = max_acc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0, m = 0.0;
        for (int w = 0; w < BUILD_WARPS; w++) { c += s_chi[w]; m = fmax(m, s_max[w]); }
        O.part_chi2[blockIdx.x] = c;
        if (O.mode == 0) O.part_maxdiag[blockIdx.x] = m;
    }
}


// Packed lanes (stage_kernel in schur_pairs.cuh, update_packed_kernel below): lane = observation, a task = a run of whole
// landmarks with <= 32 observations in total, segmented butterfly reductions per landmark.
#define ST_THREADS 256
#ifndef ST_MINB
#define ST_MINB 2                          // resident CTAs per SM asked of the compiler for stage / update_packed
#endif  // ST_MINB
#define ST_WARPS (ST_THREADS / 32)

struct UpdateOut {
    double lambda;
    const double *xp;             // [6 n_free] pose step
    const double *pose_trial;     // poses after oplus
    double *pt_trial;             // [n_points][3]
    double *edge_chi2;            // [n_obs] (sorted order)
    double *part_chi2;            // [gridDim.x]
    double *part_scale;           // [gridDim.x]  sum over landmarks of x_l (lambda x_l + b_l)
    const int *lm_list; int n_list;   // optional: update_kernel then handles only these landmarks (the "wide" ones)
    // chained mode: lambda and the current / trial roles of the two state buffers come from the device state
    const LmDev *lm = nullptr;
    const double *pose_a = nullptr, *pose_b = nullptr; double *pt_a = nullptr, *pt_b = nullptr;
};
// resolve the chained-mode indirection: (pose, pt) = current state, O.pose_trial / O.pt_trial = trial state, O.lambda
BA_DEV bool update_resolve(UpdateOut &O, const double *&pose, const double *&pt) {
    if (!O.lm) return true;
    if (O.lm->done) return false;
    O.lambda = O.lm->lambda;
    const bool f = O.lm->flip != 0;
    pose = f ? O.pose_b : O.pose_a; pt = f ? O.pt_b : O.pt_a;
    O.pose_trial = f ? O.pose_a : O.pose_b; O.pt_trial = f ? O.pt_a : O.pt_b;
    return true;
}

__global__ void __launch_bounds__(BUILD_THREADS) update_kernel(BaDev D, const double *__restrict__ pose_in,
                                                              const double *__restrict__ pt_in, UpdateOut O) {
    __shared__ double s_chi[BUILD_WARPS], s_sc[BUILD_WARPS];
    const double *pose = pose_in, *pt = pt_in;
    if (!update_resolve(O, pose, pt)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * BUILD_WARPS + warp, nw = gridDim.x * BUILD_WARPS;
    double chi_acc = 0.0, sc_acc = 0.0;
    const int n_lm = O.lm_list ? O.n_list : D.n_points;
    for (int jj = gw; jj < n_lm; jj += nw) {
        const int j = O.lm_list ? __ldg(O.lm_list + jj) : jj;
        const int e0 = __ldg(D.lm_ptr + j), k = __ldg(D.lm_ptr + j + 1) - e0;
        const double X = __ldg(pt + 3 * (size_t)j), Y = __ldg(pt + 3 * (size_t)j + 1), Z = __ldg(pt + 3 * (size_t)j + 2);
        if (k == 0) {
            if (lane == 0) { O.pt_trial[3 * (size_t)j] = X; O.pt_trial[3 * (size_t)j + 1] = Y; O.pt_trial[3 * (size_t)j + 2] = Z; }
            continue;
        }
        const int nchunk = (k + 31) >> 5;
        double h[6] = {0, 0, 0, 0, 0, 0}, cl[3] = {0, 0, 0}, bl[3] = {0, 0, 0};
        int nact = 0;
        LaneEdge E;
        for (int c = 0; c < nchunk; c++) {
            lane_linearize(D, pose, X, Y, Z, e0 + 32 * c + lane, 32 * c + lane < k, E);
            if (E.valid) {
                nact++;
                const double *A = E.L.A, *B = E.L.B;
                h[0] += E.wgt * (A[0] * A[0] + A[3] * A[3] + A[6] * A[6]);
                h[1] += E.wgt * (A[0] * A[1] + A[3] * A[4] + A[6] * A[7]);
                h[2] += E.wgt * (A[0] * A[2] + A[3] * A[5] + A[6] * A[8]);
                h[3] += E.wgt * (A[1] * A[1] + A[4] * A[4] + A[7] * A[7]);
                h[4] += E.wgt * (A[1] * A[2] + A[4] * A[5] + A[7] * A[8]);
                h[5] += E.wgt * (A[2] * A[2] + A[5] * A[5] + A[8] * A[8]);
                bl[0] += A[0] * E.g0 + A[3] * E.g1 + A[6] * E.g2;
                bl[1] += A[1] * E.g0 + A[4] * E.g1 + A[7] * E.g2;
                bl[2] += A[2] * E.g0 + A[5] * E.g1 + A[8] * E.g2;
                if (E.hidx >= 0) {
                    // W^T xp = wgt A^T (B xp)
                    const double *x = O.xp + 6 * E.hidx;
                    double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) { const double xa = __ldg(x + a); s0 += B[a] * xa; s1 += B[6 + a] * xa; s2 += B[12 + a] * xa; }
                    cl[0] -= E.wgt * (A[0] * s0 + A[3] * s1 + A[6] * s2);
                    cl[1] -= E.wgt * (A[1] * s0 + A[4] * s1 + A[7] * s2);
                    cl[2] -= E.wgt * (A[2] * s0 + A[5] * s1 + A[8] * s2);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) h[i] = warp_allsum(h[i]);
#pragma unroll
        for (int i = 0; i < 3; i++) { bl[i] = warp_allsum(bl[i]); cl[i] = warp_allsum(cl[i]); }
        nact = __reduce_add_sync(0xffffffffu, nact);
        double nx = X, ny = Y, nz = Z;
        if (nact > 0) {
            h[0] += O.lambda; h[3] += O.lambda; h[5] += O.lambda;
            double di[6];
            sym3_inverse(h, di);
            const double c0 = bl[0] + cl[0], c1 = bl[1] + cl[1], c2 = bl[2] + cl[2];
            const double x0 = di[0] * c0 + di[1] * c1 + di[2] * c2;
            const double x1 = di[1] * c0 + di[3] * c1 + di[4] * c2;
            const double x2 = di[2] * c0 + di[4] * c1 + di[5] * c2;
            nx = X + x0; ny = Y + x1; nz = Z + x2;
            if (lane == 0) sc_acc += x0 * (O.lambda * x0 + bl[0]) + x1 * (O.lambda * x1 + bl[1]) + x2 * (O.lambda * x2 + bl[2]);
        }
        if (lane == 0) { O.pt_trial[3 * (size_t)j] = nx; O.pt_trial[3 * (size_t)j + 1] = ny; O.pt_trial[3 * (size_t)j + 2] = nz; }
        // evaluation at the trial state
        double chi = 0.0;
        for (int c = 0; c < nchunk; c++) {
            const int e = e0 + 32 * c + lane;
            if (32 * c + lane < k) {
                const uint32_t m = D.o_meta[e];
                if (!(m & META_LEVEL1)) {
                    const int ip = __ldg(D.o_pose + e);
                    const int kind = META_KIND(m);
                    const Cam cam = load_cam(D.cams + META_CAM(m));
                    const Pose T = load_pose(O.pose_trial + 7 * (size_t)ip);
                    Pose Trl;
                    if (kind == BAGPU_EDGE_BODY) Trl = load_pose(D.rigs + 7 * META_RIG(m));
                    const double om = __ldg(D.o_w + e);
                    double r0, r1, r2;
                    edge_residual(kind, cam, T, &Trl, nx, ny, nz, __ldg(D.o_u + e), __ldg(D.o_v + e),
                                  (kind == BAGPU_EDGE_STEREO) ? __ldg(D.o_ur + e) : 0.0, false, r0, r1, r2);
                    const double chi2 = r0 * (om * r0) + r1 * (om * r1) + r2 * (om * r2);
                    O.edge_chi2[e] = chi2;
                    double rho0 = chi2, rho1;
                    if (m & META_ROBUST) huber(chi2, kind == BAGPU_EDGE_STEREO ? D.delta_stereo : D.delta_mono, rho0, rho1);
                    chi += rho0;
                }
            }
        }
        chi_acc += warp_allsum(chi);
    }
    if (lane == 0) { s_chi[warp] = chi_acc; s_sc[warp] = sc_acc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0, s = 0.0;
        for (int w = 0; w < BUILD_WARPS; w++) { c += s_chi[w]; s += s_sc[w]; }
        O.part_chi2[blockIdx.x] = c;
        O.part_scale[blockIdx.x] = s;
    }
}

// update_packed_kernel: update_kernel with the lanes packed like stage_kernel (lane = observation, a task = a run of whole
// landmarks with <= 32 observations, segmented reductions per landmark). Landmarks the tasks do not cover ("wide") go
// through update_kernel with a mask.
struct UpdateTasks { const int2 *tasks; int n_tasks; };

__global__ void __launch_bounds__(ST_THREADS, ST_MINB) update_packed_kernel(BaDev D, const double *__restrict__ pose, const double *__restrict__ pt,
                                                                  UpdateOut O, UpdateTasks K) {
    __shared__ double s_chi[ST_WARPS], s_sc[ST_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double chi_acc = 0.0, sc_acc = 0.0;
    for (int t = blockIdx.x * ST_WARPS + warp; t < K.n_tasks; t += gridDim.x * ST_WARPS) {
        const int2 tk = K.tasks[t];
        const int e_first = __ldg(D.lm_ptr + tk.x), nobs = __ldg(D.lm_ptr + tk.y) - e_first;
        const bool in = lane < nobs;
        const int e = e_first + lane;
        const int j = in ? __ldg(D.o_point + e) : -1;
        int head = lane, segl = 1;
        double X = 0, Y = 0, Z = 1;
        if (in) {
            const int p0 = __ldg(D.lm_ptr + j);
            head = lane - (e - p0); segl = __ldg(D.lm_ptr + j + 1) - p0;
            X = __ldg(pt + 3 * (size_t)j); Y = __ldg(pt + 3 * (size_t)j + 1); Z = __ldg(pt + 3 * (size_t)j + 2);
        }
        const int seg_last = head + segl - 1;
        LaneEdge E;
        lane_linearize(D, pose, X, Y, Z, e, in, E);
        double r[12];
#pragma unroll
        for (int i = 0; i < 12; i++) r[i] = 0.0;
        if (E.valid) {
            const double *A = E.L.A, *B = E.L.B;
            r[0] = E.wgt * (A[0] * A[0] + A[3] * A[3] + A[6] * A[6]);
            r[1] = E.wgt * (A[0] * A[1] + A[3] * A[4] + A[6] * A[7]);
            r[2] = E.wgt * (A[0] * A[2] + A[3] * A[5] + A[6] * A[8]);
            r[3] = E.wgt * (A[1] * A[1] + A[4] * A[4] + A[7] * A[7]);
            r[4] = E.wgt * (A[1] * A[2] + A[4] * A[5] + A[7] * A[8]);
            r[5] = E.wgt * (A[2] * A[2] + A[5] * A[5] + A[8] * A[8]);
            r[6] = A[0] * E.g0 + A[3] * E.g1 + A[6] * E.g2;
            r[7] = A[1] * E.g0 + A[4] * E.g1 + A[7] * E.g2;
            r[8] = A[2] * E.g0 + A[5] * E.g1 + A[8] * E.g2;
            if (E.hidx >= 0) {
                const double *x = O.xp + 6 * E.hidx;                    // W^T xp = wgt A^T (B xp)
                double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
                for (int a = 0; a < 6; a++) { const double xa = __ldg(x + a); s0 += B[a] * xa; s1 += B[6 + a] * xa; s2 += B[12 + a] * xa; }
                r[9] = -E.wgt * (A[0] * s0 + A[3] * s1 + A[6] * s2);
                r[10] = -E.wgt * (A[1] * s0 + A[4] * s1 + A[7] * s2);
                r[11] = -E.wgt * (A[2] * s0 + A[5] * s1 + A[8] * s2);
            }
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const bool take = in && lane + off <= seg_last;
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const double o = __shfl_down_sync(0xffffffffu, r[i], off);
                if (take) r[i] += o;
            }
        }
#pragma unroll
        for (int i = 0; i < 12; i++) r[i] = __shfl_sync(0xffffffffu, r[i], head);
        const unsigned vmask = __ballot_sync(0xffffffffu, E.valid);
        const unsigned segmask = (segl >= 32 ? 0xffffffffu : ((1u << segl) - 1u)) << head;
        const bool any_act = (vmask & segmask) != 0u;
        double nx = X, ny = Y, nz = Z;
        if (in && any_act) {
            double h[6] = {r[0] + O.lambda, r[1], r[2], r[3] + O.lambda, r[4], r[5] + O.lambda}, di[6];
            sym3_inverse(h, di);
            const double c0 = r[6] + r[9], c1 = r[7] + r[10], c2 = r[8] + r[11];
            const double x0 = di[0] * c0 + di[1] * c1 + di[2] * c2;
            const double x1 = di[1] * c0 + di[3] * c1 + di[4] * c2;
            const double x2 = di[2] * c0 + di[4] * c1 + di[5] * c2;
            nx = X + x0; ny = Y + x1; nz = Z + x2;
            if (lane == head) sc_acc += x0 * (O.lambda * x0 + r[6]) + x1 * (O.lambda * x1 + r[7]) + x2 * (O.lambda * x2 + r[8]);
        }
        if (in && lane == head) { O.pt_trial[3 * (size_t)j] = nx; O.pt_trial[3 * (size_t)j + 1] = ny; O.pt_trial[3 * (size_t)j + 2] = nz; }
        // evaluation at the trial state
        if (E.valid) {
            const uint32_t m = D.o_meta[e];
            const int ip = __ldg(D.o_pose + e);
            const int kind = META_KIND(m);
            const Cam cam = load_cam(D.cams + META_CAM(m));
            const Pose T = load_pose(O.pose_trial + 7 * (size_t)ip);
            Pose Trl;
            if (kind == BAGPU_EDGE_BODY) Trl = load_pose(D.rigs + 7 * META_RIG(m));
            const double om = __ldg(D.o_w + e);
            double r0, r1, r2;
            edge_residual(kind, cam, T, &Trl, nx, ny, nz, __ldg(D.o_u + e), __ldg(D.o_v + e),
                          (kind == BAGPU_EDGE_STEREO) ? __ldg(D.o_ur + e) : 0.0, false, r0, r1, r2);
            const double chi2 = r0 * (om * r0) + r1 * (om * r1) + r2 * (om * r2);
            O.edge_chi2[e] = chi2;
            double rho0 = chi2, rho1;
            if (m & META_ROBUST) huber(chi2, kind == BAGPU_EDGE_STEREO ? D.delta_stereo : D.delta_mono, rho0, rho1);
            chi_acc += rho0;
        }
    }
    chi_acc = warp_allsum(chi_acc);
    sc_acc = warp_allsum(sc_acc);
    if (lane == 0) { s_chi[warp] = chi_acc; s_sc[warp] = sc_acc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0, s = 0.0;
        for (int w = 0; w < ST_WARPS; w++) { c += s_chi[w]; s += s_sc[w]; }
        O.part_chi2[blockIdx.x] = c;
        O.part_scale[blockIdx.x] = s;
    }
}

// Per-edge pass used between rounds and at the end of a call:
//   mode GATE_LBA : chi2 > gate || !isDepthPositive -> level 1                  (src/Optimizer.cc:3745-3776)
//   mode GATE_POSE: level-1 edges recompute their error, (float)chi2 > (float)gate -> level 1 else 0 (:1013-1100)
//   always        : depth_pos[e] = isDepthPositive() on the current estimates     (OptimizableTypes.h:107-111,135-139)
//   drop_kernel   : setRobustKernel(0)
__global__ void gate_kernel(BaDev D, const double *__restrict__ pose, const double *__restrict__ pt, double *edge_chi2,
                            uint8_t *depth_pos, int gate_mode, double gate_mono, double gate_stereo, int drop_kernel) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= D.n_obs) return;
    uint32_t m = D.o_meta[e];
    const int ip = D.o_pose[e], jp = D.o_point[e];
    const int kind = META_KIND(m);
    const Pose T = load_pose(pose + 7 * (size_t)ip);
    Pose Trl;
    if (kind == BAGPU_EDGE_BODY) Trl = load_pose(D.rigs + 7 * META_RIG(m));
    const Cam cam = load_cam(D.cams + META_CAM(m));
    const double X = pt[3 * (size_t)jp], Y = pt[3 * (size_t)jp + 1], Z = pt[3 * (size_t)jp + 2];
    double r0, r1, r2;
    const double z = edge_residual(kind, cam, T, &Trl, X, Y, Z, D.o_u[e], D.o_v[e],
                                   (kind == BAGPU_EDGE_STEREO) ? D.o_ur[e] : 0.0, false, r0, r1, r2);
    const bool dp = z > 0.0;
    if (depth_pos) depth_pos[e] = dp ? 1 : 0;
    if (gate_mode == BAGPU_GATE_LBA) {
        const double th = (kind == BAGPU_EDGE_STEREO) ? gate_stereo : gate_mono;
        if (edge_chi2[e] > th || !dp) m |= META_LEVEL1;
    } else if (gate_mode == BAGPU_GATE_POSE) {
        double c = edge_chi2[e];
        if (m & META_LEVEL1) {
            const double om = D.o_w[e];
            c = r0 * (om * r0) + r1 * (om * r1) + r2 * (om * r2);
            edge_chi2[e] = c;
        }
        const float th = (kind == BAGPU_EDGE_STEREO) ? (float)gate_stereo : (float)gate_mono;
        if ((float)c > th) m |= META_LEVEL1; else m &= ~META_LEVEL1;
    }
    if (drop_kernel) m &= ~META_ROBUST;
    D.o_meta[e] = m;
}

// count active (level-0) edges; one block-wide atomic per block
__global__ void count_active_kernel(const uint32_t *meta, int64_t n, unsigned long long *out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int a = (e < n && !(meta[e] & META_LEVEL1)) ? 1 : 0;
    const int s = __reduce_add_sync(0xffffffffu, a);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, (unsigned long long)s);
}

// Pose step: T_trial = exp(x_p) * T (types_six_dof_expmap.h:73-76) for free poses, copy for fixed ones;
// pose part of computeScale. Single block, fixed-order reduction.
__global__ void pose_update_kernel(int n_poses, const int *__restrict__ hidx, const double *__restrict__ pose_a,
                                   double *pose_b, const double *__restrict__ xp, const double *__restrict__ bp,
                                   double lambda, double *scale_out, const LmDev *lm = nullptr) {
    __shared__ double sh[1024];
    if (lm_done(lm)) return;
    if (lm) lambda = lm->lambda;
    const double *pose = (lm && lm->flip) ? pose_b : pose_a;
    double *pose_trial = (lm && lm->flip) ? const_cast<double *>(pose_a) : pose_b;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_poses; i += blockDim.x) {
        const Pose T = load_pose(pose + 7 * (size_t)i);
        const int h = hidx[i];
        if (h < 0) { store_pose(pose_trial + 7 * (size_t)i, T); continue; }
        double u[6];
#pragma unroll
        for (int a = 0; a < 6; a++) { u[a] = xp[6 * h + a]; acc += u[a] * (lambda * u[a] + bp[6 * h + a]); }
        store_pose(pose_trial + 7 * (size_t)i, pose_oplus(T, u));
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *scale_out = sh[0];
}

// Fixed-order reduction of per-block partials into the small status record the host reads.
//   out[0] = sum part_a, out[1] = sum part_b (+ add_b), out[2] = max part_c
__global__ void reduce_partials_kernel(int n, const double *a, const double *b, const double *c, const double *add_b,
                                       const double *hpp_diag, int n_hpp, double *out, const double *a2 = nullptr, int n2 = 0) {
    __shared__ double sa[256], sb[256], sc[256];
    double va = 0, vb = 0, vc = 0;
    for (int i = threadIdx.x; i < n; i += 256) {
        if (a) va += a[i];
        if (b) vb += b[i];
        if (c) vc = fmax(vc, c[i]);
    }
    if (hpp_diag) for (int i = threadIdx.x; i < n_hpp; i += 256) vc = fmax(vc, fabs(hpp_diag[i]));
    if (a2) for (int i = threadIdx.x; i < n2; i += 256) va += a2[i];
    sa[threadIdx.x] = va; sb[threadIdx.x] = vb; sc[threadIdx.x] = vc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            sa[threadIdx.x] += sa[threadIdx.x + s];
            sb[threadIdx.x] += sb[threadIdx.x + s];
            sc[threadIdx.x] = fmax(sc[threadIdx.x], sc[threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = sa[0]; out[1] = sb[0] + (add_b ? *add_b : 0.0); out[2] = sc[0]; }
}

// End of an LM trial: the three sums the host decides on, in one launch (block 0: chi2 at the linearisation point, block 1:
// chi2 at the trial state, block 2: landmark part of computeScale), each a fixed-order reduction of per-CTA partials
// (+ the partials of the long-track kernels), and the failure flag next to them.
struct TrialSums { const double *a[3]; int na[3]; const double *b[3]; int nb[3]; const int *fail; double *out; };
__global__ void finish_trial_kernel(TrialSums T) {
    __shared__ double sh[256];
    const int q = blockIdx.x;
    double v = 0.0;
    for (int i = threadIdx.x; i < T.na[q]; i += 256) v += T.a[q][i];
    if (T.b[q]) for (int i = threadIdx.x; i < T.nb[q]; i += 256) v += T.b[q][i];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        T.out[q] = sh[0];
        if (q == 0) { const int wd = T.fail[1]; T.out[5] = wd ? (double)wd : (double)T.fail[0]; }   // fail[1]: starved wait (2 / 3), never a numerical verdict
    }
}

// out[perm[i]] = in[i]  (sorted order -> caller's edge order)
// ---- chained mode (no host round trip per trial)
// computeLambdaInit on the device: lambda = tau * max diagonal (dstat[2], from reduce_partials_kernel) or the caller's value
__global__ void lm_init_kernel(LmDev *lm, const double *dstat, double lambda_user, int iterations, int round, int max_trace) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    lm->lambda = (lambda_user > 0) ? lambda_user : 1e-5 * dstat[2];
    lm->ni = 2.0; lm->currentChi = 0.0; lm->iniChi = 0.0; lm->rho = 0.0;
    lm->flip = 0; lm->qmax = 0; lm->it = 0; lm->nBad = 0; lm->first = 1; lm->done = (iterations <= 0) ? 1 : 0; lm->status = BAGPU_OK;
    lm->iterations = iterations; lm->round = round; lm->n_trace = 0; lm->max_trace = max_trace; lm->trials = 0; lm->starved = 0;
}
// End of a trial: the three sums (as finish_trial_kernel, same fixed order) and the decision of
// OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:99-166) + the stop rules of
// SparseOptimizer::optimize (sparse_optimizer.cpp:380-416) as bagpu.cu's host loop applies them.
__global__ void lm_decide_kernel(TrialSums T, const double *pose_scale, LmDev *lm, LmTraceDev *trace) {
    __shared__ double sh[256];
    __shared__ double sums[3];
    if (lm->done) return;
    for (int q = 0; q < 3; q++) {
        double v = 0.0;
        for (int i = threadIdx.x; i < T.na[q]; i += 256) v += T.a[q][i];
        if (T.b[q]) for (int i = threadIdx.x; i < T.nb[q]; i += 256) v += T.b[q][i];
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) sums[q] = sh[0];
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    T.out[0] = sums[0]; T.out[1] = sums[1]; T.out[2] = sums[2];
    const int failflag = T.fail[1] ? T.fail[1] : T.fail[0];
    if (failflag >= 2) { lm->starved = 1; lm->done = 1; return; }          // cannot happen without the overlap; the host reports it
    const bool ok2 = failflag == 0;
    double currentChi = lm->currentChi, iniChi = lm->iniChi, lambda = lm->lambda, ni = lm->ni;
    if (lm->first) { currentChi = sums[0]; iniChi = currentChi; lm->first = 0; }
    const double tempChi = ok2 ? sums[1] : DBL_MAX;
    double rho = currentChi - tempChi;
    double scale = sums[2] + *pose_scale;
    scale += 1e-3;
    rho /= scale;
    if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - pow((2 * rho - 1), 3.0);
        alpha = fmin(alpha, 2. / 3.);
        const double scaleFactor = fmax(1. / 3., alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
        lm->flip ^= 1;                                                     // discardTop(): the trial state becomes current
    } else {
        lambda *= ni;
        ni *= 2;                                                           // pop()
    }
    int qmax = lm->qmax + 1;
    lm->trials++;
    lm->rho = rho;
    if (!(rho < 0 && qmax < 10)) {                                         // the iteration ends
        int stt = BAGPU_OK, nBad = lm->nBad;
        if (qmax == 10 || rho == 0) stt = BAGPU_TERMINATE_TRIALS;
        else {
            if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
            if (nBad >= 3) stt = BAGPU_TERMINATE_NBAD;
        }
        lm->nBad = nBad;
        if (lm->n_trace < lm->max_trace) {
            LmTraceDev &t = trace[lm->n_trace++];
            t.round = lm->round; t.iteration = lm->it; t.trials = qmax; t.status = stt; t.chi2_before = iniChi; t.chi2_after = currentChi; t.lambda = lambda;
        }
        lm->it++;
        lm->status = stt;
        lm->first = 1;
        qmax = 0;
        if (stt != BAGPU_OK || lm->it >= lm->iterations) lm->done = 1;
    }
    lm->qmax = qmax; lm->lambda = lambda; lm->ni = ni; lm->currentChi = currentChi; lm->iniChi = iniChi;
}

template <typename T>
__global__ void scatter_perm_kernel(int64_t n, const int *__restrict__ perm, const T *__restrict__ in, T *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[perm[i]] = in[i];
}
__global__ void level_from_meta_kernel(int64_t n, const int *__restrict__ perm, const uint32_t *__restrict__ meta, uint8_t *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[perm ? perm[i] : i] = (meta[i] & META_LEVEL1) ? 1 : 0;
}
