// Linearise + Schur pass, pair-major (sm_100a).
//
// Replaces BlockSolver::buildSystem + setLambda + the landmark loop of BlockSolver::solve
// (Thirdparty/g2o/g2o/core/block_solver.hpp:501-560, 373-444) by two kernels that never read-modify-write the reduced
// camera system while they accumulate:
//
//   stage_kernel / stage_wide_kernel   landmark-major (lane = observation). Linearise, reduce the landmark's Hll / bl,
//       factor Hll + lambda I = L L^T (3x3 Cholesky, in registers) and write per observation
//         Z  = W L^-T = [P^T Y; Y]         (6x3; W = rho1 w B^T A is the Hpl block; stored as Y (3x3) and X_l, P = -[X_l]x)
//         Dr = [ w M^T M (6) | M^T g (3) | Y L^-1 bl (3) | X_l (3) ], B = M [-[X_l]x | I]   -> the per-camera sums (Hpp, bp, bs)
//       and per landmark the factor L, L^-1 bl and bl (update_z_kernel back-substitutes from these, no second linearisation).
//   pair_kernel    camera-pair-major. At upload the device lists, for every upper block (a, b) of the reduced system,
//       the observation pairs (e_a, e_b) of the landmarks both cameras see, sorted by block (pair_plan_*). A warp takes
//       one chunk of one block's list, each lane accumulates Z_a Z_b^T of its entries in 36 registers (G = Y_a Y_b^T, then
//       [P_a^T G P_b, P_a^T G; G P_b, G]), the warp sums the lanes once and one total per block reaches S. The block is a
//       contraction over its shared landmarks, so the work per entry is two 96-byte gathers (L2: neighbouring blocks share
//       the records) and 81 + 27 MAC. Diagonal blocks also sum the Dr records of their camera (Hpp, bp, bs).
//
// Any track length, any number of edges on one (pose, point) pair and any camera span go through the same path.
#pragma once
#include "ba_kernels.cuh"

#define ZR_STRIDE 12                       // Y (9) | X_l (3): 96-byte records = three 32-byte sectors, one 256-bit access each
#define DR_STRIDE 16                       // 15 used: [w M^T M (6) | M^T g (3) | Y L^-1 b_l (3) | X_l (3)]: 128 bytes = exactly two 64-byte L2 fetches per gather
                                           // (a 96-byte record + the X_l sector of the Z record costs 128 + 64 bytes of DRAM reads at the L2's 64-byte fetch granularity)
#define LM_STRIDE 12
#ifndef UZ_MINB
#define UZ_MINB 3                          // resident CTAs per SM asked of the compiler for update_z_kernel
#endif
#ifndef PK_CHUNK
#define PK_CHUNK 512                       // entries per work item (per-item reductions vs the working set of the chunks in flight)
#endif
#ifndef PK_MINB
#define PK_MINB 3
#endif
#define PK_THREADS 128
#define PK_WARPS (PK_THREADS / 32)
#define PK_RED_LD 33
#define PK_SMEM_BYTES (PK_WARPS * 36 * PK_RED_LD * 8)

struct PairItem { int a, b, begin, end, first, nit; };   // [begin, end) entries; the block's items are first .. first + nit - 1

// 256-bit global accesses (sm_100a): one lane moves one whole 32-byte sector per instruction
BA_DEV void ldg256(const double *p, double *v) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
BA_DEV void stg256(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// plan (upload): observations are sorted landmark-major / pose-ascending, free-pose indices are monotone in the pose index,
// so inside a landmark e <= e2 implies hidx(e) <= hidx(e2): every pair lands in the upper triangle.
// Block id = row_pos[a] * bw1 + (b - a) with bw1 = band_blocks + 1 (the envelope agreed at upload bounds b - a); row_pos is the
// order in which pair_kernel takes the camera rows (identity, or from both ends when the factorisation is two-way).
__global__ void pair_count_kernel(int64_t n_obs, const int *__restrict__ lm_ptr, const int *__restrict__ o_pose,
                                  const int *__restrict__ o_point, const int *__restrict__ hidx, const int *__restrict__ row_pos, int bw1,
                                  unsigned *__restrict__ npairs, unsigned *blk_cnt) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_obs) return;
    const int ha = hidx[o_pose[e]];
    unsigned cnt = 0;
    if (ha >= 0) {
        const int end = lm_ptr[o_point[e] + 1];
        for (int e2 = (int)e; e2 < end; e2++) {
            const int hb = hidx[o_pose[e2]];
            if (hb < 0) continue;
            cnt++;
            atomicAdd(blk_cnt + (size_t)row_pos[ha] * bw1 + (hb - ha), 1u);
        }
    }
    npairs[e] = cnt;
}

__global__ void pair_gen_kernel(int64_t n_obs, const int *__restrict__ lm_ptr, const int *__restrict__ o_pose,
                                const int *__restrict__ o_point, const int *__restrict__ hidx, const int *__restrict__ row_pos, int bw1,
                                const unsigned *__restrict__ pair_off, unsigned *__restrict__ keys, int2 *__restrict__ vals) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_obs) return;
    const int ha = hidx[o_pose[e]];
    if (ha < 0) return;
    const int end = lm_ptr[o_point[e] + 1];
    const int pa = row_pos[ha];
    unsigned pos = pair_off[e];
    for (int e2 = (int)e; e2 < end; e2++) {
        const int hb = hidx[o_pose[e2]];
        if (hb < 0) continue;
        keys[pos] = (unsigned)pa * (unsigned)bw1 + (unsigned)(hb - ha);
        vals[pos] = make_int2((int)e, e2);
        pos++;
    }
}

__global__ void pair_item_count_kernel(int nblk, const unsigned *__restrict__ blk_cnt, unsigned *__restrict__ item_cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nblk) item_cnt[i] = (blk_cnt[i] + PK_CHUNK - 1) / PK_CHUNK;
}

__global__ void pair_item_fill_kernel(int nblk, int bw1, const int *__restrict__ row_of_pos, const unsigned *__restrict__ blk_off, const unsigned *__restrict__ blk_cnt,
                                      const unsigned *__restrict__ item_off, PairItem *__restrict__ items) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblk) return;
    const unsigned cnt = blk_cnt[i];
    if (!cnt) return;
    const int pa = i / bw1, a = row_of_pos[pa], b = a + (i - pa * bw1);
    const unsigned off = blk_off[i];
    unsigned io = item_off[i];
    for (unsigned c = 0; c < cnt; c += PK_CHUNK) {
        PairItem I; I.a = a; I.b = b; I.begin = (int)(off + c); I.end = (int)(off + min(cnt, c + PK_CHUNK));
        I.first = (int)item_off[i]; I.nit = (int)((cnt + PK_CHUNK - 1) / PK_CHUNK);
        items[io++] = I;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage
struct LmFactor { double i00, l10, l20, i11, l21, i22, c0, c1, c2; };

// r = [Hll 00 01 02 11 12 22 | bl 0 1 2]; Hll + lambda I = L L^T, c = L^-1 bl. false if a pivot is not positive.
BA_DEV bool lm_factor(const double *r, double lambda, LmFactor &F) {
    const double h0 = r[0] + lambda, h3 = r[3] + lambda, h5 = r[5] + lambda;
    bool ok = h0 > 0.0;
    F.i00 = 1.0 / sqrt(h0);
    F.l10 = r[1] * F.i00; F.l20 = r[2] * F.i00;
    const double p1 = h3 - F.l10 * F.l10;
    ok = ok && p1 > 0.0;
    F.i11 = 1.0 / sqrt(p1);
    F.l21 = (r[4] - F.l20 * F.l10) * F.i11;
    const double p2 = h5 - F.l20 * F.l20 - F.l21 * F.l21;
    ok = ok && p2 > 0.0;
    F.i22 = 1.0 / sqrt(p2);
    if (!ok) { F.i00 = F.i11 = F.i22 = 0.0; F.l10 = F.l20 = F.l21 = 0.0; }
    F.c0 = r[6] * F.i00;
    F.c1 = (r[7] - F.l10 * F.c0) * F.i11;
    F.c2 = (r[8] - F.l20 * F.c0 - F.l21 * F.c1) * F.i22;
    return ok;
}

// the landmark's contribution to Hll (6), bl (3) and the robustified cost (1) of one lane
BA_DEV void lane_hll(const LaneEdge &E, double *r) {
    const double *A = E.L.A;
    const double w = E.wgt;
    r[0] = E.valid ? w * (A[0] * A[0] + A[3] * A[3] + A[6] * A[6]) : 0.0;
    r[1] = E.valid ? w * (A[0] * A[1] + A[3] * A[4] + A[6] * A[7]) : 0.0;
    r[2] = E.valid ? w * (A[0] * A[2] + A[3] * A[5] + A[6] * A[8]) : 0.0;
    r[3] = E.valid ? w * (A[1] * A[1] + A[4] * A[4] + A[7] * A[7]) : 0.0;
    r[4] = E.valid ? w * (A[1] * A[2] + A[4] * A[5] + A[7] * A[8]) : 0.0;
    r[5] = E.valid ? w * (A[2] * A[2] + A[5] * A[5] + A[8] * A[8]) : 0.0;
    r[6] = E.valid ? A[0] * E.g0 + A[3] * E.g1 + A[6] * E.g2 : 0.0;
    r[7] = E.valid ? A[1] * E.g0 + A[4] * E.g1 + A[7] * E.g2 : 0.0;
    r[8] = E.valid ? A[2] * E.g0 + A[5] * E.g1 + A[8] * E.g2 : 0.0;
    r[9] = E.valid ? E.rho0 : 0.0;
}

// write the Z and Dr records of observation e (free pose). Inactive (level-1) edges write exact zeros.
BA_DEV void lane_emit(const LaneEdge &E, const LmFactor &F, int64_t e, double *__restrict__ Zp, double *__restrict__ Dp) {
    // B = M [P | I] with M = columns 3..5 of B (= Jn R_rl) and P = -[X_l]x, so the Hpl block is W = w B^T A = [P^T V; V] with the
    // 3x3 V = w M^T A, and Z = W L^-T = [P^T Y; Y] with Y = V L^-T: 9 + 3 numbers per observation instead of 18. In the same way
    // w B^T B = [P | I]^T N [P | I] with N = w M^T M (6), B^T g = [P^T m; m] with m = M^T g (3), and the Schur right-hand side
    // term -Z L^-1 b_l = -[P^T q; q] with q = Y L^-1 b_l (3). pair_kernel / update_z_kernel rebuild what they need.
    const bool act = E.valid;
    const double *A = E.L.A, *B = E.L.B;
    double *zo = Zp + ZR_STRIDE * e, *dout = Dp + DR_STRIDE * e;
    double y[9], q[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double v0 = act ? E.wgt * (B[3 + i] * A[0] + B[9 + i] * A[3] + B[15 + i] * A[6]) : 0.0;
        const double v1 = act ? E.wgt * (B[3 + i] * A[1] + B[9 + i] * A[4] + B[15 + i] * A[7]) : 0.0;
        const double v2 = act ? E.wgt * (B[3 + i] * A[2] + B[9 + i] * A[5] + B[15 + i] * A[8]) : 0.0;
        const double y0 = v0 * F.i00;
        const double y1 = (v1 - y0 * F.l10) * F.i11;
        const double y2 = (v2 - y0 * F.l20 - y1 * F.l21) * F.i22;
        y[3 * i] = y0; y[3 * i + 1] = y1; y[3 * i + 2] = y2;
        q[i] = y0 * F.c0 + y1 * F.c1 + y2 * F.c2;
    }
    const double X0 = act ? E.L.xl[0] : 0.0, X1 = act ? E.L.xl[1] : 0.0, X2 = act ? E.L.xl[2] : 0.0;
    stg256(zo, y[0], y[1], y[2], y[3]);
    stg256(zo + 4, y[4], y[5], y[6], y[7]);
    stg256(zo + 8, y[8], X0, X1, X2);
    const double m00 = B[3], m01 = B[4], m02 = B[5], m10 = B[9], m11 = B[10], m12 = B[11], m20 = B[15], m21 = B[16], m22 = B[17];
    const double n00 = act ? E.wgt * (m00 * m00 + m10 * m10 + m20 * m20) : 0.0;
    const double n01 = act ? E.wgt * (m00 * m01 + m10 * m11 + m20 * m21) : 0.0;
    const double n02 = act ? E.wgt * (m00 * m02 + m10 * m12 + m20 * m22) : 0.0;
    const double n11 = act ? E.wgt * (m01 * m01 + m11 * m11 + m21 * m21) : 0.0;
    const double n12 = act ? E.wgt * (m01 * m02 + m11 * m12 + m21 * m22) : 0.0;
    const double n22 = act ? E.wgt * (m02 * m02 + m12 * m12 + m22 * m22) : 0.0;
    const double g0 = act ? (m00 * E.g0 + m10 * E.g1 + m20 * E.g2) : 0.0;
    const double g1 = act ? (m01 * E.g0 + m11 * E.g1 + m21 * E.g2) : 0.0;
    const double g2 = act ? (m02 * E.g0 + m12 * E.g1 + m22 * E.g2) : 0.0;
    stg256(dout, n00, n01, n02, n11);
    stg256(dout + 4, n12, n22, g0, g1);
    stg256(dout + 8, g2, q[0], q[1], q[2]);
    stg256(dout + 12, X0, X1, X2, 0.0);
}

struct StageArgs {
    const int2 *tasks;            // [n_tasks] landmark range [x, y): whole landmarks, <= 32 observations in total
    int n_tasks;
    const int *lm_list; int n_list;   // stage_wide_kernel: landmarks with more than 32 observations
    double *Z, *Dr;
    double *Lm;                   // [n_points][LM_STRIDE] landmark records: L^-1 diag / L off-diag (6) | c = L^-1 bl (3) | bl (3)
    double lambda;
    double *part_chi2;            // [gridDim.x]
    double *part_maxdiag;         // optional [gridDim.x]: max |Hll diagonal| (computeLambdaInit)
    int *fail;
    const LmDev *lm = nullptr; const double *pose_b = nullptr, *pt_b = nullptr;   // chained mode: lambda / current buffers from the device state
};
BA_DEV bool stage_resolve(StageArgs &S, const double *&pose, const double *&pt) {
    if (!S.lm) return true;
    if (S.lm->done) return false;
    S.lambda = S.lm->lambda;
    if (S.lm->flip) { pose = S.pose_b; pt = S.pt_b; }
    return true;
}

__global__ void __launch_bounds__(ST_THREADS, ST_MINB) stage_kernel(BaDev D, const double *__restrict__ pose_in, const double *__restrict__ pt_in, StageArgs S) {
    __shared__ double s_chi[ST_WARPS], s_max[ST_WARPS];
    const double *pose = pose_in, *pt = pt_in;
    if (!stage_resolve(S, pose, pt)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double chi_acc = 0.0, max_acc = 0.0;
    bool bad = false;
    for (int t = blockIdx.x * ST_WARPS + warp; t < S.n_tasks; t += gridDim.x * ST_WARPS) {
        const int2 tk = S.tasks[t];
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
        double r[10];
        lane_hll(E, r);
        // segmented reduction (fixed order): after the 5 steps the head lane of every landmark holds its sums
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const bool take = in && lane + off <= seg_last;
#pragma unroll
            for (int i = 0; i < 10; i++) {
                const double o = __shfl_down_sync(0xffffffffu, r[i], off);
                if (take) r[i] += o;
            }
        }
#pragma unroll
        for (int i = 0; i < 10; i++) r[i] = __shfl_sync(0xffffffffu, r[i], head);
        if (in && lane == head) { chi_acc += r[9]; max_acc = fmax(max_acc, fmax(fabs(r[0]), fmax(fabs(r[3]), fabs(r[5])))); }
        LmFactor F;
        const bool ok = lm_factor(r, S.lambda, F);
        if (in && !ok) bad = true;
        // level-1 edges on a free pose still own a record (the pair lists do not change between rounds): exact zeros
        const int hx = in ? __ldg(D.pose_hidx + __ldg(D.o_pose + e)) : -1;
        if (hx >= 0) lane_emit(E, F, e, S.Z, S.Dr);
        if (in && lane == head) {                              // what update_z_kernel needs to back-substitute this landmark
            double *lm = S.Lm + LM_STRIDE * (size_t)j;
            stg256(lm, F.i00, F.l10, F.l20, F.i11);
            stg256(lm + 4, F.l21, F.i22, F.c0, F.c1);
            stg256(lm + 8, F.c2, r[6], r[7], r[8]);
        }
    }
    if (bad) atomicOr(S.fail, 1);
    chi_acc = warp_allsum(chi_acc);
    max_acc = warp_allmax(max_acc);
    if (lane == 0) { s_chi[warp] = chi_acc; s_max[warp] = max_acc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0, m = 0.0;
        for (int w = 0; w < ST_WARPS; w++) { c += s_chi[w]; m = fmax(m, s_max[w]); }
        S.part_chi2[blockIdx.x] = c;
        if (S.part_maxdiag) S.part_maxdiag[blockIdx.x] = m;
    }
}

// the same for landmarks with more than 32 observations: warp = landmark, two passes over its chunks
__global__ void __launch_bounds__(ST_THREADS, ST_MINB) stage_wide_kernel(BaDev D, const double *__restrict__ pose_in, const double *__restrict__ pt_in, StageArgs S) {
    __shared__ double s_chi[ST_WARPS], s_max[ST_WARPS];
    const double *pose = pose_in, *pt = pt_in;
    if (!stage_resolve(S, pose, pt)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double chi_acc = 0.0, max_acc = 0.0;
    bool bad = false;
    for (int jj = blockIdx.x * ST_WARPS + warp; jj < S.n_list; jj += gridDim.x * ST_WARPS) {
        const int j = __ldg(S.lm_list + jj);
        const int e0 = __ldg(D.lm_ptr + j), k = __ldg(D.lm_ptr + j + 1) - e0;
        if (k == 0) continue;
        const double X = __ldg(pt + 3 * (size_t)j), Y = __ldg(pt + 3 * (size_t)j + 1), Z = __ldg(pt + 3 * (size_t)j + 2);
        const int nchunk = (k + 31) >> 5;
        double acc[10];
#pragma unroll
        for (int i = 0; i < 10; i++) acc[i] = 0.0;
        LaneEdge E;
        for (int c = 0; c < nchunk; c++) {
            lane_linearize(D, pose, X, Y, Z, e0 + 32 * c + lane, 32 * c + lane < k, E);
            double r[10];
            lane_hll(E, r);
#pragma unroll
            for (int i = 0; i < 10; i++) acc[i] += r[i];
        }
#pragma unroll
        for (int i = 0; i < 10; i++) acc[i] = warp_allsum(acc[i]);
        chi_acc += acc[9];
        max_acc = fmax(max_acc, fmax(fabs(acc[0]), fmax(fabs(acc[3]), fabs(acc[5]))));
        LmFactor F;
        if (!lm_factor(acc, S.lambda, F)) bad = true;
        for (int c = 0; c < nchunk; c++) {
            const bool in = 32 * c + lane < k;
            lane_linearize(D, pose, X, Y, Z, e0 + 32 * c + lane, in, E);
            const int hx = in ? __ldg(D.pose_hidx + __ldg(D.o_pose + e0 + 32 * c + lane)) : -1;
            if (hx >= 0) lane_emit(E, F, e0 + 32 * c + lane, S.Z, S.Dr);
        }
    }
    if (bad) atomicOr(S.fail, 1);
    if (lane == 0) { s_chi[warp] = chi_acc; s_max[warp] = max_acc; }          // warp_allsum left the same value in every lane
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0, m = 0.0;
        for (int w = 0; w < ST_WARPS; w++) { c += s_chi[w]; m = fmax(m, s_max[w]); }
        S.part_chi2[blockIdx.x] = c;
        if (S.part_maxdiag) S.part_maxdiag[blockIdx.x] = m;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// pair
#define PK_PART 72                         // doubles per item in the partial buffer: Z Z^T (36) | Dr sums (33) | pad
struct PairArgs {
    const PairItem *items; int n_items;
    const int2 *entries;
    const double *Z, *Dr;
    double *S; int ld; double *bp, *bs;
    double *part;                 // [n_items][PK_PART] partial sums of blocks cut into several items
    unsigned *blk_done; int bw1;  // finished items per block (never reset: the last finisher is the one that completes a multiple of nit)
    unsigned *row_done;           // optional: finished items per camera row (chol_band_kernel runs beside this kernel)
    double *S2; int n_tot, n1;    // optional (two-way factorisation): elements with column >= n1 go to the mirrored system S2
    double *hpp_diag;             // optional [6 n_free]: diagonal-only pass of computeLambdaInit (only the Dr sums of the diagonal blocks)
};

// P^T u for P = -[X]x: (-z u1 + y u2, z u0 - x u2, -y u0 + x u1)
#define PT0(x, y, z, u0, u1, u2) (-(z) * (u1) + (y) * (u2))
#define PT1(x, y, z, u0, u1, u2) ((z) * (u0) - (x) * (u2))
#define PT2(x, y, z, u0, u1, u2) (-(y) * (u0) + (x) * (u1))

// acc (6x6, row-major) += Z_x Z_y^T with Z = [P^T Y; Y]:  G = Y_x Y_y^T,  Z_x Z_y^T = [P_x^T G P_y, P_x^T G; G P_y, G]
BA_DEV void pair_accumulate(const double *__restrict__ Z, int ex, int ey, double *acc) {
    const double *px = Z + ZR_STRIDE * (size_t)ex, *py = Z + ZR_STRIDE * (size_t)ey;
    double a[ZR_STRIDE], b[ZR_STRIDE];
#pragma unroll
    for (int i = 0; i < ZR_STRIDE / 4; i++) ldg256(py + 4 * i, b + 4 * i);
#pragma unroll
    for (int i = 0; i < ZR_STRIDE / 4; i++) ldg256(px + 4 * i, a + 4 * i);
    const double ax = a[9], ay = a[10], az = a[11], bx = b[9], by = b[10], bz = b[11];
    double G[9], H[9];                                   // G = Y_a Y_b^T,  H = G P_b  (H[i][c] = (P_b^T applied to row i of G)[c])
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) G[3 * i + k] = a[3 * i] * b[3 * k] + a[3 * i + 1] * b[3 * k + 1] + a[3 * i + 2] * b[3 * k + 2];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        H[3 * i + 0] = PT0(bx, by, bz, G[3 * i], G[3 * i + 1], G[3 * i + 2]);
        H[3 * i + 1] = PT1(bx, by, bz, G[3 * i], G[3 * i + 1], G[3 * i + 2]);
        H[3 * i + 2] = PT2(bx, by, bz, G[3 * i], G[3 * i + 1], G[3 * i + 2]);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        // rotation rows (0..2): P_a^T applied to the columns of H (left block) and of G (right block)
        acc[0 * 6 + c] += PT0(ax, ay, az, H[c], H[3 + c], H[6 + c]);
        acc[1 * 6 + c] += PT1(ax, ay, az, H[c], H[3 + c], H[6 + c]);
        acc[2 * 6 + c] += PT2(ax, ay, az, H[c], H[3 + c], H[6 + c]);
        acc[0 * 6 + 3 + c] += PT0(ax, ay, az, G[c], G[3 + c], G[6 + c]);
        acc[1 * 6 + 3 + c] += PT1(ax, ay, az, G[c], G[3 + c], G[6 + c]);
        acc[2 * 6 + 3 + c] += PT2(ax, ay, az, G[c], G[3 + c], G[6 + c]);
        // translation rows (3..5)
        acc[3 * 6 + c] += H[c]; acc[4 * 6 + c] += H[3 + c]; acc[5 * 6 + c] += H[6 + c];
        acc[3 * 6 + 3 + c] += G[c]; acc[4 * 6 + 3 + c] += G[3 + c]; acc[5 * 6 + 3 + c] += G[6 + c];
    }
}

__global__ void __launch_bounds__(PK_THREADS, PK_MINB) pair_kernel(PairArgs P) {
    extern __shared__ double pk_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *red = pk_sm + warp * 36 * PK_RED_LD;
    for (int it = blockIdx.x * PK_WARPS + warp; it < P.n_items; it += gridDim.x * PK_WARPS) {
        const PairItem I = P.items[it];
        const bool diag = I.a == I.b;
        if (P.hpp_diag && !diag) continue;
        // ---- this item's share of sum Z_a Z_b^T: lane l ends with elements l and l + 32 (sz0, sz1)
        double sz0 = 0.0, sz1 = 0.0, sd0 = 0.0, sd1 = 0.0;
        if (!P.hpp_diag) {
            double acc[36];
#pragma unroll
            for (int i = 0; i < 36; i++) acc[i] = 0.0;
            for (int i = I.begin + lane; i < I.end; i += 32) {
                const int2 en = __ldg(P.entries + i);
                pair_accumulate(P.Z, en.x, en.y, acc);
                if (diag && en.x != en.y) pair_accumulate(P.Z, en.y, en.x, acc);     // two edges on one (pose, point): M + M^T
            }
#pragma unroll
            for (int i = 0; i < 36; i++) red[i * PK_RED_LD + lane] = acc[i];
            __syncwarp();
#pragma unroll 8
            for (int l = 0; l < 32; l++) sz0 += red[lane * PK_RED_LD + l];
            if (lane < 4) {
#pragma unroll 8
                for (int l = 0; l < 32; l++) sz1 += red[(32 + lane) * PK_RED_LD + l];
            }
            __syncwarp();
        }
        if (diag) {
            // Hpp, bp, bs of camera a: sum of the Dr records of its observations (the (e, e) entries)
            double d[33];
#pragma unroll
            for (int i = 0; i < 33; i++) d[i] = 0.0;
            for (int i = I.begin + lane; i < I.end; i += 32) {
                const int2 en = __ldg(P.entries + i);
                if (en.x != en.y) continue;
                const double *pd = P.Dr + DR_STRIDE * (size_t)en.x;
                double v[DR_STRIDE];
#pragma unroll
                for (int q = 0; q < DR_STRIDE / 4; q++) ldg256(pd + 4 * q, v + 4 * q);
                // v = N (6) | m (3) | q (3) | X (3).  w B^T B = [P | I]^T N [P | I] with P = -[X]x; T = N P, TL = P^T T
                const double n00 = v[0], n01 = v[1], n02 = v[2], n11 = v[3], n12 = v[4], n22 = v[5], x = v[12], y = v[13], z = v[14];
                const double t00 = -z * n01 + y * n02, t01 = z * n00 - x * n02, t02 = -y * n00 + x * n01;
                const double t10 = -z * n11 + y * n12, t11 = z * n01 - x * n12, t12 = -y * n01 + x * n11;
                const double t20 = -z * n12 + y * n22, t21 = z * n02 - x * n22, t22 = -y * n02 + x * n12;
                d[0] += -z * t10 + y * t20; d[1] += -z * t11 + y * t21; d[2] += -z * t12 + y * t22;      // TL row 0
                d[3] += t00; d[4] += t10; d[5] += t20;                                                    // H(0, 3..5) = T(.,0)
                d[6] += z * t01 - x * t21; d[7] += z * t02 - x * t22;                                     // TL(1,1), TL(1,2)
                d[8] += t01; d[9] += t11; d[10] += t21;
                d[11] += -y * t02 + x * t12;                                                              // TL(2,2)
                d[12] += t02; d[13] += t12; d[14] += t22;
                d[15] += n00; d[16] += n01; d[17] += n02; d[18] += n11; d[19] += n12; d[20] += n22;
                d[21] += PT0(x, y, z, v[6], v[7], v[8]); d[22] += PT1(x, y, z, v[6], v[7], v[8]); d[23] += PT2(x, y, z, v[6], v[7], v[8]);   // B^T g = [P^T m; m]
                d[24] += v[6]; d[25] += v[7]; d[26] += v[8];
                d[27] -= PT0(x, y, z, v[9], v[10], v[11]); d[28] -= PT1(x, y, z, v[9], v[10], v[11]); d[29] -= PT2(x, y, z, v[9], v[10], v[11]);   // -Z L^-1 b_l = -[P^T q; q]
                d[30] -= v[9]; d[31] -= v[10]; d[32] -= v[11];
            }
#pragma unroll
            for (int i = 0; i < 33; i++) red[i * PK_RED_LD + lane] = d[i];
            __syncwarp();
#pragma unroll 8
            for (int l = 0; l < 32; l++) sd0 += red[lane * PK_RED_LD + l];
            if (lane == 0) {
#pragma unroll 8
                for (int l = 0; l < 32; l++) sd1 += red[32 * PK_RED_LD + l];
            }
            __syncwarp();
        }
        // ---- blocks cut into several items: partials go to memory, the LAST finisher adds them up in item order, so the
        // reduced system does not depend on which warp ran when (one total per block reaches S)
        if (I.nit > 1) {
            double *mine = P.part + (size_t)it * PK_PART;
            __stcg(mine + lane, sz0);
            if (lane < 4) __stcg(mine + 32 + lane, sz1);
            if (diag) { __stcg(mine + 36 + lane, sd0); if (lane == 0) __stcg(mine + 68, sd1); }
            __threadfence();
            __syncwarp();
            unsigned prev = 0;
            if (lane == 0) prev = atomicAdd(P.blk_done + (size_t)I.a * P.bw1 + (I.b - I.a), 1u);
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if ((prev + 1u) % (unsigned)I.nit != 0u) continue;                       // somebody else finishes the block
            __threadfence();
            sz0 = sz1 = sd0 = sd1 = 0.0;
            for (int k = 0; k < I.nit; k++) {
                const double *pk = P.part + (size_t)(I.first + k) * PK_PART;
                sz0 += __ldcg(pk + lane);
                if (lane < 4) sz1 += __ldcg(pk + 32 + lane);
                if (diag) { sd0 += __ldcg(pk + 36 + lane); if (lane == 0) sd1 += __ldcg(pk + 68); }
            }
        }
        if (P.hpp_diag) {                                                            // computeLambdaInit: diagonal of Hpp only
            int r = 0, t = lane;
            while (r < 6 && t >= 6 - r) { t -= 6 - r; r++; }
            if (lane < 21 && t == 0) P.hpp_diag[6 * I.a + r] = sd0;
            continue;
        }
        // ---- the block's total: at most two adds per element of S onto the zeroed system (exact in any order)
        auto elem = [&](int r, int c) -> double * {                                  // element (r, c) of block (a, b), r <= c on the diagonal block
            const int R = 6 * I.a + r, C = 6 * I.b + c;
            if (P.S2 && C >= P.n1) return P.S2 + (size_t)(P.n_tot - 1 - C) * P.ld + (P.n_tot - 1 - R);
            return P.S + (size_t)R * P.ld + C;
        };
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int o = lane + 32 * h;
            const double s = h ? sz1 : sz0;
            if (o < 36) {
                const int r = o / 6, c = o - 6 * r;
                if ((!diag || c >= r) && s != 0.0) atomicAdd(elem(r, c), -s);
            }
        }
        if (diag) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int o = lane + 32 * h;
                const double s = h ? sd1 : sd0;
                if (o < 33 && s != 0.0) {
                    if (o < 21) {
                        int r = 0, t = o;
                        while (t >= 6 - r) { t -= 6 - r; r++; }
                        atomicAdd(elem(r, r + t), s);
                    } else if (o < 27) atomicAdd(P.bp + 6 * I.a + (o - 21), s);
                    else atomicAdd(P.bs + 6 * I.a + (o - 27), s);
                }
            }
        }
        if (P.row_done) {
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(P.row_done + I.a, (unsigned)I.nit);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// update_z_kernel: update_packed_kernel without the second linearisation. With Hll + lambda I = L L^T and Z = W L^-T from the
// stage, the back-substitution x_l = (Hll + lambda I)^-1 (b_l - W^T x_p) (block_solver.hpp:459-483) is
// x_l = L^-T (L^-1 b_l - sum_obs Z^T x_p): one 160-byte record and 18 FMA per observation, three sums per landmark. Then the
// same evaluation of every edge at the trial state as update_packed_kernel (computeActiveErrors + activeRobustChi2) and the
// landmark part of computeScale.
__global__ void __launch_bounds__(ST_THREADS, UZ_MINB) update_z_kernel(BaDev D, const double *__restrict__ pt_in, UpdateOut O, UpdateTasks K,
                                                                 const double *__restrict__ Zr, const double *__restrict__ Lm) {
    __shared__ double s_chi[ST_WARPS], s_sc[ST_WARPS];
    const double *pt = pt_in, *pose_unused = nullptr;
    if (!update_resolve(O, pose_unused, pt)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double chi_acc = 0.0, sc_acc = 0.0;
    for (int t = blockIdx.x * ST_WARPS + warp; t < K.n_tasks; t += gridDim.x * ST_WARPS) {
        const int2 tk = K.tasks[t];
        const int e_first = __ldg(D.lm_ptr + tk.x), nobs = __ldg(D.lm_ptr + tk.y) - e_first;
        const bool in = lane < nobs;
        const int e = e_first + lane;
        const int j = in ? __ldg(D.o_point + e) : -1;
        int head = lane, segl = 1;
        if (in) {
            const int p0 = __ldg(D.lm_ptr + j);
            head = lane - (e - p0); segl = __ldg(D.lm_ptr + j + 1) - p0;
        }
        const int seg_last = head + segl - 1;
        const uint32_t m = in ? D.o_meta[e] : META_LEVEL1;
        const bool valid = !(m & META_LEVEL1);
        const int ip = in ? __ldg(D.o_pose + e) : 0;
        const int hx = valid ? __ldg(D.pose_hidx + ip) : -1;
        double r[3] = {0.0, 0.0, 0.0};
        if (hx >= 0) {                                         // Z^T x_p = Y^T (P x_rot + x_trans), P = -[X_l]x
            double z[ZR_STRIDE];
#pragma unroll
            for (int q = 0; q < ZR_STRIDE / 4; q++) ldg256(Zr + ZR_STRIDE * (size_t)e + 4 * q, z + 4 * q);
            const double *x = O.xp + 6 * hx;
            const double x0 = __ldg(x), x1 = __ldg(x + 1), x2 = __ldg(x + 2);
            const double X = z[9], Y = z[10], Zc = z[11];
            const double v0 = Zc * x1 - Y * x2 + __ldg(x + 3);
            const double v1 = -Zc * x0 + X * x2 + __ldg(x + 4);
            const double v2 = Y * x0 - X * x1 + __ldg(x + 5);
            r[0] = z[0] * v0 + z[3] * v1 + z[6] * v2;
            r[1] = z[1] * v0 + z[4] * v1 + z[7] * v2;
            r[2] = z[2] * v0 + z[5] * v1 + z[8] * v2;
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const bool take = in && lane + off <= seg_last;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double o = __shfl_down_sync(0xffffffffu, r[i], off);
                if (take) r[i] += o;
            }
        }
#pragma unroll
        for (int i = 0; i < 3; i++) r[i] = __shfl_sync(0xffffffffu, r[i], head);
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        const unsigned segmask = (segl >= 32 ? 0xffffffffu : ((1u << segl) - 1u)) << head;
        const bool any_act = (vmask & segmask) != 0u;
        double nx = 0.0, ny = 0.0, nz = 1.0;
        if (in) {
            const double X = __ldg(pt + 3 * (size_t)j), Y = __ldg(pt + 3 * (size_t)j + 1), Z = __ldg(pt + 3 * (size_t)j + 2);
            nx = X; ny = Y; nz = Z;
            if (any_act) {
                double f[LM_STRIDE];
#pragma unroll
                for (int q = 0; q < LM_STRIDE / 4; q++) ldg256(Lm + LM_STRIDE * (size_t)j + 4 * q, f + 4 * q);
                // f = i00 l10 l20 i11 | l21 i22 c0 c1 | c2 bl0 bl1 bl2 ;  L^T x = c - Z^T x_p
                const double u0 = f[6] - r[0], u1 = f[7] - r[1], u2 = f[8] - r[2];
                const double x2 = u2 * f[5];
                const double x1 = (u1 - f[4] * x2) * f[3];
                const double x0 = (u0 - f[1] * x1 - f[2] * x2) * f[0];
                nx = X + x0; ny = Y + x1; nz = Z + x2;
                if (lane == head) sc_acc += x0 * (O.lambda * x0 + f[9]) + x1 * (O.lambda * x1 + f[10]) + x2 * (O.lambda * x2 + f[11]);
            }
            if (lane == head) { O.pt_trial[3 * (size_t)j] = nx; O.pt_trial[3 * (size_t)j + 1] = ny; O.pt_trial[3 * (size_t)j + 2] = nz; }
        }
        // evaluation at the trial state
        if (in && valid) {
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
