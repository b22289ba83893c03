// Reduced camera system solve  (Hschur + lambda I) x = bschur  by FP64 Cholesky inside the ENVELOPE of Hschur,
// one cooperative launch. Replaces LinearSolverEigen::solve / LinearSolverDense::solve
// (Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124, linear_solver_dense.h:64-111).
//
// Storage: the buffer the build kernel scatters into, element (R,C), R<=C at S[R*ld + C] (row-major upper), is read
// here as a column-major LOWER matrix Lm(i,j) = S[j*ld + i], i>=j. With ld = n this is a dense matrix; with
// ld = band-1 it is LAPACK-style band storage (only entries with i-j < band exist) -- same indexing, so neither
// the build kernel nor this one cares which it is.
//
// Structure: covisibility in a SLAM map is local, so Hschur has a profile: col_end[j] is the last row that can be
// nonzero in column j (monotone, so it bounds the fill of L too). A panel of CH_NB columns only touches rows up to
// col_end of its last column: the TRSM and the trailing update skip everything below, which turns the O(n^3) dense
// factorisation into O(n * envelope^2).
//
// Right-looking blocked algorithm. Per panel: every CTA factors the 32x32 diagonal block redundantly with ONE warp
// (8-wide register sub-panels, shuffles, no block barriers), the worker CTAs TRSM their share of the rows below,
// grid sync, workers update their share of the trailing tiles, grid sync. The LAST CTA of the grid owns the right-hand
// side: it keeps it in shared memory, forward-substitutes along the panels while the workers compute, and runs the
// backward substitution at the end. Measured phase costs: profiles/README.md.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

#define CH_NB 32
#define CH_TB 64          // trailing-update tile
#define CH_THREADS 256
#define CH_TR 128         // rows per TRSM pass (one thread per row)
#define CH_MAX_SMEM_N 8192  // right-hand side lives in shared memory up to this many unknowns
static_assert(32 * 128 <= 2 * 32 * (64 + 2), "Xs must fit in the Pi|Pj buffer");

struct CholArgs {
    double *S; int n; int ld;
    double lambda;
    const double *bp, *bs;    // bschur = bp + bs
    const int *col_end;       // [n] last possibly-nonzero row of column j (monotone non-decreasing, >= j)
    double *y;                // [n] scratch (used when n > CH_MAX_SMEM_N)
    double *dinv;             // [n] scratch: 1 / L(j,j)
    double *x;                // [n] out
    int *fail;                // set to 1 on a non-positive pivot
    long long *prof;          // optional [8] cycle counters (potrf, trsm, sync1, writeback+y, update, sync2, backward)
};

// One warp factors the 32x32 SPD block in shared memory (Ld[r][c], lower part), 8 columns at a time:
//   1. the 8x8 diagonal sub-block in registers of lanes 0..7 (lane = row), pivots exchanged by shuffles;
//   2. the rows below it solved against that 8x8 (lane = row, registers);
//   3. rank-8 update of the remaining lower triangle (lane = row, rolled loop over columns).
// The pivot chain is the critical path of the whole solve, so it carries no FP64 division or sqrt: one rsqrt per pivot,
// L(k,k) = d * rsqrt(d), L(i,k) = a(i,k) * rsqrt(d). invd[k] = 1/L(k,k). Rows/columns >= nb hold the identity.
__device__ __forceinline__ bool warp_potrf32_blocked(double (*Ld)[CH_NB + 1], double *invd, int lane) {
    bool ok = true;
    for (int kb = 0; kb < CH_NB; kb += 8) {
        double a[8];
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = Ld[kb + (lane & 7)][kb + j];
        double myrs = 1.0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double dkk = __shfl_sync(0xffffffffu, a[k], k);
            if (!(dkk > 0.0)) ok = false;
            const double rs = rsqrt(dkk);
            if (lane == k) { a[k] = dkk * rs; myrs = rs; }
            else if (lane > k) a[k] *= rs;
#pragma unroll
            for (int j = k + 1; j < 8; j++) {
                const double ljk = __shfl_sync(0xffffffffu, a[k], j);
                if (lane >= j) a[j] -= a[k] * ljk;
            }
        }
        if (lane < 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) if (j <= lane) Ld[kb + lane][kb + j] = a[j];
            invd[kb + lane] = myrs;
        }
        __syncwarp();
        const int nrows = CH_NB - kb - 8;                 // rows below the sub-block
        const int i = kb + 8 + lane;
        double x[8];
        if (lane < nrows) {
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = Ld[i][kb + j];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double xk = x[k] * invd[kb + k];
                x[k] = xk;
#pragma unroll
                for (int j = k + 1; j < 8; j++) x[j] -= xk * Ld[kb + j][kb + k];
            }
#pragma unroll
            for (int j = 0; j < 8; j++) Ld[i][kb + j] = x[j];
        }
        __syncwarp();
        if (lane < nrows) {
            for (int j = kb + 8; j <= i; j++) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 8; k++) s += x[k] * Ld[j][kb + k];
                Ld[i][j] -= s;
            }
        }
        __syncwarp();
    }
    return ok;
}

// CL = false: cooperative launch, grid.sync() between phases (any grid up to the device).
// CL = true : the whole grid is ONE thread-block cluster (<= 16 CTAs) and the phases are separated by the hardware
//             cluster barrier (a few hundred cycles instead of a software grid barrier) -- the narrow-band case, where a
//             panel never has more trailing tiles than a cluster has CTAs.
template <bool CL> struct CholSync {
    __device__ __forceinline__ static void sync() {
        if (CL) cg::this_cluster().sync(); else cg::this_grid().sync();
    }
};

template <bool CL>
__global__ void __launch_bounds__(CH_THREADS) chol_solve_kernel(CholArgs a) {
    extern __shared__ double ysm[];              // [n] right-hand side of the y-CTA (when it fits)
    __shared__ double Ld[CH_NB][CH_NB + 1];      // factored diagonal block, Ld[r][c] = L(r,c)
    __shared__ double invd[CH_NB];
    __shared__ double PP[2 * CH_NB * (CH_TB + 2)];             // phase A: Xs[32][128]; phase B: Pi | Pj
    double (*Xs)[CH_TR] = reinterpret_cast<double (*)[CH_TR]>(PP);
    double (*Pi)[CH_TB + 2] = reinterpret_cast<double (*)[CH_TB + 2]>(PP);
    double (*Pj)[CH_TB + 2] = reinterpret_cast<double (*)[CH_TB + 2]>(PP + CH_NB * (CH_TB + 2));
    __shared__ int s_fail;
    double *S = a.S;
    const int n = a.n, ld = a.ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + tid, gthreads = (size_t)gridDim.x * blockDim.x;
    const bool ycta = blockIdx.x == gridDim.x - 1;               // owns the right-hand side
    const int nwork = max(1, (int)gridDim.x - 1);                // CTAs 0..nwork-1 do TRSM rows and update tiles
    const bool worker = (int)blockIdx.x < nwork;
    double *yv = (n <= CH_MAX_SMEM_N) ? ysm : a.y;

    for (size_t j = gtid; j < (size_t)n; j += gthreads) S[j * ld + j] += a.lambda;
    if (ycta) for (int j = tid; j < n; j += CH_THREADS) yv[j] = a.bp[j] + a.bs[j];
    CholSync<CL>::sync();

    bool failed = false;
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0, t1 = 0;
#define PROF_TICK(i) do { if (a.prof) { t1 = clock64(); pc[i] += t1 - t0; t0 = t1; } } while (0)
    if (a.prof) t0 = clock64();
    for (int p0 = 0; p0 < n; p0 += CH_NB) {
        const int nb = min(CH_NB, n - p0);
        const int rbase = p0 + nb;
        const int rend = min(n - 1, a.col_end[p0 + nb - 1]);     // last row this panel can touch
        const int nbelow = max(0, rend - rbase + 1);
        // ---- phase A: diagonal block (warp 0 of every CTA), then TRSM of the rows below
        if (tid == 0) s_fail = 0;
        __syncthreads();
        if (warp == 0) {
            for (int j = 0; j < CH_NB; j++)
                Ld[lane][j] = (lane < nb && j < nb) ? ((j <= lane) ? S[(size_t)(p0 + j) * ld + p0 + lane] : 0.0) : ((j == lane) ? 1.0 : 0.0);
            __syncwarp();
            const bool ok = warp_potrf32_blocked(Ld, invd, lane);
            if (!ok && lane == 0) s_fail = 1;
        }
        __syncthreads();
        PROF_TICK(0);
        if (s_fail) failed = true;
        if (failed) break;                                        // every CTA factors the same block: uniform exit
        if (ycta && warp == 1) {
            // forward substitution rides along: y_p = L11^{-1} y_p (lane = row; sequential over columns)
            double v = (lane < nb) ? yv[p0 + lane] : 0.0;
            for (int k = 0; k < nb; k++) {
                const double yk = __shfl_sync(0xffffffffu, v, k) * invd[k];
                if (lane == k) v = yk;
                else if (lane > k) v -= Ld[lane][k] * yk;
            }
            if (lane < nb) yv[p0 + lane] = v;
        }
        if (worker)
        for (int r0 = blockIdx.x * CH_TR; r0 < nbelow; r0 += nwork * CH_TR) {
            // X L11^T = A21, one thread per row. Columns are solved in chunks of 8 held in registers; finished chunks sit
            // in shared memory as Xs[k][thread] so the rolled update loop reads them without bank conflicts.
            const int r = rbase + r0 + tid;
            const bool act = tid < CH_TR && r <= rend;
            if (tid < CH_TR) {
                double xin[CH_NB / 8][8];
#pragma unroll
                for (int c = 0; c < CH_NB / 8; c++)
#pragma unroll
                    for (int j = 0; j < 8; j++) xin[c][j] = (act && 8 * c + j < nb) ? S[(size_t)(p0 + 8 * c + j) * ld + r] : 0.0;
#pragma unroll
                for (int c = 0; c < CH_NB / 8; c++) {
                    const int kb = 8 * c;
                    double xr[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) xr[j] = xin[c][j];
                    for (int m = 0; m < kb; m++) {
                        const double xm = Xs[m][tid];
#pragma unroll
                        for (int j = 0; j < 8; j++) xr[j] -= xm * Ld[kb + j][m];
                    }
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const double xk = xr[k] * invd[kb + k];
                        xr[k] = xk;
#pragma unroll
                        for (int j = k + 1; j < 8; j++) xr[j] -= xk * Ld[kb + j][kb + k];
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        Xs[kb + j][tid] = xr[j];
                        if (act && kb + j < nb) S[(size_t)(p0 + kb + j) * ld + r] = xr[j];
                    }
                }
            }
        }
        __syncthreads();
        PROF_TICK(1);
        CholSync<CL>::sync();
        PROF_TICK(2);
        // the factored diagonal block goes back only now: during phase A other CTAs were still reading the original
        if (ycta) {
            for (int i = tid; i < CH_NB * CH_NB; i += CH_THREADS) {
                const int r = i % CH_NB, c = i / CH_NB;
                if (r < nb && c < nb && r >= c) S[(size_t)(p0 + c) * ld + p0 + r] = Ld[r][c];
            }
            if (tid < nb) a.dinv[p0 + tid] = invd[tid];
            // y_below -= L21 y_p
            for (int r = rbase + tid; r <= rend; r += CH_THREADS) {
                double s = 0.0;
#pragma unroll 8
                for (int k = 0; k < CH_NB; k++) if (k < nb) s += S[(size_t)(p0 + k) * ld + r] * yv[p0 + k];
                yv[r] -= s;
            }
        }
        PROF_TICK(3);
        // ---- phase B: trailing update C(i,j) -= sum_k P(i,k) P(j,k) over the lower triangle inside the envelope
        if (worker && nbelow > 0) {
            const int nt = (nbelow + CH_TB - 1) / CH_TB;
            const int ntri = nt * (nt + 1) / 2;
            const int ty = tid / 16, tx = tid % 16;
            for (int t = blockIdx.x; t < ntri; t += nwork) {
                int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
                while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
                while (ti * (ti + 1) / 2 > t) ti--;
                const int tj = t - ti * (ti + 1) / 2;
                const int ri = rbase + ti * CH_TB, cj = rbase + tj * CH_TB;
                __syncthreads();
                for (int idx = tid; idx < CH_NB * CH_TB; idx += CH_THREADS) {
                    const int k = idx / CH_TB, i = idx % CH_TB;
                    Pi[k][i] = (k < nb && ri + i <= rend) ? S[(size_t)(p0 + k) * ld + ri + i] : 0.0;
                    Pj[k][i] = (k < nb && cj + i <= rend) ? S[(size_t)(p0 + k) * ld + cj + i] : 0.0;
                }
                __syncthreads();
                double acc[4][4];
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int v = 0; v < 4; v++) acc[u][v] = 0.0;
#pragma unroll 8
                for (int k = 0; k < CH_NB; k++) {
                    double av[4], bv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { av[u] = Pi[k][ty * 4 + u]; bv[u] = Pj[k][tx * 4 + u]; }
#pragma unroll
                    for (int u = 0; u < 4; u++)
#pragma unroll
                        for (int v = 0; v < 4; v++) acc[u][v] += av[u] * bv[v];
                }
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    const int c = cj + tx * 4 + v;
                    if (c > rend) continue;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int r = ri + ty * 4 + u;
                        if (r <= rend && r >= c) S[(size_t)c * ld + r] -= acc[u][v];
                    }
                }
            }
        }
        __syncthreads();
        PROF_TICK(4);
        CholSync<CL>::sync();
        PROF_TICK(5);
    }
    if (failed) {
        if (gtid == 0) *a.fail = 1;
        return;
    }
    // ---- backward substitution L^T x = y in place (x overwrites y), left-looking, y-CTA alone
    if (!ycta) return;
    double *part = PP;                           // [32] column sums
    const int last = ((n - 1) / CH_NB) * CH_NB;
    for (int p0 = last; p0 >= 0; p0 -= CH_NB) {
        const int nb = min(CH_NB, n - p0);
        const int rbase = p0 + nb;
        const int rend = min(n - 1, a.col_end[p0 + nb - 1]);
        __syncthreads();
        // s_j = sum_{i=rbase..rend} L(i, p0+j) x_i : column-major, so warp w owns columns j = w, w+8, w+16, w+24 and its
        // lanes stride over the rows (coalesced); butterfly-reduce per column.
#pragma unroll
        for (int c = 0; c < CH_NB / (CH_THREADS / 32); c++) {
            const int j = warp + c * (CH_THREADS / 32);
            double sacc = 0.0;
            if (j < nb) {
                const double *col = S + (size_t)(p0 + j) * ld;
                for (int i = rbase + lane; i <= rend; i += 32) sacc += col[i] * yv[i];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (lane == 0) part[j] = sacc;
        }
        if (warp == 0) {
            for (int j = 0; j < CH_NB; j++)
                Ld[lane][j] = (lane < nb && j < nb && j <= lane) ? S[(size_t)(p0 + j) * ld + p0 + lane] : 0.0;
            invd[lane] = (lane < nb) ? a.dinv[p0 + lane] : 1.0;
        }
        __syncthreads();
        if (warp == 0) {
            double v = (lane < nb) ? yv[p0 + lane] - part[lane] : 0.0;
            // x_p = L11^{-T} v : lane = column, sequential over rows from the bottom, multiplications only
            for (int i = nb - 1; i >= 0; i--) {
                const double xi = __shfl_sync(0xffffffffu, v, i) * invd[i];
                if (lane == i) v = xi;
                else if (lane < i) v -= Ld[i][lane] * xi;
            }
            if (lane < nb) yv[p0 + lane] = v;
        }
    }
    __syncthreads();
    for (int j = tid; j < n; j += CH_THREADS) a.x[j] = yv[j];
    PROF_TICK(6);
    if (a.prof && tid == 0) for (int i = 0; i < 8; i++) a.prof[i] = pc[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// Narrow-band variant: ONE thread-block cluster, the active window of the factorisation lives in shared memory.
//
// A bundle-adjustment window has local covisibility, so a panel of 32 columns reaches only R <= maxr rows (a few hundred).
// The matrix is cut into 32-wide block columns; block column c is owned by cluster CTA (c mod NC) and sits in that CTA's
// shared memory from the moment the CTA takes it (loaded from the band buffer, + lambda) until it is factored:
//   owner of c : potrf of the diagonal block (whole CTA), TRSM of the rows below (thread = row, registers), L written back
//                to the band buffer (the final result, also how the other CTAs get it), forward substitution of y rides along;
//   cluster barrier (hardware, a few hundred cycles);
//   every CTA whose block column is reached by panel c stages the rows of L it needs from L2 and applies the rank-32
//   update to its own block in shared memory -- no global read-modify-write anywhere; the owner meanwhile takes block c+NC.
// NC * 32 >= maxr guarantees a CTA is done with block c before panel c+NC's first update arrives. The critical path per
// panel is update(own block) + potrf + TRSM + barrier; the other updates overlap with it. CTA 0 runs the backward
// substitution at the end.
#define CB_THREADS 512
#define CB_WARPS (CB_THREADS / 32)
#define CB_LD 34             // shared-memory row stride (doubles) of a block column: 16-byte aligned rows, 4-bank row skew
#define CB_MAXR 384          // rows (diagonal block included) a panel may reach
#define CB_PAD 4

__device__ __forceinline__ size_t chol_band_smem_doubles(int maxr) { return 2 * (size_t)(maxr + CB_PAD) * CB_LD; }

// 32x32 Cholesky in shared memory by the whole CTA. Per 8 columns: warp 0 factors the 32x8 panel in registers
// (lane = row; the rows below the 8x8 block are solved by the same instruction stream), then all threads apply the rank-8
// update to the remaining lower triangle, one element per thread.
__device__ __forceinline__ bool cta_potrf32(double (*Ld)[CH_NB + 1], double *invd, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    bool ok = true;
    for (int kb = 0; kb < CH_NB; kb += 8) {
        if (warp == 0) {
            double a[8];
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = Ld[lane][kb + j];
            double myrs = 1.0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double dkk = __shfl_sync(0xffffffffu, a[k], kb + k);
                if (!(dkk > 0.0)) ok = false;
                const double rs = rsqrt(dkk);
                if (lane == kb + k) { a[k] = dkk * rs; myrs = rs; }
                else if (lane > kb + k) a[k] *= rs;
#pragma unroll
                for (int j = k + 1; j < 8; j++) {
                    const double ljk = __shfl_sync(0xffffffffu, a[k], kb + j);
                    if (lane >= kb + j) a[j] -= a[k] * ljk;
                }
            }
            if (lane >= kb) {
#pragma unroll
                for (int j = 0; j < 8; j++) if (kb + j <= lane) Ld[lane][kb + j] = a[j];
                if (lane < kb + 8) invd[lane] = myrs;
            }
        }
        __syncthreads();
        for (int e = tid; e < CH_NB * CH_NB; e += CB_THREADS) {
            const int i = e >> 5, j = e & 31;
            if (j >= kb + 8 && i >= j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 8; k++) s += Ld[i][kb + k] * Ld[j][kb + k];
                Ld[i][j] -= s;
            }
        }
        __syncthreads();
    }
    return ok;   // meaningful in warp 0
}

__global__ void __launch_bounds__(CB_THREADS, 1) chol_band_kernel(CholArgs a, int maxr) {
    cg::cluster_group cl = cg::this_cluster();
    const int NC = (int)cl.num_blocks(), o = (int)cl.block_rank();
    extern __shared__ double cb_sm[];
    double *Cb = cb_sm;                                        // own block column [maxr + pad][CB_LD]
    double *Ls = cb_sm + (size_t)(maxr + CB_PAD) * CB_LD;      // staged rows of the panel being applied
    __shared__ double Ld[CH_NB][CH_NB + 1];
    __shared__ double invd[CH_NB], yc[CH_NB];
    __shared__ int s_fail;
    double *S = a.S;
    const int n = a.n, ld = a.ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = (n + CH_NB - 1) / CH_NB;

    if (tid == 0) s_fail = 0;
    for (int j = o * CB_THREADS + tid; j < n; j += NC * CB_THREADS) a.y[j] = a.bp[j] + a.bs[j];

    // block column c of (S + lambda I) -> Cb (rows p0 .. rend, lower part; the strict upper part of the diagonal block = 0)
    auto load_block = [&](int c) {
        const int p0 = c * CH_NB, nb = min(CH_NB, n - p0);
        const int R = min(n - 1, a.col_end[p0 + nb - 1]) - p0 + 1;
        for (int j = warp; j < CH_NB; j += CB_WARPS) {
            const double *col = S + (size_t)(p0 + j) * ld + p0;
#pragma unroll 4
            for (int i = lane; i < R; i += 32) {
                double v = 0.0;
                if (j < nb && i >= j) { v = __ldcg(col + i); if (i == j) v += a.lambda; }
                Cb[i * CB_LD + j] = v;
            }
        }
    };
    int cur = o;
    if (cur < P) load_block(cur);
    cl.sync();

    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t0 = 0, t1 = 0;
    if (a.prof) t0 = clock64();
    bool failed = false;
    for (int c = 0; c < P; c++) {
        const int p0 = c * CH_NB, nb = min(CH_NB, n - p0);
        const int rend = min(n - 1, a.col_end[p0 + nb - 1]);
        const int R = rend - p0 + 1;
        const bool owner = (c == cur);
        if (owner) {
            double yreg = 0.0;
            if (warp == CB_WARPS - 1 && lane < nb) yreg = __ldcg(a.y + p0 + lane);      // in flight during the potrf
            for (int e = tid; e < CH_NB * CH_NB; e += CB_THREADS) {
                const int i = e >> 5, j = e & 31;
                Ld[i][j] = (i < nb && j < nb) ? ((j <= i) ? Cb[i * CB_LD + j] : 0.0) : ((i == j) ? 1.0 : 0.0);
            }
            __syncthreads();
            const bool ok = cta_potrf32(Ld, invd, tid);
            if (warp == 0 && !ok && lane < NC) *cl.map_shared_rank(&s_fail, lane) = 1;
            PROF_TICK(0);
            if (warp == CB_WARPS - 1) {
                // forward substitution rides along: y_p = L11^{-1} y_p
                double v = yreg;
                for (int k = 0; k < nb; k++) {
                    const double yk = __shfl_sync(0xffffffffu, v, k) * invd[k];
                    if (lane == k) v = yk;
                    else if (lane > k) v -= Ld[lane][k] * yk;
                }
                yc[lane] = (lane < nb) ? v : 0.0;
                if (lane < nb) __stcg(a.y + p0 + lane, v);
            }
            // TRSM  X L11^T = A21 : thread = row, the whole row in registers
            const int i = CH_NB + tid;
            const bool act = i < R;
            double x[CH_NB];
            if (act) {
#pragma unroll
                for (int k = 0; k < CH_NB; k += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(Cb + i * CB_LD + k);
                    x[k] = v.x; x[k + 1] = v.y;
                }
#pragma unroll
                for (int k = 0; k < CH_NB; k++) {
                    const double xk = x[k] * invd[k];
                    x[k] = xk;
#pragma unroll
                    for (int j = k + 1; j < CH_NB; j++) x[j] -= xk * Ld[j][k];
                }
#pragma unroll
                for (int k = 0; k < CH_NB; k++) if (k < nb) __stcg(S + (size_t)(p0 + k) * ld + p0 + i, x[k]);
            }
            PROF_TICK(1);
            // factored diagonal block back to the band buffer
            for (int e = tid; e < CH_NB * CH_NB; e += CB_THREADS) {
                const int i2 = e & 31, j = e >> 5;
                if (i2 < nb && j <= i2) __stcg(S + (size_t)(p0 + j) * ld + p0 + i2, Ld[i2][j]);
            }
            if (tid < nb) a.dinv[p0 + tid] = invd[tid];
            __syncthreads();                                   // yc complete
            if (act) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < CH_NB; k++) s += x[k] * yc[k];
                __stcg(a.y + p0 + i, __ldcg(a.y + p0 + i) - s);
            }
            PROF_TICK(2);
        }
        cl.sync();
        PROF_TICK(3);
        if (s_fail) { failed = true; break; }
        if (owner) {
            cur += NC;
            if (cur < P) load_block(cur);
            __syncthreads();
            PROF_TICK(5);
        } else if (cur < P && CH_NB * cur <= rend) {
            // rows 32 cur .. rend of panel c -> Ls[i][k]; the first 32 staged rows are also the column operand
            const int r0 = CH_NB * cur, nr = rend - r0 + 1;
            for (int k = warp; k < CH_NB; k += CB_WARPS) {
                const double *col = S + (size_t)(p0 + k) * ld + r0;
#pragma unroll 4
                for (int i = lane; i < nr; i += 32) Ls[i * CB_LD + k] = __ldcg(col + i);
            }
            __syncthreads();
            // Cb[i][j] -= sum_k Ls[i][k] Ls[j][k]; thread tile: rows rg + 64 u, columns cg + 8 v (conflict-free rows)
            const int cgc = tid & 7, rg = tid >> 3;
            for (int rb = 0; rb < nr; rb += 256) {
                if (rb + rg >= nr) break;
                double acc[4][4];
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int v = 0; v < 4; v++) acc[u][v] = 0.0;
                int ri[4];
#pragma unroll
                for (int u = 0; u < 4; u++) ri[u] = min(rb + rg + 64 * u, nr - 1);
#pragma unroll 4
                for (int k = 0; k < CH_NB; k += 2) {
                    double2 av[4], bv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) av[u] = *reinterpret_cast<const double2 *>(Ls + ri[u] * CB_LD + k);
#pragma unroll
                    for (int v = 0; v < 4; v++) bv[v] = *reinterpret_cast<const double2 *>(Ls + (cgc + 8 * v) * CB_LD + k);
#pragma unroll
                    for (int u = 0; u < 4; u++)
#pragma unroll
                        for (int v = 0; v < 4; v++) acc[u][v] += av[u].x * bv[v].x + av[u].y * bv[v].y;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (rb + rg + 64 * u >= nr) continue;
#pragma unroll
                    for (int v = 0; v < 4; v++) Cb[ri[u] * CB_LD + cgc + 8 * v] -= acc[u][v];
                }
            }
            __syncthreads();
            PROF_TICK(4);
        }
    }
    if (failed) {
        if (o == 0 && tid == 0) *a.fail = 1;
        return;
    }
    if (o != 0) return;
    // ---- backward substitution L^T x = y (CTA 0), right-hand side in shared memory when it fits
    double *yv = cb_sm;
    const bool ysm = (size_t)n <= chol_band_smem_doubles(maxr);
    if (!ysm) yv = a.y;
    __syncthreads();
    if (ysm) for (int j = tid; j < n; j += CB_THREADS) yv[j] = __ldcg(a.y + j);
    double *part = yc;
    for (int c = P - 1; c >= 0; c--) {
        const int p0 = c * CH_NB, nb = min(CH_NB, n - p0);
        const int rbase = p0 + nb;
        const int rend = min(n - 1, a.col_end[p0 + nb - 1]);
        __syncthreads();
        if (warp == CB_WARPS - 1) {
            for (int j = 0; j < CH_NB; j++)
                Ld[lane][j] = (lane < nb && j < nb && j <= lane) ? __ldcg(S + (size_t)(p0 + j) * ld + p0 + lane) : 0.0;
            invd[lane] = (lane < nb) ? __ldcg(a.dinv + p0 + lane) : 1.0;
        }
#pragma unroll
        for (int q = 0; q < CH_NB / CB_WARPS; q++) {
            const int j = warp + q * CB_WARPS;
            double sacc = 0.0;
            if (j < nb) {
                const double *col = S + (size_t)(p0 + j) * ld;
#pragma unroll 4
                for (int i = rbase + lane; i <= rend; i += 32) sacc += __ldcg(col + i) * yv[i];
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, off);
            if (lane == 0) part[j] = sacc;
        }
        __syncthreads();
        if (warp == 0) {
            double v = (lane < nb) ? yv[p0 + lane] - part[lane] : 0.0;
            for (int i = nb - 1; i >= 0; i--) {
                const double xi = __shfl_sync(0xffffffffu, v, i) * invd[i];
                if (lane == i) v = xi;
                else if (lane < i) v -= Ld[i][lane] * xi;
            }
            if (lane < nb) yv[p0 + lane] = v;
        }
    }
    __syncthreads();
    for (int j = tid; j < n; j += CB_THREADS) a.x[j] = yv[j];
    PROF_TICK(6);
    if (a.prof && tid == 0) for (int i = 0; i < 8; i++) a.prof[i] = pc[i];
}
