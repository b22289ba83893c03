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

__global__ void __launch_bounds__(CH_THREADS) chol_solve_kernel(CholArgs a) {
    cg::grid_group grid = cg::this_grid();
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
    grid.sync();

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
        grid.sync();
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
        grid.sync();
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
