// Reduced camera system solve  (Hschur + lambda I) x = bschur  by dense FP64 Cholesky, one cooperative launch.
// Replaces LinearSolverEigen::solve / LinearSolverDense::solve (Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124,
// linear_solver_dense.h:64-111) for systems small/dense enough to factor directly.
//
// Storage: the buffer the build kernel scatters into, element (R,C), R<=C at S[R*ld + C] (row-major upper), is read
// here as a column-major LOWER matrix Lm(i,j) = S[j*ld + i], i>=j. Row n of that lower matrix (= "column n" of the
// row-major view) carries the right-hand side, so the forward substitution L y = b falls out of the panel TRSMs.
// Right-looking blocked algorithm, panel width CH_NB, two grid syncs per panel; backward substitution is
// right-looking too (one grid sync per panel).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

#define CH_NB 32
#define CH_RT 64          // rows per TRSM tile (one thread per row)
#define CH_TB 64          // trailing-update tile
#define CH_THREADS 256

struct CholArgs {
    double *S; int n; int ld;
    double lambda;
    const double *bp, *bs;    // bschur = bp + bs
    double *x;                // [n] out
    int *fail;                // set to 1 on a non-positive pivot
};

__global__ void __launch_bounds__(CH_THREADS) chol_solve_kernel(CholArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double Ld[CH_NB][CH_NB + 1];
    __shared__ double Pi[CH_NB][CH_TB + 2];
    __shared__ double Pj[CH_NB][CH_TB + 2];
    __shared__ int s_fail;
    double *S = a.S;
    const int n = a.n, ld = a.ld, nrows = a.n + 1;
    const int tid = threadIdx.x;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + tid, gthreads = (size_t)gridDim.x * blockDim.x;

    for (size_t j = gtid; j < (size_t)n; j += gthreads) {
        S[j * ld + j] += a.lambda;
        S[j * ld + n] = a.bp[j] + a.bs[j];
    }
    grid.sync();

    bool failed = false;
    for (int p0 = 0; p0 < n; p0 += CH_NB) {
        const int nb = min(CH_NB, n - p0);
        // ---- phase A: factor the diagonal block (redundantly per CTA), TRSM the rows below
        for (int i = tid; i < CH_NB * CH_NB; i += CH_THREADS) {
            const int r = i % CH_NB, c = i / CH_NB;
            Ld[r][c] = (r < nb && c < nb && r >= c) ? S[(size_t)(p0 + c) * ld + p0 + r] : 0.0;
        }
        if (tid == 0) s_fail = 0;
        __syncthreads();
        for (int k = 0; k < nb; k++) {
            if (tid == 0) {
                const double d = Ld[k][k];
                if (!(d > 0.0)) s_fail = 1;
                Ld[k][k] = sqrt(d);
            }
            __syncthreads();
            const double dk = Ld[k][k];
            for (int i = k + 1 + tid; i < nb; i += CH_THREADS) Ld[i][k] /= dk;
            __syncthreads();
            // trailing update of the small block: element (i,j), k < j <= i
            for (int idx = tid; idx < CH_NB * CH_NB; idx += CH_THREADS) {
                const int i = idx / CH_NB, j = idx % CH_NB;
                if (j > k && i >= j && i < nb) Ld[i][j] -= Ld[i][k] * Ld[j][k];
            }
            __syncthreads();
        }
        if (s_fail) { failed = true; }
        if (failed) break;        // every CTA factors the same block: uniform exit
        const int rbase = p0 + nb;
        const int nbelow = nrows - rbase;
        const int ntile = (nbelow + CH_RT - 1) / CH_RT;
        for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
            const int r = rbase + t * CH_RT + tid;
            if (tid < CH_RT && r < nrows) {
                double xr[CH_NB];
#pragma unroll
                for (int j = 0; j < CH_NB; j++) xr[j] = (j < nb) ? S[(size_t)(p0 + j) * ld + r] : 0.0;
#pragma unroll
                for (int j = 0; j < CH_NB; j++) {
                    if (j < nb) {
                        double s = xr[j];
#pragma unroll
                        for (int k = 0; k < j; k++) s -= xr[k] * Ld[j][k];
                        xr[j] = s / Ld[j][j];
                    }
                }
#pragma unroll
                for (int j = 0; j < CH_NB; j++) if (j < nb) S[(size_t)(p0 + j) * ld + r] = xr[j];
            }
        }
        grid.sync();
        // the factored diagonal block goes back only now: during phase A other CTAs were still reading the original
        if (blockIdx.x == 0) {
            for (int i = tid; i < CH_NB * CH_NB; i += CH_THREADS) {
                const int r = i % CH_NB, c = i / CH_NB;
                if (r < nb && c < nb && r >= c) S[(size_t)(p0 + c) * ld + p0 + r] = Ld[r][c];
            }
        }
        // ---- phase B: trailing update C(i,j) -= sum_k P(i,k) P(j,k) over the lower triangle (incl. the RHS row)
        if (nbelow > 0) {
            const int nt = (nbelow + CH_TB - 1) / CH_TB;
            const int ntri = nt * (nt + 1) / 2;
            const int ty = tid / 16, tx = tid % 16;
            for (int t = blockIdx.x; t < ntri; t += gridDim.x) {
                int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
                while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
                while (ti * (ti + 1) / 2 > t) ti--;
                const int tj = t - ti * (ti + 1) / 2;
                const int ri = rbase + ti * CH_TB, cj = rbase + tj * CH_TB;
                __syncthreads();
                for (int idx = tid; idx < CH_NB * CH_TB; idx += CH_THREADS) {
                    const int k = idx / CH_TB, i = idx % CH_TB;
                    Pi[k][i] = (k < nb && ri + i < nrows) ? S[(size_t)(p0 + k) * ld + ri + i] : 0.0;
                    Pj[k][i] = (k < nb && cj + i < nrows) ? S[(size_t)(p0 + k) * ld + cj + i] : 0.0;
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
                    if (c >= n) continue;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int r = ri + ty * 4 + u;
                        if (r < nrows && r >= c) S[(size_t)c * ld + r] -= acc[u][v];
                    }
                }
            }
        }
        grid.sync();
    }
    if (failed) {
        if (gtid == 0) *a.fail = 1;
        return;
    }
    // ---- backward substitution L^T x = y, y = row n of the lower matrix
    // y lives in the RHS row S[j*ld + n] (updated in place); final x goes to a.x (written by CTA 0 only).
    const int last = ((n - 1) / CH_NB) * CH_NB;
    double *xs = &Pi[0][0];     // reuse shared memory: x of the current block
    for (int p0 = last; p0 >= 0; p0 -= CH_NB) {
        const int nb = min(CH_NB, n - p0);
        __syncthreads();
        for (int i = tid; i < CH_NB * CH_NB; i += CH_THREADS) {
            const int r = i % CH_NB, c = i / CH_NB;
            Ld[r][c] = (r < nb && c < nb && r >= c) ? S[(size_t)(p0 + c) * ld + p0 + r] : 0.0;
        }
        if (tid < CH_NB) xs[tid] = (tid < nb) ? S[(size_t)(p0 + tid) * ld + n] : 0.0;
        __syncthreads();
        if (tid < 32) {
            // warp 0: x_p = D^{-T} y_p, lane = column; sequential over rows from the bottom
            double v = xs[tid];
            for (int i = nb - 1; i >= 0; i--) {
                const double xi = __shfl_sync(0xffffffffu, v, i) / Ld[i][i];
                if (tid == i) v = xi;
                else if (tid < i) v -= Ld[i][tid] * xi;
            }
            xs[tid] = v;
        }
        __syncthreads();
        if (blockIdx.x == 0 && tid < nb) a.x[p0 + tid] = xs[tid];
        // y_j -= sum_{i in block} Lm(p0+i, j) x_i  for j < p0
        for (size_t j = gtid; j < (size_t)p0; j += gthreads) {
            const double *col = S + j * ld + p0;
            double s = 0.0;
#pragma unroll 8
            for (int i = 0; i < CH_NB; i++) if (i < nb) s += col[i] * xs[i];
            S[j * ld + n] -= s;
        }
        grid.sync();
    }
}
