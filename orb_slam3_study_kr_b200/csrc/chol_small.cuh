// Reduced camera system solve for SMALL systems (local BA: 20-40 free keyframes), one CTA, everything in shared memory (sm_100a).
//
// chol_band_kernel (chol.cuh) pays 13-16 us per 32-column panel for its cluster pipeline: 66 us for the n = 120 system of BASELINE
// config 1 (4 panels), 40 % of a chained LM trial there. A system of n <= 224 unknowns fits one SM's shared memory as a packed lower
// triangle, and a column-by-column Cholesky over 1024 threads needs ONE barrier per column:
//   * the right-hand side b_p + b_s is appended as row n of the matrix, so the forward substitution L y = b is the factorisation's own
//     trailing update (row n of L is y^T);
//   * column k is used UNSCALED in the rank-1 update (a_ij -= a_ik a_jk / d_k), so no barrier separates "scale the column" from "update
//     the trailing matrix"; the column is scaled by 1 / sqrt(d_k) one step later, when nobody reads it any more;
//   * the backward substitution L^T x = y is column-oriented too (one barrier per unknown).
// Every element is updated by exactly one thread in a fixed order: the solve is reproducible. A non-positive pivot raises the failure
// flag = a rejected LM trial, as in the other solvers. Used by the chained LM loop (no overlap with the accumulation there).
// MEASURED (B200, chained local BA): 130 us per solve for n = 120 and 268 us for n = 180 against 62 / 90 us of chol_band_kernel -- a column
// step costs ~1 us here (the FP64 reciprocal and the packed index arithmetic sit on the chain between two barriers of 1024 threads), so
// the kernel is OPT-IN (BAGPU_SMALL_CHOL=1) and the cluster kernel stays the default. Kept as the evidence of that experiment.
// Same seam: LinearSolverEigen::solve (Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124).
#pragma once
#include "ba_kernels.cuh"

#define CS_THREADS 1024
#define CS_MAX_N 224
inline size_t chol_small_smem(int n) { return sizeof(double) * ((size_t)(n + 1) * (n + 2) / 2 + 2 * (size_t)n); }

// S: upper band storage, element (R, C), R <= C <= R + ld, at S[R * ld + C]; x: [n] out
__global__ void __launch_bounds__(CS_THREADS, 1) chol_small_kernel(const double *__restrict__ S, int n, int ld, double lambda, const double *__restrict__ bp,
                                                                   const double *__restrict__ bs, double *__restrict__ x, int *fail, const LmDev *lm) {
    extern __shared__ double cs_sm[];
    double *L = cs_sm;                                          // packed rows 0 .. n: element (i, j), j <= i, at i (i + 1) / 2 + j; row n = right-hand side
    double *rd = L + (size_t)(n + 1) * (n + 2) / 2;             // [n] 1 / sqrt(pivot)
    double *xs = rd + n;                                        // [n] solution
    if (lm) { if (lm->done) return; lambda = lm->lambda; }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto at = [](int i, int j) { return i * (i + 1) / 2 + j; };
    // lower triangle of S + lambda I: column j of the lower part = row j of the stored upper part (coalesced reads)
    for (int j = warp; j < n; j += CS_THREADS / 32)
        for (int i = j + lane; i < n; i += 32) {
            double v = (i - j <= ld) ? __ldcg(S + (size_t)j * ld + i) : 0.0;
            if (i == j) v += lambda;
            L[at(i, j)] = v;
        }
    for (int j = tid; j < n; j += CS_THREADS) L[at(n, j)] = __ldcg(bp + j) + __ldcg(bs + j);
    __syncthreads();
    const int ti = tid >> 5, tj = tid & 31;
    for (int k = 0; k < n; k++) {
        const double d = L[at(k, k)];                           // the same value in every thread: uniform control flow
        if (!(d > 0.0)) { if (tid == 0) atomicExch(fail, 1); return; }
        const double inv = 1.0 / d;
        if (tid == 0) rd[k] = rsqrt(d);
        if (k > 0) {                                            // column k - 1 is final now: scale it (rows k .. n)
            const double r = rd[k - 1];
            for (int i = k + tid; i <= n; i += CS_THREADS) L[at(i, k - 1)] *= r;
        }
        // a_ij -= a_ik a_jk / d for k < j <= i <= n (j < n): warp = row(s), lane = column(s)
        for (int i = k + 1 + ti; i <= n; i += 32) {
            const double lik = L[at(i, k)] * inv;
            const int jend = min(i, n - 1);
            double *row = L + at(i, 0);
            for (int j = k + 1 + tj; j <= jend; j += 32) row[j] -= lik * L[at(j, k)];
        }
        __syncthreads();
    }
    if (tid == 0) L[at(n, n - 1)] *= rd[n - 1];                 // the last column has one row below its diagonal: y_{n-1}
    __syncthreads();
    // L^T x = y, y = row n; L(k, k) = 1 / rd[k]
    for (int k = n - 1; k >= 0; k--) {
        const double xk = L[at(n, k)] * rd[k];
        if (tid == 0) xs[k] = xk;
        const double *rowk = L + at(k, 0);
        double *y = L + at(n, 0);
        for (int j = tid; j < k; j += CS_THREADS) y[j] -= rowk[j] * xk;
        __syncthreads();
    }
    for (int j = tid; j < n; j += CS_THREADS) x[j] = xs[j];
}
