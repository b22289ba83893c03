// Block-Jacobi preconditioned conjugate gradients on the reduced camera system (BAGPU_SOLVER_PCG).
//
// north_star names PCG as the reduced-system solver; the direct band solvers of chol.cuh / chol_parts.cuh are the default
// because the keyframe chains of this path have a narrow envelope (DESIGN.md holds the measured comparison). This is the
// iterative alternative behind the same seam (g2o::LinearSolver::solve, Thirdparty/g2o/g2o/core/linear_solver.h:50-58):
//     (S + lambda I) x = b_p + b_s,   S symmetric, upper band storage (element (R, C), R <= C, at S[R * ld + C]),
// preconditioner M = blockdiag(S_aa + lambda I)^-1 over the 6x6 camera blocks.
// Three kernels per iteration, every dot product reduced in a fixed order (per-CTA partials, then every CTA sums the partials
// in the same order), so the solve is reproducible:
//     pcg_spmv_kernel      q = (S + lambda I) p,                 partials of p.q
//     pcg_update_kernel    alpha = rz / pq; x += alpha p; r -= alpha q; z = M r;   partials of r.z and r.r
//     pcg_dir_kernel       beta = rz' / rz; p = z + beta p
#pragma once
#include <cuda_runtime.h>

#define PCG_THREADS 256
#define PCG_MAXB 1024            // CTAs (= partial sums) per vector kernel

struct PcgArgs {
    const double *S; int n, ld, band; double lambda;
    const double *bp, *bs;
    double *x, *r, *z, *p, *q, *Minv;          // Minv: [n / 6][36]
    double *part;                                // [3][PCG_MAXB] partial sums: p.q | r.z | r.r
    double *scal;                                // [8]: rz, pq, rz_new, rr, bb, alpha, beta, iterations done
    int nblk;                                    // CTAs of the vector kernels
};

__device__ __forceinline__ double pcg_block_sum(double v, double *sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < PCG_THREADS / 32; w++) s += sh[w];
    __syncthreads();
    return s;                                    // valid in thread 0
}
// every CTA adds the same partials in the same order: a deterministic grid-wide sum without a second launch
__device__ __forceinline__ double pcg_total(const double *part, int nblk, double *sh) {
    double v = 0.0;
    for (int i = threadIdx.x; i < nblk; i += PCG_THREADS) v += part[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < PCG_THREADS / 32; w++) s += sh[w];
    __syncthreads();
    return s;                                    // valid in every thread
}

// Minv_a = (S_aa + lambda I)^-1 by Cholesky of the 6x6 block (thread = camera); a non-positive pivot raises the failure flag
__global__ void pcg_prec_kernel(PcgArgs A, int *fail) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (6 * a >= A.n) return;
    double L[6][6], Li[6][6];
    for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++) L[j][i] = A.S[(size_t)(6 * a + i) * A.ld + 6 * a + j] + ((i == j) ? A.lambda : 0.0);
    bool ok = true;
    for (int k = 0; k < 6; k++) {
        double d = L[k][k];
        for (int m = 0; m < k; m++) d -= L[k][m] * L[k][m];
        if (!(d > 0.0)) { ok = false; d = 1.0; }
        const double lk = sqrt(d);
        L[k][k] = lk;
        for (int i = k + 1; i < 6; i++) {
            double v = L[i][k];
            for (int m = 0; m < k; m++) v -= L[i][m] * L[k][m];
            L[i][k] = v / lk;
        }
    }
    for (int c = 0; c < 6; c++)                 // Li = L^-1
        for (int r = 0; r < 6; r++) {
            double v = (r == c) ? 1.0 : 0.0;
            for (int m = c; m < r; m++) v -= L[r][m] * Li[m][c];
            Li[r][c] = (r >= c) ? v / L[r][r] : 0.0;
        }
    for (int i = 0; i < 6; i++)                 // Minv = Li^T Li
        for (int j = 0; j < 6; j++) {
            double v = 0.0;
            for (int m = (i > j ? i : j); m < 6; m++) v += Li[m][i] * Li[m][j];
            A.Minv[(size_t)a * 36 + 6 * i + j] = v;
        }
    if (!ok) atomicOr(fail, 1);
}

// x = 0, r = b, z = M r, p = z; partials of r.z and r.r (= b.b)
__global__ void __launch_bounds__(PCG_THREADS) pcg_init_kernel(PcgArgs A) {
    __shared__ double sh[PCG_THREADS / 32];
    double rz = 0.0, rr = 0.0;
    for (int a = blockIdx.x * PCG_THREADS + threadIdx.x; 6 * a < A.n; a += gridDim.x * PCG_THREADS) {
        double r[6], z[6];
        for (int i = 0; i < 6; i++) { r[i] = A.bp[6 * a + i] + A.bs[6 * a + i]; A.x[6 * a + i] = 0.0; A.r[6 * a + i] = r[i]; }
        for (int i = 0; i < 6; i++) {
            double v = 0.0;
            for (int j = 0; j < 6; j++) v += A.Minv[(size_t)a * 36 + 6 * i + j] * r[j];
            z[i] = v; A.z[6 * a + i] = v; A.p[6 * a + i] = v;
            rz += r[i] * v; rr += r[i] * r[i];
        }
    }
    const double s1 = pcg_block_sum(rz, sh), s2 = pcg_block_sum(rr, sh);
    if (threadIdx.x == 0) { A.part[PCG_MAXB + blockIdx.x] = s1; A.part[2 * PCG_MAXB + blockIdx.x] = s2; }
}
// scal[0] = rz, scal[4] = bb after the init (one CTA)
__global__ void __launch_bounds__(PCG_THREADS) pcg_init_finish_kernel(PcgArgs A) {
    __shared__ double sh[PCG_THREADS / 32];
    const double rz = pcg_total(A.part + PCG_MAXB, A.nblk, sh), rr = pcg_total(A.part + 2 * PCG_MAXB, A.nblk, sh);
    if (threadIdx.x == 0) { A.scal[0] = rz; A.scal[3] = rr; A.scal[4] = rr; A.scal[7] = 0.0; }
}

// q = (S + lambda I) p: warp = row R; upper part S[R][R + k] (coalesced), lower part S[R - k][R]; partials of p.q
__global__ void __launch_bounds__(PCG_THREADS) pcg_spmv_kernel(PcgArgs A) {
    __shared__ double sh[PCG_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double pq = 0.0;
    for (int R = blockIdx.x * (PCG_THREADS / 32) + warp; R < A.n; R += gridDim.x * (PCG_THREADS / 32)) {
        double s = 0.0;
        const double *row = A.S + (size_t)R * A.ld + R;
        const int ku = min(A.band, A.n - 1 - R), kl = min(A.band, R);
        for (int k = lane; k <= ku; k += 32) s += __ldg(row + k) * A.p[R + k];
        for (int k = 1 + lane; k <= kl; k += 32) s += __ldg(A.S + (size_t)(R - k) * A.ld + R) * A.p[R - k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) { const double pr = A.p[R]; s += A.lambda * pr; A.q[R] = s; pq += pr * s; }
    }
    const double tot = pcg_block_sum(pq, sh);
    if (threadIdx.x == 0) A.part[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(PCG_THREADS) pcg_update_kernel(PcgArgs A, int nblk_spmv) {
    __shared__ double sh[PCG_THREADS / 32];
    const double pq = pcg_total(A.part, nblk_spmv, sh);
    const double rz = A.scal[0];
    const double alpha = (pq > 0.0) ? rz / pq : 0.0;
    double rzn = 0.0, rr = 0.0;
    for (int a = blockIdx.x * PCG_THREADS + threadIdx.x; 6 * a < A.n; a += gridDim.x * PCG_THREADS) {
        double r[6];
        for (int i = 0; i < 6; i++) {
            A.x[6 * a + i] += alpha * A.p[6 * a + i];
            r[i] = A.r[6 * a + i] - alpha * A.q[6 * a + i];
            A.r[6 * a + i] = r[i];
        }
        for (int i = 0; i < 6; i++) {
            double v = 0.0;
            for (int j = 0; j < 6; j++) v += A.Minv[(size_t)a * 36 + 6 * i + j] * r[j];
            A.z[6 * a + i] = v;
            rzn += r[i] * v; rr += r[i] * r[i];
        }
    }
    const double s1 = pcg_block_sum(rzn, sh), s2 = pcg_block_sum(rr, sh);
    if (threadIdx.x == 0) {
        A.part[PCG_MAXB + blockIdx.x] = s1; A.part[2 * PCG_MAXB + blockIdx.x] = s2;
        if (blockIdx.x == 0) { A.scal[1] = pq; A.scal[5] = alpha; }
    }
}

__global__ void __launch_bounds__(PCG_THREADS) pcg_dir_kernel(PcgArgs A) {
    __shared__ double sh[PCG_THREADS / 32];
    const double rzn = pcg_total(A.part + PCG_MAXB, A.nblk, sh), rr = pcg_total(A.part + 2 * PCG_MAXB, A.nblk, sh);
    const double rz = A.scal[0];
    const double beta = (rz != 0.0) ? rzn / rz : 0.0;
    for (int i = blockIdx.x * PCG_THREADS + threadIdx.x; i < A.n; i += gridDim.x * PCG_THREADS) A.p[i] = A.z[i] + beta * A.p[i];
    // the scalars change hands only after every CTA has read them: the LAST CTA to arrive publishes (scal[6] counts arrivals)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(reinterpret_cast<unsigned *>(A.scal + 6), 1u);
        if (prev + 1 == gridDim.x) {
            *reinterpret_cast<unsigned *>(A.scal + 6) = 0u;
            A.scal[0] = rzn; A.scal[3] = rr; A.scal[2] = beta; A.scal[7] += 1.0;
        }
    }
}
