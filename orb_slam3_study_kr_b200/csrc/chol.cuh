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
    int *fail;                // [2]: fail[0] = 1 on a non-positive pivot; fail[1] = 2 / 3 when a wait for pair_kernel timed out
    long long *prof;          // optional [8] cycle counters (potrf, trsm, sync1, writeback+y, update, sync2, backward)
    // chol_band_kernel running BESIDE pair_kernel (single GPU): block column c is read once every work item of the camera
    // rows it covers has been accumulated. row_done[a] counts finished items of row a, item_off[a * bw1] is the first item
    // of row a in the plan. nullptr: the system is complete at launch.
    const unsigned *row_done = nullptr; const unsigned *item_off = nullptr; int bw1 = 0;
    const int *row_pos = nullptr;   // position of camera row a in the item order (items of row a: item_off[row_pos[a] * bw1 ...])
    // Two-way factorisation (chol_band_kernel): the band is factored from both ends by two clusters that stop at a common
    // separator block M. p_stop > 0: factor panels [0, p_stop) only, write the partially reduced block columns >= p_stop
    // back (without lambda) and return; mirror_n > 0: this system is the mirror image (index j <-> mirror_n - 1 - j of the
    // original ordering: right-hand side, x, camera rows to wait for); back_from > 0: backward substitution only, for panels
    // back_from - 1 .. 0, with x of the rows >= 32 back_from already in y.
    int p_stop = 0, mirror_n = 0, back_from = 0, wait_band = 0;
    // Partitioned factorisation (chol_parts.cuh): this system is the sub-system that starts at global row row_base (the camera
    // rows to wait for are counted from there); no_back = 1: factor every panel and return (the backward substitution runs later,
    // once the separator unknowns are known).
    int row_base = 0, no_back = 0;
    // chained mode: lambda from the device-side LM state (and nothing to do once the loop has ended); ba_kernels.cuh's LmDev
    const struct LmDev *lm = nullptr;
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
    if (a.lm) { if (a.lm->done) return; a.lambda = a.lm->lambda; }      // uniform over the grid: before any grid / cluster barrier

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
// shared memory from the moment the CTA takes it (loaded from the band buffer, + lambda) until it is factored. There is
// no global read-modify-write anywhere. Step c of the pipeline (B_c = hardware cluster barrier "panel c is in L2"):
//   owner(c)      TRSM of the rows below the (already factored) diagonal block, thread = row, registers. Warp 0 holds
//                 the 32 rows that form the NEXT diagonal block: it pushes them straight into owner(c+1)'s shared memory
//                 (DSMEM) and raises a flag. All rows go to the band buffer (the result, and how the others get them);
//                 the forward substitution of y rides along. Arrive/wait B_c, then take block column c+NC.
//   owner(c+1)    warps 0-3: wait for the flag, subtract the pushed rows from their diagonal block and factor it (potrf),
//                 i.e. the next panel's critical path starts before panel c is even complete;
//                 warps 4-15: wait B_c, stage the remaining rows from L2, rank-32 update of the rows under the diagonal.
//   other CTAs    wait B_c, stage the rows that reach their block column, rank-32 update of the whole block.
// NC * 32 >= maxr guarantees a CTA is done with block c before panel c+NC's first update arrives. CTA 0 runs the backward
// substitution at the end with the panels double-buffered into shared memory by cp.async.
#define CB_THREADS 512
#define CB_WARPS (CB_THREADS / 32)
#define CB_PG 128            // threads of the potrf group (warps 0-3)
#define CB_LD 34             // shared-memory row stride (doubles) of a block column: 16-byte aligned rows, 4-bank row skew
#define CB_MAXR 352          // rows (diagonal block included) a panel may reach
#define CB_PAD 4
#define CB_MAXP 512          // panels whose extent is cached in shared memory

__device__ __forceinline__ void bar_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_release_cluster(int *p, int v) { asm volatile("st.release.cluster.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_cluster(const int *p) {
    int v;
    asm volatile("ld.acquire.cluster.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

// 32x32 Cholesky in shared memory by a group of NT threads (whole warps, named barrier 1), 8 columns at a time:
//  * warp 0 ("chain"): lane = row, 8 pivots; pivot and column entries travel by shuffle (FP64 instruction issue, not
//    shuffle latency, is the scarce resource: ~5 cycles per DP instruction per warp);
//  * right after, the whole group updates only the NEXT 8 columns (what the next chain needs);
//  * the update of the columns further right is deferred: warps 1.. apply it while warp 0 already runs the next chain.
// tri[e] = (row << 8 | col) enumerates a lower triangle row by row (a smaller triangle is a prefix). Returns ok in warp 0.
template <int NT>
__device__ __forceinline__ bool group_potrf32(double (*Ld)[CH_NB + 1], double *invd, const unsigned short *tri, int t, long long *chain_cyc) {
    const int lane = t & 31, warp = t >> 5;
    bool ok = true;
    for (int kb = 0; kb < CH_NB; kb += 8) {
        const long long tc0 = chain_cyc ? clock64() : 0;
        if (warp == 0) {
            // lane = row (all 32 rows: the rows below the 8x8 block are solved by the same instruction stream); the pivot
            // and the column entries travel by shuffle. DP instructions are what costs here, so nothing is replicated.
            double a8[8];
#pragma unroll
            for (int q = 0; q < 8; q++) a8[q] = Ld[lane][kb + q];
            double myrs = 1.0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double dkk = __shfl_sync(0xffffffffu, a8[k], kb + k);
                if (!(dkk > 0.0)) ok = false;
                const double rs = rsqrt(dkk);
                if (lane == kb + k) { a8[k] = dkk * rs; myrs = rs; }
                else a8[k] *= rs;
#pragma unroll
                for (int q = k + 1; q < 8; q++) {
                    const double lqk = __shfl_sync(0xffffffffu, a8[k], kb + q);
                    a8[q] -= a8[k] * lqk;
                }
            }
            if (lane >= kb) {
#pragma unroll
                for (int q = 0; q < 8; q++) if (kb + q <= lane) Ld[lane][kb + q] = a8[q];
                if (lane < kb + 8) invd[lane] = myrs;
            }
            if (chain_cyc) *chain_cyc += clock64() - tc0;
        } else if (kb >= 8 && kb + 8 < CH_NB) {
            // deferred: columns >= kb + 8 get the rank-8 update of the PREVIOUS sub-block (kb - 8)
            const int base = kb + 8, m = CH_NB - base, cnt = m * (m + 1) / 2, pk = kb - 8;
            for (int e = t - 32; e < cnt; e += NT - 32) {
                const int rc = tri[e], i = base + (rc >> 8), jj = base + (rc & 255);
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int k = 0; k < 8; k += 2) { s0 += Ld[i][pk + k] * Ld[jj][pk + k]; s1 += Ld[i][pk + k + 1] * Ld[jj][pk + k + 1]; }
                Ld[i][jj] -= s0 + s1;
            }
        }
        bar_named(1, NT);
        if (kb + 8 < CH_NB) {
            // the next 8 columns (rows >= column) get this sub-block's update now
            const int base = kb + 8, cnt = (CH_NB - base) * 8;
            for (int e = t; e < cnt; e += NT) {
                const int i = base + (e >> 3), jj = base + (e & 7);
                if (i >= jj) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int k = 0; k < 8; k += 2) { s0 += Ld[i][kb + k] * Ld[jj][kb + k]; s1 += Ld[i][kb + k + 1] * Ld[jj][kb + k + 1]; }
                    Ld[i][jj] -= s0 + s1;
                }
            }
            bar_named(1, NT);
        }
    }
    return ok;
}

// Cb[i][j] -= sum_k Lr[i][k] Lc[j][k] for rows row_lo <= i < nr, all 32 columns j, by NT threads (t = 0..NT-1).
// Thread tile: U rows rg + (NT/8) u, 4 columns cg + 8 v, so neighbouring lanes touch neighbouring rows (conflict-free
// LDS.128).
template <int NT, int U>
__device__ __forceinline__ void rank32_update(double *Cb, const double *Lr, const double *Lc, int row_lo, int nr, int t) {
    constexpr int RP = NT / 8;
    const int cgc = t & 7, rg = t >> 3;
    for (int rb = row_lo; rb + rg < nr; rb += U * RP) {
        double acc[U][4];
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int v = 0; v < 4; v++) acc[u][v] = 0.0;
        int ri[U];
#pragma unroll
        for (int u = 0; u < U; u++) ri[u] = min(rb + rg + RP * u, nr - 1);
#pragma unroll 4
        for (int k = 0; k < CH_NB; k += 2) {
            double2 av[U], bv[4];
#pragma unroll
            for (int u = 0; u < U; u++) av[u] = *reinterpret_cast<const double2 *>(Lr + ri[u] * CB_LD + k);
#pragma unroll
            for (int v = 0; v < 4; v++) bv[v] = *reinterpret_cast<const double2 *>(Lc + (cgc + 8 * v) * CB_LD + k);
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int v = 0; v < 4; v++) acc[u][v] += av[u].x * bv[v].x + av[u].y * bv[v].y;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (rb + rg + RP * u >= nr) continue;
#pragma unroll
            for (int v = 0; v < 4; v++) Cb[ri[u] * CB_LD + cgc + 8 * v] -= acc[u][v];
        }
    }
}

// table != nullptr: the grid holds one cluster per entry of `table` (the partitions of chol_parts.cuh, all with the same cluster
// size and maxr); cluster q works on table[q] with the launch's lambda.
__global__ void __launch_bounds__(CB_THREADS, 1) chol_band_kernel(CholArgs a, int maxr, int dyn_doubles, const CholArgs *__restrict__ table) {
    cg::cluster_group cl = cg::this_cluster();
    const int NC = (int)cl.num_blocks(), o = (int)cl.block_rank();
    if (a.lm) { if (a.lm->done) return; a.lambda = a.lm->lambda; }      // uniform over the grid: before any cluster barrier
    if (table) {      // the launch carries lambda and whether to wait for pair_kernel's row counters; everything else is per partition
        const double lam = a.lambda; const bool wait = a.row_done != nullptr;
        a = table[blockIdx.x / NC]; a.lambda = lam;
        if (!wait) a.row_done = nullptr;
    }
    extern __shared__ double cb_sm[];
    double *Cb = cb_sm;                                        // own block column [maxr + pad][CB_LD]
    double *Ls = cb_sm + (size_t)(maxr + CB_PAD) * CB_LD;      // staged rows of the panel being applied
    __shared__ double Ld[CH_NB][CH_NB + 1];
    __shared__ __align__(16) double Lt[CH_NB * CB_LD];         // rows of the previous panel that make up my diagonal block
    __shared__ __align__(16) double LdT[CH_NB][CH_NB];        // LdT[m][j] = L11(j, m), j > m: what the TRSM streams
    __shared__ double invd[CH_NB], yc[CH_NB];
    __shared__ unsigned short tri[CH_NB * (CH_NB + 1) / 2];
    __shared__ int s_fail, s_flag;
    __shared__ int s_rend[CB_MAXP];                            // last row of every panel (col_end of its last column)
    double *S = a.S;
    const int n = a.n, ld = a.ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int P = (n + CH_NB - 1) / CH_NB;
    for (int c = tid; c < min(P, CB_MAXP); c += CB_THREADS) s_rend[c] = min(n - 1, a.col_end[min(n, (c + 1) * CH_NB) - 1]);
    auto rend_of = [&](int c) { return c < CB_MAXP ? s_rend[c] : min(n - 1, a.col_end[min(n, (c + 1) * CH_NB) - 1]); };
    for (int e = tid; e < CH_NB * (CH_NB + 1) / 2; e += CB_THREADS) {
        int r = (int)((sqrtf(8.f * e + 1.f) - 1.f) * 0.5f);
        while ((r + 1) * (r + 2) / 2 <= e) r++;
        while (r * (r + 1) / 2 > e) r--;
        tri[e] = (unsigned short)((r << 8) | (e - r * (r + 1) / 2));
    }

    long long pc[24], t0 = 0, t1 = 0;
#pragma unroll
    for (int q = 0; q < 24; q++) pc[q] = 0;
#define BT(i) do { if (a.prof) { t1 = clock64(); pc[i] += t1 - t0; t0 = t1; } } while (0)
    if (tid == 0) { s_fail = 0; s_flag = 0; }
    // a.y (zeroed by the host) accumulates the forward-substitution updates; bp + bs join when a panel is solved

    // block column c of (S + lambda I) -> Cb (rows p0 .. rend, lower part; the strict upper part of the diagonal block = 0)
    auto load_block = [&](int c) {
        const int p0 = c * CH_NB, nb = min(CH_NB, n - p0);
        if (a.row_done) {                                  // wait for the camera rows of this block column (pair_kernel)
            int a0 = (a.row_base + p0) / 6, a1 = (a.row_base + p0 + nb - 1) / 6;
            if (a.mirror_n) {                                  // mirrored system: original columns C_lo..C_hi, rows from C_lo - band up
                const int c_hi = a.mirror_n - 1 - p0, c_lo = a.mirror_n - 1 - (p0 + nb - 1);
                a0 = max(0, c_lo - a.wait_band) / 6; a1 = c_hi / 6;
            }
            for (int cam = a0 + tid; cam <= a1; cam += CB_THREADS) {
                const int pos = a.row_pos ? a.row_pos[cam] : cam;
                const unsigned need = a.item_off[(size_t)(pos + 1) * a.bw1] - a.item_off[(size_t)pos * a.bw1];
                // pair_kernel runs beside this kernel; if it cannot (a profiler serialising kernels, a starved device) give up after
                // 2 s and raise the starvation word fail[1] instead of hanging: the host re-runs the trial with the solve after pair_kernel
                unsigned long long t_start = 0, t_now = 0;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
                while (ld_acquire_gpu(a.row_done + cam) < need) {
                    __nanosleep(100);
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
                    if (t_now - t_start > 2000000000ull) { atomicExch(a.fail + 1, 2 + (a.mirror_n ? 1 : 0)); break; }   // own word: a later pivot failure on the incomplete system must not hide it
                }
            }
            __syncthreads();
        }
        const int R = rend_of(c) - p0 + 1;
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
    // Ld <- diagonal block of Cb (minus Lt Lt^T when the previous panel reaches it), identity beyond nb; potrf; flag a failure
    auto factor_diag = [&](int c, bool sub) {                  // potrf group only, t = tid < CB_PG
        const int nb = min(CH_NB, n - c * CH_NB);
        // Ld <- lower part of (diagonal block - Lt Lt^T), identity beyond nb
        for (int e = tid; e < CH_NB * (CH_NB + 1) / 2; e += CB_PG) {
            const int rc = tri[e], i = rc >> 8, j = rc & 255;
            double v = (i == j) ? 1.0 : 0.0;
            if (i < nb) {
                v = Cb[i * CB_LD + j];
                if (sub) {
                    double s0 = 0.0, s1 = 0.0;
                    const double2 *li = reinterpret_cast<const double2 *>(Lt + i * CB_LD), *lj = reinterpret_cast<const double2 *>(Lt + j * CB_LD);
#pragma unroll 4
                    for (int k = 0; k < CH_NB / 2; k++) { const double2 p = li[k], q = lj[k]; s0 += p.x * q.x; s1 += p.y * q.y; }
                    v -= s0 + s1;
                }
            }
            Ld[i][j] = v;
        }
        bar_named(1, CB_PG);
        BT(5);
        const bool ok = group_potrf32<CB_PG>(Ld, invd, tri, tid, a.prof ? &pc[22] : nullptr);
        BT(6);
        if (warp == 0 && !ok && lane < NC) atomicCAS(cl.map_shared_rank(&s_fail, lane), 0, c + 1);   // the first failure wins
        for (int e = tid; e < CH_NB * CH_NB; e += CB_PG) {
            const int m = e >> 5, j = e & 31;
            LdT[m][j] = (j > m) ? Ld[j][m] : 0.0;
        }
    };

    __syncthreads();                                           // s_rend, tri
    const int Pend = a.p_stop > 0 ? min(a.p_stop, P) : P;      // panels to factor
    auto ridx = [&](int j) { return a.mirror_n ? a.mirror_n - 1 - j : j; };   // index into bp / bs / x (original ordering)
    if (a.back_from == 0) {
    int cur = o;
    if (cur < P) load_block(cur);
    cl.sync();
    if (o == 0 && tid < CB_PG && Pend > 0) factor_diag(0, false);
    __syncthreads();

    if (a.prof) t0 = clock64();
    bool failed = false;
    for (int c = 0; c < Pend; c++) {
        const int p0 = c * CH_NB, nb = min(CH_NB, n - p0);
        const int rend = rend_of(c);
        const int R = rend - p0 + 1;
        const bool owner = (c == cur), nextowner = (c + 1 == cur && cur < P);
        if (owner) {
            BT(20);
            if (warp == CB_WARPS - 1) {
                // forward substitution rides along: y_p = L11^{-1} y_p
                const double di = invd[lane];
                double v = (lane < nb) ? (__ldcg(a.bp + ridx(p0 + lane)) + __ldcg(a.bs + ridx(p0 + lane)) + __ldcg(a.y + p0 + lane)) * di : 0.0;
#pragma unroll
                for (int k = 0; k < CH_NB - 1; k++) {
                    const double yk = __shfl_sync(0xffffffffu, v, k);
                    const double l = (lane > k && lane < nb) ? Ld[lane][k] * di : 0.0;
                    v -= l * yk;
                }
                yc[lane] = (lane < nb) ? v : 0.0;
                if (lane < nb) __stcg(a.y + p0 + lane, v);
            }
            // TRSM  X L11^T = A21 : thread = row, the whole row in registers
            // warp 0 holds the rows of the next diagonal block (the critical path); the others skip warps 4, 8, 12 so that
            // warp 0 has its scheduler to itself
            const int rw = (warp == 0) ? 0 : ((warp & 3) ? warp - (warp >> 2) : 12 + (warp >> 2));   // 0 | 1..9 | 13..15 spare
            const int i = CH_NB + 32 * rw + lane;
            const bool act = (warp != CB_WARPS - 1) && i < R;
            const bool push = (warp == 0) && (c + 1 < P) && (R > CH_NB);
            const int r = 32 * rw + lane;                      // row slot; solved values go to Xs[k][r] (the staging buffer is idle)
            const int XR = (R - CH_NB + 31) & ~31;
            double *Xs = Ls;
            if (act || push) {
                double *rt = push ? cl.map_shared_rank(Lt, (o + 1) % NC) + lane * CB_LD : nullptr;
#pragma unroll 1
                for (int kb = 0; kb < CH_NB; kb += 8) {
                    double xr[8];
#pragma unroll
                    for (int q = 0; q < 8; q += 2) {
                        double2 v = make_double2(0.0, 0.0);
                        if (act) v = *reinterpret_cast<const double2 *>(Cb + i * CB_LD + kb + q);
                        xr[q] = v.x; xr[q + 1] = v.y;
                    }
#pragma unroll 2
                    for (int m = 0; m < kb; m++) {
                        const double xm = Xs[m * XR + r];
                        const double2 *lt = reinterpret_cast<const double2 *>(&LdT[m][kb]);
#pragma unroll
                        for (int q = 0; q < 4; q++) { const double2 l = lt[q]; xr[2 * q] -= xm * l.x; xr[2 * q + 1] -= xm * l.y; }
                    }
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const double xk = xr[k] * invd[kb + k];
                        xr[k] = xk;
#pragma unroll
                        for (int q = k + 1; q < 8; q++) xr[q] -= xk * LdT[kb + k][kb + q];
                    }
#pragma unroll
                    for (int q = 0; q < 8; q++) Xs[(kb + q) * XR + r] = xr[q];
                    if (push) {
#pragma unroll
                        for (int q = 0; q < 8; q += 2) *reinterpret_cast<double2 *>(rt + kb + q) = make_double2(xr[q], xr[q + 1]);
                    }
                    if (act) {
#pragma unroll
                        for (int q = 0; q < 8; q++) if (kb + q < nb) __stcg(S + (size_t)(p0 + kb + q) * ld + p0 + i, xr[q]);
                    }
                }
                BT(21);
                if (push) {
                    asm volatile("fence.acq_rel.cluster;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) st_release_cluster(cl.map_shared_rank(&s_flag, (o + 1) % NC), c + 1);
                }
            }
            BT(0);
            // factored diagonal block back to the band buffer
            for (int e = tid; e < CH_NB * CH_NB; e += CB_THREADS) {
                const int i2 = e & 31, j = e >> 5;
                if (i2 < nb && j <= i2) __stcg(S + (size_t)(p0 + j) * ld + p0 + i2, (j == i2) ? invd[j] : Ld[i2][j]);   // diagonal: 1 / L(j,j)
            }
            __syncthreads();                                   // yc complete
            if (act) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
                for (int k = 0; k < CH_NB; k += 2) { s0 += Xs[k * XR + r] * yc[k]; s1 += Xs[(k + 1) * XR + r] * yc[k + 1]; }
                __stcg(a.y + p0 + i, __ldcg(a.y + p0 + i) - (s0 + s1));
            }
            BT(1);
            cluster_arrive();
            cluster_wait();
            BT(2);
            { const int sf = s_fail; if (sf != 0 && sf - 1 <= c) failed = true; }
            cur += NC;
            if (cur < P && !failed) load_block(cur);
            BT(3);
        } else if (nextowner) {
            const int r0 = CH_NB * cur, nr = rend - r0 + 1;    // rows of panel c inside my block column (may be <= 0)
            if (tid < CB_PG) {
                cluster_arrive();
                if (R > CH_NB) {
                    if (lane == 0) while (ld_acquire_cluster(&s_flag) != c + 1) { }
                    __syncwarp();
                }
                BT(4);
                if (c + 1 < Pend) factor_diag(c + 1, R > CH_NB);   // the separator's diagonal block is not factored here
                BT(7);
                cluster_wait();
                BT(8);
            } else {
                cluster_arrive();
                cluster_wait();
                if (nr > CH_NB) {
                    const int t = tid - CB_PG;
                    for (int e = t >> 5; e < CH_NB; e += (CB_THREADS - CB_PG) / 32) {
                        const double *col = S + (size_t)(p0 + e) * ld + r0;
#pragma unroll 4
                        for (int i = CH_NB + lane; i < nr; i += 32) Ls[i * CB_LD + e] = __ldcg(col + i);
                    }
                    bar_named(2, CB_THREADS - CB_PG);
                    rank32_update<CB_THREADS - CB_PG, 4>(Cb, Ls, Lt, CH_NB, nr, t);
                }
            }
            { const int sf = s_fail; if (sf != 0 && sf - 1 <= c) failed = true; }
        } else {
            cluster_arrive();
            cluster_wait();
            BT(10);
            { const int sf = s_fail; if (sf != 0 && sf - 1 <= c) failed = true; }
            if (cur < P && CH_NB * cur <= rend && !failed) {
                // rows 32 cur .. rend of panel c -> Ls[i][k]; the first 32 staged rows are also the column operand
                const int r0 = CH_NB * cur, nr = rend - r0 + 1;
                for (int k = warp; k < CH_NB; k += CB_WARPS) {
                    const double *col = S + (size_t)(p0 + k) * ld + r0;
#pragma unroll 4
                    for (int i = lane; i < nr; i += 32) Ls[i * CB_LD + k] = __ldcg(col + i);
                }
                __syncthreads();
                BT(11);
                rank32_update<CB_THREADS, 4>(Cb, Ls, Ls, 0, nr, tid);
                BT(12);
            }
        }
        __syncthreads();
        BT(owner ? 13 : (nextowner ? 9 : 14));
        if (failed) break;
    }
    cl.sync();                                                 // nobody leaves while a remote shared-memory write may be in flight
    if (failed) {
        if (o == 0 && tid == 0) { *a.fail = 1; if (a.prof) a.prof[23] = s_fail; }
        return;
    }
    if (a.p_stop > 0) {
        // two-way: the block columns of the separator that were loaded (and reduced by the panels factored here) go back to
        // the band buffer WITHOUT lambda; the merge kernel adds the two partial Schur complements
        if (cur < P) {
            const int p0 = cur * CH_NB, nb = min(CH_NB, n - p0);
            const int R = rend_of(cur) - p0 + 1;
            if (cur == Pend && Pend > 0 && rend_of(Pend - 1) >= p0) {
                // the diagonal block still lacks the update of the last panel (its rows were pushed into Lt)
                for (int e = tid; e < CH_NB * (CH_NB + 1) / 2; e += CB_THREADS) {
                    const int rc = tri[e], i = rc >> 8, j = rc & 255;
                    double s0 = 0.0;
                    for (int k = 0; k < CH_NB; k++) s0 += Lt[i * CB_LD + k] * Lt[j * CB_LD + k];
                    Cb[i * CB_LD + j] -= s0;
                }
                __syncthreads();
            }
            for (int j = warp; j < nb; j += CB_WARPS) {
                double *col = S + (size_t)(p0 + j) * ld + p0;
                for (int i = j + lane; i < R; i += 32) __stcg(col + i, Cb[i * CB_LD + j] - ((i == j) ? a.lambda : 0.0));
            }
        }
        return;
    }
    if (o != 0 || a.no_back) return;
    }   // a.back_from == 0
    else if (o != 0) return;
    BT(15);
    // ---- backward substitution L^T x = y (CTA 0); the band buffer holds 1 / L(j,j) on the diagonal
    const int bs = (maxr + 1) | 1;                             // odd column stride of a prefetched panel
    const bool ysm = n <= dyn_doubles;                         // right-hand side in shared memory
    const bool pref = n + 2 * CH_NB * bs <= dyn_doubles;       // ... and two panel buffers behind it
    double *yv = ysm ? cb_sm : a.y;
    double *pb = cb_sm + n;
    __syncthreads();
    const int cstart = a.back_from > 0 ? min(a.back_from, P) - 1 : P - 1;
    if (ysm) for (int j = tid; j < n; j += CB_THREADS) yv[j] = __ldcg(a.y + j);
    double *part = yc;
    auto prefetch = [&](int c, double *dst) {                  // columns of panel c, rows p0 .. rend -> dst[j * bs + i]
        const int p0 = c * CH_NB, nb = min(CH_NB, n - p0);
        const int R = rend_of(c) - p0 + 1;
        for (int j = warp; j < nb; j += CB_WARPS) {
            const double *col = S + (size_t)(p0 + j) * ld + p0;
            for (int i = j + lane; i < R; i += 32) cp_async8(dst + j * bs + i, col + i);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const bool pref3 = cstart >= 0 && n + 3 * CH_NB * bs <= dyn_doubles;   // three panel buffers: software-pipelined variant
    if (pref3) {
        // While warp 0 runs the triangular chain of panel c, warps 1..15 form the dot products of panel c-1 with everything
        // that is already known (the rows from its second block down) and prefetch panel c-2; the chain of panel c-1 then only
        // adds the 32 rows panel c has just produced. One barrier per panel.
        __shared__ double partA[2][CH_NB];
        auto bufof = [&](int c) { return pb + (c % 3) * (CH_NB * bs); };
        auto prefetch_w = [&](int c) {
            const int p0 = c * CH_NB, nb = min(CH_NB, n - p0);
            const int R = rend_of(c) - p0 + 1;
            double *dst = bufof(c);
            for (int j = warp - 1; j < nb; j += CB_WARPS - 1) {
                const double *col = S + (size_t)(p0 + j) * ld + p0;
                for (int i = j + lane; i < R; i += 32) cp_async8(dst + j * bs + i, col + i);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto old_dots = [&](int c) {
            const int p0 = c * CH_NB, nb = min(CH_NB, n - p0), rbase = p0 + nb, rend = rend_of(c);
            const double *cb = bufof(c);
            for (int j = warp - 1; j < CH_NB; j += CB_WARPS - 1) {
                double sacc = 0.0;
                if (j < nb) {
                    const double *col = cb + j * bs - p0;
#pragma unroll 4
                    for (int i = rbase + CH_NB + lane; i <= rend; i += 32) sacc += col[i] * yv[i];
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, off);
                if (lane == 0) partA[c & 1][j] = sacc;
            }
        };
        if (warp > 0) {
            prefetch_w(cstart);
            if (cstart > 0) prefetch_w(cstart - 1);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                                       // yv loaded, the first two panels landed
        if (warp > 0) old_dots(cstart);
        if (a.prof) { t1 = clock64(); pc[16] += t1 - t0; t0 = t1; }
        for (int c = cstart; c >= 0; c--) {
            if (warp > 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                                   // x of panel c+1 in yv, partA of panel c complete, panel c-1 landed
            if (warp == 0) {
                const int p0 = c * CH_NB, nb = min(CH_NB, n - p0), rbase = p0 + nb, rend = rend_of(c);
                const double *cb = bufof(c);
                const double dv = (lane < nb) ? cb[lane * bs + lane] : 1.0;
                const int nnew = max(0, min(CH_NB, rend - rbase + 1));
                double nd0 = 0.0, nd1 = 0.0;
                if (lane < nb) {
                    const double *col = cb + lane * bs + nb;
                    int i = 0;
                    for (; i + 1 < nnew; i += 2) { nd0 += col[i] * yv[rbase + i]; nd1 += col[i + 1] * yv[rbase + i + 1]; }
                    if (i < nnew) nd0 += col[i] * yv[rbase + i];
                }
                const double *lp = cb + lane * bs;
                double v = (lane < nb) ? (yv[p0 + lane] - partA[c & 1][lane] - (nd0 + nd1)) * dv : 0.0;
#pragma unroll
                for (int i = CH_NB - 1; i >= 1; i--) {
                    const double xi = __shfl_sync(0xffffffffu, v, i);
                    const double l = (lane < i && i < nb) ? lp[i] * dv : 0.0;
                    v -= l * xi;
                }
                if (lane < nb) yv[p0 + lane] = v;
            } else {
                if (c >= 2) prefetch_w(c - 2);
                if (c >= 1) old_dots(c - 1);
            }
        }
        if (a.prof) { t1 = clock64(); pc[19] += t1 - t0; t0 = t1; }
    } else {
    if (pref && cstart >= 0) prefetch(cstart, pb);
    for (int c = cstart; c >= 0; c--) {
        const int p0 = c * CH_NB, nb = min(CH_NB, n - p0);
        const int rbase = p0 + nb;
        const int rend = rend_of(c);
        double *cbuf = pb + ((cstart - c) & 1) * (CH_NB * bs);
        double dv = 1.0;
        if (!pref && warp == 0 && lane < nb) dv = __ldcg(S + (size_t)(p0 + lane) * ld + p0 + lane);
        __syncthreads();                                       // previous panel's x is in yv; its buffer is free
        if (a.prof) { t1 = clock64(); pc[16] += t1 - t0; t0 = t1; }
        if (pref) {
            if (c > 0) { prefetch(c - 1, pb + ((cstart + 1 - c) & 1) * (CH_NB * bs)); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        } else if (warp == CB_WARPS - 1) {
            for (int j = 0; j < CH_NB; j++)
                Ld[lane][j] = (lane < nb && j < nb && j <= lane) ? __ldcg(S + (size_t)(p0 + j) * ld + p0 + lane) : 0.0;
        }
        if (a.prof) { t1 = clock64(); pc[17] += t1 - t0; t0 = t1; }
        if (pref && warp == 0 && lane < nb) dv = cbuf[lane * bs + lane];
#pragma unroll
        for (int q = 0; q < CH_NB / CB_WARPS; q++) {
            const int j = warp + q * CB_WARPS;
            double sacc = 0.0;
            if (j < nb) {
                if (pref) {
                    const double *col = cbuf + j * bs - p0;
#pragma unroll 4
                    for (int i = rbase + lane; i <= rend; i += 32) sacc += col[i] * yv[i];
                } else {
                    const double *col = S + (size_t)(p0 + j) * ld;
#pragma unroll 4
                    for (int i = rbase + lane; i <= rend; i += 32) sacc += __ldcg(col + i) * yv[i];
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, off);
            if (lane == 0) part[j] = sacc;
        }
        __syncthreads();
        if (a.prof) { t1 = clock64(); pc[18] += t1 - t0; t0 = t1; }
        if (warp == 0) {
            // x_p = L11^{-T} v, lane = column. With column `lane` pre-scaled by 1 / L(lane,lane) the chain per row is one
            // shuffle and one fma: x_i = v'_i - sum_{k>i} L'(k,i) x_k
            const double *lp = pref ? cbuf + lane * bs : &Ld[0][lane];
            const int lstep = pref ? 1 : CH_NB + 1;
            double v = (lane < nb) ? (yv[p0 + lane] - part[lane]) * dv : 0.0;
#pragma unroll
            for (int i = CH_NB - 1; i >= 1; i--) {
                const double xi = __shfl_sync(0xffffffffu, v, i);
                const double l = (lane < i && i < nb) ? lp[i * lstep] * dv : 0.0;
                v -= l * xi;
            }
            if (lane < nb) yv[p0 + lane] = v;
        }
        if (a.prof) { t1 = clock64(); pc[19] += t1 - t0; t0 = t1; }
    }
    }   // !pref3
    __syncthreads();
    {   // two-way: only the rows solved here are written (the separator's x is already in place), at their original index
        const int nx = a.back_from > 0 ? min(n, CH_NB * a.back_from) : n;
        for (int j = tid; j < nx; j += CB_THREADS) a.x[ridx(j)] = yv[j];
    }
    if (a.prof && tid == 0) for (int i = 0; i < 24; i++) a.prof[i] = pc[i];
}
