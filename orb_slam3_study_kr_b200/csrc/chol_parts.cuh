// Partitioned band Cholesky: P factorisation fronts instead of one or two (sm_100a).
//
// The chain of 32-column panels is the serial part of the reduced-camera-system solve (chol.cuh): 12 us per panel, so the
// 5000-keyframe system of BASELINE config 5 (n = 29 994, 938 panels) costs 8.8 ms even when it is factored from both ends.
// Here the band is cut into P interiors I_0 .. I_{P-1} with P-1 separators T_1 .. T_{P-1} between them, each at least as wide
// as the band, so that no interior touches another one:
//
//        I_0 | T_1 | I_1 | T_2 | ... | T_{P-1} | I_{P-1}          (natural order, all boundaries multiples of 96 rows)
//
//  1. ONE launch of chol_band_kernel with a table of P sub-systems (one cluster each): cluster i factors I_i top-down and
//     carries the updates into its BOTTOM separator T_{i+1} like into any trailing column (p_stop), exactly as the two-way
//     solver does: T_{i+1} <- T_{i+1} - G_i G_i^T with G_i = L(T_{i+1}, I_i).
//  2. The TOP separator T_i couples to the first rows of I_i. With the separators ordered last its row block of L is the
//     "spike" V_i^T, V_i = L_i^{-1} A(I_i, T_i) (m_i x w, dense): spike_forward_kernel, a forward substitution with w right-hand
//     sides. The columns are independent, so every CTA takes 8 of them and walks the panels alone (no barriers between CTAs);
//     the 32x32 diagonal blocks are inverted beforehand (panel_inverse_kernel) so a panel step is two small products.
//  3. spike_gram_kernel: the Schur complement the interior leaves on its separators,
//         T_i <- T_i - V_i^T V_i,      E_i = -V_i^T G_i^T  (block T_i x T_{i+1}),      rhs(T_i) <- rhs(T_i) - V_i^T y_i.
//  4. sep_assemble_kernel gathers the block-tridiagonal separator system ((P-1) w rows), the production solver factors it
//     (one cluster / cooperative grid), sep_scatter_kernel puts x_T in place and spike_apply_kernel corrects the interiors'
//     right-hand sides, y_i <- y_i - V_i x_{T_i}.
//  5. ONE launch of chol_band_kernel (P single-CTA clusters, back_from) runs the P backward substitutions side by side.
//
// Every sum is formed in a fixed order (no atomics), so the solve stays bit-for-bit reproducible.
// Replaces LinearSolverEigen::solve (Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124) for long keyframe chains.
#pragma once
#include "chol.cuh"

#define PS_MAX_PARTS 32
#define PS_NCOL 8                 // right-hand sides per CTA of spike_forward_kernel
#define PS_THREADS 256
#define PS_WIN (CB_MAXR + 32)     // rows of the rolling window (a panel reaches at most CB_MAXR rows)

struct PartDesc {
    int r0;                       // first global row of the interior
    int m;                        // interior rows
    int k;                        // interior panels = ceil(m / 32)
    int wT, wB;                   // rows of the top / bottom separator (0: none); the top one sits at [r0 - wT, r0)
    int sep;                      // index of the top separator in the separator system (partition i >= 1: i - 1)
    long long v_off;              // V_i (m x wT, row-major) in the spike buffer
    long long linv_off;           // inverted diagonal blocks [k][32][32] (row-major, lower)
    int ce_off;                   // col_end of the sub-system (relative rows) in the sub-system col_end array
    int cta0;                     // first CTA of this partition in spike_forward_kernel's grid
    int pan0;                     // first panel of this partition in panel_inverse_kernel's grid
    int tile0;                    // first CTA of this partition in spike_gram_kernel's grid
};
struct PartTable { int P; int w; PartDesc d[PS_MAX_PARTS]; };

// L^-1 of every 32x32 diagonal block of the interiors that carry a spike (partitions >= 1). warp = panel.
// The band buffer holds L below the diagonal and 1 / L(j,j) on it (chol_band_kernel).
#define PS_INV_WARPS 2
__global__ void __launch_bounds__(32 * PS_INV_WARPS) panel_inverse_kernel(PartTable T, const double *__restrict__ S, int ld, double *__restrict__ linv, int n_panels) {
    __shared__ double Ls[PS_INV_WARPS][32][33];
    __shared__ double Xs[PS_INV_WARPS][32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gp = blockIdx.x * PS_INV_WARPS + warp;
    if (gp >= n_panels) return;
    int i = 1;
    while (i + 1 < T.P && T.d[i + 1].pan0 <= gp) i++;
    const PartDesc &D = T.d[i];
    const int c = gp - D.pan0, p0 = 32 * c, nb = min(32, D.m - p0);
    // Ls[r][j] = L(p0 + r, p0 + j) (lane = r)
    for (int j = 0; j < 32; j++) {
        double v = (lane == j) ? 1.0 : 0.0;
        if (lane < nb && j < nb && j <= lane) v = __ldcg(S + (size_t)(D.r0 + p0 + j) * ld + D.r0 + p0 + lane);
        Ls[warp][lane][j] = v;
    }
    __syncwarp();
    // lane = column c of the inverse: x_i = (delta_ic - sum_{j<i} L_ij x_j) * (1 / L_ii)
    for (int r = 0; r < 32; r++) {
        double s = (r == lane) ? 1.0 : 0.0;
        for (int j = lane; j < r; j++) s -= Ls[warp][r][j] * Xs[warp][j][lane];
        Xs[warp][r][lane] = (r >= lane) ? s * Ls[warp][r][r] : 0.0;
    }
    __syncwarp();
    double *out = linv + D.linv_off + (size_t)c * 1024;
    for (int r = 0; r < 32; r++) out[r * 32 + lane] = Xs[warp][r][lane];
}

// V_i = L_i^-1 A(I_i, T_i) for PS_NCOL columns of T_i per CTA.
__global__ void __launch_bounds__(PS_THREADS) spike_forward_kernel(PartTable T, const double *__restrict__ S, int ld, const int *__restrict__ sub_colend,
                                                                   const double *__restrict__ linv, double *__restrict__ V) {
    __shared__ double win[PS_WIN][PS_NCOL];          // rows p0 .. p0 + PS_WIN - 1 of the right-hand sides, circular in the row index
    __shared__ double Li[32][33];
    __shared__ double vp[32][PS_NCOL];
    int i = 1;
    while (i + 1 < T.P && T.d[i + 1].cta0 <= (int)blockIdx.x) i++;
    const PartDesc &D = T.d[i];
    const int t0 = ((int)blockIdx.x - D.cta0) * PS_NCOL;
    const int tid = threadIdx.x;
    const int *ce = sub_colend + D.ce_off;
    const int rT = D.r0 - D.wT;                       // first global row of the top separator
    // A(I, T)(r, t) = S[(rT + t) * ld + r0 + r] while the column stays inside the stored band of that row
    for (int e = tid; e < PS_WIN * PS_NCOL; e += PS_THREADS) {
        const int r = e / PS_NCOL, j = e - r * PS_NCOL, t = t0 + j;
        double v = 0.0;
        if (r < D.m && t < D.wT && (D.r0 + r) - (rT + t) <= ld) v = __ldcg(S + (size_t)(rT + t) * ld + D.r0 + r);
        win[r][j] = v;
    }
    double *Vp = V + D.v_off;
    for (int c = 0; c < D.k; c++) {
        const int p0 = 32 * c, nb = min(32, D.m - p0);
        const double *lp = linv + D.linv_off + (size_t)c * 1024;
        for (int e = tid; e < 1024; e += PS_THREADS) Li[e >> 5][e & 31] = __ldcg(lp + e);
        __syncthreads();
        {   // v_p = L_pp^-1 win[p0 .. p0 + 31]
            const int r = tid >> 3, j = tid & 7;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
            for (int q = 0; q < 32; q += 2) {
                s0 += Li[r][q] * win[(p0 + q) % PS_WIN][j];
                s1 += Li[r][q + 1] * win[(p0 + q + 1) % PS_WIN][j];
            }
            const double v = (r < nb) ? s0 + s1 : 0.0;
            vp[r][j] = v;
            if (r < nb && t0 + j < D.wT) Vp[(size_t)(p0 + r) * D.wT + t0 + j] = v;
        }
        __syncthreads();
        // rows below the panel inside the interior: win[row] -= L(row, p0 .. p0 + 31) v_p
        const int rend = min(D.m - 1, ce[p0 + nb - 1]);
        for (int row = p0 + 32 + tid; row <= rend; row += PS_THREADS) {
            double acc[PS_NCOL];
            const int slot = row % PS_WIN;
#pragma unroll
            for (int j = 0; j < PS_NCOL; j++) acc[j] = win[slot][j];
            const double *lrow = S + (size_t)(D.r0 + p0) * ld + D.r0 + row;
#pragma unroll 8
            for (int q = 0; q < 32; q++) {
                const double l = (q < nb) ? __ldcg(lrow + (size_t)q * ld) : 0.0;
#pragma unroll
                for (int j = 0; j < PS_NCOL; j++) acc[j] -= l * vp[q][j];
            }
#pragma unroll
            for (int j = 0; j < PS_NCOL; j++) win[slot][j] = acc[j];
        }
        // the 32 slots of this panel become rows p0 + PS_WIN ..: beyond the reach of A(I, T) (PS_WIN > band), so they start at zero
        if (tid < 32 * PS_NCOL) win[(p0 + (tid >> 3)) % PS_WIN][tid & 7] = 0.0;
        __syncthreads();
    }
}

// spike_forward_kernel with the panel's operands prefetched (sm_100a: cp.async double buffering). The kernel above reads the inverted
// diagonal block and the panel's rows of L straight from L2 inside every panel step, so a step is two or three exposed L2 round trips
// (8.8 us per panel on config 5: 0.64 ms for 72 panels). Here panel c + 1's operands (32 x 32 inverse, up to maxr - 32 rows x 32
// columns of L) stream into shared memory while panel c is applied, a CTA takes 16 right-hand sides (half as many re-reads of L), and
// the trailing update uses two threads per row.  Dynamic shared memory: win [PS_WIN][17] | Li [2][32][33] | Lb [2][32][ldr].
#define PS2_NCOL 16
#define PS2_THREADS 512
#define PS2_WP 17
__device__ __forceinline__ void ps_cp_async8(void *smem_dst, const void *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ps_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ps_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
inline size_t spike_forward2_smem(int ldr) { return sizeof(double) * ((size_t)PS_WIN * PS2_WP + 2 * 32 * 33 + 2 * 32 * (size_t)ldr); }

__global__ void __launch_bounds__(PS2_THREADS, 1) spike_forward2_kernel(PartTable T, const double *__restrict__ S, int ld, const int *__restrict__ sub_colend,
                                                                        const double *__restrict__ linv, double *__restrict__ V, int ldr, int ctas_per_part) {
    extern __shared__ double ps2_sm[];
    double *win = ps2_sm;                                   // [PS_WIN][PS2_WP]: rows p0 .. of the right-hand sides, circular in the row index
    double *Li = win + PS_WIN * PS2_WP;                     // [2][32][33]
    double *Lb = Li + 2 * 32 * 33;                          // [2][32][ldr]: Lb[q][r] = L(p0 + 32 + r, p0 + q)
    __shared__ double vp[32][PS2_NCOL];
    const int i = 1 + (int)blockIdx.x / ctas_per_part;
    const PartDesc &D = T.d[i];
    const int t0 = ((int)blockIdx.x - (i - 1) * ctas_per_part) * PS2_NCOL;
    const int tid = threadIdx.x;
    const int *ce = sub_colend + D.ce_off;
    const int rT = D.r0 - D.wT;
    for (int e = tid; e < PS_WIN * PS2_NCOL; e += PS2_THREADS) {
        const int r = e / PS2_NCOL, j = e - r * PS2_NCOL, t = t0 + j;
        double v = 0.0;
        if (r < D.m && t < D.wT && (D.r0 + r) - (rT + t) <= ld) v = __ldcg(S + (size_t)(rT + t) * ld + D.r0 + r);
        win[r * PS2_WP + j] = v;
    }
    auto rows_below = [&](int c) { const int p0 = 32 * c, nb = min(32, D.m - p0); return max(0, min(D.m - 1, ce[p0 + nb - 1]) - (p0 + 32) + 1); };
    auto prefetch = [&](int c, int buf) {
        const int p0 = 32 * c, nb = min(32, D.m - p0);
        const double *lp = linv + D.linv_off + (size_t)c * 1024;
        double *li = Li + buf * 32 * 33;
        for (int e = tid; e < 1024; e += PS2_THREADS) ps_cp_async8(li + (e >> 5) * 33 + (e & 31), lp + e);
        const int R = rows_below(c);
        double *lb = Lb + (size_t)buf * 32 * ldr;
        for (int e = tid; e < 32 * R; e += PS2_THREADS) {
            const int q = e / R, r = e - q * R;
            if (q < nb) ps_cp_async8(lb + (size_t)q * ldr + r, S + (size_t)(D.r0 + p0 + q) * ld + D.r0 + p0 + 32 + r);
            else lb[(size_t)q * ldr + r] = 0.0;
        }
        ps_cp_commit();
    };
    double *Vp = V + D.v_off;
    prefetch(0, 0);
    for (int c = 0; c < D.k; c++) {
        const int p0 = 32 * c, nb = min(32, D.m - p0), buf = c & 1;
        ps_cp_wait_all();
        __syncthreads();                                   // panel c's operands are in place; everybody is done with the other buffer
        if (c + 1 < D.k) prefetch(c + 1, buf ^ 1);
        {   // v_p = L_pp^-1 win[p0 .. p0 + 31]: thread = (row, column)
            const int r = tid >> 4, j = tid & 15;
            const double *li = Li + buf * 32 * 33 + r * 33;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
            for (int q = 0; q < 32; q += 2) {
                s0 += li[q] * win[((p0 + q) % PS_WIN) * PS2_WP + j];
                s1 += li[q + 1] * win[((p0 + q + 1) % PS_WIN) * PS2_WP + j];
            }
            const double v = (r < nb) ? s0 + s1 : 0.0;
            vp[r][j] = v;
            if (r < nb && t0 + j < D.wT) Vp[(size_t)(p0 + r) * D.wT + t0 + j] = v;
        }
        __syncthreads();
        {   // rows below the panel: win[row] -= L(row, p0 .. p0 + 31) v_p: thread = (row, half of the columns)
            const int R = rows_below(c);
            const double *lb = Lb + (size_t)buf * 32 * ldr;
            for (int rr = tid >> 1; rr < R; rr += PS2_THREADS / 2) {
                const int h8 = 8 * (tid & 1);
                double *w = win + ((p0 + 32 + rr) % PS_WIN) * PS2_WP + h8;
                double acc[8];
#pragma unroll
                for (int j = 0; j < 8; j++) acc[j] = w[j];
#pragma unroll 8
                for (int q = 0; q < 32; q++) {
                    const double l = lb[(size_t)q * ldr + rr];
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[j] -= l * vp[q][h8 + j];
                }
#pragma unroll
                for (int j = 0; j < 8; j++) w[j] = acc[j];
            }
            // the 32 slots of this panel become rows p0 + PS_WIN ..: beyond the reach of A(I, T) (PS_WIN > band), so they start at zero
            // (v_p has read them before the barrier above; the update does not touch them: R <= PS_WIN - 64)
            if (tid < 32 * PS2_NCOL) win[((p0 + (tid >> 4)) % PS_WIN) * PS2_WP + (tid & 15)] = 0.0;
        }
    }
}

// ---- FP64 tensor-core tile product (mma.sync.m8n8k4.f64 = DMMA): the spike products V^T V and F^T F are dense contractions over
// thousands of rows, the one place on this path where the FP64 MMA pays (measured on this B200: 37.0 TFLOP/s against 33.7 for DFMA,
// and one instruction per 256 multiply-adds instead of 8 per thread, so the loop is no longer bound by shared-memory loads).
// out (32x32) += X[rows, xa .. xa + 31]^T  Y[rows, yb .. yb + 31], rows = [r_begin, r_end), X and Y row-major with leading dimension ldxy.
// 256 threads: warp v owns the 8x8 sub-tiles (v >> 1, 2 (v & 1)) and (v >> 1, 2 (v & 1) + 1); a lane ends with elements
// (lane / 4, 2 (lane % 4) + {0, 1}) of each. Shared tiles use a row stride of 36 doubles: the fragment loads are conflict-free.
#define PS_TLD 36
#define CR_MAXW 384                    // widest separator the block kernels (and the gp product) are sized for
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// The 32-row slabs of X and Y are double-buffered: slab s + 1 streams in (cp.async) while slab s is multiplied.
__device__ __forceinline__ void tile_xty_dmma(const double *__restrict__ X, const double *__restrict__ Y, int ldxy, int r_begin, int r_end, int xa, int yb,
                                              double (*As)[32][PS_TLD], double (*Bs)[32][PS_TLD], double c[2][2]) {
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int si = wv >> 1, sj = 2 * (wv & 1), fr = lane >> 2, fk = lane & 3;
    auto prefetch = [&](int r0, int b) {
        for (int e = tid; e < 1024; e += 256) {
            const int rr = e >> 5, cc = e & 31;
            if (r0 + rr < r_end) {
                ps_cp_async8(&As[b][rr][cc], X + (size_t)(r0 + rr) * ldxy + xa + cc);
                ps_cp_async8(&Bs[b][rr][cc], Y + (size_t)(r0 + rr) * ldxy + yb + cc);
            } else { As[b][rr][cc] = 0.0; Bs[b][rr][cc] = 0.0; }
        }
        ps_cp_commit();
    };
    __syncthreads();                                       // a previous call's last slab has been read
    if (r_begin < r_end) prefetch(r_begin, 0);
    int b = 0;
    for (int r0 = r_begin; r0 < r_end; r0 += 32, b ^= 1) {
        ps_cp_wait_all();
        __syncthreads();
        if (r0 + 32 < r_end) prefetch(r0 + 32, b ^ 1);
#pragma unroll
        for (int k0 = 0; k0 < 32; k0 += 4) {
            const double a = As[b][k0 + fk][8 * si + fr];
            const double b0 = Bs[b][k0 + fk][8 * sj + fr], b1 = Bs[b][k0 + fk][8 * sj + 8 + fr];
            dmma_m8n8k4(c[0][0], c[0][1], a, b0);
            dmma_m8n8k4(c[1][0], c[1][1], a, b1);
        }
    }
}
// where the lane's four results live inside the 32x32 tile: row, and column of c[q][0] (c[q][1] is the next column)
#define PS_FRAG_ROW(tid) (8 * ((tid) >> 6) + (((tid) & 31) >> 2))
#define PS_FRAG_COL(tid, q) (8 * (2 * (((tid) >> 5) & 1) + (q)) + 2 * ((tid) & 3))

// Per partition i >= 1:  Dp_i = V_i^T V_i (upper 32x32 tiles),  Ep_i = V_i^T G_i^T (all tiles; G_i = L(T_{i+1}, I_i), only the last
// rows of the interior reach it),  gp_i = V_i^T y_i.  CTA = one output tile, the whole sum in a fixed order.
__global__ void __launch_bounds__(256) spike_gram_kernel(PartTable T, const double *__restrict__ S, int ld, const double *__restrict__ V,
                                                         const double *__restrict__ y, double *__restrict__ Dp, double *__restrict__ Ep, double *__restrict__ gp) {
    __shared__ double sg_sm[4 * 32 * PS_TLD];              // As[2] | Bs[2]; the gp branch uses the same bytes for its partial sums
    double (*As)[32][PS_TLD] = reinterpret_cast<double (*)[32][PS_TLD]>(sg_sm);
    double (*Bs)[32][PS_TLD] = reinterpret_cast<double (*)[32][PS_TLD]>(sg_sm + 2 * 32 * PS_TLD);
    static_assert(8 * CR_MAXW <= 4 * 32 * PS_TLD, "gp partial sums must fit in the tile buffers");
    int i = 1;
    while (i + 1 < T.P && T.d[i + 1].tile0 <= (int)blockIdx.x) i++;
    const PartDesc &D = T.d[i];
    const int w = D.wT, nt = w / 32, ntri = nt * (nt + 1) / 2;
    int tile = (int)blockIdx.x - D.tile0;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const double *Vp = V + D.v_off;
    const size_t so = (size_t)D.sep * w * w;
    if (tile < ntri) {                                   // ---- Dp: tile (a, b), a <= b, on the FP64 tensor pipe
        int a = 0;
        while (tile >= nt - a) { tile -= nt - a; a++; }
        const int b = a + tile;
        double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        tile_xty_dmma(Vp, Vp, w, 0, D.m, 32 * a, 32 * b, As, Bs, c);
#pragma unroll
        for (int q = 0; q < 2; q++) {
            double *o = Dp + so + (size_t)(32 * a + PS_FRAG_ROW(tid)) * w + 32 * b + PS_FRAG_COL(tid, q);
            o[0] = c[q][0]; o[1] = c[q][1];
        }
        return;
    }
    tile -= ntri;
    if (tile < nt * nt && D.wB > 0) {                    // ---- Ep: tile (a, b) of V^T G^T, G(t, c) = L(m + t, c) = S[(r0 + c) * ld + r0 + m + t]
        const int a = tile / nt, b = tile - a * nt;
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        // interior columns c that reach separator row m + 32 b (and beyond): c >= m + 32 b - ld
        const int cfirst = max(0, (D.m + 32 * b - ld) & ~31);
        for (int r0 = cfirst; r0 < D.m; r0 += 32) {
            __syncthreads();
            for (int e = tid; e < 1024; e += 256) {
                const int rr = e >> 5, cc = e & 31;
                const int c = r0 + rr, trow = D.m + 32 * b + cc;                 // Bs[0][rr][cc] = G(32 b + cc, c)
                const bool in = c < D.m;
                As[0][rr][cc] = in ? __ldcg(Vp + (size_t)c * w + 32 * a + cc) : 0.0;
                Bs[0][rr][cc] = (in && trow - c <= ld) ? __ldcg(S + (size_t)(D.r0 + c) * ld + D.r0 + trow) : 0.0;
            }
            __syncthreads();
#pragma unroll 8
            for (int rr = 0; rr < 32; rr++) {
                const double a0 = As[0][rr][ty], a1 = As[0][rr][ty + 16], b0 = Bs[0][rr][tx], b1 = Bs[0][rr][tx + 16];
                acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int v = 0; v < 2; v++) Ep[so + (size_t)(32 * a + ty + 16 * u) * w + 32 * b + tx + 16 * v] = acc[u][v];
        return;
    }
    if (tile == nt * nt) {                               // ---- gp = V^T y: warp v takes the rows r = v (mod 8), lane the columns t = lane (mod 32)
        double (*gpart)[CR_MAXW] = reinterpret_cast<double (*)[CR_MAXW]>(sg_sm);   // (one thread per column walking all m rows is a chain of m exposed L2 loads)
        const double *yi = y + D.r0;
        const int lane = tid & 31, wv = tid >> 5, ng = w / 32;
        double acc[CR_MAXW / 32];
#pragma unroll
        for (int g = 0; g < CR_MAXW / 32; g++) acc[g] = 0.0;
#pragma unroll 4
        for (int r = wv; r < D.m; r += 8) {
            const double yr = __ldcg(yi + r);
            const double *vr = Vp + (size_t)r * w + lane;
#pragma unroll
            for (int g = 0; g < CR_MAXW / 32; g++) if (g < ng) acc[g] += __ldcg(vr + 32 * g) * yr;
        }
#pragma unroll
        for (int g = 0; g < CR_MAXW / 32; g++) if (g < ng) gpart[wv][32 * g + lane] = acc[g];
        __syncthreads();
        for (int t = tid; t < w; t += 256) {
            double sum = 0.0;
#pragma unroll
            for (int v = 0; v < 8; v++) sum += gpart[v][t];
            gp[(size_t)D.sep * w + t] = sum;
        }
    }
}

// The separator system in band storage (element (R, C), R <= C, at SM[R * ldM + C]; lambda is added by the solver):
//   diagonal block j:   S(T_j, T_j) as partition j-1... left it (A - G G^T, no lambda)  -  Dp_j
//   block (j, j + 1):   -Ep_{j}'s partition: the interior between T_j and T_{j+1} couples them
//   rhs:                b_p + b_s + y (the forward updates of the interior above)  -  gp_j
// Separator j (0-based) sits at global rows sep_row[j] .. sep_row[j] + w - 1 and is the TOP separator of partition j + 1.
__global__ void sep_assemble_kernel(int nsep, int w, const int *__restrict__ sep_row, const double *__restrict__ S, int ld, const double *__restrict__ bp,
                                    const double *__restrict__ bs, const double *__restrict__ y, const double *__restrict__ Dp, const double *__restrict__ Ep,
                                    const double *__restrict__ gp, double *__restrict__ SM, int ldM, double *__restrict__ rhsM) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)w * 2 * w;            // per separator row block: w rows x (own block + next block) columns
    if (idx >= (long long)nsep * per) return;
    const int j = (int)(idx / per);
    const int rem = (int)(idx - (long long)j * per);
    const int tr = rem / (2 * w), tc2 = rem - tr * 2 * w;
    const int R = j * w + tr;
    if (tc2 < w) {
        const int tc = tc2;
        if (tc < tr) return;
        const int gr = sep_row[j] + tr, gc = sep_row[j] + tc;
        double v = (gc - gr <= ld) ? S[(size_t)gr * ld + gc] : 0.0;
        v -= Dp[(size_t)j * w * w + (size_t)tr * w + tc];
        SM[(size_t)R * ldM + (j * w + tc)] = v;
        if (tc == tr) rhsM[R] = bp[gr] + bs[gr] + y[gr] - gp[(size_t)j * w + tr];
    } else if (j + 1 < nsep) {
        const int tc = tc2 - w;
        SM[(size_t)R * ldM + ((j + 1) * w + tc)] = -Ep[(size_t)j * w * w + (size_t)tr * w + tc];
    }
}

// x_T into the final x and into y (the backward substitutions read the separator below their interior from y)
__global__ void sep_scatter_kernel(int nsep, int w, const int *__restrict__ sep_row, const double *__restrict__ xM, double *__restrict__ y, double *__restrict__ x) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nsep * w) return;
    const int j = idx / w, t = idx - j * w;
    const double v = xM[idx];
    y[sep_row[j] + t] = v; x[sep_row[j] + t] = v;
}

// y_i <- y_i - V_i x_{T_i} for the interiors with a top separator. warp = row.
__global__ void __launch_bounds__(256) spike_apply_kernel(PartTable T, const double *__restrict__ V, const double *__restrict__ xM, double *__restrict__ y, int rows_total) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= rows_total) return;
    int i = 1, base = 0;
    while (i + 1 < T.P && base + T.d[i].m <= g) { base += T.d[i].m; i++; }
    const PartDesc &D = T.d[i];
    const int r = g - base;
    if (r >= D.m) return;
    const double *vr = V + D.v_off + (size_t)r * D.wT;
    const double *xt = xM + (size_t)D.sep * D.wT;
    double s = 0.0;
    for (int t = lane; t < D.wT; t += 32) s += __ldcg(vr + t) * __ldcg(xt + t);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[D.r0 + r] -= s;
}

// order in which pair_kernel takes the camera rows when the factorisation has P fronts: round-robin over the sub-systems
// (interior + bottom separator), each from its first camera down, so that every cluster is fed from the start.
__global__ void row_order_parts_kernel(int nf, PartTable T, int *row_pos, int *row_of_pos) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nf) return;
    // sub-system q owns the cameras of its interior and of its bottom separator: [r0_q / 6, r0_{q+1} / 6)
    int owner = 0;
    while (owner + 1 < T.P && T.d[owner + 1].r0 / 6 <= a) owner++;
    const int first = (owner == 0) ? 0 : T.d[owner].r0 / 6;
    const int t = a - first;
    int pos = 0;
    for (int q = 0; q < T.P; q++) {
        const int f = (q == 0) ? 0 : T.d[q].r0 / 6;
        const int e = (q + 1 < T.P) ? T.d[q + 1].r0 / 6 : nf;                         // one past the last camera of sub-system q (interior + bottom separator)
        const int len = e - f;
        pos += min(t, len) + ((q < owner && len > t) ? 1 : 0);
    }
    row_pos[a] = pos; row_of_pos[pos] = a;
}

// ---------------------------------------------------------------------------------------------------------------------
// Separator system by block cyclic reduction.
//
// The separator system is block tridiagonal: K = P - 1 dense w x w diagonal blocks D_j, couplings C_j = block (j, j + 1), right-hand
// sides g_j. Factoring it as one band system is again a serial chain (6 K panels, and with a reach of 2 w rows it does not fit the
// shared-memory window of chol_band_kernel). Level l (stride s = 2^l) eliminates every second block that is still active
// (j = s - 1 + 2 s t), all of them AT ONCE: their neighbours j - s and j + s are not eliminated at this level, so the blocks are
// independent. Per level:
//   factor       D_j + lambda I = L_j L_j^T, y_j = L_j^-1 g_j        chol_band_kernel, one cluster per block (table launch, no_back)
//   spikes       Fa_j = L_j^-1 C_{j-s}^T,  Fb_j = L_j^-1 C_j         block_inverse_kernel + block_spike_kernel
//   reduce       D_t -= Fb_{t-s}^T Fb_{t-s} + Fa_{t+s}^T Fa_{t+s},   g_t -= Fb_{t-s}^T y_{t-s} + Fa_{t+s}^T y_{t+s},
//                C_t (now block (t, t + 2 s)) = -Fa_{t+s}^T Fb_{t+s}  for the blocks t that stay      block_gram_kernel
// and on the way back, level by level:  y_j -= Fa_j x_{j-s} + Fb_j x_{j+s}  (block_apply_kernel),  x_j = L_j^-T y_j  (table launch).
// log2(K) levels of one dense w x w factorisation each instead of K of them in a row; every sum in a fixed order.
struct CrPlan { int K, w; double *Dd, *Cc, *gg, *yy, *xx, *Fa, *Fb, *linv; };

// level-0 blocks from the band buffer and the interiors' products (same content as sep_assemble_kernel, dense block layout)
__global__ void cr_assemble_kernel(CrPlan C, const int *__restrict__ sep_row, const double *__restrict__ S, int ld, const double *__restrict__ bp,
                                   const double *__restrict__ bs, const double *__restrict__ y, const double *__restrict__ Dp, const double *__restrict__ Ep,
                                   const double *__restrict__ gp) {
    const int w = C.w;
    const long long per = (long long)w * 2 * w;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)C.K * per) return;
    const int j = (int)(idx / per);
    const int rem = (int)(idx - (long long)j * per);
    const int tr = rem / (2 * w), tc2 = rem - tr * 2 * w;
    const size_t bo = (size_t)j * w * w;
    if (tc2 < w) {
        const int tc = tc2;
        if (tc < tr) return;
        const int gr = sep_row[j] + tr, gc = sep_row[j] + tc;
        const double v = ((gc - gr <= ld) ? S[(size_t)gr * ld + gc] : 0.0) - Dp[bo + (size_t)tr * w + tc];
        C.Dd[bo + (size_t)tr * w + tc] = v;
        if (tc == tr) C.gg[(size_t)j * w + tr] = bp[gr] + bs[gr] + y[gr] - gp[(size_t)j * w + tr];
    } else if (j + 1 < C.K) {
        const int tc = tc2 - w;
        C.Cc[bo + (size_t)tr * w + tc] = -Ep[bo + (size_t)tr * w + tc];
    }
}

// inverse of the 32x32 diagonal blocks of the factored D_j of this level. warp = (block, panel)
__global__ void __launch_bounds__(32 * PS_INV_WARPS) block_inverse_kernel(CrPlan C, int s, int cnt) {
    __shared__ double Ls[PS_INV_WARPS][32][33];
    __shared__ double Xs[PS_INV_WARPS][32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int np = C.w / 32;
    const int g = blockIdx.x * PS_INV_WARPS + warp;
    if (g >= cnt * np) return;
    const int j = s - 1 + 2 * s * (g / np), c = g % np, p0 = 32 * c;
    const double *L = C.Dd + (size_t)j * C.w * C.w;
    for (int q = 0; q < 32; q++) Ls[warp][lane][q] = (q <= lane) ? __ldcg(L + (size_t)(p0 + q) * C.w + p0 + lane) : 0.0;
    __syncwarp();
    for (int r = 0; r < 32; r++) {
        double sum = (r == lane) ? 1.0 : 0.0;
        for (int q = lane; q < r; q++) sum -= Ls[warp][r][q] * Xs[warp][q][lane];
        Xs[warp][r][lane] = (r >= lane) ? sum * Ls[warp][r][r] : 0.0;
    }
    __syncwarp();
    double *out = C.linv + ((size_t)j * np + c) * 1024;
    for (int r = 0; r < 32; r++) out[r * 32 + lane] = Xs[warp][r][lane];
}

// Fa_j = L_j^-1 C_{j-s}^T (side 0), Fb_j = L_j^-1 C_j (side 1), PS_NCOL columns per CTA; blockIdx = ((t * 2 + side) * (w / PS_NCOL) + group)
__global__ void __launch_bounds__(PS_THREADS) block_spike_kernel(CrPlan C, int s, int cnt) {
    __shared__ double win[CR_MAXW][PS_NCOL];
    __shared__ double Li[32][33];
    __shared__ double vp[32][PS_NCOL];
    const int w = C.w, ng = w / PS_NCOL, tid = threadIdx.x;
    const int grp = blockIdx.x % ng, side = (blockIdx.x / ng) & 1, t = blockIdx.x / (2 * ng);
    const int j = s - 1 + 2 * s * t;
    if (side == 0 ? (j - s < 0) : (j + s >= C.K)) return;
    const int t0 = grp * PS_NCOL;
    const double *src = C.Cc + (size_t)(side == 0 ? j - s : j) * w * w;
    for (int e = tid; e < w * PS_NCOL; e += PS_THREADS) {
        const int r = e / PS_NCOL, q = e - r * PS_NCOL;
        win[r][q] = (side == 0) ? __ldcg(src + (size_t)(t0 + q) * w + r) : __ldcg(src + (size_t)r * w + t0 + q);
    }
    const double *L = C.Dd + (size_t)j * w * w;
    double *F = (side == 0 ? C.Fa : C.Fb) + (size_t)j * w * w;
    const int np = w / 32;
    for (int c = 0; c < np; c++) {
        const int p0 = 32 * c;
        const double *lp = C.linv + ((size_t)j * np + c) * 1024;
        for (int e = tid; e < 1024; e += PS_THREADS) Li[e >> 5][e & 31] = __ldcg(lp + e);
        __syncthreads();
        {
            const int r = tid >> 3, q = tid & 7;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
            for (int k = 0; k < 32; k += 2) { s0 += Li[r][k] * win[p0 + k][q]; s1 += Li[r][k + 1] * win[p0 + k + 1][q]; }
            vp[r][q] = s0 + s1;
            F[(size_t)(p0 + r) * w + t0 + q] = s0 + s1;
        }
        __syncthreads();
        for (int row = p0 + 32 + tid; row < w; row += PS_THREADS) {
            double acc[PS_NCOL];
#pragma unroll
            for (int q = 0; q < PS_NCOL; q++) acc[q] = win[row][q];
            const double *lrow = L + (size_t)p0 * w + row;
#pragma unroll 8
            for (int k = 0; k < 32; k++) {
                const double l = __ldcg(lrow + (size_t)k * w);
#pragma unroll
                for (int q = 0; q < PS_NCOL; q++) acc[q] -= l * vp[k][q];
            }
#pragma unroll
            for (int q = 0; q < PS_NCOL; q++) win[row][q] = acc[q];
        }
        __syncthreads();
    }
}

// the blocks that stay at this level: t = 2 s - 1 + 2 s u. blockIdx = u * (2 nt^2 + 1) + tile
__global__ void __launch_bounds__(256) block_gram_kernel(CrPlan C, int s, int n_keep) {
    __shared__ double As[2][32][PS_TLD], Bs[2][32][PS_TLD];
    const int w = C.w, nt = w / 32, per = 2 * nt * nt + 1;
    const int u = blockIdx.x / per;
    int tile = blockIdx.x - u * per;
    const int t = 2 * s - 1 + 2 * s * u, jl = t - s, jr = t + s;
    const bool has_r = jr < C.K;
    const size_t ww = (size_t)w * w;
    const int tid = threadIdx.x;
    if (tile < nt * nt) {                                   // D_t, upper tiles only
        const int a = tile / nt, b = tile - a * nt;
        if (a > b) return;
        double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        tile_xty_dmma(C.Fb + jl * ww, C.Fb + jl * ww, w, 0, w, 32 * a, 32 * b, As, Bs, c);
        if (has_r) tile_xty_dmma(C.Fa + jr * ww, C.Fa + jr * ww, w, 0, w, 32 * a, 32 * b, As, Bs, c);
#pragma unroll
        for (int q = 0; q < 2; q++) {
            double *o = C.Dd + t * ww + (size_t)(32 * a + PS_FRAG_ROW(tid)) * w + 32 * b + PS_FRAG_COL(tid, q);
            o[0] -= c[q][0]; o[1] -= c[q][1];
        }
        return;
    }
    tile -= nt * nt;
    if (tile < nt * nt) {                                   // new coupling (t, t + 2 s) = -Fa_{jr}^T Fb_{jr}
        if (!has_r || t + 2 * s >= C.K) return;
        const int a = tile / nt, b = tile - a * nt;
        double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        tile_xty_dmma(C.Fa + jr * ww, C.Fb + jr * ww, w, 0, w, 32 * a, 32 * b, As, Bs, c);
#pragma unroll
        for (int q = 0; q < 2; q++) {
            double *o = C.Cc + t * ww + (size_t)(32 * a + PS_FRAG_ROW(tid)) * w + 32 * b + PS_FRAG_COL(tid, q);
            o[0] = -c[q][0]; o[1] = -c[q][1];
        }
        return;
    }
    for (int c = tid; c < w; c += 256) {                    // g_t
        double s0 = 0.0, s1 = 0.0;
        const double *F = C.Fb + jl * ww, *yv = C.yy + (size_t)jl * w;
#pragma unroll 8
        for (int r = 0; r < w; r++) s0 += __ldcg(F + (size_t)r * w + c) * __ldcg(yv + r);
        if (has_r) {
            F = C.Fa + jr * ww; yv = C.yy + (size_t)jr * w;
#pragma unroll 8
            for (int r = 0; r < w; r++) s1 += __ldcg(F + (size_t)r * w + c) * __ldcg(yv + r);
        }
        C.gg[(size_t)t * w + c] -= s0 + s1;
    }
}

// back phase: y_j -= Fa_j x_{j-s} + Fb_j x_{j+s} for the blocks eliminated at this level. warp = row
__global__ void __launch_bounds__(256) block_apply_kernel(CrPlan C, int s, int cnt) {
    const int lane = threadIdx.x & 31, w = C.w;
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= cnt * w) return;
    const int j = s - 1 + 2 * s * (g / w), r = g % w;
    double sum = 0.0;
    if (j - s >= 0) {
        const double *f = C.Fa + (size_t)j * w * w + (size_t)r * w, *xv = C.xx + (size_t)(j - s) * w;
        for (int c = lane; c < w; c += 32) sum += __ldcg(f + c) * __ldcg(xv + c);
    }
    if (j + s < C.K) {
        const double *f = C.Fb + (size_t)j * w * w + (size_t)r * w, *xv = C.xx + (size_t)(j + s) * w;
        for (int c = lane; c < w; c += 32) sum += __ldcg(f + c) * __ldcg(xv + c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) C.yy[(size_t)j * w + r] -= sum;
}
