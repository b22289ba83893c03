// Camera half of the Schur complement, tile-major (sm_100a): pair_tile_kernel.
//
// pair_kernel (schur_pairs.cuh) walks one list entry per observation pair and gathers two 96-byte Z records for each of them:
// 2.4 GB of L2 -> SM traffic for the 12.3 M pairs of BASELINE config 4, 25 GB (33 GB of DRAM reads) for config 5 -- the kernel is
// bound by the misses it can keep in flight. Covisibility is local: a landmark's track covers a run of neighbouring keyframes, so the
// pairs of one landmark fall into a few TP_T x TP_T tiles of camera blocks. Here the unit of the plan is one (tile, landmark) record
//     { observation of camera 4 tA + i (i = 0..3) | observation of camera 4 tB + i }        (32 bytes, -1 = not observed)
// and a warp holds the 16 blocks of a tile in registers: lane = (half h, ia, ib) accumulates block (4 tA + ia, 4 tB + ib) over
// every second record. The four lanes that share ia read the same Z record (one L1 request), so a record is fetched once per
// tile instead of once per pair: 0.62 records per pair instead of 2 on config 4, and the 98 MB pair list becomes 53 MB of records.
// The block contraction is the one of pair_kernel (G = Y_a Y_b^T, then [P_a^T G P_b, P_a^T G; G P_b, G]) written as chained FMAs:
// 99 FP64 instructions per pair.
//
// Several edges on one (pose, point) pair (left + right eye of a rig keyframe): the k-th edge of a pose is "layer" k, a group is the
// set of a landmark's observations with the same (tile, layer), and the plan holds one record per ordered pair of groups
// (for tA == tB: all ordered pairs, the lanes with ia <= ib only), which enumerates every (e_a, e_b), a <= b, exactly once and the
// (e, e') / (e', e) couples of one camera both ways (M + M^T on the diagonal block), like the pair list did.
// Diagonal tiles also sum the Dr records of their cameras (Hpp, b_p, b_s): a second pass over the same records, lane = (record, camera).
//
// Everything that reaches the reduced system is summed in a fixed order (records in landmark order, items of a tile in item order by
// the last finisher), so two solves of a map agree bit for bit.
// Replaces the landmark loop of BlockSolver::solve (Thirdparty/g2o/g2o/core/block_solver.hpp:400-444) and the Hpp / b gather of
// buildSystem (block_solver.hpp:501-560).
#pragma once
#include "schur_pairs.cuh"

#define TP_T 4
#ifndef TP_CHUNK
#define TP_CHUNK 128                       // records per work item
#endif
#define TP_THREADS 128
#define TP_WARPS (TP_THREADS / 32)
#ifndef TP_MINB
#define TP_MINB 3
#endif
#define TP_PART 720                        // doubles per item in the partial buffer: [36][16] block sums | [33][4] Dr sums | pad
#define TP_MAX_LAYERS 32                   // edges on one (pose, point) pair the tile plan can express (more: pair_kernel path)

struct __align__(16) TileRec { int ea[TP_T]; int eb[TP_T]; };
struct TileItem { int ta, tb, begin, end, first, nit; };

// ---------------------------------------------------------------------------------------------------------------------
// plan (upload). Observations are sorted landmark-major / pose-ascending, free indices are monotone in the pose index, so the
// free observations of a landmark that fall into one tile are a contiguous run (fixed poses may sit in between).

// order in which the tiles' rows are taken: the order of their first camera in row_pos (identity / both ends inwards / round-robin
// over the fronts of the partitioned solver). cam_tpos[c] = position of camera c's tile row (what chol_band_kernel waits on).
__global__ void tile_order_kernel(int ntile, int nf, const int *__restrict__ row_pos, int *__restrict__ tile_pos, int *__restrict__ tile_of_pos, int *__restrict__ cam_tpos) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntile) return;
    const int key = row_pos[TP_T * t];
    int rank = 0;
    for (int u = 0; u < ntile; u++) rank += (row_pos[TP_T * u] < key) ? 1 : 0;
    tile_pos[t] = rank; tile_of_pos[rank] = t;
    for (int c = TP_T * t; c < min(nf, TP_T * t + TP_T); c++) cam_tpos[c] = rank;
}

// One thread per observation e. If e is the first observation of its group (tile tA, layer l), it owns the records (G, G') for every
// group G' of the same landmark in a tile >= tA (all groups of tA itself included).  GEN = false: count them (nrec[e], blk_cnt);
// GEN = true: write key / record / index at rec_off[e] ..
template <bool GEN>
__global__ void tile_plan_kernel(int64_t n_obs, const int *__restrict__ lm_ptr, const int *__restrict__ o_pose, const int *__restrict__ o_point,
                                 const int *__restrict__ hidx, const int *__restrict__ tile_pos, int tbw1, unsigned *__restrict__ nrec, unsigned *blk_cnt,
                                 const unsigned *__restrict__ rec_off, unsigned *__restrict__ keys, unsigned *__restrict__ idx, TileRec *__restrict__ recs, int *overflow) {
    const int64_t e64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e64 >= n_obs) return;
    const int e = (int)e64;
    const int pe = o_pose[e];
    const int ha = hidx[pe];
    if (ha < 0) { if (!GEN) nrec[e] = 0; return; }
    const int j = o_point[e];
    const int p0 = lm_ptr[j], p1 = lm_ptr[j + 1];
    const int ta = ha / TP_T;
    int le = 0;
    for (int q = e - 1; q >= p0 && o_pose[q] == pe; q--) le++;
    if (le >= TP_MAX_LAYERS) { if (!GEN) { nrec[e] = 0; atomicExch(overflow, 1); } return; }
    // leader of (ta, le)? no earlier observation of this tile has the same layer
    {
        int prev = -1, lay = 0;
        for (int q = p0; q < e; q++) {
            const int pq = o_pose[q];
            lay = (pq == prev) ? lay + 1 : 0; prev = pq;
            const int hq = hidx[pq];
            if (hq >= 0 && hq / TP_T == ta && lay == le) { if (!GEN) nrec[e] = 0; return; }
        }
    }
    TileRec R;
    if (GEN) {                                             // members of my group: same tile, same layer
#pragma unroll
        for (int i = 0; i < TP_T; i++) R.ea[i] = -1;
        int prev = -1, lay = 0;
        for (int q = p0; q < p1; q++) {
            const int pq = o_pose[q];
            lay = (pq == prev) ? lay + 1 : 0; prev = pq;
            const int hq = hidx[pq];
            if (hq < 0) continue;
            if (hq / TP_T > ta) break;
            if (hq / TP_T == ta && lay == le) R.ea[hq - TP_T * ta] = q;
        }
    }
    // every group leader of the landmark in a tile >= ta
    unsigned cnt = 0;
    unsigned pos = GEN ? rec_off[e] : 0u;
    int prev = -1, lay = 0, cur_tile = -1;
    unsigned seen = 0u;
    const unsigned row = (unsigned)tile_pos[ta] * (unsigned)tbw1;
    for (int q = p0; q < p1; q++) {
        const int pq = o_pose[q];
        lay = (pq == prev) ? lay + 1 : 0; prev = pq;
        const int hq = hidx[pq];
        if (hq < 0) continue;
        const int tq = hq / TP_T;
        if (tq != cur_tile) { cur_tile = tq; seen = 0u; }
        if (lay >= TP_MAX_LAYERS) continue;                // flagged by that observation's own thread
        const bool lead = !((seen >> lay) & 1u);
        seen |= 1u << lay;
        if (!lead || tq < ta) continue;
        if (!GEN) { cnt++; atomicAdd(blk_cnt + row + (unsigned)(tq - ta), 1u); continue; }
#pragma unroll
        for (int i = 0; i < TP_T; i++) R.eb[i] = -1;
        int prev2 = pq, lay2 = lay;
        for (int q2 = q; q2 < p1; q2++) {                  // members of the group led by q
            const int p2 = o_pose[q2];
            if (q2 > q) { lay2 = (p2 == prev2) ? lay2 + 1 : 0; prev2 = p2; }
            const int h2 = hidx[p2];
            if (h2 < 0) continue;
            if (h2 / TP_T != tq) break;
            if (lay2 == lay) R.eb[h2 - TP_T * tq] = q2;
        }
        keys[pos] = row + (unsigned)(tq - ta);
        idx[pos] = pos;
        recs[pos] = R;
        pos++;
    }
    if (!GEN) nrec[e] = cnt;
}

__global__ void tile_gather_kernel(unsigned n, const unsigned *__restrict__ idx, const TileRec *__restrict__ in, TileRec *__restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 *s = reinterpret_cast<const int4 *>(in + idx[i]);
    int4 *d = reinterpret_cast<int4 *>(out + i);
    d[0] = s[0]; d[1] = s[1];
}

__global__ void tile_item_count_kernel(int nblk, const unsigned *__restrict__ blk_cnt, unsigned *__restrict__ item_cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nblk) item_cnt[i] = (blk_cnt[i] + TP_CHUNK - 1) / TP_CHUNK;
}

__global__ void tile_item_fill_kernel(int nblk, int tbw1, const int *__restrict__ tile_of_pos, const unsigned *__restrict__ blk_off, const unsigned *__restrict__ blk_cnt,
                                      const unsigned *__restrict__ item_off, TileItem *__restrict__ items) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblk) return;
    const unsigned cnt = blk_cnt[i];
    if (!cnt) return;
    const int tp = i / tbw1, ta = tile_of_pos[tp], tb = ta + (i - tp * tbw1);
    const unsigned off = blk_off[i];
    unsigned io = item_off[i];
    const int nit = (int)((cnt + TP_CHUNK - 1) / TP_CHUNK);
    for (unsigned c = 0; c < cnt; c += TP_CHUNK) {
        TileItem I; I.ta = ta; I.tb = tb; I.begin = (int)(off + c); I.end = (int)(off + min(cnt, c + TP_CHUNK));
        I.first = (int)item_off[i]; I.nit = nit;
        items[io++] = I;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
struct TileArgs {
    const TileItem *items; int n_items;
    const TileRec *recs;
    const double *Z, *Dr;
    double *S; int ld; double *bp, *bs;
    double *part;                 // [n_items][TP_PART] partial sums of tiles cut into several items
    unsigned *blk_done; int tbw1; const int *tile_pos;   // finished items per tile (never reset, see pair_kernel)
    unsigned *row_done; int n_free;                      // optional: finished items per camera row (chol_band_kernel runs beside this kernel)
    double *S2; int n_tot, n1;    // optional (two-way factorisation): elements with column >= n1 go to the mirrored system S2
    double *hpp_diag;             // optional [6 n_free]: diagonal-only pass of computeLambdaInit
};

// acc (6x6, row-major) += Z_x Z_y^T with Z = [P^T Y; Y], P = -[X]x:  G = Y_x Y_y^T,  H = G P_y,  Z_x Z_y^T = [P_x^T H, P_x^T G; H, G]
BA_DEV void tile_accumulate(const double *__restrict__ Z, int ex, int ey, double *acc) {
    const double *px = Z + ZR_STRIDE * (size_t)ex, *py = Z + ZR_STRIDE * (size_t)ey;
    double a[ZR_STRIDE], b[ZR_STRIDE];
#pragma unroll
    for (int i = 0; i < ZR_STRIDE / 4; i++) ldg256(py + 4 * i, b + 4 * i);
#pragma unroll
    for (int i = 0; i < ZR_STRIDE / 4; i++) ldg256(px + 4 * i, a + 4 * i);
    const double ax = a[9], ay = a[10], az = a[11], bx = b[9], by = b[10], bz = b[11];
    double G[9], H[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) G[3 * i + k] = fma(a[3 * i + 2], b[3 * k + 2], fma(a[3 * i + 1], b[3 * k + 1], a[3 * i] * b[3 * k]));
    // H[i][.] = P_y^T applied to row i of G: (-z u1 + y u2, z u0 - x u2, -y u0 + x u1)
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double u0 = G[3 * i], u1 = G[3 * i + 1], u2 = G[3 * i + 2];
        H[3 * i + 0] = fma(-bz, u1, by * u2);
        H[3 * i + 1] = fma(bz, u0, -(bx * u2));
        H[3 * i + 2] = fma(-by, u0, bx * u1);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        // rotation rows (0..2): P_x^T applied to the columns of H (left block) and of G (right block)
        acc[0 * 6 + c] = fma(ay, H[6 + c], fma(-az, H[3 + c], acc[0 * 6 + c]));
        acc[1 * 6 + c] = fma(-ax, H[6 + c], fma(az, H[c], acc[1 * 6 + c]));
        acc[2 * 6 + c] = fma(ax, H[3 + c], fma(-ay, H[c], acc[2 * 6 + c]));
        acc[0 * 6 + 3 + c] = fma(ay, G[6 + c], fma(-az, G[3 + c], acc[0 * 6 + 3 + c]));
        acc[1 * 6 + 3 + c] = fma(-ax, G[6 + c], fma(az, G[c], acc[1 * 6 + 3 + c]));
        acc[2 * 6 + 3 + c] = fma(ax, G[3 + c], fma(-ay, G[c], acc[2 * 6 + 3 + c]));
        // translation rows (3..5)
        acc[3 * 6 + c] += H[c]; acc[4 * 6 + c] += H[3 + c]; acc[5 * 6 + c] += H[6 + c];
        acc[3 * 6 + 3 + c] += G[c]; acc[4 * 6 + 3 + c] += G[3 + c]; acc[5 * 6 + 3 + c] += G[6 + c];
    }
}

// d (33) += the Hpp / b_p / b_s terms of observation e:  Dr = [N (6) | m (3) | q (3)], X_l from the Z record.
//   w B^T B = [P | I]^T N [P | I],  B^T g = [P^T m; m],  -Z L^-1 b_l = -[P^T q; q]
BA_DEV void tile_diag_accumulate(const double *__restrict__ Dr, const double *__restrict__ Z, int e, double *d) {
    const double *pd = Dr + DR_STRIDE * (size_t)e;
    double v[DR_STRIDE], zx[4];
#pragma unroll
    for (int q = 0; q < DR_STRIDE / 4; q++) ldg256(pd + 4 * q, v + 4 * q);
    ldg256(Z + ZR_STRIDE * (size_t)e + 8, zx);
    const double n00 = v[0], n01 = v[1], n02 = v[2], n11 = v[3], n12 = v[4], n22 = v[5], x = zx[1], y = zx[2], z = zx[3];
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
    d[21] += PT0(x, y, z, v[6], v[7], v[8]); d[22] += PT1(x, y, z, v[6], v[7], v[8]); d[23] += PT2(x, y, z, v[6], v[7], v[8]);
    d[24] += v[6]; d[25] += v[7]; d[26] += v[8];
    d[27] -= PT0(x, y, z, v[9], v[10], v[11]); d[28] -= PT1(x, y, z, v[9], v[10], v[11]); d[29] -= PT2(x, y, z, v[9], v[10], v[11]);
    d[30] -= v[9]; d[31] -= v[10]; d[32] -= v[11];
}

__global__ void __launch_bounds__(TP_THREADS, TP_MINB) pair_tile_kernel(TileArgs P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane >> 4, ia = (lane >> 2) & 3, ib = lane & 3;
    for (int it = blockIdx.x * TP_WARPS + warp; it < P.n_items; it += gridDim.x * TP_WARPS) {
        const TileItem I = P.items[it];
        const bool dt = I.ta == I.tb;
        if (P.hpp_diag && !dt) continue;
        const int ca = TP_T * I.ta + ia, cb = TP_T * I.tb + ib;                    // my block (lanes of the first half own the output)
        const bool blk_on = ca < P.n_free && cb < P.n_free && (!dt || ia <= ib);
        const bool multi = I.nit > 1;
        double *mine = P.part + (size_t)it * TP_PART;
        // ---- phase 1: the 16 blocks of the tile
        if (!P.hpp_diag) {
            double acc[36];
#pragma unroll
            for (int i = 0; i < 36; i++) acc[i] = 0.0;
            const int *ra = &P.recs[I.begin].ea[ia], *rb = &P.recs[I.begin].eb[ib];
            const int nr = I.end - I.begin;
            for (int r = h; r < nr; r += 2) {
                const int ea = __ldg(ra + 8 * r), eb = __ldg(rb + 8 * r);
                if (blk_on && ea >= 0 && eb >= 0) tile_accumulate(P.Z, ea, eb, acc);
            }
#pragma unroll
            for (int i = 0; i < 36; i++) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
            if (multi) {
                if (h == 0) {
#pragma unroll
                    for (int i = 0; i < 36; i++) __stcg(mine + 16 * i + (lane & 15), acc[i]);
                }
            } else if (h == 0 && blk_on) {
                const bool dg = ca == cb;
#pragma unroll
                for (int i = 0; i < 36; i++) {
                    const int r = i / 6, c = i - 6 * r;
                    if ((dg && c < r) || acc[i] == 0.0) continue;
                    const int R = 6 * ca + r, C = 6 * cb + c;
                    double *dst = (P.S2 && C >= P.n1) ? P.S2 + (size_t)(P.n_tot - 1 - C) * P.ld + (P.n_tot - 1 - R) : P.S + (size_t)R * P.ld + C;
                    atomicAdd(dst, -acc[i]);
                }
            }
        }
        // ---- phase 2 (diagonal tiles): Hpp, b_p, b_s of the tile's cameras. lane = (record of 8, camera)
        double d[33];
        const int dc = lane & 3, dcam = TP_T * I.ta + dc;
        if (dt) {
#pragma unroll
            for (int i = 0; i < 33; i++) d[i] = 0.0;
            const int *ra = &P.recs[I.begin].ea[dc], *rb = &P.recs[I.begin].eb[dc];
            const int nr = I.end - I.begin;
            for (int r = lane >> 2; r < nr; r += 8) {
                const int ea = __ldg(ra + 8 * r), eb = __ldg(rb + 8 * r);
                if (ea >= 0 && ea == eb) tile_diag_accumulate(P.Dr, P.Z, ea, d);          // the (G, G) record of a group holds each observation once
            }
#pragma unroll
            for (int i = 0; i < 33; i++) {
                d[i] += __shfl_xor_sync(0xffffffffu, d[i], 4);
                d[i] += __shfl_xor_sync(0xffffffffu, d[i], 8);
                d[i] += __shfl_xor_sync(0xffffffffu, d[i], 16);
            }
            if (multi && lane < 4) {
#pragma unroll
                for (int i = 0; i < 33; i++) __stcg(mine + 576 + 4 * i + lane, d[i]);
            }
        }
        // ---- tiles cut into several items: the LAST finisher adds the partials in item order (one total per element reaches S)
        if (multi) {
            __threadfence();
            __syncwarp();
            unsigned prev = 0;
            if (lane == 0) prev = atomicAdd(P.blk_done + (size_t)P.tile_pos[I.ta] * P.tbw1 + (I.tb - I.ta), 1u);
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if ((prev + 1u) % (unsigned)I.nit != 0u) continue;
            __threadfence();
            if (!P.hpp_diag && lane < 16) {
                const int ja = lane >> 2, jb = lane & 3;
                const int a2 = TP_T * I.ta + ja, b2 = TP_T * I.tb + jb;
                if (a2 < P.n_free && b2 < P.n_free && (!dt || ja <= jb)) {
                    double s[36];
#pragma unroll
                    for (int i = 0; i < 36; i++) s[i] = 0.0;
                    for (int k = 0; k < I.nit; k++) {
                        const double *pk = P.part + (size_t)(I.first + k) * TP_PART + lane;
#pragma unroll
                        for (int i = 0; i < 36; i++) s[i] += __ldcg(pk + 16 * i);
                    }
                    const bool dg = a2 == b2;
#pragma unroll
                    for (int i = 0; i < 36; i++) {
                        const int r = i / 6, c = i - 6 * r;
                        if ((dg && c < r) || s[i] == 0.0) continue;
                        const int R = 6 * a2 + r, C = 6 * b2 + c;
                        double *dst = (P.S2 && C >= P.n1) ? P.S2 + (size_t)(P.n_tot - 1 - C) * P.ld + (P.n_tot - 1 - R) : P.S + (size_t)R * P.ld + C;
                        atomicAdd(dst, -s[i]);
                    }
                }
            }
            if (dt && lane < 4) {
#pragma unroll
                for (int i = 0; i < 33; i++) d[i] = 0.0;
                for (int k = 0; k < I.nit; k++) {
                    const double *pk = P.part + (size_t)(I.first + k) * TP_PART + 576 + lane;
#pragma unroll
                    for (int i = 0; i < 33; i++) d[i] += __ldcg(pk + 4 * i);
                }
            }
        }
        if (dt && lane < 4 && dcam < P.n_free) {
            if (P.hpp_diag) {                                                        // computeLambdaInit: diagonal of Hpp only
                P.hpp_diag[6 * dcam + 0] = d[0]; P.hpp_diag[6 * dcam + 1] = d[6]; P.hpp_diag[6 * dcam + 2] = d[11];
                P.hpp_diag[6 * dcam + 3] = d[15]; P.hpp_diag[6 * dcam + 4] = d[18]; P.hpp_diag[6 * dcam + 5] = d[20];
            } else {
#pragma unroll
                for (int i = 0; i < 33; i++) {
                    if (d[i] == 0.0) continue;
                    if (i < 21) {
                        int r = 0, t = i;
                        while (t >= 6 - r) { t -= 6 - r; r++; }
                        const int R = 6 * dcam + r, C = R + t;
                        double *dst = (P.S2 && C >= P.n1) ? P.S2 + (size_t)(P.n_tot - 1 - C) * P.ld + (P.n_tot - 1 - R) : P.S + (size_t)R * P.ld + C;
                        atomicAdd(dst, d[i]);
                    } else if (i < 27) atomicAdd(P.bp + 6 * dcam + (i - 21), d[i]);
                    else atomicAdd(P.bs + 6 * dcam + (i - 27), d[i]);
                }
            }
        }
        if (P.row_done && !P.hpp_diag) {
            __threadfence();
            __syncwarp();
            if (lane < 4 && dcam < P.n_free) atomicAdd(P.row_done + dcam, (unsigned)I.nit);
        }
    }
}
