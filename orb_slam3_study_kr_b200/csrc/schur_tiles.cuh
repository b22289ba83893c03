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
#define TP_CHUNK 256                       // most records per work item (the plan uses 128 for small maps: more items, shorter chains)
#endif
#define TP_THREADS 128
#define TP_WARPS (TP_THREADS / 32)
#ifndef TP_MINB
#define TP_MINB 3
#endif
#ifndef TP_STAGES
#define TP_STAGES 6                        // depth of the per-warp cp.async ring (steps in flight)
#endif
#define TP_WARP_SMEM (TP_CHUNK * 32 + TP_STAGES * 16 * 96)     // the item's records | TP_STAGES x 16 Z records
#define TP_SMEM_BYTES (TP_WARPS * TP_WARP_SMEM)
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

__global__ void tile_item_count_kernel(int nblk, const unsigned *__restrict__ blk_cnt, unsigned *__restrict__ item_cnt, unsigned chunk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nblk) item_cnt[i] = (blk_cnt[i] + chunk - 1) / chunk;
}

__global__ void tile_item_fill_kernel(int nblk, int tbw1, const int *__restrict__ tile_of_pos, const unsigned *__restrict__ blk_off, const unsigned *__restrict__ blk_cnt,
                                      const unsigned *__restrict__ item_off, TileItem *__restrict__ items, unsigned chunk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblk) return;
    const unsigned cnt = blk_cnt[i];
    if (!cnt) return;
    const int tp = i / tbw1, ta = tile_of_pos[tp], tb = ta + (i - tp * tbw1);
    const unsigned off = blk_off[i];
    unsigned io = item_off[i];
    const int nit = (int)((cnt + chunk - 1) / chunk);
    for (unsigned c = 0; c < cnt; c += chunk) {
        TileItem I; I.ta = ta; I.tb = tb; I.begin = (int)(off + c); I.end = (int)(off + min(cnt, c + chunk));
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
    unsigned *work;               // [4] next item / warps that ran out of items, per kernel (zero at launch; the last warp out resets them)
    const LmDev *lm = nullptr;    // chained mode: nothing to do once the LM loop has ended
    int item_begin = 0, item_end = -1;   // pair_tile_mma_kernel: the items of this launch (-1: all); a multi-GPU trial launches the rows chunk by chunk
};

BA_DEV void cp_async16(void *smem_dst, const void *gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
BA_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> BA_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// acc (6x6, row-major) += Z_x Z_y^T with Z = [P^T Y; Y], P = -[X]x:  G = Y_x Y_y^T,  H = G P_y,  Z_x Z_y^T = [P_x^T H, P_x^T G; H, G]
// a, b: the two 96-byte records (shared memory)
BA_DEV void tile_accumulate(const double *__restrict__ pa, const double *__restrict__ pb, double *acc) {
    double a[ZR_STRIDE], b[ZR_STRIDE];
#pragma unroll
    for (int i = 0; i < ZR_STRIDE / 2; i++) {
        const double2 va = reinterpret_cast<const double2 *>(pa)[i], vb = reinterpret_cast<const double2 *>(pb)[i];
        a[2 * i] = va.x; a[2 * i + 1] = va.y; b[2 * i] = vb.x; b[2 * i + 1] = vb.y;
    }
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

// d (33) += the Hpp / b_p / b_s terms of observation e:  Dr = [N (6) | m (3) | q (3) | X_l (3)]
//   w B^T B = [P | I]^T N [P | I],  B^T g = [P^T m; m],  -Z L^-1 b_l = -[P^T q; q]
BA_DEV void tile_diag_add(const double *v, double *d);
BA_DEV void tile_diag_accumulate(const double *__restrict__ Dr, int e, double *d) {
    const double *pd = Dr + DR_STRIDE * (size_t)e;
    double v[DR_STRIDE];
#pragma unroll
    for (int q = 0; q < DR_STRIDE / 4; q++) ldg256(pd + 4 * q, v + 4 * q);
    tile_diag_add(v, d);
}
// v = [N (6) | m (3) | q (3) | X_l (3) | -]
BA_DEV void tile_diag_add(const double *v, double *d) {
    const double *zx = v + DR_STRIDE - 5;                      // zx[1..3] = X_l
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
    extern __shared__ __align__(16) unsigned char tp_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = lane >> 4, ia = (lane >> 2) & 3, ib = lane & 3;
    int *srec = reinterpret_cast<int *>(tp_sm + (size_t)warp * TP_WARP_SMEM);                 // the item's records: [record][ea 0..3 | eb 0..3]
    double *ring = reinterpret_cast<double *>(tp_sm + (size_t)warp * TP_WARP_SMEM + TP_CHUNK * 32);   // [stage][row: half h, ea 0..3 | eb 0..3][12]
    // Items are taken from a counter, in plan order: the items of a diagonal tile (full chunks + the Dr pass) cost several times what a
    // sparse off-diagonal tile's do, and a fixed stride can resonate with the period of a tile row (config 4: a few warps ran 3x longer
    // than the average).
    for (;;) {
        int it = 0;
        if (lane == 0) it = (int)atomicAdd(P.work, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= P.n_items) break;
        const TileItem I = P.items[it];
        const bool dt = I.ta == I.tb;
        if (P.hpp_diag && !dt) continue;
        const int ca = TP_T * I.ta + ia, cb = TP_T * I.tb + ib;                    // my block (lanes of the first half own the output)
        const bool blk_on = ca < P.n_free && cb < P.n_free && (!dt || ia <= ib);
        const bool multi = I.nit > 1;
        const int nr = I.end - I.begin;
        double *mine = P.part + (size_t)it * TP_PART;
        __syncwarp();
        {
            const int4 *src = reinterpret_cast<const int4 *>(P.recs + I.begin);
            int4 *dst = reinterpret_cast<int4 *>(srec);
            for (int k = lane; k < 2 * nr; k += 32) dst[k] = __ldg(src + k);
        }
        __syncwarp();
        // ---- phase 1: the 16 blocks of the tile. Two records per step; the 16 Z records a step needs are fetched once, by
        // cp.async, TP_STAGES steps ahead of their use (two lanes per record), so the bytes in flight are not bounded by registers.
        if (!P.hpp_diag) {
            double acc[36];
#pragma unroll
            for (int i = 0; i < 36; i++) acc[i] = 0.0;
            const int nsteps = (nr + 1) >> 1;
            const int q = lane >> 1, hf = lane & 1;
            auto issue = [&](int s, int stage) {
                if (s < nsteps && 2 * s + (q >> 3) < nr) {
                    const int idx = srec[16 * s + q];
                    if (idx >= 0) {
                        const double *src = P.Z + ZR_STRIDE * (size_t)idx + 6 * hf;
                        double *dst = ring + (size_t)(stage * 16 + q) * ZR_STRIDE + 6 * hf;
                        cp_async16(dst, src); cp_async16(dst + 2, src + 2); cp_async16(dst + 4, src + 4);
                    }
                }
                cp_async_commit();
            };
#pragma unroll
            for (int s = 0; s < TP_STAGES - 1; s++) issue(s, s);
            int st_use = 0, st_fill = TP_STAGES - 1;
            for (int s = 0; s < nsteps; s++) {
                issue(s + TP_STAGES - 1, st_fill);
                cp_async_wait<TP_STAGES - 1>();
                __syncwarp();
                const int r = 2 * s + h;
                if (r < nr) {
                    const int ea = srec[8 * r + ia], eb = srec[8 * r + 4 + ib];
                    if (blk_on && ea >= 0 && eb >= 0)
                        tile_accumulate(ring + (size_t)(st_use * 16 + 8 * h + ia) * ZR_STRIDE, ring + (size_t)(st_use * 16 + 8 * h + 4 + ib) * ZR_STRIDE, acc);
                }
                __syncwarp();
                st_use = (st_use + 1 == TP_STAGES) ? 0 : st_use + 1;
                st_fill = (st_fill + 1 == TP_STAGES) ? 0 : st_fill + 1;
            }
            cp_async_wait<0>();
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
            for (int r = lane >> 2; r < nr; r += 8) {
                const int ea = srec[8 * r + dc], eb = srec[8 * r + 4 + dc];
                if (ea >= 0 && ea == eb) tile_diag_accumulate(P.Dr, ea, d);          // the (G, G) record of a group holds each observation once
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
    if (lane == 0) {
        const unsigned out = atomicAdd(P.work + 1, 1u);
        if (out == gridDim.x * TP_WARPS - 1) { P.work[0] = 0u; P.work[1] = 0u; __threadfence(); }     // every warp has made its last grab
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// pair_tile_mma_kernel: the same tile contraction on the FP64 tensor pipe (mma.sync.m8n8k4.f64 = DMMA).
//
// Seen as a matrix product, a record adds  Z_A Z_B^T  to the 24 x 24 tile, with Z_A = the 6 x 3 blocks [P^T Y; Y] of the tile's four
// cameras stacked (zero rows for a camera that does not see the landmark) and K = 3. Four records make K = 12 = three k-steps of
// nine 8x8x4 DMMAs (six on a diagonal tile: the lower 8x8 tiles are not needed). Per group of four records a lane
//   * fetches ONE raw Z record (its slot: record, side, camera) two groups ahead, into one of three register buffers,
//   * expands it to the 6 x 3 block (18 FMAs) and stores it into the group's operand buffer in shared memory (double-buffered),
//   * issues 27 (18) DMMAs on fragments read back from that buffer.
// Against the FMA kernel above: 6.75 DMMA + 4.5 FP64 instructions per record instead of ~70 per record and half warp, 18 accumulator
// registers instead of 72 (more warps per SM), and no idle lanes: an absent observation is a zero row, not a predicated-off lane.
// The tensor pipe adds the k products of an instruction in a fixed order, so the result is reproducible like every other sum here.
// The Dr sums of the diagonal tiles (Hpp, b_p, b_s) are a kernel of their own (tile_diag_kernel), launched before this one.
//
// Operand buffer: element (side, k, row) at side * TM_SIDE + k * TM_LDK + row + row / 6 (k-major, one pad row per camera): the
// expansion stores (lanes = record x side x camera) and the fragment loads (lanes = row x k) are both free of bank conflicts
// except for the half warps whose four rows straddle a camera boundary.
#define TM_LDK 28
#define TM_SIDE (12 * TM_LDK + 8)
#define TM_BUF (2 * TM_SIDE)
#define TM_WARP_SMEM (2 * TM_BUF * 8)      // two buffers
#define TM_SMEM_BYTES (TP_WARPS * TM_WARP_SMEM)
#ifndef TM_MINB
#define TM_MINB 3                          // 168 registers, no spills (4 CTAs at 128 registers spill and run slower)
#endif

BA_DEV void tp_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(TP_THREADS, TM_MINB) pair_tile_mma_kernel(TileArgs P) {
    extern __shared__ __align__(16) unsigned char tp_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *E = reinterpret_cast<double *>(tp_sm + (size_t)warp * TM_WARP_SMEM);
    const int s_rec = lane >> 3, s_col = lane & 7, s_side = s_col >> 2, s_i = s_col & 3;     // my slot of a group: record, ea[i] / eb[i]
    const int fr = lane >> 2, fk = lane & 3;                                                  // fragment row / k (and output row / column pair)
    const int st_off = s_side * TM_SIDE + 3 * s_rec * TM_LDK + 7 * s_i;                        // my 6 x 3 block: + d * TM_LDK + comp
    const int f0 = fk * TM_LDK + fr + fr / 6, f1 = fk * TM_LDK + (8 + fr) + (8 + fr) / 6, f2 = fk * TM_LDK + (16 + fr) + (16 + fr) / 6;
    const int *rec_int = reinterpret_cast<const int *>(P.recs);
    if (lm_done(P.lm)) return;
    for (;;) {
        int it = 0;
        if (lane == 0) it = P.item_begin + (int)atomicAdd(P.work, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= (P.item_end < 0 ? P.n_items : P.item_end)) break;
        const TileItem I = P.items[it];
        const bool dt = I.ta == I.tb;
        const bool multi = I.nit > 1;
        const int nr = I.end - I.begin;
        double c[9][2];
        // the tile's totals: one add per element of S (two on the diagonal blocks: the Dr sums of tile_diag_kernel), exact in any order
        auto emit_tile = [&]() {
#pragma unroll
            for (int t = 0; t < 9; t++) {
                const int TA = t / 3, TB = t - 3 * TA;
                if (dt && TA > TB) continue;
                const int R24 = 8 * TA + fr, ja = R24 / 6, r = R24 - 6 * ja;
                const int a2 = TP_T * I.ta + ja;
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int C24 = 8 * TB + 2 * fk + q, jb = C24 / 6, cc = C24 - 6 * jb;
                    const int b2 = TP_T * I.tb + jb;
                    const double v = c[t][q];
                    if (a2 >= P.n_free || b2 >= P.n_free || (dt && ja > jb) || (a2 == b2 && cc < r) || v == 0.0) continue;
                    const int R = 6 * a2 + r, C = 6 * b2 + cc;
                    double *dst = (P.S2 && C >= P.n1) ? P.S2 + (size_t)(P.n_tot - 1 - C) * P.ld + (P.n_tot - 1 - R) : P.S + (size_t)R * P.ld + C;
                    atomicAdd(dst, -v);
                }
            }
        };
#pragma unroll
        for (int t = 0; t < 9; t++) { c[t][0] = 0.0; c[t][1] = 0.0; }
        const int ngroups = (nr + 3) >> 2;
        const int *my_int = rec_int + (size_t)I.begin * 8 + 8 * s_rec + s_col;
        auto load_idx = [&](int g) -> int { return (4 * g + s_rec < nr) ? __ldg(my_int + 32 * g) : -1; };
        auto load_row = [&](int idx, double *raw) {                // an absent observation is a record of zeros (its rows of the operand are zero)
            if (idx >= 0) {
                const double *src = P.Z + ZR_STRIDE * (size_t)idx;
                ldg256(src, raw); ldg256(src + 4, raw + 4); ldg256(src + 8, raw + 8);
            } else {
#pragma unroll
                for (int i = 0; i < ZR_STRIDE; i++) raw[i] = 0.0;
            }
        };
        // group g: expand `cur` (fetched two groups ago), start the fetch of group g + 2 into `fill`, 27 / 18 DMMAs
        int idx_q = -1;                                        // index of my slot in group g + 2 (loaded one group earlier)
        // (skipping the DMMAs of all-zero 8-row operand blocks -- warp-uniform branches on a ballot of the group's slots -- was measured:
        // 3.84 ms instead of 2.35 ms per launch on config 5; the predicated chain no longer overlaps fragment loads and tensor instructions)
        auto group = [&](int g, const double *cur, double *fill) {
            double *buf = E + (g & 1) * TM_BUF + st_off;
            {   // rows 7 i .. 7 i + 5 of my side, columns 3 rec .. 3 rec + 2;  [X]x y = (-z y1 + y y2, z y0 - x y2, -y y0 + x y1) per column y of Y
                const double x = cur[9], y = cur[10], z = cur[11];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double u0 = cur[d], u1 = cur[3 + d], u2 = cur[6 + d];   // column d of Y (row-major: Y[i][d] = cur[3 i + d])
                    double *o = buf + d * TM_LDK;
                    o[0] = fma(y, u2, -(z * u1));
                    o[1] = fma(z, u0, -(x * u2));
                    o[2] = fma(x, u1, -(y * u0));
                    o[3] = u0; o[4] = u1; o[5] = u2;
                }
            }
            load_row(idx_q, fill);
            idx_q = load_idx(g + 3);
            __syncwarp();
            const double *ea = E + (g & 1) * TM_BUF, *eb = ea + TM_SIDE;
#pragma unroll
            for (int ks = 0; ks < 3; ks++) {
                const int ko = 4 * ks * TM_LDK;
                const double a0 = ea[ko + f0], a1 = ea[ko + f1], a2 = ea[ko + f2];
                const double b0 = eb[ko + f0], b1 = eb[ko + f1], b2 = eb[ko + f2];
                tp_dmma(c[0][0], c[0][1], a0, b0); tp_dmma(c[1][0], c[1][1], a0, b1); tp_dmma(c[2][0], c[2][1], a0, b2);
                tp_dmma(c[4][0], c[4][1], a1, b1); tp_dmma(c[5][0], c[5][1], a1, b2); tp_dmma(c[8][0], c[8][1], a2, b2);
                if (!dt) { tp_dmma(c[3][0], c[3][1], a1, b0); tp_dmma(c[6][0], c[6][1], a2, b0); tp_dmma(c[7][0], c[7][1], a2, b1); }
            }
        };
        double r0[ZR_STRIDE], r1[ZR_STRIDE], r2[ZR_STRIDE];
        {
            const int i0 = load_idx(0), i1 = load_idx(1);
            idx_q = load_idx(2);
            load_row(i0, r0); load_row(i1, r1);
        }
        for (int g = 0; g < ngroups; g += 3) {
            group(g, r0, r2);
            if (g + 1 < ngroups) group(g + 1, r1, r0);
            if (g + 2 < ngroups) group(g + 2, r2, r1);
        }
        __syncwarp();
        if (!multi) { emit_tile(); }
        else {
            // tiles cut into several items: the LAST finisher adds the partials in item order (one total per element reaches S)
            double *mine = P.part + (size_t)it * TP_PART;
#pragma unroll
            for (int t = 0; t < 9; t++) { __stcg(mine + (2 * t) * 32 + lane, c[t][0]); __stcg(mine + (2 * t + 1) * 32 + lane, c[t][1]); }
            __threadfence();
            __syncwarp();
            unsigned prev = 0;
            if (lane == 0) prev = atomicAdd(P.blk_done + (size_t)P.tile_pos[I.ta] * P.tbw1 + (I.tb - I.ta), 1u);
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if ((prev + 1u) % (unsigned)I.nit != 0u) continue;
            __threadfence();
#pragma unroll
            for (int t = 0; t < 9; t++) { c[t][0] = 0.0; c[t][1] = 0.0; }
            // (the adds stay in item order; four items' loads are in flight at a time)
            int k = 0;
            for (; k + 4 <= I.nit; k += 4) {
                double v[4][18];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const double *pk = P.part + (size_t)(I.first + k + u) * TP_PART + lane;
#pragma unroll
                    for (int t = 0; t < 18; t++) v[u][t] = __ldcg(pk + t * 32);
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int t = 0; t < 9; t++) { c[t][0] += v[u][2 * t]; c[t][1] += v[u][2 * t + 1]; }
            }
            for (; k < I.nit; k++) {
                const double *pk = P.part + (size_t)(I.first + k) * TP_PART + lane;
#pragma unroll
                for (int t = 0; t < 9; t++) { c[t][0] += __ldcg(pk + (2 * t) * 32); c[t][1] += __ldcg(pk + (2 * t + 1) * 32); }
            }
            emit_tile();
        }
        if (P.row_done) {
            __threadfence();
            __syncwarp();
            const int cam = TP_T * I.ta + lane;
            if (lane < TP_T && cam < P.n_free) atomicAdd(P.row_done + cam, (unsigned)I.nit);
        }
    }
    if (lane == 0) {
        const unsigned out = atomicAdd(P.work + 1, 1u);
        if (out == gridDim.x * TP_WARPS - 1) { P.work[0] = 0u; P.work[1] = 0u; __threadfence(); }
    }
}

// Hpp, b_p, b_s of every camera (and, for computeLambdaInit, the diagonal of Hpp only): the Dr records of the diagonal tiles' items,
// lane = (record of 8, camera). Runs BEFORE pair_tile_mma_kernel on the same stream (both count a tile's finished items in blk_done:
// a launch adds nit per tile, so the counters stay multiples of nit between launches).
__global__ void __launch_bounds__(TP_THREADS, 3) tile_diag_kernel(TileArgs P) {
    const int lane = threadIdx.x & 31;
    const int *rec_int = reinterpret_cast<const int *>(P.recs);
    if (lm_done(P.lm)) return;
    for (;;) {
        int it = 0;
        if (lane == 0) it = (int)atomicAdd(P.work + 2, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= P.n_items) break;
        const TileItem I = P.items[it];
        if (I.ta != I.tb) continue;
        const bool multi = I.nit > 1;
        const int nr = I.end - I.begin;
        double d[33];
        const int dc = lane & 3, dcam = TP_T * I.ta + dc;
#pragma unroll
        for (int i = 0; i < 33; i++) d[i] = 0.0;
        const int *rp = rec_int + (size_t)I.begin * 8;
        // the (G, G) record of a group holds each of its observations once. Index loads run four records ahead of the Dr gathers, and
        // two gathers are in flight per lane: the kernel is a chain of dependent loads otherwise.
        auto obs_of = [&](int r) -> int {
            if (r >= nr) return -1;
            const int ea = __ldg(rp + 8 * r + dc), eb = __ldg(rp + 8 * r + 4 + dc);
            return (ea == eb) ? ea : -1;
        };
        int e0 = obs_of(lane >> 2), e1 = obs_of((lane >> 2) + 8);
        for (int r = lane >> 2; r < nr; r += 16) {
            const int n0 = obs_of(r + 16), n1 = obs_of(r + 24);
            double v0[DR_STRIDE], v1[DR_STRIDE];
            if (e0 >= 0) { const double *pd = P.Dr + DR_STRIDE * (size_t)e0; ldg256(pd, v0); ldg256(pd + 4, v0 + 4); ldg256(pd + 8, v0 + 8); ldg256(pd + 12, v0 + 12); }
            if (e1 >= 0) { const double *pd = P.Dr + DR_STRIDE * (size_t)e1; ldg256(pd, v1); ldg256(pd + 4, v1 + 4); ldg256(pd + 8, v1 + 8); ldg256(pd + 12, v1 + 12); }
            if (e0 >= 0) tile_diag_add(v0, d);
            if (e1 >= 0) tile_diag_add(v1, d);
            e0 = n0; e1 = n1;
        }
#pragma unroll
        for (int i = 0; i < 33; i++) {
            d[i] += __shfl_xor_sync(0xffffffffu, d[i], 4);
            d[i] += __shfl_xor_sync(0xffffffffu, d[i], 8);
            d[i] += __shfl_xor_sync(0xffffffffu, d[i], 16);
        }
        if (multi) {
            double *mine = P.part + (size_t)it * TP_PART;
            if (lane < 4) {
#pragma unroll
                for (int i = 0; i < 33; i++) __stcg(mine + 576 + 4 * i + lane, d[i]);
            }
            __threadfence();
            __syncwarp();
            unsigned prev = 0;
            if (lane == 0) prev = atomicAdd(P.blk_done + (size_t)P.tile_pos[I.ta] * P.tbw1, 1u);
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if ((prev + 1u) % (unsigned)I.nit != 0u) continue;
            __threadfence();
            // lane (slice, camera) adds every eighth item, then the slices are added in a fixed order (item order inside a slice)
#pragma unroll
            for (int i = 0; i < 33; i++) d[i] = 0.0;
            for (int k = lane >> 2; k < I.nit; k += 8) {
                const double *pk = P.part + (size_t)(I.first + k) * TP_PART + 576 + dc;
#pragma unroll
                for (int i = 0; i < 33; i++) d[i] += __ldcg(pk + 4 * i);
            }
#pragma unroll
            for (int i = 0; i < 33; i++) {
                d[i] += __shfl_xor_sync(0xffffffffu, d[i], 4);
                d[i] += __shfl_xor_sync(0xffffffffu, d[i], 8);
                d[i] += __shfl_xor_sync(0xffffffffu, d[i], 16);
            }
        }
        if (lane < 4 && dcam < P.n_free) {
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
    }
    if (lane == 0) {
        const unsigned out = atomicAdd(P.work + 3, 1u);
        if (out == gridDim.x * TP_WARPS - 1) { P.work[2] = 0u; P.work[3] = 0u; __threadfence(); }
    }
}
