"""CPU tests of the tile-major Schur pass (csrc/schur_tiles.cuh), restated in numpy:
  * the (tile, landmark) records of tile_plan_kernel -- groups = (4-camera tile, layer), one record per ordered pair of groups, lanes
    ia <= ib on a diagonal tile -- enumerate exactly the observation pairs the pair list of pair_plan enumerates (every (e_a, e_b) with
    camera(a) <= camera(b) once, the couples of one camera both ways), incl. several edges on one (pose, point) pair and fixed poses;
  * the tensor-pipe formulation of pair_tile_mma_kernel: Z_A Z_B^T over the 24 x 24 tile with zero rows for absent cameras, read back
    through the fragment -> block mapping of emit_tile, equals the per-block sums  sum_j Z_aj Z_bj^T  (block_solver.hpp:400-444);
  * the operand buffer layout (k-major, one pad row per camera, padded side stride): every (side, k, row) has its own address and the
    expansion stores / fragment loads of a half warp fall into distinct banks where the header says so.
No GPU needed."""
import collections

import numpy as np

from orb_slam3_study_kr_b200 import synthetic

T = 4                       # TP_T
TM_LDK, TM_SIDE = 28, 12 * 28 + 8


def sorted_view(p):
    """Observations landmark-major, pose-ascending, insertion order among equals (what bagpu_upload establishes)."""
    order = np.lexsort((np.arange(p.n_obs), p.obs_pose, p.obs_point))
    op, opt = p.obs_pose[order], p.obs_point[order]
    fixed = np.asarray(p.pose_fixed).astype(bool)
    hidx = np.cumsum(~fixed) - 1
    hidx[fixed] = -1
    ptr = np.flatnonzero(np.r_[True, opt[1:] != opt[:-1], True])
    return op, opt, hidx, ptr


def pair_list(op, hidx, p0, p1):
    """pair_count_kernel / pair_gen_kernel + pair_kernel's rule for the diagonal block (schur_pairs.cuh)."""
    out = collections.Counter()
    for e in range(p0, p1):
        ha = hidx[op[e]]
        if ha < 0:
            continue
        for e2 in range(e, p1):
            hb = hidx[op[e2]]
            if hb < 0:
                continue
            out[(e, e2)] += 1
            if ha == hb and e != e2:
                out[(e2, e)] += 1           # two edges on one (pose, point) pair: M + M^T
    return out


def tile_records(op, hidx, p0, p1):
    """tile_plan_kernel: layer = how many earlier edges the same pose has on this landmark; group = (tile, layer); the leader of
    group G owns the records (G, G') for every group G' in a tile >= tile(G) (all groups of its own tile included)."""
    lay, groups = 0, collections.OrderedDict()
    for q in range(p0, p1):
        lay = lay + 1 if q > p0 and op[q] == op[q - 1] else 0
        h = hidx[op[q]]
        if h < 0:
            continue
        groups.setdefault((h // T, lay), [-1] * T)[h % T] = q
    recs = []
    for gx, A in groups.items():
        for gy, B in groups.items():
            if gy[0] >= gx[0]:
                recs.append((gx[0], gy[0], A, B))
    return recs


def pairs_of_records(recs):
    """pair_tile_kernel / pair_tile_mma_kernel + emit_tile: lane (ia, ib), ia <= ib on a diagonal tile."""
    out = collections.Counter()
    for ta, tb, A, B in recs:
        for ia in range(T):
            for ib in range(T):
                if ta == tb and ia > ib:
                    continue
                if A[ia] >= 0 and B[ib] >= 0:
                    out[(A[ia], B[ib])] += 1
    return out


def with_duplicates(p, copies, rng):
    idx = np.concatenate([np.arange(p.n_obs), rng.choice(p.n_obs, copies)])
    from orb_slam3_study_kr_b200.problem import BAProblem
    return BAProblem(p.pose_qt, p.pose_fixed, p.points, p.cameras, p.rigs, p.obs_pose[idx], p.obs_point[idx], p.obs_cam[idx],
                     p.obs_rig[idx], p.obs_kind[idx], p.obs_flags[idx], p.obs_u[idx], p.obs_v[idx], p.obs_ur[idx], p.obs_inv_sigma2[idx])


def test_tile_records_enumerate_the_pairs_of_the_pair_list():
    rng = np.random.default_rng(3)
    cases = [synthetic.config(1, 0.04), synthetic.config(3, 0.04),                       # config 3: left + right eye of a rig keyframe
             with_duplicates(synthetic.config(2, 0.02), 400, rng)]                        # up to several edges on one (pose, point) pair
    n_rec = n_pair = 0
    for p in cases:
        op, opt, hidx, ptr = sorted_view(p)
        for li in range(len(ptr) - 1):
            p0, p1 = ptr[li], ptr[li + 1]
            recs = tile_records(op, hidx, p0, p1)
            want, got = pair_list(op, hidx, p0, p1), pairs_of_records(recs)
            assert want == got, (li, want - got, got - want)
            n_rec += len(recs); n_pair += sum(want.values())
    assert n_pair > 4 * n_rec > 0                                                        # the point of the records: several pairs per record (7.4 on config 4)


def zfull(Y, X):
    x, y, z = X
    P = np.array([[0.0, z, -y], [-z, 0.0, x], [y, -x, 0.0]])                             # P = -[X]x
    return np.vstack([P.T @ Y, Y])                                                       # [P^T Y; Y], 6 x 3


def test_tensor_pipe_formulation_gives_the_block_sums():
    """Records of one tile pair -> operands with zero rows -> 24 x 24 product -> blocks through emit_tile's mapping."""
    rng = np.random.default_rng(5)
    n_rec = 37
    for diag in (False, True):
        recA = [[int(rng.integers(0, 2)) for _ in range(T)] for _ in range(n_rec)]       # which cameras of the tile see the landmark
        recB = recA if diag else [[int(rng.integers(0, 2)) for _ in range(T)] for _ in range(n_rec)]
        ZA = [[zfull(rng.normal(size=(3, 3)), rng.normal(size=3) * 4) if on else None for on in r] for r in recA]
        ZB = ZA if diag else [[zfull(rng.normal(size=(3, 3)), rng.normal(size=3) * 4) if on else None for on in r] for r in recB]
        want = np.zeros((T, T, 6, 6))
        for j in range(n_rec):
            for ia in range(T):
                for ib in range(T):
                    if ZA[j][ia] is not None and ZB[j][ib] is not None and (not diag or ia <= ib):
                        want[ia, ib] += ZA[j][ia] @ ZB[j][ib].T
        # four records = K 12; an absent observation is a zero block of the operand
        tile = np.zeros((24, 24))
        for g0 in range(0, n_rec, 4):
            A = np.zeros((24, 12)); B = np.zeros((24, 12))
            for r in range(min(4, n_rec - g0)):
                for i in range(T):
                    if ZA[g0 + r][i] is not None:
                        A[6 * i:6 * i + 6, 3 * r:3 * r + 3] = ZA[g0 + r][i]
                    if ZB[g0 + r][i] is not None:
                        B[6 * i:6 * i + 6, 3 * r:3 * r + 3] = ZB[g0 + r][i]
            for ks in range(3):                                                          # three k-steps of 8x8x4 products over the nine 8x8 tiles
                for TA in range(3):
                    for TB in range(3):
                        if diag and TA > TB:
                            continue                                                     # the kernel skips the lower 8x8 tiles of a diagonal tile
                        tile[8 * TA:8 * TA + 8, 8 * TB:8 * TB + 8] += A[8 * TA:8 * TA + 8, 4 * ks:4 * ks + 4] @ B[8 * TB:8 * TB + 8, 4 * ks:4 * ks + 4].T
        # emit_tile: lane (fr, fk) of 8x8 tile (TA, TB) holds element (8 TA + fr, 8 TB + 2 fk + q)
        got = np.zeros((T, T, 6, 6))
        for TA in range(3):
            for TB in range(3):
                if diag and TA > TB:
                    continue
                for fr in range(8):
                    for fk in range(4):
                        for q in range(2):
                            R24, C24 = 8 * TA + fr, 8 * TB + 2 * fk + q
                            ja, r, jb, c = R24 // 6, R24 % 6, C24 // 6, C24 % 6
                            if diag and ja > jb:
                                continue
                            got[ja, jb, r, c] = tile[R24, C24]
        for ia in range(T):
            for ib in range(T):
                if diag and ia > ib:
                    continue
                if diag and ia == ib:                                                    # only the upper triangle of a diagonal block reaches S
                    assert np.allclose(np.triu(got[ia, ib]), np.triu(want[ia, ib]), rtol=1e-12, atol=1e-12)
                else:
                    assert np.allclose(got[ia, ib], want[ia, ib], rtol=1e-12, atol=1e-12), (diag, ia, ib)


def test_operand_buffer_layout():
    """Element (side, k, row) at side * TM_SIDE + k * TM_LDK + row + row // 6: all distinct; bank checks of one half warp (16 lanes,
    8-byte accesses = 2 banks each, 32 banks)."""
    addr = {}
    for side in range(2):
        for k in range(12):
            for row in range(24):
                a = side * TM_SIDE + k * TM_LDK + row + row // 6
                assert a not in addr and a < 2 * TM_SIDE
                addr[a] = (side, k, row)
    # expansion store of (d, comp): lanes = (record, side, camera); a half warp = records {0, 1} or {2, 3}
    for d in range(3):
        for comp in range(6):
            for recs in ((0, 1), (2, 3)):
                banks = [((side * TM_SIDE + (3 * rec + d) * TM_LDK + 7 * i + comp) * 2) % 32 for rec in recs for side in range(2) for i in range(4)]
                assert len(set(banks)) == 16, (d, comp, recs)
    # fragment load of (ks, IA): lanes = (fr, fk); half warps fr 0-3 / 4-7. Free of conflicts unless the four rows straddle a camera boundary
    clean = 0
    for ks in range(3):
        for IA in range(3):
            for half in range(2):
                rows = [8 * IA + 4 * half + f for f in range(4)]
                banks = [(((4 * ks + fk) * TM_LDK + r + r // 6) * 2) % 32 for r in rows for fk in range(4)]
                straddles = len({r // 6 for r in rows}) > 1
                if not straddles:
                    assert len(set(banks)) == 16, (ks, IA, half)
                    clean += 1
    assert clean == 3 * 4                                                                # 4 of the 6 half-warp loads per k-step are conflict-free
