"""Seeded synthetic maps of the BASELINE.json shapes.

Nothing here reads the reference at run time. Intrinsics are the values of the reference's
example YAMLs, written out as literals:
  Examples/Monocular/TUM1.yaml        fx 517.306408 fy 516.469215 cx 318.643040 cy 255.313989
  Examples/Stereo/EuRoC.yaml          fx 458.654 fy 457.296 cx 367.215 cy 248.375, b 0.110074
  Examples/Stereo/TUM-VI.yaml         Camera1/Camera2 KB8 + Stereo.T_c1_c2
Everything that crosses the reference's float boundary (pose quaternion/translation, MapPoint
position, keypoint pixel, mvuRight, mvInvLevelSigma2, camera parameters) is rounded through
float32 first, as Optimizer.cc does.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation

from .problem import (BAProblem, PoseBatch, CAM_KB8, CAM_PINHOLE, EDGE_BODY, EDGE_MONO, EDGE_STEREO,
                      FLAG_ROBUST, inv_level_sigma2)

TUM1 = dict(type=CAM_PINHOLE, p=[517.306408, 516.469215, 318.643040, 255.313989], bf=0.0, w=640, h=480)
EUROC = dict(type=CAM_PINHOLE, p=[458.654, 457.296, 367.215, 248.375], bf=458.654 * 0.110074, w=752, h=480)
TUMVI_L = dict(type=CAM_KB8, p=[190.97847715128717, 190.9733070521226, 254.93170605935475, 256.8974428996504,
                                0.0034823894022493434, 0.0007150348452162257, -0.0020532361418706202,
                                0.00020293673591811182], bf=0.0, w=512, h=512)
TUMVI_R = dict(type=CAM_KB8, p=[190.44236969414825, 190.4344384721956, 252.59949716835982, 254.91723064636983,
                                0.0034003170790442797, 0.001766278153469831, -0.00266312569781606,
                                0.0003299517423931039], bf=0.0, w=512, h=512)
TUMVI_T_C1_C2 = np.array([[0.999999445773493, 0.000791687752817, 0.000694034010224, 0.101063427414194],
                          [-0.000823363992158, 0.998899461915674, 0.046895490788700, 0.001946204678584],
                          [-0.000656143613644, -0.046896036240590, 0.998899560146304, 0.001015350132563],
                          [0.0, 0.0, 0.0, 1.0]])


def f32(a):
    return np.asarray(a, np.float32).astype(np.float64)


# --------------------------------------------------------------------------- geometry helpers
def _project(cam: dict, X: np.ndarray):
    """Double-precision projection used only to synthesise measurements. Returns (u, v, valid)."""
    p = cam["p"]
    x, y, z = X[..., 0], X[..., 1], X[..., 2]
    if cam["type"] == CAM_PINHOLE:
        zz = np.where(z > 1e-6, z, 1.0)
        u = p[0] * x / zz + p[2]
        v = p[1] * y / zz + p[3]
        valid = z > 0.3
    else:
        r = np.sqrt(x * x + y * y)
        th = np.arctan2(r, z)
        psi = np.arctan2(y, x)
        t2 = th * th
        rd = th * (1 + t2 * (p[4] + t2 * (p[5] + t2 * (p[6] + t2 * p[7]))))
        u = p[0] * rd * np.cos(psi) + p[2]
        v = p[1] * rd * np.sin(psi) + p[3]
        valid = (th < 1.25) & (np.sqrt(r * r + z * z) > 0.3)
    m = 8.0
    valid = valid & (u > m) & (u < cam["w"] - m) & (v > m) & (v < cam["h"] - m)
    return u, v, valid


def _trajectory(n_kf: int, step: float, dyaw_deg: float, rng, look_deg: float = -90.0):
    """Planar arc: `step` metres and `dyaw_deg` per keyframe; camera z looks `look_deg` from the heading."""
    k = np.arange(n_kf)
    th = np.deg2rad(dyaw_deg) * k
    d = np.stack([np.cos(th), np.sin(th), np.zeros(n_kf)], 1) * step
    pos = np.cumsum(d, 0)
    pos[:, 2] = 0.05 * np.sin(0.21 * k)
    a = th + np.deg2rad(look_deg) + np.deg2rad(2.0) * np.sin(0.37 * k)
    zc = np.stack([np.cos(a), np.sin(a), np.zeros(n_kf)], 1)
    yc = np.tile(np.array([0.0, 0.0, -1.0]), (n_kf, 1))
    xc = np.cross(yc, zc)
    R_wc = np.stack([xc, yc, zc], 2)                  # columns = camera axes in world
    R_cw = np.transpose(R_wc, (0, 2, 1))
    t_cw = -np.einsum("nij,nj->ni", R_cw, pos)
    return R_cw, t_cw


def _to_qt(R_cw: np.ndarray, t_cw: np.ndarray) -> np.ndarray:
    q = Rotation.from_matrix(R_cw).as_quat()          # x y z w
    q = np.where(q[:, 3:4] < 0, -q, q)
    return np.concatenate([t_cw, q], 1)


def _perturb(R_cw, t_cw, rng, rot_deg, trans_m):
    n = len(R_cw)
    w = rng.normal(0, np.deg2rad(rot_deg), (n, 3))
    v = rng.normal(0, trans_m, (n, 3))
    dR = Rotation.from_rotvec(w).as_matrix()
    return np.einsum("nij,njk->nik", dR, R_cw), np.einsum("nij,nj->ni", dR, t_cw) + v


def _levels(rng, n, n_levels=8):
    wts = 1.2 ** (-2.0 * np.arange(n_levels))
    return rng.choice(n_levels, size=n, p=wts / wts.sum())


# --------------------------------------------------------------------------- BA maps
def make_ba_problem(*, seed: int, n_kf: int, n_fixed: int, n_points: int, mean_track: float, sensor: str,
                    step: float = 0.15, dyaw_deg: float = 3.0, window: int = 40, outlier_frac: float = 0.05,
                    robust: bool = True, stereo_frac: float = 0.7, name: str = "", chunk: int = 20000,
                    shard: int = 0) -> BAProblem:
    """sensor: 'mono' (TUM1 pinhole), 'stereo' (EuRoC pinhole, mixed mono/stereo edges), 'fisheye' (TUM-VI KB8 rig:
    left mono edges + right ToBody edges). The first `n_fixed` keyframes (oldest ids) are fixed.
    Keyframes (trajectory and initial-pose perturbation) depend on `seed` only; landmarks and observations depend on
    (`seed`, `shard`), so rank r of a multi-GPU run can generate ITS landmark shard of one common map directly."""
    rng_kf = np.random.default_rng(seed)
    rng = np.random.default_rng([seed, 7919 + shard])
    R_cw, t_cw = _trajectory(n_kf, step, dyaw_deg, rng_kf)
    camL = {"mono": TUM1, "stereo": EUROC, "fisheye": TUMVI_L}[sensor]
    cams = [camL]
    rigs = None
    if sensor == "fisheye":
        cams = [TUMVI_L, TUMVI_R]
        T_lr = TUMVI_T_C1_C2
        R_rl = T_lr[:3, :3].T
        t_rl = -R_rl @ T_lr[:3, 3]
        q_rl = Rotation.from_matrix(R_rl).as_quat()
        if q_rl[3] < 0:
            q_rl = -q_rl
        rigs = f32(np.concatenate([t_rl, q_rl])[None, :])
        # geometry below uses the float-rounded extrinsics, like the real boundary
        R_rl = Rotation.from_quat(rigs[0, 3:] / np.linalg.norm(rigs[0, 3:])).as_matrix()
        t_rl = rigs[0, :3]
    sig_tab = inv_level_sigma2().astype(np.float64)

    lo_anchor = n_fixed if n_fixed < n_kf else 0
    offs_all = np.arange(-window, window + 1)
    P_all, cols = [], {k: [] for k in ("pose", "point", "cam", "rig", "kind", "u", "v", "ur", "lvl")}
    n_done = 0
    # MapPoint ids grow with the keyframe that created them: landmarks come in creation (anchor keyframe) order
    k0_all = np.sort(rng.integers(lo_anchor, n_kf, n_points))
    while n_done < n_points:
        pc = min(chunk, n_points - n_done)
        k0 = k0_all[n_done:n_done + pc]
        d = rng.uniform(2.0, 12.0, pc)
        if camL["type"] == CAM_PINHOLE:
            uu = rng.uniform(20, camL["w"] - 20, pc)
            vv = rng.uniform(20, camL["h"] - 20, pc)
            Xc = np.stack([(uu - camL["p"][2]) / camL["p"][0] * d, (vv - camL["p"][3]) / camL["p"][1] * d, d], 1)
        else:
            th = rng.uniform(0.05, 1.15, pc)
            ps = rng.uniform(-np.pi, np.pi, pc)
            Xc = np.stack([d * np.sin(th) * np.cos(ps), d * np.sin(th) * np.sin(ps), d * np.cos(th)], 1)
        Xw = np.einsum("nji,nj->ni", R_cw[k0], Xc - t_cw[k0])          # R_wc (Xc - t)
        Xw = f32(Xw)                                                      # truth lives on the float grid too
        kk = k0[:, None] + offs_all[None, :]
        inb = (kk >= 0) & (kk < n_kf)
        kk = np.clip(kk, 0, n_kf - 1)
        Xl = np.einsum("pwij,pj->pwi", R_cw[kk], Xw) + t_cw[kk]
        uL, vL, visL = _project(camL, Xl)
        visL &= inb
        if sensor == "fisheye":
            Xr = np.einsum("ij,pwj->pwi", R_rl, Xl) + t_rl
            uR, vR, visR = _project(TUMVI_R, Xr)
            visR &= inb
            vis_any = visL | visR
        else:
            vis_any = visL
        # A MapPoint is tracked over a CONTIGUOUS run of the keyframes that can see it (with a few missed
        # detections), around the keyframe that created it -- not over a random subset of them.
        rank_w = np.cumsum(vis_any, 1)                                     # 1-based rank among visible keyframes
        ra = rank_w[:, window]
        n_kf_mean = mean_track / (0.9 * (1.6 if sensor == "fisheye" else 1.0))
        run = 2 + rng.poisson(max(n_kf_mean - 2.0, 0.5), pc)
        left = (rng.random(pc) * run).astype(np.int64)
        lo_r = (ra - left)[:, None]
        hi_r = lo_r + run[:, None] - 1
        keep = vis_any & (rank_w >= lo_r) & (rank_w <= hi_r) & (rng.random(vis_any.shape) < 0.9)
        keep[np.arange(pc), window] |= vis_any[:, window]                  # the anchor keyframe always observes it
        pi, wi = np.nonzero(keep)                                          # point-major, keyframe ascending
        kf = kk[pi, wi]
        if sensor == "mono":
            n = len(pi)
            cols["pose"].append(kf); cols["point"].append(pi + n_done)
            cols["cam"].append(np.zeros(n, np.int16)); cols["rig"].append(-np.ones(n, np.int16))
            cols["kind"].append(np.full(n, EDGE_MONO, np.uint8))
            cols["u"].append(uL[pi, wi]); cols["v"].append(vL[pi, wi]); cols["ur"].append(np.zeros(n))
        elif sensor == "stereo":
            n = len(pi)
            z = Xl[pi, wi, 2]
            ur = uL[pi, wi] - EUROC["bf"] / z
            is_st = (rng.random(n) < stereo_frac) & (ur > 1.0)
            cols["pose"].append(kf); cols["point"].append(pi + n_done)
            cols["cam"].append(np.zeros(n, np.int16)); cols["rig"].append(-np.ones(n, np.int16))
            cols["kind"].append(np.where(is_st, EDGE_STEREO, EDGE_MONO).astype(np.uint8))
            cols["u"].append(uL[pi, wi]); cols["v"].append(vL[pi, wi]); cols["ur"].append(np.where(is_st, ur, 0.0))
        else:
            # per kept (point, keyframe): left mono edge if visible left, then right body edge if visible right
            hasL = visL[pi, wi]
            hasR = visR[pi, wi]
            n2 = len(pi) * 2
            order = np.arange(n2)
            sel = np.empty(n2, bool); sel[0::2] = hasL; sel[1::2] = hasR
            rep = lambda a: np.repeat(a, 2)
            is_r = (order % 2) == 1
            cols["pose"].append(rep(kf)[sel]); cols["point"].append(rep(pi + n_done)[sel])
            cols["cam"].append(is_r.astype(np.int16)[sel]); cols["rig"].append(np.where(is_r, 0, -1).astype(np.int16)[sel])
            cols["kind"].append(np.where(is_r, EDGE_BODY, EDGE_MONO).astype(np.uint8)[sel])
            u2 = np.empty(n2); u2[0::2] = uL[pi, wi]; u2[1::2] = uR[pi, wi]
            v2 = np.empty(n2); v2[0::2] = vL[pi, wi]; v2[1::2] = vR[pi, wi]
            cols["u"].append(u2[sel]); cols["v"].append(v2[sel]); cols["ur"].append(np.zeros(int(sel.sum())))
        P_all.append(Xw)
        n_done += pc
    obs = {k: np.concatenate(v) for k, v in cols.items() if k != "lvl"}
    ne = len(obs["pose"])
    lvl = _levels(rng, ne)
    sigma = 1.2 ** lvl
    nu, nv, nr = (rng.normal(0, 1, ne) * sigma for _ in range(3))
    is_out = rng.random(ne) < outlier_frac
    gu, gv, gr = (rng.uniform(-30, 30, ne) for _ in range(3))
    u = obs["u"] + nu + np.where(is_out, gu, 0.0)
    v = obs["v"] + nv + np.where(is_out, gv, 0.0)
    ur = np.where(obs["kind"] == EDGE_STEREO, obs["ur"] + nr + np.where(is_out, gr, 0.0), 0.0)
    ur = np.where((obs["kind"] == EDGE_STEREO) & (ur < 0), 0.0, ur)

    Pw = np.concatenate(P_all)
    Ri, ti = _perturb(R_cw, t_cw, rng_kf, 0.5, 0.01)
    fixed = np.zeros(n_kf, np.uint8)
    fixed[:n_fixed] = 1
    Ri[:n_fixed], ti[:n_fixed] = R_cw[:n_fixed], t_cw[:n_fixed]
    pose_qt = f32(_to_qt(Ri, ti))
    pts0 = f32(Pw + rng.normal(0, 0.02, Pw.shape))
    flags = np.full(ne, FLAG_ROBUST if robust else 0, np.uint8)
    flags[obs["kind"] == EDGE_BODY] = FLAG_ROBUST            # ToBody edges always get Huber (Optimizer.cc:247-249)
    return BAProblem(pose_qt, fixed, pts0, cams, rigs, obs["pose"], obs["point"], obs["cam"], obs["rig"], obs["kind"],
                     flags, f32(u), f32(v), f32(ur), sig_tab[lvl], name=name,
                     truth=dict(pose_qt=_to_qt(R_cw, t_cw), points=Pw, outlier=is_out))


def config(n: int, scale: float = 1.0, robust: bool = True) -> BAProblem:
    """BASELINE.json configs 1-5 (index n). `scale` < 1 shrinks points/observations (parity-test sizes);
    keyframe counts shrink only for the global-BA configs. `mean_track` below is the generator's run-length
    parameter, calibrated so that the OBSERVATION counts hit BASELINE's (~40k, ~120k, ~110k, ~2M, ~20M): runs are
    clipped by the window's ends and by visibility, so the realised mean track is shorter than the parameter."""
    s = scale
    if n == 1:
        return make_ba_problem(seed=1, n_kf=30, n_fixed=10, n_points=max(50, int(5000 * s)), mean_track=14.0,
                               sensor="mono", window=30, name="C1 mono pinhole local BA")
    if n == 2:
        return make_ba_problem(seed=2, n_kf=70, n_fixed=25, n_points=max(50, int(12000 * s)), mean_track=12.8,
                               sensor="stereo", window=70, dyaw_deg=1.5, name="C2 stereo pinhole local BA")
    if n == 3:
        return make_ba_problem(seed=3, n_kf=45, n_fixed=15, n_points=max(50, int(8000 * s)), mean_track=12.0,
                               sensor="fisheye", window=45, dyaw_deg=2.0, name="C3 fisheye KB8 rig local BA")
    if n == 4:
        nk = max(12, int(500 * s))
        return make_ba_problem(seed=4, n_kf=nk, n_fixed=1, n_points=max(100, int(200000 * s)), mean_track=11.2,
                               sensor="stereo", window=40, dyaw_deg=0.6, robust=robust, name="C4 global BA")
    if n == 5:
        nk = max(24, int(5000 * s))
        return make_ba_problem(seed=5, n_kf=nk, n_fixed=1, n_points=max(200, int(2000000 * s)), mean_track=11.2,
                               sensor="stereo", window=40, dyaw_deg=0.06, robust=robust, name="C5 large global BA")
    raise ValueError(n)


def global_ba_shard(rank: int, n_ranks: int, seed: int = 4, n_kf: int = 500, points_per_rank: int = 200000,
                    robust: bool = False) -> BAProblem:
    """Weak-scaling global BA: the C4 keyframes (`n_kf`, one fixed) are common to all ranks; rank r owns
    `points_per_rank` landmarks of its own (= C4's per-GPU work) and all of their observations. Rank 0 of a
    1-rank run is exactly config(4)."""
    return make_ba_problem(seed=seed, n_kf=n_kf, n_fixed=1, n_points=points_per_rank, mean_track=11.2,
                           sensor="stereo", window=40, dyaw_deg=0.6, robust=robust, shard=rank,
                           name=f"C4 global BA, landmark shard {rank}/{n_ranks}")


def concat_shards(shards) -> BAProblem:
    """The single-GPU view of a sharded map: same keyframes, landmarks and observations concatenated."""
    p0 = shards[0]
    off = np.cumsum([0] + [s.n_points for s in shards[:-1]])
    cat = lambda k: np.concatenate([getattr(s, k) for s in shards])
    return BAProblem(p0.pose_qt, p0.pose_fixed, cat("points"), p0.cameras, p0.rigs, cat("obs_pose"),
                     np.concatenate([s.obs_point + o for s, o in zip(shards, off)]), cat("obs_cam"), cat("obs_rig"),
                     cat("obs_kind"), cat("obs_flags"), cat("obs_u"), cat("obs_v"), cat("obs_ur"), cat("obs_inv_sigma2"),
                     name="concat of %d shards" % len(shards))


# --------------------------------------------------------------------------- PoseOptimization batch
def make_pose_batch(seed: int = 3, n_frames: int = 1000, n_matches: int = 300, outlier_frac: float = 0.10) -> PoseBatch:
    """Independent frames for Optimizer::PoseOptimization: frame f uses sensor f%3 in
    (TUM1 mono, EuRoC stereo+mono, TUM-VI left mono + right body)."""
    rng = np.random.default_rng(seed)
    cams = [TUM1, EUROC, TUMVI_L, TUMVI_R]
    T_lr = TUMVI_T_C1_C2
    R_rl = T_lr[:3, :3].T
    t_rl = -R_rl @ T_lr[:3, 3]
    q_rl = Rotation.from_matrix(R_rl).as_quat()
    if q_rl[3] < 0:
        q_rl = -q_rl
    rigs = f32(np.concatenate([t_rl, q_rl])[None, :])
    R_rl = Rotation.from_quat(rigs[0, 3:] / np.linalg.norm(rigs[0, 3:])).as_matrix()
    t_rl = rigs[0, :3]
    R_cw, t_cw = _trajectory(n_frames, 0.15, 1.0, rng)
    sig_tab = inv_level_sigma2().astype(np.float64)
    ptr = [0]
    out = {k: [] for k in ("xw", "cam", "rig", "kind", "u", "v", "ur", "is2")}
    for f in range(n_frames):
        sensor = f % 3
        n = int(rng.integers(int(0.7 * n_matches), int(1.3 * n_matches) + 1))
        if f == 7:
            n = 2                                  # < 3 correspondences: returns 0 (Optimizer.cc:996-997)
        if f == 11:
            n = 8                                  # < 10 edges: one round only (Optimizer.cc:1102)
        d = rng.uniform(1.5, 15.0, n)
        if sensor < 2:
            cam = cams[sensor]
            uu = rng.uniform(20, cam["w"] - 20, n)
            vv = rng.uniform(20, cam["h"] - 20, n)
            Xc = np.stack([(uu - cam["p"][2]) / cam["p"][0] * d, (vv - cam["p"][3]) / cam["p"][1] * d, d], 1)
        else:
            th = rng.uniform(0.05, 1.1, n)
            ps = rng.uniform(-np.pi, np.pi, n)
            Xc = np.stack([d * np.sin(th) * np.cos(ps), d * np.sin(th) * np.sin(ps), d * np.cos(th)], 1)
        Xw = f32((Xc - t_cw[f]) @ R_cw[f])          # R_wc (Xc - t)
        Xl = Xw @ R_cw[f].T + t_cw[f]
        if sensor == 0:
            u, v, _ = _project(TUM1, Xl)
            kind = np.full(n, EDGE_MONO, np.uint8); cam_i = np.zeros(n, np.int16); rig_i = -np.ones(n, np.int16)
            ur = np.zeros(n)
        elif sensor == 1:
            u, v, _ = _project(EUROC, Xl)
            ur = u - EUROC["bf"] / Xl[:, 2]
            st = (rng.random(n) < 0.7) & (ur > 1.0)
            kind = np.where(st, EDGE_STEREO, EDGE_MONO).astype(np.uint8)
            cam_i = np.ones(n, np.int16); rig_i = -np.ones(n, np.int16)
            ur = np.where(st, ur, 0.0)
        else:
            right = rng.random(n) < 0.45            # keypoints i >= Nleft come from the right image
            Xr = Xl @ R_rl.T + t_rl
            uL, vL, _ = _project(TUMVI_L, Xl)
            uR, vR, _ = _project(TUMVI_R, Xr)
            u = np.where(right, uR, uL); v = np.where(right, vR, vL)
            kind = np.where(right, EDGE_BODY, EDGE_MONO).astype(np.uint8)
            cam_i = np.where(right, 3, 2).astype(np.int16); rig_i = np.where(right, 0, -1).astype(np.int16)
            ur = np.zeros(n)
        out["xw"].append(Xw); out["cam"].append(cam_i); out["rig"].append(rig_i); out["kind"].append(kind)
        out["u"].append(u); out["v"].append(v); out["ur"].append(ur)
        ptr.append(ptr[-1] + n)
    cat = {k: np.concatenate(v) for k, v in out.items() if k != "is2"}
    ne = len(cat["u"])
    lvl = _levels(rng, ne)
    sigma = 1.2 ** lvl
    is_out = rng.random(ne) < outlier_frac
    u = cat["u"] + rng.normal(0, 1, ne) * sigma + np.where(is_out, rng.uniform(-50, 50, ne), 0.0)
    v = cat["v"] + rng.normal(0, 1, ne) * sigma + np.where(is_out, rng.uniform(-50, 50, ne), 0.0)
    st = cat["kind"] == EDGE_STEREO
    ur = np.where(st, cat["ur"] + rng.normal(0, 1, ne) * sigma + np.where(is_out, rng.uniform(-50, 50, ne), 0.0), 0.0)
    ur = np.where(st & (ur < 0), 0.0, ur)
    Ri, ti = _perturb(R_cw, t_cw, rng, 1.0, 0.03)
    return PoseBatch(f32(_to_qt(Ri, ti)), np.asarray(ptr, np.int64), cams, rigs, cat["xw"], cat["cam"], cat["rig"],
                     cat["kind"], f32(u), f32(v), f32(ur), sig_tab[lvl],
                     truth=dict(pose_qt=_to_qt(R_cw, t_cw), outlier=is_out))
