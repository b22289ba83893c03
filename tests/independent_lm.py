"""A SECOND, independent implementation of the reference's bundle-adjustment loop, in numpy (test infrastructure).

Purpose: pin the oracle's CONTROL FLOW (accept / reject, lambda and nu updates, the ORB-SLAM `_nBad` stop rule, the gate
between rounds and the re-structuring of the active set) with code the oracle did not write. It shares nothing with
oracle/ba_ref.cpp: residuals and Jacobians are vectorised numpy (own derivation, `A = -Jh R`, `B = -Jh [-[Xc]x | I]`),
the damped normal equations of the FULL problem (poses and points, no Schur complement) are assembled densely and
solved with numpy.linalg.solve, and the LM loop is written from the reference sources:

  OptimizationAlgorithmLevenberg::solve       Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-169
  computeLambdaInit / computeScale            ... :171-194
  SparseOptimizer::optimize / initializeOptimization   Thirdparty/g2o/g2o/core/sparse_optimizer.cpp:354-419, 199-267
  RobustKernelHuber::robustify                Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:78-91
  EdgeSE3ProjectXYZ / Pinhole                 src/OptimizableTypes.cpp:139-160, src/CameraModels/Pinhole.cpp:35-41,71-81
  EdgeStereoSE3ProjectXYZ (float invz, float bf)   Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:190-197,228-274
  SE3Quat::exp, operator*                     Thirdparty/g2o/g2o/types/se3quat.h:104-110,223-257
  merge-LBA gate                              src/Optimizer.cc:3745-3776

Pinhole mono and stereo edges only (configs 1, 2, 4, 5); no KB8, no body edges.
"""
from __future__ import annotations

import numpy as np

EDGE_MONO, EDGE_STEREO = 0, 1
DBL_MAX = np.finfo(np.float64).max


def _skew(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def _quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _R_to_quat(R):
    """Shepperd's method (what Eigen's Quaterniond(Matrix3d) does), then w >= 0 and unit norm."""
    tr = np.trace(R)
    q = np.zeros(4)
    if tr > 0:
        s = np.sqrt(tr + 1.0)
        q[3] = 0.5 * s
        s = 0.5 / s
        q[0], q[1], q[2] = (R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = 0.5 * s
        s = 0.5 / s
        q[3] = (R[k, j] - R[j, k]) * s
        q[j] = (R[j, i] + R[i, j]) * s
        q[k] = (R[k, i] + R[i, k]) * s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def _quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def pose_oplus(qt, d):
    """T <- exp(d) T with d = [omega, upsilon] (se3quat.h:223-257: theta < 1e-5 uses R = I + W + W W and V = R)."""
    w, u = d[:3], d[3:]
    th = np.linalg.norm(w)
    W = _skew(w)
    if th < 1e-5:
        R = np.eye(3) + W + W @ W
        V = R
    else:
        W2 = W @ W
        R = np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th ** 2 * W2
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * W + (th - np.sin(th)) / th ** 3 * W2
    qe, te = _R_to_quat(R), V @ u
    t = te + _quat_to_R(qe) @ qt[:3]
    q = _quat_mul(qe, qt[3:])
    if q[3] < 0:
        q = -q
    q = q / np.linalg.norm(q)
    return np.concatenate([t, q])


class IndependentBA:
    def __init__(self, problem):
        p = problem
        assert all(c["type"] == 0 for c in p.cameras) and (p.obs_kind <= EDGE_STEREO).all()
        self.p = p
        q = p.pose_qt[:, 3:].copy()
        q[q[:, 3] < 0] *= -1
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        self.pose = np.concatenate([p.pose_qt[:, :3], q], 1)
        self.pts = p.points.copy()
        self.level = np.zeros(p.n_obs, np.uint8)
        self.robust = (p.obs_flags & 1).astype(bool)
        self.edge_chi2 = np.zeros(p.n_obs)
        cam = np.array([[np.float32(v) for v in c["p"][:4]] + [np.float32(c.get("bf", 0.0))] for c in p.cameras], np.float64)
        self.K = cam[p.obs_cam]                      # fx fy cx cy bf per edge (float values widened)
        self.stereo = p.obs_kind == EDGE_STEREO

    # -- residuals (and optionally Jacobians) of the edges in `idx` at (pose, pts)
    def _edges(self, idx, pose, pts, jac):
        p = self.p
        R = np.stack([_quat_to_R(q) for q in pose[:, 3:]])
        Re = R[p.obs_pose[idx]]
        Xc = np.einsum("eij,ej->ei", Re, pts[p.obs_point[idx]]) + pose[p.obs_pose[idx], :3]
        x, y, z = Xc.T
        fx, fy, cx, cy, bf = self.K[idx].T
        st = self.stereo[idx]
        invz_f = (1.0 / z).astype(np.float32)                                     # const float invz = 1.0f / z
        invz = np.where(st, invz_f.astype(np.float64), 0.0)
        u = np.where(st, x * invz * fx + cx, (fx * x) / z + cx)
        v = np.where(st, y * invz * fy + cy, (fy * y) / z + cy)
        bf_invz = (bf.astype(np.float32) * invz_f).astype(np.float64)             # float * float
        r = np.zeros((len(idx), 3))
        r[:, 0] = p.obs_u[idx] - u
        r[:, 1] = p.obs_v[idx] - v
        r[:, 2] = np.where(st, p.obs_ur[idx] - (u - bf_invz), 0.0)
        chi2 = p.obs_inv_sigma2[idx] * (r * r).sum(1)
        if not jac:
            return r, chi2, z, None, None
        Jh = np.zeros((len(idx), 3, 3))
        Jh[:, 0, 0] = fx / z; Jh[:, 0, 2] = -fx * x / (z * z)
        Jh[:, 1, 1] = fy / z; Jh[:, 1, 2] = -fy * y / (z * z)
        Jh[:, 2, 0] = np.where(st, fx / z, 0.0)
        Jh[:, 2, 2] = np.where(st, -fx * x / (z * z) + bf / (z * z), 0.0)
        D = np.zeros((len(idx), 3, 6))
        D[:, 0, 1] = z; D[:, 0, 2] = -y; D[:, 1, 0] = -z; D[:, 1, 2] = x; D[:, 2, 0] = y; D[:, 2, 1] = -x
        D[:, 0, 3] = D[:, 1, 4] = D[:, 2, 5] = 1.0
        A = -np.einsum("eij,ejk->eik", Jh, Re)
        B = -np.einsum("eij,ejk->eik", Jh, D)
        return r, chi2, z, A, B

    def _rho(self, idx, chi2, delta_mono, delta_stereo):
        d = np.where(self.stereo[idx], delta_stereo, delta_mono)
        out = self.robust[idx] & (chi2 > d * d)
        sq = np.sqrt(np.where(out, chi2, 1.0))
        rho0 = np.where(out, 2 * d * sq - d * d, chi2)
        rho1 = np.where(out, d / sq, 1.0)
        return rho0, rho1

    def optimize(self, iterations, delta_mono, delta_stereo, lambda_user, trace, rnd, stop=lambda: False):
        p = self.p
        idx = np.flatnonzero(self.level == 0)
        free_pose = (p.pose_fixed == 0)
        # active set: edges at level 0 whose vertices are not all fixed (points are never fixed); vertices with >= 1 such edge
        used_pose = np.zeros(p.n_poses, bool); used_pose[p.obs_pose[idx]] = True
        pose_ids = np.flatnonzero(free_pose & used_pose)
        pt_ids = np.unique(p.obs_point[idx])
        hp = -np.ones(p.n_poses, int); hp[pose_ids] = np.arange(len(pose_ids))
        hl = -np.ones(p.n_points, int); hl[pt_ids] = np.arange(len(pt_ids))
        npz, nl = 6 * len(pose_ids), 3 * len(pt_ids)
        n = npz + nl
        ep, el = hp[p.obs_pose[idx]], hl[p.obs_point[idx]]
        lam, nu, n_bad, status = -1.0, 2.0, 0, 0
        ok = True
        it = 0
        while it < iterations and not stop() and ok:
            r, chi2, _, A, B = self._edges(idx, self.pose, self.pts, True)
            self.edge_chi2[idx] = chi2
            rho0, rho1 = self._rho(idx, chi2, delta_mono, delta_stereo)
            cur = ini = rho0.sum()
            w = rho1 * p.obs_inv_sigma2[idx]
            H = np.zeros((n, n)); b = np.zeros(n)
            g = -(w[:, None] * r)
            # dense accumulation of the per-edge blocks (np.add.at: unbuffered, duplicates add up)
            cl = npz + 3 * el
            r3, r6 = np.arange(3), np.arange(6)
            Hll = w[:, None, None] * np.einsum("eki,ekj->eij", A, A)
            np.add.at(H, (cl[:, None, None] + r3[None, :, None], cl[:, None, None] + r3[None, None, :]), Hll)
            np.add.at(b, cl[:, None] + r3[None, :], np.einsum("eki,ek->ei", A, g))
            fp = ep >= 0
            cp = 6 * ep[fp]
            Bf, Af, wf = B[fp], A[fp], w[fp]
            np.add.at(H, (cp[:, None, None] + r6[None, :, None], cp[:, None, None] + r6[None, None, :]), wf[:, None, None] * np.einsum("eki,ekj->eij", Bf, Bf))
            np.add.at(b, cp[:, None] + r6[None, :], np.einsum("eki,ek->ei", Bf, g[fp]))
            Wb = wf[:, None, None] * np.einsum("eki,ekj->eij", Bf, Af)
            np.add.at(H, (cp[:, None, None] + r6[None, :, None], cl[fp][:, None, None] + r3[None, None, :]), Wb)
            np.add.at(H, (cl[fp][:, None, None] + r3[None, :, None], cp[:, None, None] + r6[None, None, :]), np.transpose(Wb, (0, 2, 1)))
            if it == 0:
                lam = lambda_user if lambda_user > 0 else 1e-5 * np.abs(np.diag(H)).max()
                nu, n_bad = 2.0, 0
            q, rho = 0, 0.0
            while True:
                try:
                    x = np.linalg.solve(H + lam * np.eye(n), b)
                    solved = np.isfinite(x).all()
                except np.linalg.LinAlgError:
                    x, solved = np.zeros(n), False
                pose_t, pts_t = self.pose.copy(), self.pts.copy()
                for k, i in enumerate(pose_ids):
                    pose_t[i] = pose_oplus(self.pose[i], x[6 * k:6 * k + 6])
                pts_t[pt_ids] += x[npz:].reshape(-1, 3)
                _, chi2_t, _, _, _ = self._edges(idx, pose_t, pts_t, False)
                self.edge_chi2[idx] = chi2_t                 # errors are NOT recomputed after a rejected trial
                f_t = self._rho(idx, chi2_t, delta_mono, delta_stereo)[0].sum() if solved else DBL_MAX
                scale = float(x @ (lam * x + b)) + 1e-3
                rho = (cur - f_t) / scale
                if rho > 0 and np.isfinite(f_t):
                    alpha = min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)
                    lam *= max(1.0 / 3.0, alpha)
                    nu = 2.0
                    cur = f_t
                    self.pose, self.pts = pose_t, pts_t
                else:
                    lam *= nu
                    nu *= 2
                q += 1
                if not (rho < 0 and q < 10 and not stop()):
                    break
            st = 0
            if q == 10 or rho == 0:
                st = 1
            else:
                n_bad = n_bad + 1 if (ini - cur) * 1e3 < ini else 0
                if n_bad >= 3:
                    st = 2
            trace.append(dict(round=rnd, iteration=it, chi2_before=ini, chi2_after=cur, lambda_=lam, trials=q, status=st))
            status, ok = st, st == 0
            it += 1
        if stop() and status == 0:
            status = 3
        return status

    def gate(self, gate_mono, gate_stereo, drop_kernel):
        idx = np.flatnonzero(self.level == 0)
        _, _, z, _, _ = self._edges(idx, self.pose, self.pts, False)
        th = np.where(self.stereo[idx], gate_stereo, gate_mono)
        self.level[idx[(self.edge_chi2[idx] > th) | ~(z > 0)]] = 1
        if drop_kernel:
            self.robust[:] = False


def solve(problem, schedule):
    """The schedule's rounds (optimize, gate, drop kernels) on `problem`; returns (trace, pose_qt, points, edge_level, status)."""
    ba = IndependentBA(problem)
    trace, status = [], 0
    for k, rd in enumerate(schedule.rounds):
        if (ba.level == 0).any():
            status = ba.optimize(rd.iterations, schedule.delta_mono, schedule.delta_stereo, schedule.lambda_init, trace, k)
        if rd.gate_after == 1:
            ba.gate(rd.gate_mono, rd.gate_stereo, rd.drop_kernel_after)
        elif rd.drop_kernel_after:
            ba.robust[:] = False
    return trace, ba.pose, ba.pts, ba.level, status
