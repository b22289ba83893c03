"""CPU tests: pin the oracle (constants fixed by the reference source, g2o-style numeric Jacobians, an independent
dense solve of one LM step, LDLT vs numpy), the committed golden traces, and the fdlibm atan2f port vs libm."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from orb_slam3_study_kr_b200 import problem, synthetic
from orb_slam3_study_kr_b200.problem import (CAM_KB8, CAM_PINHOLE, EDGE_BODY, EDGE_MONO, EDGE_STEREO, Round, Schedule)
from oracle import ba_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- constants the reference source fixes
def test_reference_constants():
    assert problem.DELTA_MONO_LBA == 2.4476518630981445          # float(sqrt(5.991)), Optimizer.cc:1275
    assert problem.DELTA_MONO_LBA ** 2 == pytest.approx(5.990999642927818, abs=1e-15)
    assert problem.DELTA_MONO_GBA == 2.4474475383758545          # float(sqrt(5.99)),  Optimizer.cc:132
    assert problem.DELTA_STEREO == 2.7955322265625               # float(sqrt(7.815)), Optimizer.cc:133
    assert float(np.float32(5.991)) == 5.991000175476074 and float(np.float32(7.815)) == 7.815000057220459
    tab = problem.inv_level_sigma2().astype(np.float64)          # ORBextractor.cc:414-429
    want = [1, 0.6944444179534912, 0.4822530746459961, 0.33489790558815, 0.23256798088550568, 0.1615055352449417,
            0.11215661466121674, 0.07788652926683426]
    assert np.allclose(tab, want, rtol=0, atol=1e-16)


def test_huber_kernel():
    d = problem.DELTA_MONO_LBA
    assert ba_ref.huber(1.0, d) == (1.0, 1.0)
    r0, r1 = ba_ref.huber(100.0, d)
    assert r0 == pytest.approx(2 * 10.0 * d - d * d, rel=1e-15) and r1 == pytest.approx(d / 10.0, rel=1e-15)
    # boundary: e == delta^2 is an inlier (robust_kernel_impl.cpp:81 "e <= dsqr")
    assert ba_ref.huber(d * d, d) == (d * d, 1.0)


# ---------------------------------------------------------------- edges: residuals and Jacobians
def _model_residual(kind, cam, trl, pose_qt, X, obs):
    """Plain double-precision model of the three residuals (no float fossils)."""
    R = Rotation.from_quat(pose_qt[3:] / np.linalg.norm(pose_qt[3:])).as_matrix()
    Xc = R @ X + pose_qt[:3]
    if kind == EDGE_BODY:
        Rr = Rotation.from_quat(trl[3:] / np.linalg.norm(trl[3:])).as_matrix()
        Xc = Rr @ Xc + trl[:3]
    p = [float(np.float32(v)) for v in cam["p"]] + [0.0] * 8
    x, y, z = Xc
    if kind == EDGE_STEREO:
        bf = float(np.float32(cam["bf"]))
        u = p[0] * x / z + p[2]
        return np.array([obs[0] - u, obs[1] - (p[1] * y / z + p[3]), obs[2] - (u - bf / z)])
    if cam["type"] == CAM_PINHOLE:
        return np.array([obs[0] - (p[0] * x / z + p[2]), obs[1] - (p[1] * y / z + p[3]), 0.0])
    th = np.arctan2(np.hypot(x, y), z)
    psi = np.arctan2(y, x)
    r = th + p[4] * th ** 3 + p[5] * th ** 5 + p[6] * th ** 7 + p[7] * th ** 9
    return np.array([obs[0] - (p[0] * r * np.cos(psi) + p[2]), obs[1] - (p[1] * r * np.sin(psi) + p[3]), 0.0])


def _cases():
    rng = np.random.default_rng(0)
    trl = synthetic.make_pose_batch(n_frames=3, n_matches=5).rigs[0]
    out = []
    for kind, cam in [(EDGE_MONO, synthetic.TUM1), (EDGE_MONO, synthetic.TUMVI_L), (EDGE_STEREO, synthetic.EUROC),
                      (EDGE_BODY, synthetic.TUMVI_R), (EDGE_BODY, synthetic.TUM1)]:
        for _ in range(6):
            q = Rotation.from_rotvec(rng.normal(0, 0.4, 3)).as_quat()
            pose = np.concatenate([rng.normal(0, 0.5, 3), q if q[3] > 0 else -q])
            Xc = np.array([rng.uniform(-1.5, 1.5), rng.uniform(-1.0, 1.0), rng.uniform(2.0, 9.0)])
            R = Rotation.from_quat(pose[3:]).as_matrix()
            X = R.T @ (Xc - pose[:3])
            obs = np.array([rng.uniform(50, 450), rng.uniform(50, 400), rng.uniform(20, 400)])
            out.append((kind, cam, trl if kind == EDGE_BODY else None, pose, X, obs))
    return out


@pytest.mark.parametrize("unary", [False, True])
def test_edge_residual_and_jacobians_vs_numeric(unary):
    """Analytic Jacobians vs central differences through the vertex oplus, the construction g2o itself uses as its
    fallback (base_binary_edge.hpp:130-205, base_unary_edge.hpp:81-123). The differences are taken on the
    double-precision model so that the float fossils (atan2f, float invz) do not quantise them."""
    for kind, cam, trl, pose, X, obs in _cases():
        err, A, B, dp = ba_ref.edge(kind, cam, trl, pose, X, obs, unary)
        dim = 3 if kind == EDGE_STEREO else 2
        model = _model_residual(kind, cam, trl, pose, X, obs)
        assert np.abs(err[:dim] - model[:dim]).max() < 2e-3      # fossils move the residual by < 2e-3 px
        assert dp is True
        h = 1e-6
        numA = np.zeros((3, 3)); numB = np.zeros((3, 6))
        for i in range(3):
            d = np.zeros(3); d[i] = h
            numA[:, i] = (_model_residual(kind, cam, trl, pose, X + d, obs) - _model_residual(kind, cam, trl, pose, X - d, obs)) / (2 * h)
        for i in range(6):
            d = np.zeros(6); d[i] = h
            numB[:, i] = (_model_residual(kind, cam, trl, ba_ref.oplus(pose, d), X, obs)
                          - _model_residual(kind, cam, trl, ba_ref.oplus(pose, -d), X, obs)) / (2 * h)
        scale = max(1.0, np.abs(numB).max())
        if not unary:                                             # unary (OnlyPose) edges have no point Jacobian
            assert np.abs(A[:dim] - numA[:dim]).max() < 2e-5 * scale, (kind, cam["type"])
        assert np.abs(B[:dim] - numB[:dim]).max() < 2e-5 * scale, (kind, cam["type"])


def test_mono_pinhole_numeric_jacobian_g2o_delta():
    """For the fossil-free edge (mono pinhole) use g2o's own delta = 1e-9 on the oracle's residual itself."""
    kind, cam = EDGE_MONO, synthetic.TUM1
    for k, c, trl, pose, X, obs in _cases():
        if k != kind or c is not cam:
            continue
        err, A, B, _ = ba_ref.edge(kind, cam, None, pose, X, obs, False)
        d9 = 1e-9
        for i in range(6):
            d = np.zeros(6); d[i] = d9
            ep = ba_ref.edge(kind, cam, None, ba_ref.oplus(pose, d), X, obs, False)[0]
            em = ba_ref.edge(kind, cam, None, ba_ref.oplus(pose, -d), X, obs, False)[0]
            assert np.abs((ep - em)[:2] / (2 * d9) - B[:2, i]).max() < 2e-3 * max(1.0, np.abs(B).max())


def test_oplus_matches_reference_exp_conventions():
    """exp([w, v]): rotation first, then translation; theta < 1e-5 uses R = I + W + W^2 (se3quat.h:236-242)."""
    pose = np.array([0.1, -0.2, 0.3, 0, 0, 0, 1.0])
    out = ba_ref.oplus(pose, np.array([0, 0, 0, 1.0, 2.0, 3.0]))
    assert np.allclose(out, [1.1, 1.8, 3.3, 0, 0, 0, 1])
    w = np.array([0.3, -0.2, 0.1])
    out = ba_ref.oplus(pose, np.concatenate([w, np.zeros(3)]))
    Rw = Rotation.from_rotvec(w)
    assert np.allclose(out[3:], Rw.as_quat() * np.sign(Rw.as_quat()[3]), atol=1e-15)
    assert np.allclose(out[:3], Rw.apply(pose[:3]), atol=1e-15)
    tiny = ba_ref.oplus(pose, np.array([1e-7, 0, 0, 0, 0, 0]))
    assert abs(tiny[3] - 0.5e-7) < 1e-14 and tiny[6] == pytest.approx(1.0, abs=1e-13)


# ---------------------------------------------------------------- linear solvers vs numpy
def test_dense_ldlt_and_skyline_vs_numpy():
    rng = np.random.default_rng(1)
    for n in (6, 30, 121):
        M = rng.normal(size=(n, n))
        H = M @ M.T + n * np.eye(n)
        for i in range(n):                      # banded
            for j in range(n):
                if abs(i - j) > max(6, n // 3):
                    H[i, j] = 0
        H = H + n * np.eye(n)
        b = rng.normal(size=n)
        want = np.linalg.solve(H, b)
        ok, x = ba_ref.dense_ldlt(H, b)
        assert ok and np.abs(x - want).max() < 1e-12 * max(1, np.abs(want).max()) * n
        ok, x = ba_ref.skyline_solve(H, b)
        assert ok and np.abs(x - want).max() < 1e-12 * max(1, np.abs(want).max()) * n
    ok, _ = ba_ref.dense_ldlt(-np.eye(6), np.ones(6))     # LinearSolverDense: !isPositive -> fail
    assert not ok
    ok, _ = ba_ref.skyline_solve(np.zeros((6, 6)), np.ones(6))   # SimplicialLDLT: zero pivot -> fail
    assert not ok


def test_reference_failure_semantics_of_the_sparse_solver():
    """LinearSolverEigen fails only when SimplicialLDLT reports an EXACT zero pivot (linear_solver_eigen.h:104-110); an
    indefinite system with non-zero pivots is factored as L D L^T and solved. The oracle keeps that; the GPU's Cholesky
    flags any non-positive pivot (tests/test_gpu_solver.py::test_solver_flags_indefinite). Both end in the same LM
    decision for every system the ABI can produce: H = J^T W J + lambda I with W >= 0, lambda > 0 is positive definite, so
    a non-positive pivot can only come from overflow / NaN, where the reference's trial chi2 is not finite and the trial is
    rejected as well (optimization_algorithm_levenberg.cpp:126-129). DESIGN.md section 2 records the deviation."""
    rng = np.random.default_rng(3)
    M = rng.normal(size=(12, 12))
    H = M + M.T                                   # symmetric indefinite, generic (no zero pivots)
    assert np.linalg.eigvalsh(H).min() < 0 < np.linalg.eigvalsh(H).max()
    b = rng.normal(size=12)
    ok, x = ba_ref.skyline_solve(H, b)
    assert ok and np.abs(H @ x - b).max() < 1e-9   # negative pivots are NOT a failure in the reference's solver
    Z = np.diag([1.0, 2.0, 0.0, 4.0, 5.0, 6.0])   # an exact zero pivot is
    ok, _ = ba_ref.skyline_solve(Z, np.ones(6))
    assert not ok


def test_one_lm_step_vs_independent_dense_solve():
    """One LM iteration of the oracle (Schur + skyline LDLT + back-substitution) against a dense solve of the FULL
    (poses + points) damped normal equations assembled with numpy from the oracle's per-edge Jacobians."""
    p = synthetic.config(2, scale=0.01)
    lam = 10.0
    s = Schedule([Round(1)], problem.DELTA_MONO_LBA, problem.DELTA_STEREO, lambda_init=lam)
    res = ba_ref.solve(p, s)
    assert res.trace[0]["trials"] == 1
    free = np.flatnonzero(p.pose_fixed == 0)
    hidx = -np.ones(p.n_poses, int); hidx[free] = np.arange(len(free))
    n = 6 * len(free) + 3 * p.n_points
    H = np.zeros((n, n)); b = np.zeros(n)
    for e in range(p.n_obs):
        kind = int(p.obs_kind[e]); cam = p.cameras[p.obs_cam[e]]
        trl = p.rigs[p.obs_rig[e]] if kind == EDGE_BODY else None
        i, j = int(p.obs_pose[e]), int(p.obs_point[e])
        obs = np.array([p.obs_u[e], p.obs_v[e], p.obs_ur[e]])
        err, A, B, _ = ba_ref.edge(kind, cam, trl, p.pose_qt[i], p.points[j], obs, False)
        dim = 3 if kind == EDGE_STEREO else 2
        err, A, B = err[:dim], A[:dim], B[:dim]
        om = p.obs_inv_sigma2[e]
        chi2 = om * err @ err
        delta = problem.DELTA_STEREO if kind == EDGE_STEREO else problem.DELTA_MONO_LBA
        rho1 = ba_ref.huber(chi2, delta)[1]
        w = rho1 * om
        cols = [(6 * len(free) + 3 * j, A)]
        if hidx[i] >= 0:
            cols.append((6 * hidx[i], B))
        for c0, J0 in cols:
            b[c0:c0 + J0.shape[1]] += -w * J0.T @ err
            for c1, J1 in cols:
                H[c0:c0 + J0.shape[1], c1:c1 + J1.shape[1]] += w * J0.T @ J1
    x = np.linalg.solve(H + lam * np.eye(n), b)
    want_pts = p.points + x[6 * len(free):].reshape(-1, 3)
    assert np.abs(res.points - want_pts).max() < 1e-9
    for k, i in enumerate(free):
        assert np.abs(res.pose_qt[i] - ba_ref.oplus(p.pose_qt[i], x[6 * k:6 * k + 6])).max() < 1e-9


# ---------------------------------------------------------------- LM semantics
def test_lm_trace_semantics_and_convergence():
    p = synthetic.config(1, scale=0.1)
    res = ba_ref.solve(p, problem.schedule_local_ba())
    assert res.trace[0]["lambda_"] > 0
    for a, b in zip(res.trace[:-1], res.trace[1:]):
        assert b["chi2_before"] == pytest.approx(a["chi2_after"], rel=1e-12)     # next linearisation starts where we stopped
        assert a["chi2_after"] <= a["chi2_before"]
    # a good step shrinks lambda by at least 1/3 and at most 2/3 (optimization_algorithm_levenberg.cpp:129-133)
    l0, l1 = res.trace[0]["lambda_"], res.trace[1]["lambda_"]
    assert l0 / 3 - 1e-12 <= l1 <= l0 * 2 / 3 + 1e-12
    # the 0.5 deg / 1 cm pose perturbation is pulled back towards the truth (points start better than the data can
    # resolve them, 2 cm, so they are not a convergence signal)
    free = p.pose_fixed == 0
    q0 = p.pose_qt[free, 3:] / np.linalg.norm(p.pose_qt[free, 3:], axis=1, keepdims=True)
    qt = p.truth["pose_qt"][free, 3:]
    ang0 = 2 * np.arccos(np.clip(np.abs((q0 * qt).sum(1)), 0, 1)).mean()
    ang1 = 2 * np.arccos(np.clip(np.abs((res.pose_qt[free, 3:] * qt).sum(1)), 0, 1)).mean()
    assert ang1 < 0.5 * ang0


def test_user_lambda_and_stop_flag():
    p = synthetic.config(1, scale=0.05)
    s = problem.schedule_local_ba(inertial=True)                  # setUserLambdaInit(100), Optimizer.cc:1197-1198
    res = ba_ref.solve(p, s)
    assert res.trace[0]["lambda_"] == pytest.approx(100.0 / 3.0) or res.trace[0]["lambda_"] < 100.0
    flag = np.array([1], np.uint8)
    s2 = problem.schedule_merge_ba(); s2.stop_flag = flag
    res2 = ba_ref.solve(p, s2)
    assert res2.status == 3 and len(res2.trace) == 0
    q = p.pose_qt.copy(); q[:, 3:] /= np.linalg.norm(q[:, 3:], axis=1, keepdims=True)
    assert np.abs(res2.pose_qt - q).max() < 1e-15 and np.array_equal(res2.points, p.points)


def test_merge_schedule_gates_and_drops_kernel():
    p = synthetic.config(1, scale=0.1)
    res = ba_ref.solve(p, problem.schedule_merge_ba())
    assert {t["round"] for t in res.trace} == {0, 1}
    assert 0 < res.edge_level.sum() < p.n_obs
    # every level-1 edge failed the gate with the chi2 it had after round 0
    lvl1 = res.edge_level == 1
    assert ((res.edge_chi2[lvl1] > 5.991) | (res.edge_depth_pos[lvl1] == 0)).all()
    # most injected gross outliers are caught
    assert (lvl1 & p.truth["outlier"]).sum() > 0.9 * p.truth["outlier"].sum()


def test_pose_optimization_oracle():
    b = synthetic.make_pose_batch(n_frames=24, n_matches=120)
    r = ba_ref.pose_opt_batch(b)
    assert r.n_inliers[7] == 0                                    # < 3 correspondences
    assert np.allclose(r.pose_qt[7, :3], b.pose_qt[7, :3])
    ok = [f for f in range(b.n_frames) if f not in (7, 11)]
    terr0 = np.abs(b.pose_qt[ok, :3] - b.truth["pose_qt"][ok, :3]).mean()
    terr1 = np.abs(r.pose_qt[ok, :3] - b.truth["pose_qt"][ok, :3]).mean()
    assert terr1 < 0.25 * terr0
    caught = (r.outlier == 1) & b.truth["outlier"]
    assert caught.sum() > 0.9 * b.truth["outlier"].sum()


# ---------------------------------------------------------------- committed golden fixtures
def test_golden_traces():
    with open(os.path.join(ROOT, "tests", "golden", "oracle_traces.json")) as f:
        gold = json.load(f)
    cases = [("C1_local", 1, 0.1, problem.schedule_local_ba()), ("C1_merge", 1, 0.1, problem.schedule_merge_ba()),
             ("C2_merge", 2, 0.1, problem.schedule_merge_ba()), ("C3_local", 3, 0.1, problem.schedule_local_ba()),
             ("C4_global", 4, 0.02, problem.schedule_global_ba(10))]
    for name, n, scale, sched in cases:
        p = synthetic.config(n, scale=scale, robust=(name != "C4_global"))
        g = gold[name]
        assert (p.n_poses, p.n_points, p.n_obs) == (g["n_poses"], g["n_points"], g["n_obs"]), name
        res = ba_ref.solve(p, sched)
        assert len(res.trace) == len(g["trace"]) and res.status == g["status"], name
        for t, gt in zip(res.trace, g["trace"]):
            assert (t["round"], t["iteration"], t["trials"], t["status"]) == (gt[0], gt[1], gt[5], gt[6]), name
            assert t["chi2_after"] == pytest.approx(gt[3], rel=1e-9), name
            assert t["lambda_"] == pytest.approx(gt[4], rel=1e-9), name
        assert res.pose_qt.sum() == pytest.approx(g["pose_sum"], rel=1e-10)
        assert int(res.outliers(p).sum()) == g["n_outliers"] and int(res.edge_level.sum()) == g["n_level1"]
    b = synthetic.make_pose_batch(n_frames=24, n_matches=120)
    r = ba_ref.pose_opt_batch(b)
    assert r.n_inliers.tolist() == gold["pose_batch"]["n_inliers"]
    assert int(r.outlier.sum()) == gold["pose_batch"]["n_outliers"]


# ---------------------------------------------------------------- the float fossil: atan2f
def test_fdlibm_atan2f_port_matches_libm():
    """csrc/fdlibm_atan2f.h (the code the GPU runs) vs this host's libm atan2f, bit for bit, 20M samples incl. raw bit
    patterns (NaN/inf/denormals) and BA-like magnitudes."""
    src = r'''
#include <math.h>
#include <stdio.h>
#include "%s/orb_slam3_study_kr_b200/csrc/fdlibm_atan2f.h"
int main(){ unsigned long long s=88172645463325252ULL; long bad=0,n=20000000;
  for(long i=0;i<n;i++){ s^=s<<13; s^=s>>7; s^=s<<17; float y,x;
    if(i&1){ y=((int)(s&0xffffff)-0x800000)/65536.0f*(float)((s>>24)&0xff)/16.0f; x=((int)((s>>32)&0xffffff)-0x800000)/65536.0f; }
    else { y=baf_flt_host((unsigned)s); x=baf_flt_host((unsigned)(s>>32)); }
    float a=atan2f(y,x), b=baf_atan2f(y,x);
    if(baf_bits_host(a)!=baf_bits_host(b) && !(a!=a && b!=b)) bad++; }
  printf("%%ld\n",bad); return 0; }
''' % ROOT
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c"); exe = os.path.join(d, "t")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, c, "-lm"])
        assert int(subprocess.check_output([exe]).decode().strip()) == 0
