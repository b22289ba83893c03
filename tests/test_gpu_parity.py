"""GPU parity tests: libbagpu (through the C ABI) vs the CPU oracle on the same seeded inputs, vs the committed golden
traces, and size-independent properties at BASELINE.json's full sizes.

Tolerances are BASELINE.json's: per-iteration chi2 within 1e-6 relative, final poses/points within 1e-6, inlier sets
identical except for edges within 1e-9 (relative) of the chi2 gate."""
import json
import os

import numpy as np
import pytest

from orb_slam3_study_kr_b200 import api, problem, synthetic
from orb_slam3_study_kr_b200.problem import BAProblem, Round, Schedule
from oracle import ba_ref

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHI_RTOL, EST_ATOL, GATE_RTOL = 1e-6, 1e-6, 1e-9


def assert_parity(p, got, ref):
    assert len(got.trace) == len(ref.trace) and got.status == ref.status
    for a, b in zip(got.trace, ref.trace):
        assert (a["round"], a["iteration"], a["trials"], a["status"]) == (b["round"], b["iteration"], b["trials"], b["status"])
        assert abs(a["chi2_before"] - b["chi2_before"]) <= CHI_RTOL * abs(b["chi2_before"])
        assert abs(a["chi2_after"] - b["chi2_after"]) <= CHI_RTOL * abs(b["chi2_after"])
        assert abs(a["lambda_"] - b["lambda_"]) <= 1e-6 * abs(b["lambda_"])
    assert np.abs(got.pose_qt - ref.pose_qt).max() < EST_ATOL
    assert np.abs(got.points - ref.points).max() < EST_ATOL
    th = np.where(p.obs_kind == problem.EDGE_STEREO, problem.GATE_STEREO, problem.GATE_MONO)
    near = np.abs(ref.edge_chi2 - th) <= GATE_RTOL * th
    bad = ((got.outliers(p) != ref.outliers(p)) | (got.edge_level != ref.edge_level)) & ~near
    assert not bad.any(), int(bad.sum())
    assert np.array_equal(got.edge_depth_pos, ref.edge_depth_pos)
    assert np.abs(got.edge_chi2 - ref.edge_chi2).max() <= 1e-6 * max(1.0, np.abs(ref.edge_chi2).max())


CASES = [("C1_local", 1, 0.25, problem.schedule_local_ba, True), ("C1_merge", 1, 0.25, problem.schedule_merge_ba, True),
         ("C2_local", 2, 0.25, problem.schedule_local_ba, True), ("C2_merge", 2, 0.25, problem.schedule_merge_ba, True),
         ("C3_local", 3, 0.25, problem.schedule_local_ba, True), ("C3_merge", 3, 0.25, problem.schedule_merge_ba, True),
         ("C4_global_nonrobust", 4, 0.04, lambda: problem.schedule_global_ba(10), False),
         ("C4_global_robust", 4, 0.04, lambda: problem.schedule_global_ba(20), True),
         ("C1_inertial_lambda", 1, 0.1, lambda: problem.schedule_local_ba(True), True)]


@pytest.mark.parametrize("name,n,scale,sched,robust", CASES, ids=[c[0] for c in CASES])
def test_ba_matches_oracle(ctx, name, n, scale, sched, robust):
    p = synthetic.config(n, scale=scale, robust=robust)
    s = sched()
    assert_parity(p, ctx.solve_ba(p, s), ba_ref.solve(p, s))


def test_full_size_local_configs_match_oracle(ctx):
    """BASELINE configs 1-3 at full size (the oracle finishes each in < 1 s)."""
    for n in (1, 2, 3):
        p = synthetic.config(n)
        for s in (problem.schedule_local_ba(), problem.schedule_merge_ba()):
            assert_parity(p, ctx.solve_ba(p, s), ba_ref.solve(p, s))


def test_golden_fixture(ctx):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_traces.json")))
    for name, n, scale, sched in [("C1_merge", 1, 0.1, problem.schedule_merge_ba()), ("C3_local", 3, 0.1, problem.schedule_local_ba()),
                                  ("C4_global", 4, 0.02, problem.schedule_global_ba(10))]:
        p = synthetic.config(n, scale=scale, robust=(name != "C4_global"))
        got = ctx.solve_ba(p, sched)
        g = gold[name]
        assert len(got.trace) == len(g["trace"]) and got.status == g["status"]
        for t, gt in zip(got.trace, g["trace"]):
            assert (t["round"], t["iteration"], t["trials"], t["status"]) == (gt[0], gt[1], gt[5], gt[6])
            assert abs(t["chi2_after"] - gt[3]) <= CHI_RTOL * gt[3]
        assert int(got.outliers(p).sum()) == g["n_outliers"] and int(got.edge_level.sum()) == g["n_level1"]
        assert abs(got.pose_qt.sum() - g["pose_sum"]) < 1e-6


def test_edge_order_is_free(ctx):
    """Edges in arbitrary order go through the on-device permutation and come back in the caller's order."""
    p = synthetic.config(2, scale=0.1)
    perm = np.random.default_rng(5).permutation(p.n_obs)
    q = BAProblem(p.pose_qt, p.pose_fixed, p.points, p.cameras, p.rigs, p.obs_pose[perm], p.obs_point[perm], p.obs_cam[perm],
                  p.obs_rig[perm], p.obs_kind[perm], p.obs_flags[perm], p.obs_u[perm], p.obs_v[perm], p.obs_ur[perm], p.obs_inv_sigma2[perm])
    s = problem.schedule_merge_ba()
    a, b = ctx.solve_ba(p, s), ctx.solve_ba(q, s)
    assert np.abs(a.pose_qt - b.pose_qt).max() < 1e-8 and np.abs(a.points - b.points).max() < 1e-6
    assert (np.abs(a.edge_chi2[perm] - b.edge_chi2) <= 1e-6 * np.maximum(1.0, np.abs(b.edge_chi2))).all()
    assert np.array_equal(a.edge_level[perm], b.edge_level)
    assert_parity(q, b, ba_ref.solve(q, s))


def test_edge_cases(ctx):
    p = synthetic.config(1, scale=0.05)
    s = problem.schedule_merge_ba()
    # (a) a landmark observed more than 32 (and more than 64) times: duplicate one point's edges onto many poses
    rng = np.random.default_rng(2)
    reps = []
    for j, k in ((0, 40), (1, 70)):
        e = np.flatnonzero(p.obs_point == j)[0]
        for i in range(k):
            reps.append((i % p.n_poses, j, p.obs_u[e] + rng.normal(0, 1), p.obs_v[e] + rng.normal(0, 1)))
    reps = np.array(reps)
    cat = lambda a, b, dt: np.concatenate([a, np.asarray(b, dt)])
    q = BAProblem(p.pose_qt, p.pose_fixed, np.concatenate([p.points, [[0.0, 0.0, 5.0]]]), p.cameras, p.rigs,
                  cat(p.obs_pose, reps[:, 0], np.int32), cat(p.obs_point, reps[:, 1], np.int32),
                  cat(p.obs_cam, np.zeros(len(reps)), np.int16), cat(p.obs_rig, -np.ones(len(reps)), np.int16),
                  cat(p.obs_kind, np.zeros(len(reps)), np.uint8), cat(p.obs_flags, np.ones(len(reps)), np.uint8),
                  cat(p.obs_u, synthetic.f32(reps[:, 2]), np.float64), cat(p.obs_v, synthetic.f32(reps[:, 3]), np.float64),
                  cat(p.obs_ur, np.zeros(len(reps)), np.float64), cat(p.obs_inv_sigma2, np.ones(len(reps)), np.float64))
    # ... and the extra point (index n_points-1) has no observation at all: it must come back untouched
    got, ref = ctx.solve_ba(q, s), ba_ref.solve(q, s)
    assert_parity(q, got, ref)
    assert np.array_equal(got.points[-1], [0.0, 0.0, 5.0])
    # (b) every pose fixed: only the points move (structure-only BA)
    f = BAProblem(p.pose_qt, np.ones(p.n_poses, np.uint8), p.points, p.cameras, p.rigs, p.obs_pose, p.obs_point, p.obs_cam, p.obs_rig,
                  p.obs_kind, p.obs_flags, p.obs_u, p.obs_v, p.obs_ur, p.obs_inv_sigma2)
    got, ref = ctx.solve_ba(f, problem.schedule_local_ba()), ba_ref.solve(f, problem.schedule_local_ba())
    assert_parity(f, got, ref)
    # (c) stop flag already set: nothing runs, estimates are the (normalised) inputs
    s2 = problem.schedule_merge_ba(); s2.stop_flag = np.array([1], np.uint8)
    got = ctx.solve_ba(p, s2)
    assert got.status == 3 and len(got.trace) == 0 and np.array_equal(got.points, p.points)
    # (d) bad arguments are reported, not crashed on
    bad = BAProblem(p.pose_qt, p.pose_fixed, p.points, p.cameras, p.rigs, p.obs_pose + 1000, p.obs_point, p.obs_cam, p.obs_rig,
                    p.obs_kind, p.obs_flags, p.obs_u, p.obs_v, p.obs_ur, p.obs_inv_sigma2)
    with pytest.raises(api.BagpuError):
        ctx.solve_ba(bad, s)


def test_pose_optimization_batch_matches_oracle(ctx):
    b = synthetic.make_pose_batch(n_frames=120, n_matches=250)
    got, ref = ctx.pose_opt_batch(b), ba_ref.pose_opt_batch(b)
    assert np.abs(got.pose_qt - ref.pose_qt).max() < EST_ATOL
    assert np.array_equal(got.n_inliers, ref.n_inliers)
    assert np.array_equal(got.outlier, ref.outlier)
    assert np.abs(got.final_chi2 - ref.final_chi2).max() <= CHI_RTOL * np.abs(ref.final_chi2).max()
    assert got.n_inliers[7] == 0                      # < 3 correspondences (Optimizer.cc:996-997)


def test_device_atan2f_is_libm_exact(ctx):
    rng = np.random.default_rng(0)
    y = np.concatenate([rng.normal(0, 5, 2_000_000), rng.integers(0, 2**32, 2_000_000, dtype=np.uint64).astype(np.uint32).view(np.float32)]).astype(np.float32)
    x = np.concatenate([rng.normal(0, 5, 2_000_000), rng.integers(0, 2**32, 2_000_000, dtype=np.uint64).astype(np.uint32).view(np.float32)]).astype(np.float32)
    d, h = ctx.device_atan2f(y, x), ba_ref.atan2f(y, x)
    same = (d.view(np.uint32) == h.view(np.uint32)) | (np.isnan(d) & np.isnan(h))
    assert same.all(), int((~same).sum())


def test_full_size_global_ba_properties(ctx):
    """BASELINE config 4 at full size (500 KFs, 200k points, ~2M observations, 20 iterations): properties that do not
    need the oracle -- chi2 never increases, the LM bookkeeping is consistent, solving the resident problem twice gives
    the same trajectory, and the estimate moves towards the ground truth."""
    p = synthetic.config(4, robust=False)
    assert p.n_poses == 500 and p.n_points == 200000 and 1.8e6 < p.n_obs < 2.3e6
    s = problem.schedule_global_ba(20)
    ctx.upload(p)
    a = ctx.solve_resident(s)
    t = ctx.timing()
    assert t["lm_iterations"] == len(a.trace) and t["lm_trials"] == sum(x["trials"] for x in a.trace)
    for x, y in zip(a.trace[:-1], a.trace[1:]):
        assert x["chi2_after"] <= x["chi2_before"] and abs(y["chi2_before"] - x["chi2_after"]) <= 1e-9 * x["chi2_after"]
    assert a.trace[-1]["chi2_after"] < 0.5 * a.trace[0]["chi2_before"]
    ctx.upload(p)
    b = ctx.solve_resident(s)
    assert [x["trials"] for x in a.trace] == [x["trials"] for x in b.trace]
    assert abs(a.trace[-1]["chi2_after"] - b.trace[-1]["chi2_after"]) <= 1e-9 * a.trace[-1]["chi2_after"]
    assert np.abs(a.pose_qt - b.pose_qt).max() < 1e-8
    free = p.pose_fixed == 0
    e0 = np.abs(p.pose_qt[free, :3] - p.truth["pose_qt"][free, :3]).mean()
    e1 = np.abs(a.pose_qt[free, :3] - p.truth["pose_qt"][free, :3]).mean()
    assert e1 < 0.5 * e0


@pytest.mark.parametrize("robust,iters", [(False, 20), (True, 20), (False, 10)], ids=["nonrobust_20", "robust_20_Tracking2603", "nonrobust_10_LoopClosing2289"])
def test_full_size_global_ba_matches_oracle(ctx, robust, iters):
    """BASELINE config 4 at FULL size (500 KFs, 200k points, ~2.0M observations) against the oracle, non-robust and robust
    (Tracking.cc:2603: GlobalBundleAdjustemnt(map, 20), bRobust; LoopClosing.cc:2289: 10 iterations, non-robust). The oracle
    needs ~0.75 s per LM iteration on one core, so each case costs 10-20 s."""
    p = synthetic.config(4, robust=robust)
    assert p.n_poses == 500 and p.n_points == 200000 and 1.8e6 < p.n_obs < 2.3e6
    s = problem.schedule_global_ba(iters)
    assert_parity(p, ctx.solve_ba(p, s), ba_ref.solve(p, s))


def test_quarter_size_global_ba_matches_oracle(ctx):
    """Config 4 at 1/4 size (125 KFs, 50k points, ~0.5M observations)."""
    p = synthetic.config(4, scale=0.25, robust=False)
    s = problem.schedule_global_ba(10)
    assert_parity(p, ctx.solve_ba(p, s), ba_ref.solve(p, s))


def test_solve_is_bitwise_reproducible(ctx):
    """The reduced camera system is summed in a fixed order (pair_kernel: the last finisher of a block adds the item partials in
    item order), so two solves of the same map agree bit for bit - also with the Cholesky running beside pair_kernel."""
    p = synthetic.config(4, scale=0.1, robust=False)
    s = problem.schedule_global_ba(6)
    a = ctx.solve_ba(p, s)
    b = ctx.solve_ba(p, s)
    assert [t["chi2_after"] for t in a.trace] == [t["chi2_after"] for t in b.trace]
    assert np.array_equal(a.pose_qt, b.pose_qt) and np.array_equal(a.points, b.points)
    assert np.array_equal(a.edge_chi2, b.edge_chi2)


def test_solver_variants_agree():
    """The same map through the other code paths of the library (what the multi-GPU trial, a profiler run or a wide envelope
    use): Cholesky after pair_kernel instead of beside it, one-way instead of two-way factorisation, re-linearising update,
    v1 global-atomic build. Trajectories agree to 1e-9 relative; the trial counts are identical."""
    import json
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpu_variant_worker.py")

    def run(extra):
        env = dict(os.environ)
        env.update(extra)
        out = subprocess.run([sys.executable, worker], capture_output=True, text=True, timeout=300, env=env)
        assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-2000:]
        return json.loads(out.stdout.strip().splitlines()[-1])

    base = run({})
    for extra in ({"BAGPU_NO_OVERLAP": "1"}, {"BAGPU_NO_TWOWAY": "1"}, {"BAGPU_NO_OVERLAP": "1", "BAGPU_NO_TWOWAY": "1"},
                  {"BAGPU_UPDATE_RELIN": "1"}, {"BAGPU_NO_TILES": "1"}, {"BAGPU_STAGE_FIRST": "0"}, {"BAGPU_PAIR_LIST": "1"},
                  {"BAGPU_PAIR_LIST": "1", "BAGPU_NO_OVERLAP": "1"}, {"BAGPU_TILE_FMA": "1"}):
        got = run(extra)
        assert got["trials"] == base["trials"], (extra, got["trials"], base["trials"])
        for a, b in zip(got["chi2"], base["chi2"]):
            assert abs(a - b) <= 1e-9 * abs(b), (extra, a, b)
        assert abs(got["pose_sum"] - base["pose_sum"]) <= 1e-9 * base["pose_sum"], extra
        assert abs(got["point_sum"] - base["point_sum"]) <= 1e-9 * base["point_sum"], extra


def test_chained_loop_equals_host_stepped_loop():
    """Small maps run the LM loop chained on the device (lm_decide_kernel: the decisions of optimization_algorithm_levenberg.cpp:99-166
    without a host round trip per trial); BAGPU_NO_CHAIN=1 runs the same map through the host-stepped loop. Same trials, levels and
    status, chi2 to 1e-10 (the device's pow against libm's), on the rig map (config 3: two edges per (pose, point) pair) and on config 2."""
    import json
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpu_variant_worker.py")

    def run(extra):
        env = dict(os.environ)
        env.update(extra)
        out = subprocess.run([sys.executable, worker], capture_output=True, text=True, timeout=300, env=env)
        assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-2000:]
        return json.loads(out.stdout.strip().splitlines()[-1])

    for cfg in ("3", "2"):
        chained = run({"VARIANT_CONFIG": cfg})
        stepped = run({"VARIANT_CONFIG": cfg, "BAGPU_NO_CHAIN": "1"})
        listed = run({"VARIANT_CONFIG": cfg, "BAGPU_PAIR_LIST": "1"})          # pair-list path (always host-stepped)
        for other in (stepped, listed):
            assert chained["trials"] == other["trials"] and chained["levels"] == other["levels"] and chained["status"] == other["status"]
            for a, b in zip(chained["chi2"], other["chi2"]):
                assert abs(a - b) <= 1e-10 * abs(b), (cfg, a, b)
            assert abs(chained["pose_sum"] - other["pose_sum"]) <= 1e-10 * other["pose_sum"]
        assert chained["launches"] != stepped["launches"]                          # the two loops really are different code paths


def test_many_edges_on_one_pose_point_pair(ctx):
    """More edges on one (pose, point) pair than the tile plan expresses (TP_MAX_LAYERS = 32): the upload falls back to the pair list,
    which takes any number; fewer go through the tile records as extra layers. Both against the oracle."""
    base = synthetic.config(1, scale=0.05)
    for copies in (3, 40):
        e = int(np.flatnonzero(~base.pose_fixed[base.obs_pose])[7])
        idx = np.concatenate([np.arange(base.n_obs), np.full(copies, e)])
        jitter = np.concatenate([np.zeros(base.n_obs), 0.25 * np.arange(1, copies + 1)])
        p = problem.BAProblem(base.pose_qt, base.pose_fixed, base.points, base.cameras, base.rigs, base.obs_pose[idx], base.obs_point[idx],
                              base.obs_cam[idx], base.obs_rig[idx], base.obs_kind[idx], base.obs_flags[idx], base.obs_u[idx] + jitter,
                              base.obs_v[idx], base.obs_ur[idx], base.obs_inv_sigma2[idx])
        s = problem.schedule_local_ba()
        ref = ba_ref.solve(p, s)
        got = ctx.solve_ba(p, s)
        assert [t["trials"] for t in got.trace] == [t["trials"] for t in ref.trace]
        for a, b in zip(got.trace, ref.trace):
            assert abs(a["chi2_after"] - b["chi2_after"]) <= 1e-6 * abs(b["chi2_after"])
        assert np.abs(got.pose_qt - ref.pose_qt).max() < 1e-6 and np.abs(got.points - ref.points).max() < 1e-6


def test_partitioned_solver_in_the_lm_loop():
    """Config 4 at 1/2 size (250 keyframes, n = 1494) and at full size with the partitioned solver forced (BAGPU_PARTS): the LM
    trajectory equals the default solver's, beside pair_kernel (row counters per sub-system) and after it."""
    import json
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpu_variant_worker.py")

    def run(extra):
        env = dict(os.environ)
        env.update(extra)
        out = subprocess.run([sys.executable, worker], capture_output=True, text=True, timeout=300, env=env)
        assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-2000:]
        return json.loads(out.stdout.strip().splitlines()[-1])

    for scale, variants in (("0.5", ({"BAGPU_PARTS": "3"}, {"BAGPU_PARTS": "3", "BAGPU_NO_OVERLAP": "1"})),
                            ("1.0", ({"BAGPU_PARTS": "4"}, {"BAGPU_PARTS": "5", "BAGPU_NO_OVERLAP": "1"}))):
        base = run({"VARIANT_SCALE": scale})
        assert base["solver_parts"] == 2
        for extra in variants:
            got = run(dict(extra, VARIANT_SCALE=scale))
            assert got["solver_parts"] == int(extra["BAGPU_PARTS"]), (extra, got["solver_parts"])
            assert got["trials"] == base["trials"], (extra, got["trials"], base["trials"])
            for a, b in zip(got["chi2"], base["chi2"]):
                assert abs(a - b) <= 1e-9 * abs(b), (extra, a, b)
            assert abs(got["pose_sum"] - base["pose_sum"]) <= 1e-9 * base["pose_sum"], extra
            assert abs(got["point_sum"] - base["point_sum"]) <= 1e-9 * base["point_sum"], extra
