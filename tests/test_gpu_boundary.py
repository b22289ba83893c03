"""GPU tests of the boundary itself: re-entrancy across threads and contexts, the stop flag written by another
thread while a solve runs, the starved-overlap path, and a compiled C++ consumer of the C ABI."""
import json
import subprocess
import threading
import time

import numpy as np
import pytest

from orb_slam3_study_kr_b200 import api, problem, synthetic

pytestmark = pytest.mark.gpu


def _same(a, b):
    assert [t["trials"] for t in a.trace] == [t["trials"] for t in b.trace] and a.status == b.status
    for x, y in zip(a.trace, b.trace):
        assert abs(x["chi2_after"] - y["chi2_after"]) <= 1e-12 * abs(y["chi2_after"])
    assert np.abs(a.pose_qt - b.pose_qt).max() < 1e-12 and np.abs(a.points - b.points).max() < 1e-10
    assert np.array_equal(a.edge_level, b.edge_level)


def test_four_threads_four_contexts_match_serial_runs():
    """Tracking (pose batch), LocalMapping (LBA), LoopClosing (merge LBA) and the GBA thread call concurrently, one context
    each (SURVEY 8b "Threading"); every result equals the one the same call gives when it runs alone."""
    jobs = {"lba": (synthetic.config(1), problem.schedule_local_ba()),
            "merge": (synthetic.config(2, scale=0.5), problem.schedule_merge_ba()),
            "gba": (synthetic.config(4, scale=0.25, robust=False), problem.schedule_global_ba(10))}
    batch = synthetic.make_pose_batch(n_frames=200, n_matches=200)
    serial = {}
    for k, (p, s) in jobs.items():
        c = api.Context(0); serial[k] = c.solve_ba(p, s); c.close()
    c = api.Context(0); serial["pose"] = c.pose_opt_batch(batch); c.close()

    got, errs = {}, []

    def run(k):
        try:
            ctx = api.Context(0)
            for _ in range(3):                                  # several calls per thread, so that the calls really interleave
                got[k] = ctx.pose_opt_batch(batch) if k == "pose" else ctx.solve_ba(*jobs[k])
            got[k + "_retries"] = ctx.timing()["solve_retries"]
            ctx.close()
        except Exception as e:                                  # noqa: BLE001
            errs.append((k, repr(e)))

    th = [threading.Thread(target=run, args=(k,)) for k in ("lba", "merge", "gba", "pose")]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for k in jobs:
        _same(got[k], serial[k])
    assert np.array_equal(got["pose"].n_inliers, serial["pose"].n_inliers) and np.array_equal(got["pose"].outlier, serial["pose"].outlier)
    assert np.abs(got["pose"].pose_qt - serial["pose"].pose_qt).max() < 1e-12


def test_stop_flag_written_by_another_thread_mid_solve(ctx):
    """LocalMapping::InterruptBA sets the flag from another thread while the optimiser runs (src/LocalMapping.cc:929); the
    library polls it where g2o polls terminate(): the call returns BAGPU_STOPPED with fewer iterations and valid estimates."""
    p = synthetic.config(4, scale=0.5, robust=False)
    s = problem.schedule_global_ba(20)
    full = ctx.solve_ba(p, s)
    flag = np.zeros(1, np.uint8)
    s2 = problem.schedule_global_ba(20); s2.stop_flag = flag
    t = threading.Timer(0.004, lambda: flag.__setitem__(0, 1))
    t.start()
    part = ctx.solve_ba(p, s2)
    t.join()
    assert part.status == 3 and 0 <= len(part.trace) < len(full.trace)
    assert np.isfinite(part.pose_qt).all() and np.isfinite(part.points).all()
    for a, b in zip(part.trace[:-1], full.trace):               # what ran before the flag is the same trajectory
        assert a["trials"] == b["trials"] and abs(a["chi2_after"] - b["chi2_after"]) <= 1e-12 * b["chi2_after"]


def test_update_estimates_keeps_the_plan(ctx):
    """bagpu_update_estimates: new poses / points for the resident map (first step). Re-optimising the same window from
    other estimates through the kept plan gives bit for bit what a fresh bagpu_solve_ba of that problem gives, and a reset goes back to
    the NEW uploaded state."""
    p = synthetic.config(2, scale=0.3)
    rng = np.random.default_rng(11)
    pose2 = p.pose_qt.copy(); pose2[:, :3] += rng.normal(0, 0.01, (p.n_poses, 3))
    pts2 = p.points + rng.normal(0, 0.02, p.points.shape)
    q = problem.BAProblem(pose2, p.pose_fixed, pts2, p.cameras, p.rigs, p.obs_pose, p.obs_point, p.obs_cam, p.obs_rig, p.obs_kind,
                          p.obs_flags, p.obs_u, p.obs_v, p.obs_ur, p.obs_inv_sigma2)
    s = problem.schedule_merge_ba()
    fresh = ctx.solve_ba(q, s)
    ctx.upload(p)
    ctx.solve_resident(s, download=False)                       # the map has been optimised (and its edge levels changed) before the update
    ctx.update_estimates(pose2, pts2)
    kept = ctx.solve_resident(s)
    assert [t["trials"] for t in kept.trace] == [t["trials"] for t in fresh.trace]
    assert np.array_equal(kept.pose_qt, fresh.pose_qt) and np.array_equal(kept.points, fresh.points)
    assert np.array_equal(kept.edge_level, fresh.edge_level)
    ctx.reset_resident()
    again = ctx.solve_resident(s)
    assert np.array_equal(again.pose_qt, fresh.pose_qt)
    ctx.update_estimates(points=p.points)                       # points only: poses stay the updated ones
    mixed = ctx.solve_resident(s)
    assert np.isfinite(mixed.pose_qt).all() and not np.array_equal(mixed.points, fresh.points)


def test_per_iteration_phase_record(ctx):
    """bagpu_trace carries the G2OBatchStatistics view (Thirdparty/g2o/g2o/core/batch_stats.h:40-62): edges and per-phase times."""
    p = synthetic.config(2)
    got = ctx.solve_ba(p, problem.schedule_merge_ba())
    t = ctx.timing()
    assert len(got.trace) == t["lm_iterations"]
    n0 = got.trace[0]["active_edges"]
    assert n0 == p.n_obs
    second = [x for x in got.trace if x["round"] == 1]
    assert second and second[0]["active_edges"] == p.n_obs - int(got.edge_level.sum())
    for x in got.trace:
        assert x["linearise_schur_us"] > 0 and x["update_us"] > 0 and x["linear_solve_us"] > 0
        assert x["iteration_us"] >= 0.5 * max(x["linearise_schur_us"], x["update_us"])
    tot = sum(x["linearise_schur_us"] for x in got.trace) * 1e-3
    assert tot <= t["build_ms"] * 1.0001                         # the record splits the call's total (the rest: the lambda-init pass)


def test_cpp_adapter(tmp_path):
    """A compiled C++ consumer (tests/abi/adapter.cpp): gather like INTEGRATION.md, one bagpu_solve_ba, the classification loop of
    src/Optimizer.cc:1416-1460, error codes instead of exceptions."""
    from test_abi_and_host import build_cpp_adapter
    exe = build_cpp_adapter(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["rc"] in (0, 1, 2) and 1 <= r["n_trace"] <= 10 and r["version"] == 2
    assert r["chi2_last"] < 0.6 * r["chi2_first"] and r["active_edges"] == r["edges"]
    assert r["caught"] >= 0.9 * r["injected"] and r["false_pos"] <= 0.08 * r["edges"]
    assert r["rc_bad"] == -2 and r["rc_stop"] == 3 and r["stop_trace"] == 0
