"""Development script (not a test): one rank's share of config 5 (5000 keyframes, 250k landmarks, ~2.8 M observations) on one GPU."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api, synthetic, problem

t0 = time.time()
p = synthetic.global_ba_shard(0, 8, seed=5, n_kf=int(sys.argv[1]) if len(sys.argv) > 1 else 5000, points_per_rank=250000)
print(f"generated in {time.time() - t0:.1f} s: poses {p.n_poses} free {p.n_free} points {p.n_points} obs {p.n_obs}", flush=True)
ctx = api.Context(0)
s = problem.schedule_global_ba(8)
for k in range(2):
    t1 = time.time()
    got = ctx.solve_ba(p, s)
    tm = ctx.timing()
    print(f"solve {k}: wall {time.time() - t1:.3f} s solve_ms {tm['solve_ms']:.1f} build {tm['build_ms']:.1f} lin {tm['linsolve_ms']:.1f} upd {tm['update_ms']:.1f} "
          f"iters {tm['lm_iterations']} trials {tm['lm_trials']} status {got.status}")
for t in got.trace:
    print("  it", t["iteration"], "chi2", t["chi2_before"], "->", t["chi2_after"], "lambda", t["lambda_"], "trials", t["trials"], "status", t["status"])
chis = [t["chi2_after"] for t in got.trace]
assert all(b <= a * (1 + 1e-12) for a, b in zip(chis, chis[1:])), "chi2 must not increase"
assert np.isfinite(got.pose_qt).all() and np.isfinite(got.points).all()
print("c5 shard ok")
