"""Development script (not a test): margin of the quarter-size global BA against the oracle, repeated (atomics order varies)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api, synthetic, problem
from oracle import ba_ref

ctx = api.Context(0)
p = synthetic.config(4, scale=0.25, robust=False)
s = problem.schedule_global_ba(10)
ref = ba_ref.solve(p, s)
cnt = np.bincount(p.obs_point, minlength=p.n_points)
runs = []
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    got = ctx.solve_ba(p, s)
    runs.append(got)
    d = np.abs(got.points - ref.points).max(axis=1)
    j = int(d.argmax())
    rel = max(abs(a["chi2_after"] - b["chi2_after"]) / abs(b["chi2_after"]) for a, b in zip(got.trace, ref.trace))
    print(f"run {k}: pose max {np.abs(got.pose_qt - ref.pose_qt).max():.3e} point max {d.max():.3e} (landmark {j}, {cnt[j]} obs, "
          f"99.99th pct {np.quantile(d, 0.9999):.3e}) chi2 rel {rel:.3e}; step |x| of that landmark {np.abs(got.points[j] - p.points[j]).max():.3e}")
for k in range(1, len(runs)):
    print(f"run {k} vs run 0: point max {np.abs(runs[k].points - runs[0].points).max():.3e} pose max {np.abs(runs[k].pose_qt - runs[0].pose_qt).max():.3e}")
