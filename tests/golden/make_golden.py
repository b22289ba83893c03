"""Regenerates tests/golden/*.json from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference ships no golden vectors for its BA path, and cannot be built here, so these fixtures
pin (a) the synthetic generators and (b) the oracle against silent drift; the constants block is what the
reference SOURCE fixes and is asserted independently in tests/test_oracle.py.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from orb_slam3_study_kr_b200 import problem, synthetic  # noqa: E402
from oracle import ba_ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def digest(res, p):
    return dict(
        n_poses=p.n_poses, n_points=p.n_points, n_obs=p.n_obs, status=res.status,
        trace=[[t["round"], t["iteration"], t["chi2_before"], t["chi2_after"], t["lambda_"], t["trials"], t["status"]] for t in res.trace],
        pose_sum=float(res.pose_qt.sum()), point_sum=float(res.points.sum()),
        pose_first=res.pose_qt[-1].tolist(), point_first=res.points[0].tolist(),
        n_outliers=int(res.outliers(p).sum()), n_level1=int(res.edge_level.sum()),
        chi2_sum=float(res.edge_chi2.sum()))


def main():
    out = {}
    cases = [("C1_local", 1, 0.1, problem.schedule_local_ba()), ("C1_merge", 1, 0.1, problem.schedule_merge_ba()),
             ("C2_merge", 2, 0.1, problem.schedule_merge_ba()), ("C3_local", 3, 0.1, problem.schedule_local_ba()),
             ("C4_global", 4, 0.02, problem.schedule_global_ba(10))]
    for name, n, scale, sched in cases:
        p = synthetic.config(n, scale=scale, robust=(name != "C4_global"))
        out[name] = digest(ba_ref.solve(p, sched), p)
    b = synthetic.make_pose_batch(n_frames=24, n_matches=120)
    r = ba_ref.pose_opt_batch(b)
    out["pose_batch"] = dict(n_obs=b.n_obs, n_inliers=r.n_inliers.tolist(), pose_sum=float(r.pose_qt.sum()),
                             n_outliers=int(r.outlier.sum()), final_chi2=r.final_chi2.tolist())
    with open(os.path.join(HERE, "oracle_traces.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "oracle_traces.json"))


if __name__ == "__main__":
    main()
