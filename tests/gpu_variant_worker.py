"""Worker for tests/test_gpu_parity.py::test_solver_variants_agree: one quarter-size global BA through bagpu_solve_ba, prints the
chi2 trajectory and checksums. The code path is selected by BAGPU_* environment variables read by libbagpu at first use."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam3_study_kr_b200 import api, problem, synthetic  # noqa: E402

ctx = api.Context(0)
cfg = int(os.environ.get("VARIANT_CONFIG", "4"))
if cfg == 4:
    p = synthetic.config(4, scale=float(os.environ.get("VARIANT_SCALE", "0.25")), robust=False)
    sched = problem.schedule_global_ba(6)
else:                                   # a local-BA map (configs 1-3) through the merge schedule: two rounds and the gates between them
    p = synthetic.config(cfg, scale=float(os.environ.get("VARIANT_SCALE", "1.0")))
    sched = problem.schedule_merge_ba()
got = ctx.solve_ba(p, sched)
print(json.dumps({"chi2": [t["chi2_after"] for t in got.trace], "trials": [t["trials"] for t in got.trace],
                  "pose_sum": float(np.abs(got.pose_qt).sum()), "point_sum": float(np.abs(got.points).sum()),
                  "launches": int(ctx.timing()["total_launches"]), "solver_parts": int(ctx.timing()["solver_parts"]),
                  "levels": int(got.edge_level.sum()), "status": int(got.status)}))
