"""Worker for tests/test_gpu_parity.py::test_solver_variants_agree: one quarter-size global BA through bagpu_solve_ba, prints the
chi2 trajectory and checksums. The code path is selected by BAGPU_* environment variables read by libbagpu at first use."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam3_study_kr_b200 import api, problem, synthetic  # noqa: E402

ctx = api.Context(0)
p = synthetic.config(4, scale=float(os.environ.get("VARIANT_SCALE", "0.25")), robust=False)
got = ctx.solve_ba(p, problem.schedule_global_ba(6))
print(json.dumps({"chi2": [t["chi2_after"] for t in got.trace], "trials": [t["trials"] for t in got.trace],
                  "pose_sum": float(np.abs(got.pose_qt).sum()), "point_sum": float(np.abs(got.points).sum()),
                  "launches": int(ctx.timing()["total_launches"]), "solver_parts": int(ctx.timing()["solver_parts"])}))
