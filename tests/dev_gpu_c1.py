import sys, numpy as np
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api, synthetic, problem
ctx = api.Context(0)
p = synthetic.config(1, scale=0.25, robust=True)
s = problem.schedule_local_ba()
r = ctx.solve_ba(p, s)
for t in r.trace[:4]:
    print({k: t[k] for k in ("iteration", "trials", "chi2_before", "chi2_after", "lambda_", "status")})
