import sys, numpy as np
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api, synthetic, problem
from oracle import ba_ref
ctx = api.Context(0)
for scale in (0.04, 0.1, 0.2):
    p = synthetic.config(4, scale=scale, robust=False)
    s = problem.schedule_global_ba(3)
    got = ctx.solve_ba(p, s); ref = ba_ref.solve(p, s)
    print(scale, p.n_poses, [ (t['trials'], t['chi2_after'], t['status']) for t in got.trace], "| ref", [ (t['trials'], t['chi2_after']) for t in ref.trace])
