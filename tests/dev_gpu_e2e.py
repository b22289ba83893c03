import sys, time, numpy as np
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api, synthetic, problem
ctx = api.Context(0)
p = synthetic.global_ba_shard(0, 1)
s = problem.schedule_global_ba(20)
ctx.pin_problem(p)
for i in range(3):
    t0 = time.perf_counter()
    r = ctx.solve_ba(p, s)
    t1 = time.perf_counter()
    tm = ctx.timing()
    print(f"wall {1e3*(t1-t0):.1f} ms | h2d {tm['h2d_ms']:.1f} solve {tm['solve_ms']:.1f} d2h {tm['d2h_ms']:.1f}", flush=True)
