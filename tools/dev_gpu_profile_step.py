"""One resident global-BA solve of the bench workload (used under ncu for the launch list / kernel captures)."""
import sys
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api, synthetic, problem
ctx = api.Context(0)
p = synthetic.global_ba_shard(0, 1, robust=False)
ctx.upload(p)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx.solve_resident(problem.schedule_global_ba(iters), download=False)
t = ctx.timing()
print({k: t[k] for k in ("solve_ms", "build_ms", "linsolve_ms", "update_ms", "total_launches", "lm_iterations")})
