"""One resident global-BA solve (used under ncu for the launch list / kernel captures).
  python tools/dev_gpu_profile_step.py [config 4|5] [LM iterations] [BAGPU_PARTS value or 0]
Under ncu the library sees the injection variables and queues the reduced-system solve AFTER pair_kernel (no overlap); with
parts > 0 the partitioned solver is forced, which is what a multi-GPU trial runs after its all-reduce."""
import os, sys
sys.path.insert(0, ".")
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
parts = sys.argv[3] if len(sys.argv) > 3 else "0"
if parts != "0":
    os.environ["BAGPU_PARTS"] = parts
    os.environ["BAGPU_NO_OVERLAP"] = "1"
from orb_slam3_study_kr_b200 import api, synthetic, problem
ctx = api.Context(0)
p = synthetic.config(cfg, robust=False)
ctx.upload(p)
ctx.solve_resident(problem.schedule_global_ba(iters), download=False)
t = ctx.timing()
print({k: t[k] for k in ("solve_ms", "build_ms", "linsolve_ms", "update_ms", "total_launches", "lm_iterations", "lm_trials", "solver_parts")})
