"""Development script (not a test): full config 5 on one GPU with different numbers of factorisation fronts (BAGPU_PARTS)."""
import os, sys, time
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api, synthetic, problem

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
p = synthetic.config(cfg, robust=False)
print(f"poses {p.n_poses} points {p.n_points} obs {p.n_obs}", flush=True)
s = problem.schedule_global_ba(20)
base = None
NO = {"BAGPU_NO_OVERLAP": "1"}
for parts, extra in [("12", NO), ("13", NO), ("14", NO), ("15", NO), ("16", NO), ("10", NO)]:
    os.environ["BAGPU_PARTS"] = parts
    for k in ("BAGPU_NO_OVERLAP", "BAGPU_SEP_TILED"):
        os.environ.pop(k, None)
    os.environ.update(extra)
    ctx = api.Context(0)
    ctx.upload(p)
    got = ctx.solve_resident(s)
    ctx.reset_resident()
    t1 = time.time()
    got = ctx.solve_resident(s)
    tm = ctx.timing()
    chis = [t["chi2_after"] for t in got.trace]
    if base is None:
        base = chis
    dev = max(abs(a - b) / b for a, b in zip(chis, base)) if len(chis) == len(base) else -1
    print(f"parts {parts:>2} {extra}: solver_parts {tm['solver_parts']} solve_ms {tm['solve_ms']:.1f} trials {tm['lm_trials']} per-trial {tm['solve_ms'] / tm['lm_trials']:.2f} ms | "
          f"build {tm['build_ms'] / tm['build_launches']:.2f} linsolve {tm['linsolve_ms'] / max(1, tm['linsolve_launches']):.2f} update {tm['update_ms'] / tm['update_launches']:.2f} | "
          f"max rel chi2 dev vs first {dev:.2e} retries {tm['solve_retries']}", flush=True)
    ctx.close()
