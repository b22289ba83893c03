"""Development script (not a test): block-Jacobi PCG (BAGPU_SOLVER_PCG) against the direct band solvers on configs 4 and 5:
iterations and ms per solve, and how far the LM trajectory moves, per tolerance. Output feeds the table in DESIGN.md."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api, synthetic, problem

out = []
for cfg, iters in ((4, 20), (5, 6)):
    p = synthetic.config(cfg, robust=False)
    ctx = api.Context(0)
    ctx.upload(p)
    s = problem.schedule_global_ba(iters)
    ctx.solve_resident(s)
    ctx.reset_resident()
    base = ctx.solve_resident(s)
    tb = ctx.timing()
    row = {"config": cfg, "n": 6 * p.n_free, "solver": "cholesky", "solver_parts": tb["solver_parts"], "trials": tb["lm_trials"],
           "ms_per_solve": tb["linsolve_ms"] / max(1, tb["linsolve_launches"]), "ms_per_trial": tb["solve_ms"] / tb["lm_trials"]}
    print(json.dumps(row), flush=True); out.append(row)
    for tol in (1e-4, 1e-6, 1e-8, 1e-10, 1e-12):
        s2 = problem.schedule_global_ba(iters)
        s2.linear_solver = problem.SOLVER_PCG
        s2.pcg_tolerance = tol
        s2.pcg_max_iterations = 20000 if cfg == 4 else 6000
        ctx.reset_resident()
        t0 = time.time()
        got = ctx.solve_resident(s2)
        tm = ctx.timing()
        same = [a["trials"] for a in got.trace] == [b["trials"] for b in base.trace]
        m = min(len(got.trace), len(base.trace))
        dev = max(abs(a["chi2_after"] - b["chi2_after"]) / b["chi2_after"] for a, b in zip(got.trace[:m], base.trace[:m]))
        row = {"config": cfg, "solver": "pcg", "tol": tol, "trials": tm["lm_trials"], "pcg_iterations_per_solve": tm["pcg_iterations"] / max(1, tm["lm_trials"]),
               "ms_per_solve": tm["linsolve_ms"] / max(1, tm["linsolve_launches"]), "ms_per_trial": tm["solve_ms"] / max(1, tm["lm_trials"]),
               "same_trial_counts": same, "max_rel_chi2_dev": dev, "max_abs_pose_dev": float(np.abs(got.pose_qt - base.pose_qt).max()),
               "status": got.status, "wall_s": time.time() - t0}
        print(json.dumps(row), flush=True); out.append(row)
    ctx.close()
json.dump(out, open("gpurun_out/r2e_pcg_table.json", "w"), indent=1)
