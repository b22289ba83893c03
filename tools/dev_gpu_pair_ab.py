"""Development script (not a test): the linearise+Schur pass through pair_tile_kernel (default) and pair_kernel (BAGPU_PAIR_LIST=1),
with the solve beside it and after it, on config 4 or 5. The options are read per context (bagpu_init), so one process covers all variants.
  python tools/dev_gpu_pair_ab.py [config 4|5] [LM iterations]"""
import os, sys, time
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api, synthetic, problem

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
p = synthetic.config(cfg, robust=False)
print(f"config {cfg}: poses {p.n_poses} points {p.n_points} obs {p.n_obs}", flush=True)
s = problem.schedule_global_ba(iters)
base = None
KEYS = ("BAGPU_PAIR_LIST", "BAGPU_NO_OVERLAP", "BAGPU_PARTS", "BAGPU_TILE_FMA")
variants = [{}, {"BAGPU_TILE_FMA": "1"}, {"BAGPU_PAIR_LIST": "1"}, {"BAGPU_NO_OVERLAP": "1"}, {"BAGPU_TILE_FMA": "1", "BAGPU_NO_OVERLAP": "1"}, {"BAGPU_PAIR_LIST": "1", "BAGPU_NO_OVERLAP": "1"}]
if cfg == 5:
    variants = [{}, {"BAGPU_NO_OVERLAP": "1"}, {"BAGPU_TILE_FMA": "1", "BAGPU_NO_OVERLAP": "1"}, {"BAGPU_PAIR_LIST": "1", "BAGPU_NO_OVERLAP": "1"}]
for extra in variants:
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(extra)
    ctx = api.Context(0)
    t0 = time.time(); ctx.upload(p); t_up = time.time() - t0
    got = ctx.solve_resident(s)
    ctx.reset_resident()
    got = ctx.solve_resident(s)
    tm = ctx.timing()
    chis = [t["chi2_after"] for t in got.trace]
    if base is None:
        base = chis
    dev = max(abs(a - b) / b for a, b in zip(chis, base)) if len(chis) == len(base) else -1
    print(f"{str(extra):<80} upload {1e3 * t_up:7.1f} ms | solve {tm['solve_ms']:8.2f} ms, {tm['lm_trials']} trials, {tm['solve_ms'] / tm['lm_trials']:.3f} ms per trial | "
          f"pass {tm['build_ms'] / tm['build_launches']:.3f} solve {tm['linsolve_ms'] / max(1, tm['linsolve_launches']):.3f} update {tm['update_ms'] / tm['update_launches']:.3f} | "
          f"parts {tm['solver_parts']} retries {tm['solve_retries']} | max rel chi2 dev vs first {dev:.2e}", flush=True)
    ctx.close()
