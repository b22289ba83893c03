"""Development script (not a test): one local BA (config 1-3, full size) through bagpu_solve_ba; prints the phase split. Used under ncu for
the launch list of a small map.   python tools/dev_gpu_local.py [config 1|2|3] [repeats]"""
import sys, time
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api, synthetic, problem
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = api.Context(0)
p = synthetic.config(cfg)
s = problem.schedule_local_ba()
for i in range(rep):
    t0 = time.perf_counter(); got = ctx.solve_ba(p, s); t1 = time.perf_counter()
    t = ctx.timing()
    print(f"C{cfg} call {i}: wall {1e3 * (t1 - t0):.3f} ms | upload {t['h2d_ms']:.3f} solve {t['solve_ms']:.3f} download {t['d2h_ms']:.3f} | build {t['build_ms']:.3f} lin {t['linsolve_ms']:.3f} upd {t['update_ms']:.3f} | "
          f"iters {t['lm_iterations']} trials {t['lm_trials']} launches {t['total_launches']}")
