"""Development script (not a test): upload phase of bagpu_solve_ba on config 4 / 5 with the library's lap timers (BAGPU_DEBUG=1)."""
import os, sys, time
os.environ["BAGPU_DEBUG"] = "1"
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api, synthetic, problem
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 5
p = synthetic.config(cfg, robust=False)
ctx = api.Context(0)
ctx.pin_problem(p)
for i in range(3):
    t0 = time.perf_counter(); ctx.upload(p); t1 = time.perf_counter()
    print(f"upload {i}: wall {1e3 * (t1 - t0):.2f} ms, h2d_ms {ctx.timing()['h2d_ms']:.2f}, bytes {ctx.timing()['h2d_bytes'] / 1e6:.1f} MB", flush=True)
