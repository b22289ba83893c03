"""Development script (not a test): GPU vs oracle on small configs, verbose."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api, synthetic, problem
from oracle import ba_ref

ctx = api.Context(0)
y = np.random.default_rng(0).normal(size=1000000).astype(np.float32) * 5
x = np.random.default_rng(1).normal(size=1000000).astype(np.float32) * 5
d = ctx.device_atan2f(y, x); h = ba_ref.atan2f(y, x)
print("atan2f device vs libm mismatches:", int((d.view(np.uint32) != h.view(np.uint32)).sum()))

def cmp(name, p, s):
    t0 = time.time(); ref = ba_ref.solve(p, s); t1 = time.time()
    got = ctx.solve_ba(p, s); t2 = time.time()
    tm = ctx.timing()
    print(f"== {name}: poses {p.n_poses} free {p.n_free} points {p.n_points} obs {p.n_obs}  cpu {t1-t0:.3f}s gpu-call {t2-t1:.3f}s solve_ms {tm['solve_ms']:.2f} build_ms {tm['build_ms']:.2f} lin_ms {tm['linsolve_ms']:.2f} upd_ms {tm['update_ms']:.2f} status {got.status}/{ref.status}")
    n = min(len(got.trace), len(ref.trace))
    worst = 0
    for a, b in zip(got.trace[:n], ref.trace[:n]):
        rel = abs(a['chi2_after'] - b['chi2_after']) / abs(b['chi2_after'])
        worst = max(worst, rel)
        if rel > 1e-9 or a['trials'] != b['trials']:
            print("  it", a['round'], a['iteration'], "gpu", a['chi2_before'], a['chi2_after'], a['lambda_'], a['trials'], "| ref", b['chi2_before'], b['chi2_after'], b['lambda_'], b['trials'], "rel", rel)
    print(f"  trace len {len(got.trace)}/{len(ref.trace)} worst rel chi2 {worst:.3e}  pose maxdiff {np.abs(got.pose_qt-ref.pose_qt).max():.3e} point maxdiff {np.abs(got.points-ref.points).max():.3e}")
    print(f"  edge chi2 max rel diff {np.max(np.abs(got.edge_chi2-ref.edge_chi2)/(1e-9+np.abs(ref.edge_chi2))):.3e} level mismatches {(got.edge_level!=ref.edge_level).sum()} depth mismatches {(got.edge_depth_pos!=ref.edge_depth_pos).sum()} outlier-set mismatches {(got.outliers(p)!=ref.outliers(p)).sum()}")

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
for n in (1, 2, 3):
    p = synthetic.config(n, scale=scale)
    cmp(f"C{n} local", p, problem.schedule_local_ba())
    cmp(f"C{n} merge", p, problem.schedule_merge_ba())
p = synthetic.config(4, scale=scale * 0.2, robust=False)
cmp("C4 global nonrobust", p, problem.schedule_global_ba(10))
# shuffled edge order exercises the permutation path
rng = np.random.default_rng(5)
p = synthetic.config(2, scale=scale)
perm = rng.permutation(p.n_obs)
q = problem.BAProblem(p.pose_qt, p.pose_fixed, p.points, p.cameras, p.rigs, p.obs_pose[perm], p.obs_point[perm], p.obs_cam[perm], p.obs_rig[perm], p.obs_kind[perm], p.obs_flags[perm], p.obs_u[perm], p.obs_v[perm], p.obs_ur[perm], p.obs_inv_sigma2[perm])
cmp("C2 shuffled merge", q, problem.schedule_merge_ba())
b = synthetic.make_pose_batch(n_frames=60, n_matches=200)
t0 = time.time(); ref = ba_ref.pose_opt_batch(b); t1 = time.time(); got = ctx.pose_opt_batch(b); t2 = time.time()
print(f"== pose batch: cpu {t1-t0:.3f}s gpu {t2-t1:.3f}s solve_ms {ctx.timing()['solve_ms']:.3f}")
print("  pose maxdiff", np.abs(got.pose_qt - ref.pose_qt).max(), "outlier mismatches", int((got.outlier != ref.outlier).sum()), "inlier-count mismatches", int((got.n_inliers != ref.n_inliers).sum()), "chi2 rel", np.max(np.abs(got.final_chi2-ref.final_chi2)/(1e-9+np.abs(ref.final_chi2))))
