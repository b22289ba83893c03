import sys, numpy as np
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api
ctx = api.Context(0)
def case(n, h, env, scale, seed=1):
    rng = np.random.default_rng(seed)
    Bb = np.zeros((n, n + h))
    for i in range(n):
        Bb[i, i:i + h + 1] = rng.normal(size=h + 1)
    A = Bb @ Bb.T
    ce = np.minimum(n - 1, np.arange(n) + env).astype(np.int32)
    sc = 10.0 ** rng.uniform(-scale, scale, size=n)
    A = A * sc[:, None] * sc[None, :]
    b = rng.normal(size=n)
    lam = 1e-8 * np.abs(np.diag(A)).max()
    ref = np.linalg.solve(A + lam * np.eye(n), b)
    x, fail = ctx.test_solve(A, b, ce, lam)
    print(f"n={n} h={h} env={env} scale={scale} fail={fail} relerr={np.abs(x-ref).max()/np.abs(ref).max():.3e}", flush=True)
case(300, 50, 100, 3)
