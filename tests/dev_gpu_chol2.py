import sys, numpy as np
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api
ctx = api.Context(0)
rng = np.random.default_rng(0)
def case(n, bw, ragged=False):
    A = np.zeros((n, n))
    ce = np.minimum(n - 1, np.arange(n) + bw)
    if ragged:
        ce = np.minimum(n - 1, np.arange(n) + rng.integers(0, bw + 1, size=n))
        ce = np.maximum.accumulate(ce)
    for j in range(n):
        v = rng.normal(size=ce[j] - j + 1)
        A[j:ce[j] + 1, j] = v
        A[j, j:ce[j] + 1] = v
    A += np.eye(n) * (np.abs(A).sum(1).max() + 1.0)
    b = rng.normal(size=n)
    x, fail = ctx.test_solve(A, b, ce.astype(np.int32), 0.5)
    ref = np.linalg.solve(A + 0.5 * np.eye(n), b)
    print(f"n={n} bw={bw} ragged={ragged} fail={fail} err={np.abs(x - ref).max():.3e} ref={np.abs(ref).max():.3e}", flush=True)
for n, bw in [(18, 17), (32, 10), (40, 39), (100, 30), (333, 60), (1000, 170), (3000, 170), (2994, 209), (1500, 351), (700, 400)]:
    case(n, bw)
    case(n, bw, True)
# not positive definite -> fail flag
n = 200
A = np.eye(n); A[50, 50] = -1.0
x, fail = ctx.test_solve(A, np.ones(n), np.minimum(n - 1, np.arange(n) + 20).astype(np.int32), 0.0)
print("indefinite fail flag:", fail)
