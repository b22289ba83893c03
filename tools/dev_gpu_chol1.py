import sys, numpy as np
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api
ctx = api.Context(0)
rng = np.random.default_rng(0)
n, bw = 3000, 170
A = np.zeros((n, n))
for d in range(bw + 1):
    v = rng.normal(size=n - d)
    A[np.arange(d, n), np.arange(0, n - d)] = v
    A[np.arange(0, n - d), np.arange(d, n)] = v
A += np.eye(n) * (np.abs(A).sum(1).max() + 1.0)
ce = np.minimum(n - 1, np.arange(n) + bw)
b = rng.normal(size=n)
for _ in range(3):
    x, fail = ctx.test_solve(A, b, ce, 0.5)
print("fail", fail, "err", np.abs(x - np.linalg.solve(A + 0.5 * np.eye(n), b)).max())
