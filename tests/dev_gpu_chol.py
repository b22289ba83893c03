import sys, numpy as np
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api
ctx = api.Context(0)
rng = np.random.default_rng(0)
def banded_spd(n, bw_of):
    A = np.zeros((n, n))
    ce = np.zeros(n, int)
    for j in range(n):
        ce[j] = min(n - 1, j + bw_of(j))
    ce = np.maximum.accumulate(ce)
    for j in range(n):
        for i in range(j, ce[j] + 1):
            # respect the true (non-monotone) envelope for content, monotone for storage
            if i - j <= bw_of(j): A[i, j] = A[j, i] = rng.normal()
    A = A + np.eye(n) * (np.abs(A).sum(1).max() + 1.0)
    return A, ce
for n, f in [(100, lambda j: 1000), (300, lambda j: 1000), (120, lambda j: 50), (294, lambda j: 150), (594, lambda j: 150), (594, lambda j: 60 + (j * 7) % 90), (1000, lambda j: 40), (3000, lambda j: 170)]:
    A, ce = banded_spd(n, f)
    b = rng.normal(size=n)
    x, fail = ctx.test_solve(A, b, ce, 0.5)
    want = np.linalg.solve(A + 0.5 * np.eye(n), b)
    print(n, "band", int((ce - np.arange(n)).max()) + 1, "fail", fail, "err", np.abs(x - want).max() / np.abs(want).max())
