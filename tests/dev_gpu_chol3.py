import sys, numpy as np
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api
ctx = api.Context(0)
rng = np.random.default_rng(1)
for n, bw, eps in [(120, 119, 1e-6), (120, 119, 1e-10), (78, 77, 1e-8), (300, 100, 1e-8), (2994, 180, 1e-8)]:
    B = rng.normal(size=(n, n))
    A = np.zeros((n, n))
    # banded Gram matrix: rows of B restricted to a sliding window so that A = B B^T is banded with half-width bw
    h = bw // 2
    Bb = np.zeros((n, n + h))
    for i in range(n):
        Bb[i, i:i + h + 1] = rng.normal(size=h + 1)
    A = Bb @ Bb.T
    ce = np.minimum(n - 1, np.arange(n) + 2 * h).astype(np.int32)
    # squash the spectrum: scale rows/cols wildly (like metres vs radians)
    sc = 10.0 ** rng.uniform(-3, 3, size=n)
    A = A * sc[:, None] * sc[None, :]
    b = rng.normal(size=n)
    lam = eps * np.abs(np.diag(A)).max()
    ref = np.linalg.solve(A + lam * np.eye(n), b)
    x, fail = ctx.test_solve(A, b, ce, lam)
    r = (A + lam * np.eye(n)) @ x - b
    print(f"n={n} bw={bw} eps={eps} cond={np.linalg.cond(A + lam*np.eye(n)):.2e} fail={fail} relerr={np.abs(x-ref).max()/np.abs(ref).max():.3e} resid={np.abs(r).max():.3e}", flush=True)
