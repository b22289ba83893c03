"""Reduced-camera-system solver (chol_band_kernel / chol_solve_kernel) against numpy, through the C ABI's unit-test hook
bagpu_test_solve: (A + lambda I) x = b for symmetric positive definite A inside a monotone envelope col_end.
Replaces LinearSolverEigen::solve (Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-124)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _banded_spd(n, bw, rng, ragged=False):
    ce = np.minimum(n - 1, np.arange(n) + bw)
    if ragged:
        ce = np.maximum.accumulate(np.minimum(n - 1, np.arange(n) + rng.integers(0, bw + 1, size=n)))
    A = np.zeros((n, n))
    for j in range(n):
        v = rng.normal(size=ce[j] - j + 1)
        A[j:ce[j] + 1, j] = v
        A[j, j:ce[j] + 1] = v
    A += np.eye(n) * (np.abs(A).sum(1).max() + 1.0)
    return A, ce.astype(np.int32)


# (n, half-bandwidth): one panel, partial last panel, dense, 2/4/8/16-CTA clusters, envelope beyond the band kernel (tiled kernel)
SIZES = [(18, 17), (32, 10), (40, 39), (100, 30), (120, 119), (333, 60), (1000, 170), (900, 300), (700, 400)]


@pytest.mark.parametrize("n,bw", SIZES)
@pytest.mark.parametrize("ragged", [False, True])
def test_solver_matches_numpy(ctx, n, bw, ragged):
    rng = np.random.default_rng(n * 1000 + bw + int(ragged))
    A, ce = _banded_spd(n, bw, rng, ragged)
    b = rng.normal(size=n)
    x, fail = ctx.test_solve(A, b, ce, 0.5)
    ref = np.linalg.solve(A + 0.5 * np.eye(n), b)
    assert not fail
    assert np.abs(x - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("n,h,env", [(300, 50, 90), (300, 50, 100), (300, 100, 100), (300, 50, 128), (600, 50, 100), (2994, 90, 180)])
def test_solver_ill_conditioned(ctx, n, h, env):
    """Badly scaled Gram matrices (metres next to radians: condition number ~1e8) whose true band is narrower than the
    declared envelope; repeated to catch timing-dependent faults."""
    rng = np.random.default_rng(1)
    Bb = np.zeros((n, n + h))
    for i in range(n):
        Bb[i, i:i + h + 1] = rng.normal(size=h + 1)
    A = Bb @ Bb.T
    sc = 10.0 ** rng.uniform(-3, 3, size=n)
    A = A * sc[:, None] * sc[None, :]
    ce = np.minimum(n - 1, np.arange(n) + env).astype(np.int32)
    b = rng.normal(size=n)
    lam = 1e-8 * np.abs(np.diag(A)).max()
    ref = np.linalg.solve(A + lam * np.eye(n), b)
    for _ in range(3):
        x, fail = ctx.test_solve(A, b, ce, lam)
        assert not fail
        assert np.abs(x - ref).max() <= 1e-9 * np.abs(ref).max()


def test_solver_flags_indefinite(ctx):
    n = 200
    A = np.eye(n)
    A[50, 50] = -1.0
    _, fail = ctx.test_solve(A, np.ones(n), np.minimum(n - 1, np.arange(n) + 20).astype(np.int32), 0.0)
    assert fail
    # and the context keeps working afterwards
    x, fail = ctx.test_solve(np.eye(n) * 2.0, np.ones(n), np.minimum(n - 1, np.arange(n) + 20).astype(np.int32), 0.0)
    assert not fail and np.allclose(x, 0.5)


# ---------------------------------------------------------------- partitioned solver (chol_parts.cuh): P fronts, spikes, separator system
@pytest.mark.parametrize("n,bw,parts", [(1500, 60, 3), (2000, 100, 4), (3000, 180, 4), (2994, 185, 3), (4000, 90, 8), (6000, 185, 6), (2500, 30, 5)])
@pytest.mark.parametrize("ragged", [False, True])
def test_partitioned_solver_matches_numpy(ctx, n, bw, parts, ragged):
    """(A + lambda I) x = b through P factorisation fronts + spikes + the block-tridiagonal separator system, against numpy and
    against the one-front solver; n not a multiple of 32 or 96, ragged envelopes, up to 8 fronts."""
    rng = np.random.default_rng(n + 7 * bw + parts + int(ragged))
    A, ce = _banded_spd(n, bw, rng, ragged)
    b = rng.normal(size=n)
    x, fail = ctx.test_solve(A, b, ce, 0.5, parts=parts)
    assert not fail
    ref = np.linalg.solve(A + 0.5 * np.eye(n), b)
    assert np.abs(x - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    x1, _ = ctx.test_solve(A, b, ce, 0.5)
    assert np.abs(x - x1).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    # bit-for-bit reproducible: no atomics anywhere in the partitioned path
    x2, _ = ctx.test_solve(A, b, ce, 0.5, parts=parts)
    assert np.array_equal(x, x2)


def test_partitioned_solver_ill_conditioned_and_failure(ctx):
    """Badly scaled Gram matrix (condition ~1e8) through 4 fronts; an indefinite interior raises the failure flag like the
    one-front solver; a system too short for the requested fronts is refused with an error, not solved wrongly."""
    from orb_slam3_study_kr_b200 import api
    rng = np.random.default_rng(5)
    n, h, env = 2994, 90, 185
    Bb = np.zeros((n, n + h))
    for i in range(n):
        Bb[i, i:i + h + 1] = rng.normal(size=h + 1)
    A = Bb @ Bb.T
    sc = 10.0 ** rng.uniform(-3, 3, size=n)
    A = A * sc[:, None] * sc[None, :]
    ce = np.minimum(n - 1, np.arange(n) + env).astype(np.int32)
    b = rng.normal(size=n)
    lam = 1e-8 * np.abs(np.diag(A)).max()
    ref = np.linalg.solve(A + lam * np.eye(n), b)
    x, fail = ctx.test_solve(A, b, ce, lam, parts=4)
    assert not fail and np.abs(x - ref).max() <= 1e-9 * np.abs(ref).max()
    A2 = A.copy(); A2[1700, 1700] = -abs(A2[1700, 1700])
    _, fail = ctx.test_solve(A2, b, ce, lam, parts=4)
    assert fail
    with pytest.raises(api.BagpuError):
        ctx.test_solve(A[:600, :600], b[:600], np.minimum(599, ce[:600]), lam, parts=4)


def test_pcg_solver_in_the_lm_loop(ctx):
    """BAGPU_SOLVER_PCG (block-Jacobi PCG on the band-stored reduced system) behind the same seam: at a tight tolerance the LM
    trajectory equals the direct solver's; the iteration count is reported; a hopeless iteration budget is a rejected trial
    (failed linear solve), never an error."""
    from orb_slam3_study_kr_b200 import problem, synthetic
    p = synthetic.config(4, scale=0.25, robust=False)
    s = problem.schedule_global_ba(6)
    a = ctx.solve_ba(p, s)
    s2 = problem.schedule_global_ba(6)
    s2.linear_solver = problem.SOLVER_PCG
    s2.pcg_tolerance = 1e-13
    b = ctx.solve_ba(p, s2)
    t = ctx.timing()
    assert t["pcg_iterations"] > 0
    assert [x["trials"] for x in a.trace] == [x["trials"] for x in b.trace]
    for x, y in zip(a.trace, b.trace):
        assert abs(x["chi2_after"] - y["chi2_after"]) <= 1e-8 * abs(x["chi2_after"])
    assert np.abs(a.pose_qt - b.pose_qt).max() < 1e-6
    s3 = problem.schedule_global_ba(2)
    s3.linear_solver = problem.SOLVER_PCG
    s3.pcg_tolerance = 1e-13
    s3.pcg_max_iterations = 3
    c = ctx.solve_ba(p, s3)
    # every failed solve is a rejected trial: lambda grows (x nu, nu x 2) until the damped system is so diagonal that three
    # iterations suffice -- the LM loop's own answer to a failing linear solver (optimization_algorithm_levenberg.cpp:126-146)
    assert c.status in (0, 1, 2) and c.trace[0]["trials"] > 1 and c.trace[0]["lambda_"] > 1e3 * a.trace[0]["lambda_"]
    assert c.trace[0]["chi2_after"] < c.trace[0]["chi2_before"] and np.isfinite(c.pose_qt).all() and np.isfinite(c.points).all()
