"""The algebra behind the factored per-observation records of the linearise + Schur pass (csrc/schur_pairs.cuh), in numpy:
what stage_kernel writes (Y, X_l, N, m, q) reproduces what the reference accumulates (W Dinv W^T, w B^T B, B^T g, W Dinv b_l;
Thirdparty/g2o/g2o/core/block_solver.hpp:381-444, core/base_binary_edge.hpp:73-114) exactly. No GPU needed."""
import numpy as np


def skew_neg(X):
    x, y, z = X
    return np.array([[0.0, z, -y], [-z, 0.0, x], [y, -x, 0.0]])          # P = -[X]x, rows as in OptimizableTypes.cpp:154-157


def random_edge(rng, dim):
    """A = dr/dpoint (dim x 3), B = dr/dpose = M [P | I] (dim x 6), g = -rho1 w r, weight w."""
    M = rng.normal(size=(dim, 3))
    X = rng.normal(size=3) * 3 + np.array([0, 0, 6.0])
    A = rng.normal(size=(dim, 3))
    B = M @ np.hstack([skew_neg(X), np.eye(3)])
    w = abs(rng.normal()) + 0.1
    g = rng.normal(size=dim)
    return A, B, M, X, w, g


def test_factored_records_reproduce_the_schur_terms():
    rng = np.random.default_rng(7)
    k = 5                                                               # observations of one landmark
    edges = [random_edge(rng, 2 if i % 2 else 3) for i in range(k)]
    lam = 0.37
    Hll = sum(w * A.T @ A for A, B, M, X, w, g in edges) + lam * np.eye(3)
    bl = sum(A.T @ g for A, B, M, X, w, g in edges)
    L = np.linalg.cholesky(Hll)                                         # Hll + lambda I = L L^T
    Linv = np.linalg.inv(L)
    Dinv = np.linalg.inv(Hll)
    c = Linv @ bl
    recs = []
    for A, B, M, X, w, g in edges:
        W = w * B.T @ A                                                 # the Hpl block (6 x 3)
        P = skew_neg(X)
        V = w * M.T @ A
        Y = V @ Linv.T                                                  # what the Z record holds (with X)
        Z = np.vstack([P.T @ Y, Y])
        assert np.allclose(Z, W @ Linv.T, rtol=1e-12, atol=1e-12)
        N = w * M.T @ M
        m = M.T @ g
        q = Y @ c
        PI = np.hstack([P, np.eye(3)])
        assert np.allclose(PI.T @ N @ PI, w * B.T @ B, rtol=1e-12, atol=1e-12)       # Hpp contribution
        assert np.allclose(np.concatenate([P.T @ m, m]), B.T @ g, rtol=1e-12, atol=1e-12)   # bp contribution
        assert np.allclose(-np.concatenate([P.T @ q, q]), -W @ Dinv @ bl, rtol=1e-11, atol=1e-11)   # Schur rhs term
        recs.append((Y, P, W))
    # pair_kernel: block (a, b) of the Schur complement from two records
    for Ya, Pa, Wa in recs:
        for Yb, Pb, Wb in recs:
            G = Ya @ Yb.T
            blk = np.block([[Pa.T @ G @ Pb, Pa.T @ G], [G @ Pb, G]])
            assert np.allclose(blk, Wa @ Dinv @ Wb.T, rtol=1e-11, atol=1e-11)
    # update_z_kernel: x_l = Dinv (bl - sum W^T x_p) = L^-T (c - sum Y^T (P x_rot + x_trans))
    xs = [rng.normal(size=6) * 1e-2 for _ in edges]
    direct = Dinv @ (bl - sum(W.T @ x for (Y, P, W), x in zip(recs, xs)))
    t = sum(Y.T @ (P @ x[:3] + x[3:]) for (Y, P, W), x in zip(recs, xs))
    assert np.allclose(Linv.T @ (c - t), direct, rtol=1e-11, atol=1e-13)


def test_two_way_split_solves_the_band_system():
    """The two-way factorisation of chol.cuh in dense numpy: eliminate the top block top-down and the bottom block bottom-up,
    add the two partial Schur complements on the separator, solve it, substitute back."""
    rng = np.random.default_rng(3)
    n, band = 90, 12
    A = np.zeros((n, n))
    for d in range(band + 1):
        v = rng.normal(size=n - d)
        A[np.arange(d, n), np.arange(n - d)] = v
        A[np.arange(n - d), np.arange(d, n)] = v
    A += np.eye(n) * (np.abs(A).sum(1).max() + 1.0)
    b = rng.normal(size=n)
    rT, rB = 32, 64                                                     # T = [0, rT), M = [rT, rB), B = [rB, n); len(M) >= band
    T, M, Bk = slice(0, rT), slice(rT, rB), slice(rB, n)
    assert np.allclose(A[T, Bk], 0.0)
    SM = A[M, M] - A[M, T] @ np.linalg.solve(A[T, T], A[T, M]) - A[M, Bk] @ np.linalg.solve(A[Bk, Bk], A[Bk, M])
    rM = b[M] - A[M, T] @ np.linalg.solve(A[T, T], b[T]) - A[M, Bk] @ np.linalg.solve(A[Bk, Bk], b[Bk])
    x = np.zeros(n)
    x[M] = np.linalg.solve(SM, rM)
    x[T] = np.linalg.solve(A[T, T], b[T] - A[T, M] @ x[M])
    x[Bk] = np.linalg.solve(A[Bk, Bk], b[Bk] - A[Bk, M] @ x[M])
    assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-10, atol=1e-12)
