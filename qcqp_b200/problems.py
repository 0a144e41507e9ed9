"""Workload generators: the four problem families of the reference's examples/, built cvxpy-free.

Each generator returns ``(forms, info)`` where ``forms = [(P, q, r, relop), ...]`` is the quadratic
data the reference's ``get_qcqp_form`` (utilities.py:318-347) would extract -- ``forms[0]`` the objective
already in *minimise* form with ``relop=None``, then one entry per scalar constraint with relop
``'<='`` or ``'=='`` -- and ``info`` carries ``maximize`` (sign flip applied by QCQP.suggest/improve,
qcqp.py:400,416) plus the generator's raw data.

The random draws follow the example scripts line for line so that a seed reproduces their instances:
  boolean_least_squares.py:6-10, maxcut.py:6-16, secondary_user_beamforming.py:18-34, circle_packing.py:7-15.
"""
import numpy as np
import scipy.sparse as sp


def _unit_constraints(n):
    """x_i^2 == 1 for every i (cvx.square(x) == 1 expanded to scalars, utilities.py:341-345)."""
    out = []
    zero = np.zeros(n)
    for i in range(n):
        P = sp.csr_matrix(([1.0], ([i], [i])), shape=(n, n))
        out.append((P, zero, -1.0, "=="))
    return out


def boolean_least_squares(n=10, m=15, seed=1):
    """minimize ||A x - b||^2  s.t. x_i^2 == 1   (examples/boolean_least_squares.py:6-15)."""
    if seed is not None:
        np.random.seed(seed)
    A = np.random.randn(m, n)
    b = np.random.randn(m, 1)
    P0 = A.T.dot(A)
    P0 = (P0 + P0.T) / 2.0
    q0 = (-2.0 * A.T.dot(b)).ravel()
    r0 = float(b.T.dot(b)[0, 0])
    forms = [(sp.csr_matrix(P0), q0, r0, None)] + _unit_constraints(n)
    return forms, dict(maximize=False, A=A, b=b, family="boolean_least_squares")


def maxcut(n=25, p=0.2, seed=1):
    """maximize 0.25 (sum(W) - x'Wx)  s.t. x_i^2 == 1   (examples/maxcut.py:6-21), vectorised."""
    if seed is not None:
        np.random.seed(seed)
    U = np.random.uniform(low=0.0, high=1.0, size=(n, n))
    W = np.triu(U, 1)
    W = W + W.T
    np.fill_diagonal(W, 1.0)
    W = (W < p).astype(float)
    P0 = sp.csr_matrix(0.25 * W)
    forms = [(P0, np.zeros(n), -0.25 * float(W.sum()), None)] + _unit_constraints(n)
    return forms, dict(maximize=True, W=W, family="maxcut")


def beamforming(n=20, m=5, l=2, tau=20.0, eta=2.0, seed=1):
    """minimize ||x||^2  s.t. (a_i'x)^2 + (b_i'x)^2 >= tau, (c_i'x)^2 + (d_i'x)^2 <= eta
    (examples/secondary_user_beamforming.py:18-43); x in R^{2n}."""
    if seed is not None:
        np.random.seed(seed)
    HR = np.random.randn(m, n)
    HI = np.random.randn(m, n)
    A = np.hstack((HR, HI))
    B = np.hstack((-HI, HR))
    GR = np.random.randn(l, n)
    GI = np.random.randn(l, n)
    Cm = np.hstack((GR, GI))
    D = np.hstack((-GI, GR))
    N = 2 * n
    zero = np.zeros(N)
    forms = [(sp.identity(N, format="csr"), zero, 0.0, None)]
    for i in range(m):
        P = -(np.outer(A[i], A[i]) + np.outer(B[i], B[i]))
        forms.append((sp.csr_matrix((P + P.T) / 2.0), zero, float(tau), "<="))
    for i in range(l):
        P = np.outer(Cm[i], Cm[i]) + np.outer(D[i], D[i])
        forms.append((sp.csr_matrix((P + P.T) / 2.0), zero, -float(eta), "<="))
    return forms, dict(maximize=False, family="beamforming")


def circle_packing(ncirc=5, B=10.0):
    """maximize r  s.t. X >= r, X <= B - r, r >= 0, (2r)^2 <= ||X_i - X_j||^2
    (examples/circle_packing.py:7-17).  Variable order [r, X(:) column-major] -> N = 2*ncirc + 1."""
    N = 2 * ncirc + 1

    def xi(d, i):  # X[d, i], column-major
        return 1 + 2 * i + d

    forms = []
    q0 = np.zeros(N)
    q0[0] = -1.0  # maximise r -> minimise -r
    forms.append((sp.csr_matrix((N, N)), q0, 0.0, None))
    empty = sp.csr_matrix((N, N))
    for i in range(ncirc):          # r - X <= 0
        for d in range(2):
            q = np.zeros(N); q[0] = 1.0; q[xi(d, i)] = -1.0
            forms.append((empty, q, 0.0, "<="))
    for i in range(ncirc):          # X + r - B <= 0
        for d in range(2):
            q = np.zeros(N); q[0] = 1.0; q[xi(d, i)] = 1.0
            forms.append((empty, q, -float(B), "<="))
    q = np.zeros(N); q[0] = -1.0    # -r <= 0
    forms.append((empty, q, 0.0, "<="))
    zero = np.zeros(N)
    for i in range(ncirc):
        for j in range(i + 1, ncirc):
            rows, cols, vals = [0], [0], [4.0]
            for d in range(2):
                a, b = xi(d, i), xi(d, j)
                rows += [a, b, a, b]
                cols += [a, b, b, a]
                vals += [-1.0, -1.0, 1.0, 1.0]
            forms.append((sp.csr_matrix((vals, (rows, cols)), shape=(N, N)), zero, 0.0, "<="))
    return forms, dict(maximize=True, family="circle_packing", ncirc=ncirc, B=B)


def random_qcqp(n=6, m=4, seed=0, density=1.0, eq_frac=0.3):
    """Small random instance with mixed relops, indefinite P_i and nonzero q_i (property tests)."""
    rs = np.random.RandomState(seed)
    forms = []
    for j in range(m + 1):
        M = rs.randn(n, n)
        if density < 1.0:
            M = M * (rs.rand(n, n) < density)
        P = (M + M.T) / 2.0
        if j == 0:
            P = P.dot(P.T) / n  # convex objective keeps phase 2 well-posed
        q = rs.randn(n)
        r = rs.randn() - (2.0 if j > 0 else 0.0)
        relop = None if j == 0 else ("==" if rs.rand() < eq_frac else "<=")
        forms.append((sp.csr_matrix(P), q, float(r), relop))
    return forms, dict(maximize=False, family="random")


def synthetic_sdr_solution(n, rank=16, seed=5):
    """A declared stand-in for the SDP solution X* when no SDP solve is wanted:
    V = randn(n+1, rank) row-normalised, X* = V V^T  (PSD, unit diagonal, X*[-1,-1] = 1)."""
    rs = np.random.RandomState(seed)
    V = rs.randn(n + 1, rank)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    return V.dot(V.T)
