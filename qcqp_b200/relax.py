"""Host-side convex relaxations feeding the Suggest step (they stay on the host, as in the reference, where they are
CVXPY + SCS/MOSEK solves: solve_sdr qcqp.py:72-97, solve_spectral qcqp.py:41-70).  cvxpy and every SDP solver are absent
from this image, so this module carries two small NumPy solvers:

* ``_mixing_unit_diagonal`` -- for the unit-diagonal family (every constraint is x_i^2 = 1: Boolean least squares, MAXCUT)
  the SDR is  min <W0, X>  s.t.  diag(X) = 1, X >= 0,  solved in Burer-Monteiro form X = V V^T with unit rows by the
  mixing method (exact coordinate minimisation over one row at a time); rank k > sqrt(2(n+1)) makes every local optimum
  global.  O(n^2 k) per sweep -- a few seconds at n = 1000.
* ``_admm_sdp`` -- general small problems: min <C, X> s.t. <A_i, X> = b_i / <= b_i, X >= 0 by ADMM on the dual
  (Wen, Goldfarb, Yin 2010), one eigendecomposition of an (n+1) x (n+1) matrix per iteration.  Meant for example-sized
  problems (n up to a few hundred).

Both return the lifted solution X* ((n+1) x (n+1), X*[-1,-1] = 1) and the relaxation value, i.e. what QCQP.suggest
consumes as ``sdr_sol`` / ``sdr_bound``.  They are setup code, not part of the measured hot path.
"""
import numpy as np
import scipy.sparse as sp


def homogeneous_form(f):
    """[[P, q/2], [q^T/2, r]] (utilities.py:66-67), dense."""
    P = np.asarray(sp.csr_matrix(f.P).todense())
    q = np.asarray(f.qarray, dtype=np.float64).reshape(-1, 1)
    return np.block([[P, q / 2.0], [q.T / 2.0, np.array([[float(f.r)]])]])


def _is_unit_diagonal_family(form):
    n = form.n
    if form.m != n:
        return False
    seen = np.zeros(n, dtype=bool)
    for f in form.fs:
        P = sp.coo_matrix(f.P)
        if f.relop != "==" or P.nnz != 1 or np.any(f.qarray != 0) or f.r != -1.0:
            return False
        i = int(P.row[0])
        if P.col[0] != i or P.data[0] != 1.0 or seen[i]:
            return False
        seen[i] = True
    return bool(seen.all())


def _mixing_unit_diagonal(W, rank=None, sweeps=400, tol=1e-7, seed=0):
    """min <W, V V^T> over V with unit rows (W symmetric).  Returns (X = V V^T, value)."""
    N = W.shape[0]
    k = int(rank) if rank else int(np.ceil(np.sqrt(2.0 * N))) + 1
    rs = np.random.RandomState(seed)
    V = rs.randn(N, k)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    Woff = W - np.diag(np.diag(W))
    prev = np.inf
    val = float(np.sum(W * (V @ V.T)))
    for _ in range(sweeps):
        for i in range(N):
            gvec = Woff[i] @ V                      # sum_j W_ij v_j, j != i
            nrm = np.linalg.norm(gvec)
            if nrm > 0:
                V[i] = -gvec / nrm
        val = float(np.sum(W * (V @ V.T)))
        if abs(prev - val) <= tol * max(1.0, abs(val)):
            break
        prev = val
    return V @ V.T, val


def _proj_psd(M):
    w, Q = np.linalg.eigh((M + M.T) / 2.0)
    w = np.maximum(w, 0.0)
    return (Q * w) @ Q.T


def _admm_sdp(C, A_eq, b_eq, A_le, b_le, iters=100000, mu=1.0, tol=1e-7):
    """min <C, X>  s.t. <A_eq[i], X> = b_eq[i], <A_le[i], X> <= b_le[i], X >= 0  (all matrices symmetric N x N).
    Dual ADMM (SDPAD); inequalities carry a nonnegative slack.  Returns (X, primal value, info)."""
    N = C.shape[0]
    Amats = list(A_eq) + list(A_le)
    me, ml = len(A_eq), len(A_le)
    m = me + ml
    b = np.concatenate([np.asarray(b_eq, dtype=float), np.asarray(b_le, dtype=float)]) if m else np.zeros(0)
    # vectorise: variable is (X, s) with s >= 0 the slacks of the inequalities
    Amat = np.stack([Ai.ravel() for Ai in Amats]) if m else np.zeros((0, N * N))
    E = np.zeros((m, ml))
    for i in range(ml):
        E[me + i, i] = 1.0
    # scale the data: ADMM is sensitive to it
    cs = max(1.0, np.linalg.norm(C))
    Cn = C / cs
    rown = np.sqrt(np.sum(Amat * Amat, axis=1) + np.sum(E * E, axis=1))
    rown[rown == 0] = 1.0
    Amat = Amat / rown[:, None]; E = E / rown[:, None]; bn = b / rown
    M = Amat @ Amat.T + E @ E.T
    Minv = np.linalg.pinv(M)
    X = np.zeros((N, N)); s = np.zeros(ml)
    S = np.zeros((N, N)); z = np.zeros(ml)          # dual slacks of X and s
    y = np.zeros(m)
    info = {}
    for it in range(iters):
        # y-update
        rhs = mu * (bn - Amat @ X.ravel() - E @ s) + Amat @ (Cn - S).ravel() + E @ (0.0 - z)
        y = Minv @ rhs
        # (S, z)-update = projection of V onto the cone; (X, s) from the residual
        Aty = (Amat.T @ y).reshape(N, N)
        V = Cn - Aty - mu * X
        S = _proj_psd(V)
        Xn = (S - V) / mu
        vz = 0.0 - E.T @ y - mu * s
        z = np.maximum(vz, 0.0)
        sn = (z - vz) / mu
        X, s = Xn, sn
        if it % 50 == 0:
            pres = np.linalg.norm(Amat @ X.ravel() + E @ s - bn) / (1.0 + np.linalg.norm(bn))
            dres = np.linalg.norm((Cn - Aty - S).ravel()) / (1.0 + np.linalg.norm(Cn))
            gap = abs(np.sum(Cn * X) - bn @ y) / (1.0 + abs(np.sum(Cn * X)) + abs(bn @ y))
            info = dict(iters=it, pres=float(pres), dres=float(dres), gap=float(gap))
            if max(pres, dres, gap) < tol:
                break
            # residual balancing (tight: degenerate relaxations such as circle packing otherwise crawl for 10^4 iterations
            # with the two residuals a constant factor apart)
            if pres < dres / 2.0:
                mu /= 2.0
            elif dres < pres / 2.0:
                mu *= 2.0
    X = (X + X.T) / 2.0
    return X, float(np.sum(C * X)), info


def solve_sdr(form, rank=None, iters=100000, tol=1e-7, seed=0, **_cvxpy_solver_options):
    """`solver=`, `verbose=` ... of the reference's `prob.solve(*args, **kwargs)` (qcqp.py:92) are accepted and ignored.
    The SDP relaxation of the QCQP (solve_sdr, qcqp.py:72-97):
        minimize <W0, X>  s.t.  <Wi, X> <= 0 or == 0,  X[-1,-1] = 1,  X >= 0.
    Returns (X*, value)."""
    W0 = homogeneous_form(form.f0)
    N = form.n + 1
    if _is_unit_diagonal_family(form):
        # <Wi, X> = X_ii - X_NN = 0 with X_NN = 1: the unit-diagonal SDP
        X, val = _mixing_unit_diagonal(W0, rank=rank, tol=tol, seed=seed)
        return X, val
    A_eq, b_eq, A_le, b_le = [], [], [], []
    Enn = np.zeros((N, N)); Enn[-1, -1] = 1.0
    A_eq.append(Enn); b_eq.append(1.0)
    for f in form.fs:
        W = homogeneous_form(f)
        if f.relop == "==":
            A_eq.append(W); b_eq.append(0.0)
        else:
            A_le.append(W); b_le.append(0.0)
    X, val, info = _admm_sdp(W0, A_eq, b_eq, A_le, b_le, iters=iters, tol=tol)
    if max(info.get("pres", 1.0), info.get("dres", 1.0)) > 1e-3:
        raise Exception("Relaxation problem status: %s" % ("inaccurate (ADMM residuals %r)" % info))
    return X, val


def solve_spectral(form, iters=100000, tol=1e-7, **_cvxpy_solver_options):
    """(cvxpy solver options are accepted and ignored, as in solve_sdr.)
    The spectral relaxation with lambda = 1 (solve_spectral, qcqp.py:41-70): the same lifted SDP with all '<='
    constraints summed into one and all '==' constraints into one.  Returns (x, value) with x the scaled top eigenvector."""
    W0 = homogeneous_form(form.f0)
    N = form.n + 1
    Enn = np.zeros((N, N)); Enn[-1, -1] = 1.0
    A_eq, b_eq, A_le, b_le = [Enn], [1.0], [], []
    W1 = sum([homogeneous_form(f) for f in form.fs if f.relop == "<="], np.zeros((N, N)))
    W2 = sum([homogeneous_form(f) for f in form.fs if f.relop == "=="], np.zeros((N, N)))
    if np.any(W1 != 0):
        A_le.append(W1); b_le.append(0.0)
    if np.any(W2 != 0):
        A_eq.append(W2); b_eq.append(0.0)
    X, val, info = _admm_sdp(W0, A_eq, b_eq, A_le, b_le, iters=iters, tol=tol)
    if max(info.get("pres", 1.0), info.get("dres", 1.0)) > 1e-3:
        raise Exception("Relaxation problem status: %s" % ("inaccurate (ADMM residuals %r)" % info))
    w, v = np.linalg.eigh(X)
    x = np.sqrt(max(w[-1], 0.0)) * v[:-1, -1]
    return x, val
