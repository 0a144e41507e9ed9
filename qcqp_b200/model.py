"""A small modelling layer with the slice of the cvxpy 0.4 surface that the reference's front-end and examples use, so that
`QCQP(prob)` keeps its CVXPY-facing form where cvxpy itself is not installed:

    import qcqp_b200.model as cvx
    x = cvx.Variable(n)
    prob = cvx.Problem(cvx.Minimize(cvx.sum_squares(A*x - b)), [cvx.square(x) == 1])
    qcqp = QCQP(prob)

It replaces, for this path only, what the reference takes from cvxpy 0.4 + CVXcanon (setup.py:11-13, neither vendored):
`Variable`, `*` / `@` by constants and between affine expressions, `+`, `-`, indexing, `.T`, `square`, `power(., 2)`, `sum_squares`,
`quad_form`, `quad_over_lin`, `matrix_frac` (the README's list of quadratic expressions), `sum_entries`, the relations
`==`, `<=`, `>=`, `Minimize`, `Maximize`, `Problem` -- and the quadratic-coefficient extraction (`QuadCoeffExtractor`) behind
`get_qcqp_form` (utilities.py:318-347).  Conventions are cvxpy 0.4's: `.size` is the (rows, cols) pair, entries are ordered
column-major, `a >= b` is the constraint `b - a <= 0`, `prob.variables()` lists variables in order of first appearance
(objective first).  Everything here is host code that PRODUCES the pack's forms; evaluation goes through the engine.
"""
import itertools

import numpy as np
import scipy.sparse as sp

from .forms import QCQPForm, QuadraticFunction

_ids = itertools.count(1)


def _is_const(v):
    return isinstance(v, (int, float, np.integer, np.floating, np.ndarray, np.matrix, list, tuple)) or sp.issparse(v)


def _const_2d(v):
    """A constant as a dense 2-D float array (scalars become 1 x 1, 1-D arrays columns, as cvxpy 0.4 reads them)."""
    if sp.issparse(v):
        v = v.toarray()
    a = np.asarray(v, dtype=np.float64)
    if a.ndim == 0:
        return a.reshape(1, 1)
    if a.ndim == 1:
        return a.reshape(-1, 1)
    if a.ndim != 2:
        raise Exception("constants must be scalars, vectors or matrices")
    return a


class _Ctx:
    """One extraction: the variable offsets, N, and the coefficients of the nodes already visited (shared sub-expressions
    such as a Variable used by every constraint are canonicalised once)."""
    __slots__ = ("id_map", "N", "memo", "zero")

    def __init__(self, id_map, N):
        self.id_map, self.N, self.memo = id_map, N, {}
        self.zero = sp.csr_matrix((N, N))


class _Coeffs:
    """Coefficients of a (rows x cols) expression over the flattened variable x in R^N, entries column-major:
    entry i is  x' Ps[i] x + Q[i, :] x + r[i];  Ps is None for an affine expression."""
    __slots__ = ("Ps", "Q", "r")

    def __init__(self, Ps, Q, r):
        self.Ps, self.r = Ps, np.asarray(r, dtype=np.float64).ravel()
        self.Q = Q if isinstance(Q, sp.csr_matrix) else sp.csr_matrix(Q)

    @property
    def k(self):
        return self.r.size

    def promoted(self, k):
        """A scalar expression repeated k times (cvxpy promotes scalars against any shape)."""
        if self.k == k:
            return self
        assert self.k == 1
        Ps = None if self.Ps is None else [self.Ps[0]] * k
        return _Coeffs(Ps, sp.vstack([self.Q] * k, format="csr"), np.repeat(self.r, k))

    def scaled(self, w):
        """Entry i times the constant w[i]."""
        w = np.asarray(w, dtype=np.float64).ravel()
        Ps = None if self.Ps is None else [P * wi for P, wi in zip(self.Ps, w)]
        if w.size == 1 or np.all(w == w[0]):
            return _Coeffs(Ps, self.Q * float(w[0]), self.r * w)
        return _Coeffs(Ps, sp.diags(w).dot(self.Q), self.r * w)


class Expression:
    """Base of the expression tree.  Subclasses set `size` (rows, cols), `args`, `_deg` and implement `_canon`."""
    __array_ufunc__ = None          # ndarray (op) Expression defers to the reflected method below
    __hash__ = object.__hash__      # `==` builds a constraint, identity hashing is kept

    size = (1, 1)
    args = ()
    _deg = 0

    # ---- structure ----------------------------------------------------------------------------------------
    @property
    def shape(self):
        return self.size

    def variables(self):
        seen, out = set(), []
        for a in self.args:
            for v in a.variables():
                if v.id not in seen:
                    seen.add(v.id)
                    out.append(v)
        return out

    def is_quadratic(self):
        return self._deg <= 2

    def is_affine(self):
        return self._deg <= 1

    def is_constant(self):
        return self._deg == 0

    def _canon(self, ctx):
        raise NotImplementedError

    def canon(self, ctx):
        c = ctx.memo.get(id(self))
        if c is None:
            if not self.is_quadratic():
                raise Exception("expression is not quadratic")
            c = ctx.memo[id(self)] = self._canon(ctx)
        return c

    @property
    def value(self):
        """Numeric value at the variables' current values (None when one of them has none)."""
        vs = self.variables()
        if any(v.value is None for v in vs):
            return None
        id_map, N = get_id_map(vs)
        x = flatten_vars(vs, N)
        c = self.canon(_Ctx(id_map, N))
        out = c.Q.dot(x) + c.r
        if c.Ps is not None:
            out = out + np.array([x.dot(P.dot(x)) for P in c.Ps])
        return float(out[0]) if self.size == (1, 1) else out.reshape(self.size, order="F")

    @property
    def T(self):
        return _Index(self, None, transpose=True)

    # ---- arithmetic ---------------------------------------------------------------------------------------
    def __neg__(self):
        return _Scale(self, -1.0)

    def __add__(self, other):
        return _Add(self, as_expr(other))

    def __radd__(self, other):
        return _Add(as_expr(other), self)

    def __sub__(self, other):
        return _Add(self, -as_expr(other))

    def __rsub__(self, other):
        return _Add(as_expr(other), -self)

    def __mul__(self, other):       # expression * constant, or a product of two scalar expressions
        if _is_const(other):
            c = _const_2d(other)
            return _Scale(self, float(c[0, 0])) if c.size == 1 else _MatMul(None, self, c)
        return _Product(self, other)

    def __rmul__(self, other):      # constant * expression: a matrix product, as in cvxpy 0.4
        c = _const_2d(other)
        return _Scale(self, float(c[0, 0])) if c.size == 1 else _MatMul(c, self, None)

    __matmul__ = __mul__
    __rmatmul__ = __rmul__

    def __truediv__(self, other):
        c = _const_2d(other)
        if c.size != 1:
            raise Exception("can only divide by a scalar constant")
        return _Scale(self, 1.0 / float(c[0, 0]))

    __div__ = __truediv__

    def __getitem__(self, key):
        return _Index(self, key)

    # ---- relations (cvxpy 0.4: `a >= b` is LeqConstraint(b, a)) --------------------------------------------
    def __eq__(self, other):
        return Constraint(self, as_expr(other), "==")

    def __le__(self, other):
        return Constraint(self, as_expr(other), "<=")

    def __ge__(self, other):
        return Constraint(as_expr(other), self, "<=")

    def __lt__(self, other):
        raise Exception("strict inequalities are not allowed")

    __gt__ = __lt__


def as_expr(v):
    return v if isinstance(v, Expression) else Constant(v)


class Constant(Expression):
    def __init__(self, value):
        self._v = _const_2d(value)
        self.size = self._v.shape
        self._deg = 0

    def variables(self):
        return []

    def _canon(self, ctx):
        id_map, N = ctx.id_map, ctx.N
        k = self._v.size
        return _Coeffs(None, sp.csr_matrix((k, N)), self._v.ravel(order="F"))


class Variable(Expression):
    """Variable(rows=1, cols=1) as in cvxpy 0.4; Variable((rows, cols)) is accepted too.  `.value` is written by
    QCQP.suggest / improve (assign_vars, utilities.py:298-308): a float for a scalar, else a (rows, cols) array."""

    def __init__(self, rows=1, cols=1, name=None):
        if isinstance(rows, (tuple, list)):
            rows, cols = (tuple(rows) + (1,))[:2]
        self.size = (int(rows), int(cols))
        self.id = next(_ids)
        self.name = name if name is not None else "var%d" % self.id
        self._deg = 1
        self._value = None

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, val):
        if val is None:
            self._value = None
            return
        a = np.asarray(val, dtype=np.float64)
        if a.size != self.size[0] * self.size[1]:
            raise Exception("Invalid dimensions %s for Variable value." % (a.shape,))
        self._value = float(a.ravel()[0]) if self.size == (1, 1) else a.reshape(self.size).copy()

    def variables(self):
        return [self]

    def _canon(self, ctx):
        id_map, N = ctx.id_map, ctx.N
        k = self.size[0] * self.size[1]
        off = id_map[self.id]
        Q = sp.csr_matrix((np.ones(k), (np.arange(k), off + np.arange(k))), shape=(k, N))
        return _Coeffs(None, Q, np.zeros(k))

    def __repr__(self):
        return "Variable(%d, %d)" % self.size


class _Add(Expression):
    def __init__(self, a, b):
        if a.size != b.size and a.size != (1, 1) and b.size != (1, 1):
            raise Exception("Incompatible dimensions %s %s" % (a.size, b.size))
        self.args = (a, b)
        self.size = a.size if b.size == (1, 1) else b.size
        self._deg = max(a._deg, b._deg)

    def _canon(self, ctx):
        id_map, N = ctx.id_map, ctx.N
        k = self.size[0] * self.size[1]
        a = self.args[0].canon(ctx).promoted(k)
        b = self.args[1].canon(ctx).promoted(k)
        if a.Ps is None or b.Ps is None:
            Ps = a.Ps if b.Ps is None else b.Ps
        else:
            Ps = [x + y for x, y in zip(a.Ps, b.Ps)]
        return _Coeffs(Ps, a.Q + b.Q, a.r + b.r)


class _Scale(Expression):
    def __init__(self, a, w):
        self.args = (a,)
        self.size = a.size
        self._deg = a._deg
        self._w = float(w)

    def _canon(self, ctx):
        id_map, N = ctx.id_map, ctx.N
        c = self.args[0].canon(ctx)
        return c.scaled(np.full(c.k, self._w))


class _MatMul(Expression):
    """L * X (L constant, p x rows) or X * R (R constant, cols x p) for an affine X.  A scalar X against a matrix constant is
    promoted: the result is the constant's shape, entry (i, j) = c[i, j] * X."""

    def __init__(self, L, X, R):
        c = L if L is not None else R
        self.args = (X,)
        self._L, self._R = L, R
        self._deg = X._deg
        if X.size == (1, 1):
            self.size = c.shape
            self._promote = True
            return
        self._promote = False
        if L is not None:
            if L.shape[1] != X.size[0]:
                raise Exception("Incompatible dimensions %s %s" % (L.shape, X.size))
            self.size = (L.shape[0], X.size[1])
        else:
            if X.size[1] != R.shape[0]:
                raise Exception("Incompatible dimensions %s %s" % (X.size, R.shape))
            self.size = (X.size[0], R.shape[1])

    def _canon(self, ctx):
        id_map, N = ctx.id_map, ctx.N
        x = self.args[0].canon(ctx)
        c = self._L if self._L is not None else self._R
        if self._promote:
            return x.promoted(c.size).scaled(c.ravel(order="F"))
        if x.Ps is not None:
            raise Exception("a matrix product needs an affine expression")
        rows, cols = self.args[0].size
        if self._L is not None:                     # vec(L X) = (I_cols kron L) vec(X)
            K = sp.kron(sp.identity(cols), sp.csr_matrix(self._L), format="csr")
        else:                                       # vec(X R) = (R' kron I_rows) vec(X)
            K = sp.kron(sp.csr_matrix(self._R.T), sp.identity(rows), format="csr")
        return _Coeffs(None, K.dot(x.Q), K.dot(x.r))


class _Product(Expression):
    """(affine) * (affine): the matrix product of a (p x k) and a (k x q) affine expression (a scalar factor is promoted),
    entry (i, j) = sum_l A[i, l] B[l, j] with A[i, l] = a'x + alpha, B[l, j] = b'x + beta."""

    def __init__(self, a, b):
        if not isinstance(b, Expression):
            raise Exception("cannot multiply an expression by %s" % type(b))
        if a.size != (1, 1) and b.size != (1, 1) and a.size[1] != b.size[0]:
            raise Exception("Incompatible dimensions %s %s" % (a.size, b.size))
        self.args = (a, b)
        self.size = b.size if a.size == (1, 1) else a.size if b.size == (1, 1) else (a.size[0], b.size[1])
        self._deg = a._deg + b._deg

    def _canon(self, ctx):
        N = ctx.N
        a = self.args[0].canon(ctx)
        b = self.args[1].canon(ctx)
        if a.Ps is not None or b.Ps is not None:
            raise Exception("expression is not quadratic")
        (pa, ka), (kb, qb) = self.args[0].size, self.args[1].size
        Ps, Qrows, rs = [], [], []
        for j in range(self.size[1]):
            for i in range(self.size[0]):
                if (pa, ka) == (1, 1):              # scalar * matrix: entry (i, j) = a * B[i, j]
                    ra_, rb_ = [0], [i + j * kb]
                elif (kb, qb) == (1, 1):            # matrix * scalar
                    ra_, rb_ = [i + j * pa], [0]
                else:
                    ra_, rb_ = [i + l * pa for l in range(ka)], [l + j * kb for l in range(kb)]
                SA, SB = a.Q[ra_, :], b.Q[rb_, :]
                Ps.append(sp.csr_matrix(SA.T.dot(SB)))
                Qrows.append(sp.csr_matrix(SA.T.dot(b.r[rb_]) + SB.T.dot(a.r[ra_])).reshape(1, N))
                rs.append(float(a.r[ra_].dot(b.r[rb_])))
        return _Coeffs(Ps, sp.vstack(Qrows, format="csr"), rs)


class _Index(Expression):
    """X[key] with numpy semantics on the (rows, cols) grid, the result kept 2-D (X[:, i] is a column, X[i, :] a row);
    also the transpose."""

    def __init__(self, a, key, transpose=False):
        self.args = (a,)
        self._deg = a._deg
        rows, cols = a.size
        grid = np.arange(rows * cols).reshape((rows, cols), order="F")
        if transpose:
            sel = grid.T
        else:
            if not isinstance(key, tuple):
                key = (key, slice(None))
            if len(key) != 2:
                raise Exception("Invalid index/slice.")
            norm = []
            for kk, dim in zip(key, (rows, cols)):
                if isinstance(kk, (int, np.integer)):
                    kk = int(kk)
                    if kk < -dim or kk >= dim:
                        raise Exception("Index/slice out of bounds.")
                    kk = kk % dim
                    kk = slice(kk, kk + 1)
                norm.append(kk)
            sel = grid[norm[0], :][:, norm[1]]
        self.size = sel.shape
        self._sel = sel.ravel(order="F")

    def _canon(self, ctx):
        id_map, N = ctx.id_map, ctx.N
        c = self.args[0].canon(ctx)
        Ps = None if c.Ps is None else [c.Ps[i] for i in self._sel]
        return _Coeffs(Ps, c.Q[self._sel, :], c.r[self._sel])


class _Square(Expression):
    """Elementwise square of an affine expression: entry i is x'(a_i a_i')x + 2 b_i a_i'x + b_i^2."""

    def __init__(self, a):
        self.args = (a,)
        self.size = a.size
        self._deg = 2 * a._deg

    def _canon(self, ctx):
        id_map, N = ctx.id_map, ctx.N
        c = self.args[0].canon(ctx)
        if c.Ps is not None:
            raise Exception("expression is not quadratic")
        Ps = []
        for i in range(c.k):
            row = c.Q[i, :]
            Ps.append(sp.csr_matrix(row.T.dot(row)))
        return _Coeffs(Ps, sp.diags(2.0 * c.r).dot(c.Q), c.r * c.r)


class _QuadOver(Expression):
    """(Qx + r)' W (Qx + r) for an affine vector expression; W = I gives sum_squares."""

    def __init__(self, a, W=None):
        k = a.size[0] * a.size[1]
        if W is not None:
            W = _const_2d(W)
            if W.shape != (k, k) or min(a.size) != 1:
                raise Exception("Invalid dimensions for arguments.")
        self.args = (a,)
        self._W = W
        self.size = (1, 1)
        self._deg = 2 * a._deg

    def _canon(self, ctx):
        id_map, N = ctx.id_map, ctx.N
        c = self.args[0].canon(ctx)
        if c.Ps is not None:
            raise Exception("expression is not quadratic")
        if self._W is None:
            P = c.Q.T.dot(c.Q)
            q = 2.0 * c.Q.T.dot(c.r)
            r = float(c.r.dot(c.r))
        else:
            W = sp.csr_matrix(self._W)
            P = c.Q.T.dot(W.dot(c.Q))
            q = c.Q.T.dot((W + W.T).dot(c.r))
            r = float(c.r.dot(W.dot(c.r)))
        return _Coeffs([sp.csr_matrix(P)], sp.csr_matrix(np.asarray(q).reshape(1, N)), [r])


class _SumEntries(Expression):
    def __init__(self, a):
        self.args = (a,)
        self.size = (1, 1)
        self._deg = a._deg

    def _canon(self, ctx):
        id_map, N = ctx.id_map, ctx.N
        c = self.args[0].canon(ctx)
        Ps = None
        if c.Ps is not None:
            tot = sp.csr_matrix((N, N))
            for P in c.Ps:
                tot = tot + P
            Ps = [tot]
        return _Coeffs(Ps, sp.csr_matrix(c.Q.sum(axis=0)), [c.r.sum()])


# ---- atoms ------------------------------------------------------------------------------------------------------
def square(x):
    return _Square(as_expr(x))


def sum_squares(x):
    return _QuadOver(as_expr(x))


def quad_form(x, W):
    if isinstance(W, Expression):
        raise Exception("quad_form needs a constant matrix")
    return _QuadOver(as_expr(x), W)


def power(x, p):
    """power(affine, 2) (README "Quadratic expressions"); p = 1 is the expression itself."""
    if p == 2:
        return _Square(as_expr(x))
    if p == 1:
        return as_expr(x)
    raise Exception("power(x, %r) is not quadratic" % (p,))


def quad_over_lin(x, y):
    """sum_squares(x) / y for a constant y > 0."""
    if isinstance(y, Expression):
        raise Exception("quad_over_lin needs a constant denominator to stay quadratic")
    y = float(_const_2d(y)[0, 0])
    if not y > 0:
        raise Exception("quad_over_lin needs a positive denominator")
    return _Scale(_QuadOver(as_expr(x)), 1.0 / y)


def matrix_frac(x, P):
    """x' P^{-1} x for a constant positive definite P."""
    if isinstance(P, Expression):
        raise Exception("matrix_frac needs a constant matrix to stay quadratic")
    return _QuadOver(as_expr(x), np.linalg.inv(_const_2d(P)))


def sum_entries(x):
    """Sum of all entries; a plain number for constant input (maxcut.py:19 sums the adjacency matrix)."""
    if not isinstance(x, Expression):
        return float(_const_2d(x).sum())
    return _SumEntries(x)


# ---- constraints, objectives, problems --------------------------------------------------------------------------------
class Constraint:
    """lhs (OP_NAME) rhs with `_expr = lhs - rhs`, the attribute names get_qcqp_form reads (utilities.py:341-345)."""

    def __init__(self, lhs, rhs, op):
        self.args = (lhs, rhs)
        self.OP_NAME = op
        self._expr = lhs - rhs
        self.size = self._expr.size

    def variables(self):
        return self._expr.variables()

    def __bool__(self):
        raise Exception("Cannot evaluate the truth value of a constraint.")

    @property
    def violation(self):
        v = self._expr.value
        if v is None:
            return None
        return np.abs(v) if self.OP_NAME == "==" else np.maximum(v, 0)


class _Objective:
    NAME = None

    def __init__(self, expr):
        expr = as_expr(expr)
        if expr.size != (1, 1):
            raise Exception("The '%s' objective must resolve to a scalar." % self.NAME)
        self.args = [expr]

    def variables(self):
        return self.args[0].variables()

    @property
    def value(self):
        return self.args[0].value


class Minimize(_Objective):
    NAME = "minimize"


class Maximize(_Objective):
    NAME = "maximize"


class Problem:
    def __init__(self, objective, constraints=None):
        if not isinstance(objective, _Objective):
            raise Exception("Problem objective must be Minimize or Maximize.")
        self.objective = objective
        self.constraints = list(constraints) if constraints is not None else []
        for c in self.constraints:
            if not isinstance(c, Constraint):
                raise Exception("Problem has an invalid constraint of type %s" % type(c))

    def variables(self):
        """Variables in order of first appearance, the objective's first (cvxpy 0.4 Problem.variables)."""
        seen, out = set(), []
        for part in [self.objective] + self.constraints:
            for v in part.variables():
                if v.id not in seen:
                    seen.add(v.id)
                    out.append(v)
        return out


# ---- the reference's front-end helpers (utilities.py:290-347) ---------------------------------------------------------
def get_id_map(xs):
    """Offset of every variable in the flattened vector, and its length N (utilities.py:290-296)."""
    id_map, N = {}, 0
    for x in xs:
        id_map[x.id] = N
        N += x.size[0] * x.size[1]
    return id_map, N


def assign_vars(xs, vals):
    """Writes a flat vector back, column-major per variable; NaN when there is none (utilities.py:298-308)."""
    ind = 0
    for x in xs:
        size = x.size[0] * x.size[1]
        if vals is None:
            x.value = np.full(x.size, np.nan)
        else:
            x.value = np.reshape(np.asarray(vals[ind:ind + size], dtype=np.float64), x.size, order="F")
        ind += size


def flatten_vars(xs, n):
    """The variables' values as one vector (utilities.py:310-316).  The reference never advances its offset, so every
    variable lands at position 0 and the tail stays uninitialised; here the offset advances."""
    ret = np.empty(n)
    ind = 0
    for x in xs:
        size = x.size[0] * x.size[1]
        ret[ind:ind + size] = np.ravel(np.asarray(x.value, dtype=np.float64), order="F")
        ind += size
    return ret


def get_qcqp_form(prob):
    """The QCQPForm of a Problem: get_qcqp_form (utilities.py:318-347) with the coefficient extraction done here.
    A maximisation is stored negated; every entry of a constraint expression becomes one scalar constraint, column-major."""
    if not prob.objective.args[0].is_quadratic():
        raise Exception("Objective is not quadratic.")
    if not all(constr._expr.is_quadratic() for constr in prob.constraints):
        raise Exception("Not all constraints are quadratic.")
    id_map, N = get_id_map(prob.variables())
    ctx = _Ctx(id_map, N)
    zero = ctx.zero

    c0 = prob.objective.args[0].canon(ctx)
    P0 = c0.Ps[0] if c0.Ps is not None else zero
    q0 = np.asarray(c0.Q[0, :].todense()).ravel()
    r0 = float(c0.r[0])
    if prob.objective.NAME == "maximize":
        P0, q0, r0 = -P0, -q0, -r0
    f0 = QuadraticFunction(P0, q0, r0)          # symmetrises P, as utilities.py:333 does

    fs = []
    for constr in prob.constraints:
        c = constr._expr.canon(ctx)
        Qd = np.asarray(c.Q.todense())
        for i in range(c.k):
            fs.append(QuadraticFunction(c.Ps[i] if c.Ps is not None else zero, Qd[i], float(c.r[i]), constr.OP_NAME))
    return QCQPForm(f0, fs)
