"""The cvxpy-free modelling layer (qcqp_b200/model.py): the reference's four example scripts, written
against it line for line, must extract the same quadratic forms -- same variable order, same constraint order, same relops --
as the hand-built generators the parity tests use (qcqp_b200/problems.py, pinned by tests/golden/golden.json)."""
import numpy as np
import pytest
import scipy.sparse as sp

import qcqp_b200.model as cvx
from qcqp_b200 import problems as pb
from helpers import OraclePack


def _same_forms(form, forms, tol=1e-12):
    got = form.forms()
    assert len(got) == len(forms)
    for j, ((P, q, r, op), (Pw, qw, rw, opw)) in enumerate(zip(got, forms)):
        assert op == opw, j
        Pw = sp.csr_matrix(Pw)
        scale = max(1.0, float(abs(Pw).max()) if Pw.nnz else 0.0)
        assert abs(P - Pw).max() <= tol * scale, j
        assert np.allclose(q, qw, rtol=0, atol=tol * max(1.0, np.abs(qw).max())), j
        assert abs(r - rw) <= tol * max(1.0, abs(rw)), j


def test_boolean_least_squares_script():
    """examples/boolean_least_squares.py:6-15."""
    n, m = 10, 15
    np.random.seed(1)
    A = np.random.randn(m, n)
    b = np.random.randn(m, 1)
    x = cvx.Variable(n)
    obj = cvx.sum_squares(A*x - b)
    cons = [cvx.square(x) == 1]
    prob = cvx.Problem(cvx.Minimize(obj), cons)
    forms, _ = pb.boolean_least_squares(n, m, seed=1)
    _same_forms(cvx.get_qcqp_form(prob), forms)
    assert prob.variables() == [x] or [v.id for v in prob.variables()] == [x.id]


def test_maxcut_script():
    """examples/maxcut.py:6-21 (np.matrix adjacency, Maximize, float - expression)."""
    n = 25
    np.random.seed(1)
    p = 0.2
    W = np.asmatrix(np.random.uniform(low=0.0, high=1.0, size=(n, n)))
    for i in range(n):
        W[i, i] = 1
        for j in range(i+1, n):
            W[j, i] = W[i, j]
    W = (W < p).astype(float)
    x = cvx.Variable(n)
    obj = 0.25*(cvx.sum_entries(W) - cvx.quad_form(x, W))
    cons = [cvx.square(x) == 1]
    prob = cvx.Problem(cvx.Maximize(obj), cons)
    forms, info = pb.maxcut(n, p, seed=1)
    assert info["maximize"] and prob.objective.NAME == "maximize"
    _same_forms(cvx.get_qcqp_form(prob), forms)


def test_beamforming_script():
    """examples/secondary_user_beamforming.py:18-43 (sums of squares of matrix products, `>=` flipped to `<=`)."""
    n, m, l = 20, 5, 2
    tau, eta = 20, 2
    np.random.seed(1)
    HR = np.random.randn(m, n)
    HI = np.random.randn(m, n)
    A = np.hstack((HR, HI))
    B = np.hstack((-HI, HR))
    GR = np.random.randn(l, n)
    GI = np.random.randn(l, n)
    C = np.hstack((GR, GI))
    D = np.hstack((-GI, GR))
    x = cvx.Variable(2*n)
    obj = cvx.Minimize(cvx.sum_squares(x))
    cons = [
        cvx.square(A*x) + cvx.square(B*x) >= tau,
        cvx.square(C*x) + cvx.square(D*x) <= eta
    ]
    prob = cvx.Problem(obj, cons)
    forms, _ = pb.beamforming(n=n, m=m, l=l, tau=tau, eta=eta, seed=1)
    _same_forms(cvx.get_qcqp_form(prob), forms)


@pytest.mark.parametrize("n", [5, 12])
def test_circle_packing_script(n):
    """examples/circle_packing.py:7-17: two variables (the scalar r first, because the objective names it first), matrix
    variable flattened column-major, scalar promotion in `X >= r` and `X <= B-r`, indexing `X[:, i]`."""
    X = cvx.Variable(2, n)
    B = 10
    r = cvx.Variable()
    obj = cvx.Maximize(r)
    cons = [X >= r, X <= B-r, r >= 0]
    for i in range(n):
        for j in range(i+1, n):
            cons.append(cvx.square(2*r) <= cvx.sum_squares(X[:, i]-X[:, j]))
    prob = cvx.Problem(obj, cons)
    assert [v.id for v in prob.variables()] == [r.id, X.id]
    forms, _ = pb.circle_packing(n, B=10.0)
    form = cvx.get_qcqp_form(prob)
    assert form.n == 2*n + 1 and form.m == 4*n + 1 + n*(n-1)//2
    _same_forms(form, forms)


def test_values_and_write_back():
    """assign_vars / flatten_vars (utilities.py:298-316) are column-major per variable and inverse to each other; an
    expression's value at the variables' values equals the extracted form's."""
    rs = np.random.RandomState(3)
    X = cvx.Variable(2, 3)
    r = cvx.Variable()
    y = cvx.Variable(4)
    M = rs.randn(5, 2)
    expr = cvx.sum_squares(M*X[:, 1] - rs.randn(5, 1)) + 3*cvx.square(r) - cvx.sum_entries(X*rs.randn(3, 2)) + cvx.quad_form(y, rs.randn(4, 4)) + r*y[2]
    prob = cvx.Problem(cvx.Minimize(expr), [X.T*np.ones((2, 1)) <= y[0:3] + 1, y[3] == 2*r])
    xs = prob.variables()
    assert [v.id for v in xs] == [X.id, r.id, y.id]
    assert expr.value is None
    v = rs.randn(11)
    cvx.assign_vars(xs, v)
    assert X.value.shape == (2, 3) and isinstance(r.value, float) and y.value.shape == (4, 1)
    assert X.value[1, 2] == v[5] and r.value == v[6] and y.value[0, 0] == v[7]
    assert np.array_equal(cvx.flatten_vars(xs, 11), v)
    form = cvx.get_qcqp_form(prob)
    P, q, c, _ = form.f0.as_tuple()
    assert abs(expr.value - (v.dot(P.dot(v)) + q.dot(v) + c)) < 1e-10
    assert form.m == 4 and [f.relop for f in form.fs] == ["<="] * 3 + ["=="]
    for i, f in enumerate(form.fs[:3]):
        want = X.value[:, i].sum() - y.value[i, 0] - 1
        assert abs(v.dot(f.P.dot(v)) + f.qarray.dot(v) + f.r - want) < 1e-12
    viol = prob.constraints[1].violation
    assert abs(viol - abs(v[10] - 2*v[6])) < 1e-12
    cvx.assign_vars(xs, None)
    assert np.isnan(y.value).all()


def test_error_behaviour():
    x = cvx.Variable(3)
    with pytest.raises(Exception, match="Objective is not quadratic"):
        cvx.get_qcqp_form(cvx.Problem(cvx.Minimize(cvx.square(cvx.sum_squares(x)))))
    with pytest.raises(Exception, match="Not all constraints are quadratic"):
        cvx.get_qcqp_form(cvx.Problem(cvx.Minimize(cvx.sum_squares(x)), [cvx.square(cvx.square(x)) <= 1]))
    with pytest.raises(Exception, match="Incompatible dimensions"):
        np.ones((2, 4))*x
    with pytest.raises(Exception, match="scalar"):
        cvx.Minimize(x)
    with pytest.raises(Exception):
        bool(x[0] <= 1)


# ---- the facade's host logic over the modelling layer, with the engine's pack replaced by the CPU oracle ------------------
def _golden_cd(golden, name):
    return [c for c in golden["cd"] if c["name"] == name][0]


def test_facade_flows_over_the_model_layer(monkeypatch, golden):
    """The reference's call sequence QCQP(prob); suggest(RANDOM); improve(COORD_DESCENT) on the goldens G1 and G3 (whose
    recipe is exactly that sequence), and G2' from a point the user writes into x.value: results, variable write-back,
    maximize sign and the process-global np.random stream must come out as the reference's own run recorded them."""
    import qcqp_b200 as Q
    from qcqp_b200 import engine
    monkeypatch.setattr(engine, "Pack", OraclePack)

    # G1: boolean least squares, seed(7); x0 = randn(10); improve_coord_descent
    g = _golden_cd(golden, "G1")
    np.random.seed(1)
    A = np.random.randn(15, 10)
    b = np.random.randn(15, 1)
    x = cvx.Variable(10)
    qc = Q.QCQP(cvx.Problem(cvx.Minimize(cvx.sum_squares(A*x - b)), [cvx.square(x) == 1]))
    assert not qc.maximize_flag
    np.random.seed(g["seed"])
    f_s, v_s = qc.suggest(Q.RANDOM)
    assert x.value.shape == (10, 1) and np.allclose(x.value.ravel(), g["x0"], rtol=0, atol=0)
    f, v = qc.improve(Q.COORD_DESCENT)
    assert abs(f - g["f0"]) <= 1e-9 * abs(g["f0"]) and abs(v - g["maxviol"]) <= 1e-6 * g["maxviol"]
    assert np.allclose(x.value.ravel(), g["x"], rtol=1e-9, atol=1e-9)
    assert np.random.get_state()[2] == g["rng"]["pos"]

    # G2': the user supplies the point; phase1=False; the stream is not consumed
    g = _golden_cd(golden, "G2p")
    x.value = np.array(g["x0"]).reshape(10, 1)
    state = np.random.get_state()
    f, v = qc.improve(Q.COORD_DESCENT, phase1=False)
    assert abs(f - g["f0"]) <= 1e-9 * abs(g["f0"]) and abs(v - g["maxviol"]) <= 1e-6 * g["maxviol"]
    assert np.allclose(x.value.ravel(), g["x"], rtol=1e-9, atol=1e-9)
    assert np.array_equal(np.random.get_state()[1], state[1]) and np.random.get_state()[2] == state[2]

    # G3: MAXCUT (Maximize): the returned objective is the cut value, +55.0003...
    g = _golden_cd(golden, "G3")
    forms, info = pb.maxcut(25, 0.2, seed=1)
    W = info["W"]
    y = cvx.Variable(25)
    qc = Q.QCQP(cvx.Problem(cvx.Maximize(0.25*(cvx.sum_entries(W) - cvx.quad_form(y, W))), [cvx.square(y) == 1]))
    assert qc.maximize_flag
    np.random.seed(g["seed"])
    qc.suggest(Q.RANDOM)
    f, v = qc.improve(Q.COORD_DESCENT)
    assert abs(f + g["f0"]) <= 1e-9 * abs(g["f0"]) and f > 0 and abs(v - g["maxviol"]) <= 1e-6 * g["maxviol"]
    assert np.random.get_state()[2] == g["rng"]["pos"]


def test_model_objects_drive_the_reference_class(monkeypatch, golden):
    """Build container only (the reference tree is absent on the GPU box): the UNMODIFIED reference `QCQP` class
    (qcqp.py:367-432) is run on a qcqp_b200.model Problem -- its assign_vars / flatten_vars / prob.variables() /
    objective.NAME see this package's Variable and Problem objects, only its cvxpy-bound get_qcqp_form is fed the forms
    extracted here -- and this package's facade must return what it returns, call for call, global stream included."""
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present")
    import qcqp_b200 as Q
    from qcqp_b200 import engine
    u, q = rh.load()
    monkeypatch.setattr(engine, "Pack", OraclePack)
    monkeypatch.setattr(q, "get_qcqp_form", lambda prob: rh.make_form(u, cvx.get_qcqp_form(prob).forms()))

    def problems():
        np.random.seed(1)
        A = np.random.randn(15, 10)
        b = np.random.randn(15, 1)
        x = cvx.Variable(10)
        yield x, cvx.Problem(cvx.Minimize(cvx.sum_squares(A*x - b)), [cvx.square(x) == 1]), 7
        W = pb.maxcut(25, 0.2, seed=1)[1]["W"]
        y = cvx.Variable(25)
        yield y, cvx.Problem(cvx.Maximize(0.25*(cvx.sum_entries(W) - cvx.quad_form(y, W))), [cvx.square(y) == 1]), 11

    for var, prob, seed in problems():
        ref = q.QCQP(prob)
        np.random.seed(seed)
        want = [ref.suggest(q.s.RANDOM), ref.improve(q.s.COORD_DESCENT), ref.improve(q.s.COORD_DESCENT, phase1=False)]
        x_want, state_want = np.array(var.value, dtype=float).copy(), np.random.get_state()
        var.value = None
        own = Q.QCQP(prob)
        np.random.seed(seed)
        got = [own.suggest(Q.RANDOM), own.improve(Q.COORD_DESCENT), own.improve(Q.COORD_DESCENT, phase1=False)]
        for (fw, vw), (fg, vg) in zip(want, got):
            assert abs(fg - fw) <= 1e-9 * abs(fw) and abs(vg - vw) <= 1e-6 * max(abs(vw), 1e-12)
        assert np.allclose(np.asarray(var.value), x_want.reshape(var.size), rtol=1e-9, atol=1e-9)
        state = np.random.get_state()
        assert state[2] == state_want[2] and np.array_equal(state[1], state_want[1])

    # golden G4's recipe through both classes: the beamforming script, a start point written into x.value, improve(ADMM, rho)
    np.random.seed(1)
    n, m, l = 20, 5, 2
    HR = np.random.randn(m, n); HI = np.random.randn(m, n)
    A = np.hstack((HR, HI)); B = np.hstack((-HI, HR))
    GR = np.random.randn(l, n); GI = np.random.randn(l, n)
    Cm = np.hstack((GR, GI)); D = np.hstack((-GI, GR))
    x = cvx.Variable(2*n)
    prob = cvx.Problem(cvx.Minimize(cvx.sum_squares(x)),
                       [cvx.square(A*x) + cvx.square(B*x) >= 20, cvx.square(Cm*x) + cvx.square(D*x) <= 2])
    np.random.seed(4)
    x0 = 2 * np.random.randn(2*n)
    ref, own = q.QCQP(prob), Q.QCQP(prob)
    x.value = x0.reshape(2*n, 1)
    fw, vw = ref.improve(q.s.ADMM, rho=np.sqrt(m + l))
    x_want = np.array(x.value, dtype=float).ravel()
    x.value = x0.reshape(2*n, 1)
    fg, vg = own.improve(Q.ADMM, rho=np.sqrt(m + l))
    g4 = golden["admm"][0]
    assert abs(fw - g4["f0"]) <= 1e-9 * abs(fw), "the reference run here is the one the golden file recorded"
    assert abs(fg - fw) <= 1e-6 * abs(fw) and abs(vg - vw) <= 1e-6
    assert np.allclose(np.asarray(x.value).ravel(), x_want, rtol=1e-6, atol=1e-8)


def test_readme_quadratic_expressions():
    """The reference README's list of quadratic expressions (README.md "Quadratic expressions"): (affine)*(affine) as a matrix
    product, power(affine, 2), square, sum_squares, quad_over_lin(affine, constant), matrix_frac(affine, constant),
    quad_form(affine, constant) -- values of the extracted forms against NumPy."""
    rs = np.random.RandomState(0)
    x = cvx.Variable(3); Y = cvx.Variable(2, 3); t = cvx.Variable()
    A = rs.randn(2, 3); c = rs.randn(3, 1); M = rs.randn(3, 3); Pd = M @ M.T + np.eye(3)
    xv = rs.randn(3, 1); Yv = rs.randn(2, 3); tv = float(rs.randn())
    cases = [
        ((x.T + c.T)*(M*x - c), float(((xv.T + c.T) @ (M @ xv - c))[0, 0])),
        ((Y + A)*(M*x + c), (Yv + A) @ (M @ xv + c)),
        (t*(M*x + c), tv * (M @ xv + c)),
        ((M*x + c)*(t - 2), (M @ xv + c) * (tv - 2)),
        (cvx.power(A*x - 1, 2), (A @ xv - 1) ** 2),
        (cvx.quad_over_lin(A*x - 1, 4.0), float(((A @ xv - 1) ** 2).sum() / 4.0)),
        (cvx.matrix_frac(x - c, Pd), float(((xv - c).T @ np.linalg.inv(Pd) @ (xv - c))[0, 0])),
        (cvx.quad_form(M*x, Pd) - 2*cvx.sum_squares(Y[1, :]), float(((M @ xv).T @ Pd @ (M @ xv))[0, 0] - 2 * (Yv[1] ** 2).sum())),
    ]
    x.value, Y.value, t.value = xv, Yv, tv
    for e, want in cases:
        assert e.is_quadratic() and not e.is_affine()
        assert np.allclose(e.value, want, rtol=1e-12, atol=1e-12)
    # and through get_qcqp_form: one scalar constraint per entry, column-major
    prob = cvx.Problem(cvx.Minimize(cases[0][0]), [cases[1][0] <= 1, cases[6][0] == 2])
    form = cvx.get_qcqp_form(prob)
    xs = prob.variables()
    v = cvx.flatten_vars(xs, form.n)
    vals = [v.dot(f.P.dot(v)) + f.qarray.dot(v) + f.r for f in form.fs]
    assert form.m == 3 and np.allclose(vals[:2], (cases[1][1] - 1).ravel(order="F")) and abs(vals[2] - (cases[6][1] - 2)) < 1e-12
    assert not cvx.square(cvx.sum_squares(x)).is_quadratic() and not (cvx.square(x[0])*x[1]).is_quadratic()
    with pytest.raises(Exception, match="not quadratic"):
        cvx.power(x, 3)
