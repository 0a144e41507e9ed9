"""The reference's call sequence over the cvxpy-free modelling layer (qcqp_b200/model.py) on the GPU engine:
QCQP(prob); suggest(RANDOM); improve(COORD_DESCENT) is exactly the recipe of the goldens G1 / G3 minted from the reference
(tests/golden/make_golden.py), and G2' starts from a point the user writes into x.value.  Bar: 1e-6 relative on
(objective, max violation), the process-global np.random stream left where the reference leaves it."""
import numpy as np
import pytest

from helpers import rel_close

pytestmark = pytest.mark.gpu


def _case(golden, name):
    return [c for c in golden["cd"] if c["name"] == name][0]


def test_boolean_least_squares_script_g1_g2p(golden):
    import qcqp_b200 as Q
    import qcqp_b200.model as cvx
    np.random.seed(1)
    A = np.random.randn(15, 10)
    b = np.random.randn(15, 1)
    x = cvx.Variable(10)
    qc = Q.QCQP(cvx.Problem(cvx.Minimize(cvx.sum_squares(A*x - b)), [cvx.square(x) == 1]))
    g = _case(golden, "G1")
    np.random.seed(g["seed"])
    qc.suggest(Q.RANDOM)
    assert np.array_equal(x.value.ravel(), np.array(g["x0"]))
    f, v = qc.improve(Q.COORD_DESCENT)
    assert rel_close(f, g["f0"], rtol=1e-6) and rel_close(v, g["maxviol"], rtol=1e-6, atol=1e-10)
    assert rel_close(x.value.ravel(), g["x"], rtol=1e-6, atol=1e-9)
    assert np.random.get_state()[2] == g["rng"]["pos"]
    assert rel_close(cvx.sum_squares(A*x - b).value, f, rtol=1e-9)          # the expression at the written-back value

    g = _case(golden, "G2p")
    x.value = np.array(g["x0"]).reshape(10, 1)
    f, v = qc.improve(Q.COORD_DESCENT, phase1=False)
    assert rel_close(f, g["f0"], rtol=1e-6) and rel_close(v, g["maxviol"], rtol=1e-6, atol=1e-10)
    assert rel_close(x.value.ravel(), g["x"], rtol=1e-6, atol=1e-9)


def test_maxcut_script_g3(golden):
    import qcqp_b200 as Q
    import qcqp_b200.model as cvx
    from qcqp_b200 import problems as pb
    g = _case(golden, "G3")
    _forms, info = pb.maxcut(25, 0.2, seed=1)
    W = info["W"]
    x = cvx.Variable(25)
    qc = Q.QCQP(cvx.Problem(cvx.Maximize(0.25*(cvx.sum_entries(W) - cvx.quad_form(x, W))), [cvx.square(x) == 1]))
    np.random.seed(g["seed"])
    qc.suggest(Q.RANDOM)
    f, v = qc.improve(Q.COORD_DESCENT)
    assert rel_close(f, -g["f0"], rtol=1e-6) and rel_close(v, g["maxviol"], rtol=1e-6, atol=1e-10)
    assert np.random.get_state()[2] == g["rng"]["pos"]


def test_circle_packing_script_batch():
    """Two variables (scalar r, matrix X): batch of random starts, best point written back column-major."""
    import qcqp_b200 as Q
    import qcqp_b200.model as cvx
    n = 4
    X = cvx.Variable(2, n)
    r = cvx.Variable()
    cons = [X >= r, X <= 10 - r, r >= 0]
    for i in range(n):
        for j in range(i + 1, n):
            cons.append(cvx.square(2*r) <= cvx.sum_squares(X[:, i] - X[:, j]))
    qc = Q.QCQP(cvx.Problem(cvx.Maximize(r), cons))
    np.random.seed(5)
    qc.suggest(Q.RANDOM, samples=32)
    f, v = qc.improve(Q.COORD_DESCENT, seed=9, num_iters=20)
    assert isinstance(r.value, float) and X.value.shape == (2, n)
    assert rel_close(f, r.value, rtol=1e-12, atol=0) and np.array_equal(np.concatenate([[r.value], X.value.ravel(order="F")]), qc.x)
    worst = max(float(np.max(c.violation)) for c in cons)
    assert rel_close(worst, v, rtol=1e-9, atol=1e-10)
