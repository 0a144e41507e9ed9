"""Imports the UNMODIFIED reference package (read-only at /root/reference) behind a stub
``cvxpy`` so that its NumPy/SciPy hot path can be executed in the build container.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` to mint the golden vectors that
pin ``oracle/qcqp_oracle.c``.  /root/reference does not exist on the GPU box, so nothing that runs
there (``-m gpu`` tests, ``smoke()``, ``bench.py``) imports this module.

The stub is needed because ``qcqp/utilities.py:28-29`` and ``qcqp/qcqp.py:28,30`` import cvxpy
(0.4.x, not installable here); only the out-of-scope front-end (``get_qcqp_form``, ``solve_sdr``,
``solve_spectral``, ``improve_dccp``) ever touches it.
"""
import os
import sys
import tempfile
import types

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "qcqp"))


def load():
    """Returns (utilities_module, qcqp_module) of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "qcqp.qcqp" in sys.modules and getattr(sys.modules["qcqp"], "__file__", "").startswith(REFERENCE_ROOT):
        return sys.modules["qcqp.utilities"], sys.modules["qcqp.qcqp"]
    for name in ("cvxpy", "cvxpy.utilities", "cvxpy.lin_ops", "cvxpy.lin_ops.lin_utils"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = []
            sys.modules[name] = mod
    sys.modules["cvxpy.utilities"].QuadCoeffExtractor = object
    sys.modules["cvxpy"].utilities = sys.modules["cvxpy.utilities"]
    sys.modules["cvxpy"].lin_ops = sys.modules["cvxpy.lin_ops"]
    sys.modules["cvxpy.lin_ops"].lin_utils = sys.modules["cvxpy.lin_ops.lin_utils"]
    # the reference opens ./qcqp.log for writing at import time (qcqp.py:39)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="qcqp_ref_")
    os.chdir(tmp)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import qcqp.utilities as u
        import qcqp.qcqp as q
    finally:
        sys.path.remove(REFERENCE_ROOT)
        os.chdir(cwd)
    return u, q


def make_form(u, forms):
    """forms: list of (P dense/sparse n*n, q[n], r, relop) with forms[0] the objective (relop None).

    Builds the reference's own containers exactly as get_qcqp_form would hand them to the hot path
    (utilities.py:331-345): scipy *matrix* CSR for P, an n*1 CSC column for q."""
    import numpy as np
    import scipy.sparse as sp
    fs = []
    for (P, q, r, relop) in forms:
        n = len(q)
        fs.append(u.QuadraticFunction(sp.csr_matrix(P), sp.csc_matrix(np.asarray(q, dtype=float).reshape(n, 1)),
                                      float(r), relop))
    return u.QCQPForm(fs[0], fs[1:])
