"""Host containers for the quadratic data of a QCQP -- the cvxpy-free counterpart of the reference's
QuadraticFunction / QCQPForm (utilities.py:41-46, 122-131).  They only HOLD data; every evaluation goes through the
engine (qcqp_b200.engine.Pack)."""
import numpy as np
import scipy.sparse as sp


class QuadraticFunction:
    """x'Px + q'x + r with an optional relation '<=' or '==' (utilities.py:41-46). P is symmetrised on entry, as
    get_qcqp_form does (utilities.py:333,345)."""

    def __init__(self, P, q, r, relop=None):
        P = sp.csr_matrix(P, dtype=np.float64)
        self.P = sp.csr_matrix((P + P.T) / 2.0)
        qa = np.asarray(q.todense() if sp.issparse(q) else q, dtype=np.float64).ravel()
        if self.P.shape[0] != self.P.shape[1] or self.P.shape[0] != qa.size:
            raise Exception("P must be n x n and q of length n")
        if relop not in (None, "<=", "=="):
            raise Exception("relop must be None, '<=' or '=='")
        self.qarray = qa
        self.r = float(r)
        self.relop = relop

    def as_tuple(self):
        return (self.P, self.qarray, self.r, self.relop)


class QCQPForm:
    """minimize f0 subject to fs (utilities.py:122-131)."""

    def __init__(self, f0, fs):
        if f0.relop is not None:
            raise Exception("the objective carries no relation")
        if not all(f.relop is not None for f in fs):
            raise Exception("every constraint needs a relation")
        self.f0, self.fs = f0, list(fs)
        self.n = f0.P.shape[0]
        self.m = len(self.fs)

    def fi(self, i):
        return self.fs[i]

    def forms(self):
        return [self.f0.as_tuple()] + [f.as_tuple() for f in self.fs]

    @classmethod
    def from_tuples(cls, forms):
        return cls(QuadraticFunction(*forms[0]), [QuadraticFunction(*f) for f in forms[1:]])
