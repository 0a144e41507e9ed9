"""ctypes front of oracle/qcqp_oracle.c -- the CPU checker for the CUDA engine.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by qcqp_b200/.  See the header of qcqp_oracle.c for parity status.

A problem is a list of forms ``[(P, q, r, relop), ...]`` with ``forms[0]`` the objective
(``relop=None``) and ``relop`` in ``{'<=', '=='}`` for the constraints -- the same tuple layout
``oracle/ref_harness.make_form`` feeds to the reference's own QuadraticFunction/QCQPForm.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libqcqp_oracle.so")

RELOP_CODE = {None: 0, "<=": 1, "==": 2}


def build(force=False):
    src = os.path.join(_HERE, "qcqp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libqcqp_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


class RngState(C.Structure):
    """np.random.RandomState state, field for field (key[624], pos, has_gauss, cached_gaussian)."""
    _fields_ = [("key", C.c_uint32 * 624), ("pos", C.c_int32), ("has_gauss", C.c_int32), ("gauss", C.c_double)]

    @classmethod
    def from_numpy(cls, rs):
        _, key, pos, has_gauss, gauss = rs.get_state()
        st = cls()
        C.memmove(st.key, np.ascontiguousarray(key, dtype=np.uint32).ctypes.data, 624 * 4)
        st.pos, st.has_gauss, st.gauss = int(pos), int(has_gauss), float(gauss)
        return st

    @classmethod
    def from_seed(cls, seed):
        return cls.from_numpy(np.random.RandomState(seed))

    def to_numpy_state(self):
        return ("MT19937", np.frombuffer(self.key, dtype=np.uint32).copy(), int(self.pos), int(self.has_gauss), float(self.gauss))


class CdParams(C.Structure):
    _fields_ = [("num_iters", C.c_int32), ("viol_tol", C.c_double), ("tol", C.c_double), ("phase1", C.c_int32), ("fast", C.c_int32),
                ("refresh_every", C.c_int32)]


class CdStats(C.Structure):
    _fields_ = [("steps_p1", C.c_int64), ("steps_p2", C.c_int64), ("updates_p1", C.c_int64), ("updates_p2", C.c_int64),
                ("sweeps_p1", C.c_int32), ("sweeps_p2", C.c_int32), ("status", C.c_int32), ("ran_phase2", C.c_int32),
                ("steps_skipped", C.c_int64)]


class AdmmParams(C.Structure):
    _fields_ = [("num_iters", C.c_int32), ("viol_lim", C.c_double), ("tol", C.c_double), ("rho", C.c_double), ("phase1", C.c_int32)]


class AdmmStats(C.Structure):
    _fields_ = [("iters_p1", C.c_int32), ("iters_p2", C.c_int32), ("onecons_calls", C.c_int64), ("status", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_problem_create.restype = C.c_void_p
        L.orc_problem_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_problem_set_eig.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_problem_destroy.argtypes = [C.c_void_p]
        L.orc_form_eval.restype = C.c_double
        L.orc_form_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_max_violation.restype = C.c_double
        L.orc_max_violation.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_onevar_func.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_get_feasible_intervals.restype = C.c_int
        L.orc_get_feasible_intervals.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.c_void_p]
        L.orc_onevar_qcqp.restype = C.c_int
        L.orc_onevar_qcqp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_improve_cd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_cd_phase.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_improve_cd_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_eval_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_better.restype = C.c_int
        L.orc_better.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.orc_onecons.restype = C.c_int
        L.orc_onecons.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_void_p]
        L.orc_improve_admm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_improve_admm_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_sdr_sample.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sdr_sample_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_rng_uniform.restype = C.c_double
        L.orc_rng_uniform.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_rng_choice.restype = C.c_int64
        L.orc_rng_choice.argtypes = [C.c_void_p, C.c_int64]
        L.orc_rng_gauss.restype = C.c_double
        L.orc_rng_gauss.argtypes = [C.c_void_p]
        L.orc_rng_seed.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def stack_forms(forms):
    """[(P, q, r, relop)] -> stacked CSR arrays (form-major rows, sorted columns, explicit zeros dropped)."""
    n = len(np.asarray(forms[0][1]).ravel())
    mats = []
    for (P, _q, _r, _op) in forms:
        M = sp.csr_matrix(P, dtype=np.float64)
        M.sum_duplicates()
        M.eliminate_zeros()
        M.sort_indices()
        assert M.shape == (n, n)
        mats.append(M)
    S = sp.vstack(mats, format="csr")
    S.sort_indices()
    return dict(
        n=n, m=len(forms) - 1,
        indptr=np.ascontiguousarray(S.indptr, dtype=np.int64),
        indices=np.ascontiguousarray(S.indices, dtype=np.int32),
        data=np.ascontiguousarray(S.data, dtype=np.float64),
        q=np.ascontiguousarray(np.stack([np.asarray(f[1], dtype=np.float64).ravel() for f in forms])),
        r=np.ascontiguousarray([float(f[2]) for f in forms], dtype=np.float64),
        relop=np.ascontiguousarray([RELOP_CODE[f[3]] for f in forms], dtype=np.int32),
    )


class Problem:
    """QCQPForm (utilities.py:122-146) held by the C oracle."""

    def __init__(self, forms):
        self.forms = forms
        self.a = stack_forms(forms)
        self.n, self.m = self.a["n"], self.a["m"]
        a = self.a
        self.h = lib().orc_problem_create(self.n, self.m, _ptr(a["indptr"]), _ptr(a["indices"]), _ptr(a["data"]),
                                          _ptr(a["q"]), _ptr(a["r"]), _ptr(a["relop"]))
        self._eig = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_problem_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- evaluation -------------------------------------------------------------------------
    def eval(self, j, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return lib().orc_form_eval(self.h, j, _ptr(x))

    def max_violation(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return lib().orc_max_violation(self.h, _ptr(x))

    def eval_batch(self, X, want_viol=False, nthreads=1):
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, self.n)
        R = X.shape[0]
        f0 = np.empty(R); mv = np.empty(R)
        viol = np.empty((R, self.m)) if want_viol else None
        lib().orc_eval_batch(self.h, R, _ptr(X), _ptr(f0), _ptr(mv), _ptr(viol) if want_viol else None, nthreads)
        return (f0, mv, viol) if want_viol else (f0, mv)

    def better(self, x1, x2, tol=1e-4):
        x1 = np.ascontiguousarray(x1, dtype=np.float64); x2 = np.ascontiguousarray(x2, dtype=np.float64)
        return x1 if lib().orc_better(self.h, _ptr(x1), _ptr(x2), tol) == 1 else x2

    def get_onevar_func(self, j, x, k):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(3)
        lib().orc_get_onevar_func(self.h, j, _ptr(x), k, _ptr(out))
        return tuple(out)

    # -- coordinate descent -----------------------------------------------------------------
    def improve_cd(self, x0, rng, num_iters=1000, viol_tol=1e-2, tol=1e-4, phase1=True, fast=False, only_phase=None, refresh_every=0):
        """improve_coord_descent (qcqp.py:181-192). rng: RngState (advanced in place). Returns (x, CdStats)."""
        x = np.array(x0, dtype=np.float64, copy=True)
        prm = CdParams(num_iters, viol_tol, tol, int(bool(phase1)), int(bool(fast)), int(refresh_every))
        st = CdStats()
        if only_phase is None:
            lib().orc_improve_cd(self.h, C.byref(prm), _ptr(x), C.byref(rng), C.byref(st))
        else:
            lib().orc_cd_phase(self.h, C.byref(prm), int(only_phase), _ptr(x), C.byref(rng), C.byref(st))
        return x, st

    def improve_cd_batch(self, X0, rngs, num_iters=1000, viol_tol=1e-2, tol=1e-4, phase1=True, fast=False, nthreads=0, refresh_every=0):
        X = np.array(X0, dtype=np.float64, copy=True).reshape(-1, self.n)
        R = X.shape[0]
        prm = CdParams(num_iters, viol_tol, tol, int(bool(phase1)), int(bool(fast)), int(refresh_every))
        st = (CdStats * R)()
        f0 = np.empty(R); mv = np.empty(R)
        lib().orc_improve_cd_batch(self.h, C.byref(prm), R, _ptr(X), C.byref(rngs), _ptr(f0), _ptr(mv), C.byref(st), nthreads)
        return X, f0, mv, st

    # -- ADMM -------------------------------------------------------------------------------
    def set_eig(self, lmb, Q, qhat):
        self._eig = (np.ascontiguousarray(lmb, dtype=np.float64), np.ascontiguousarray(Q, dtype=np.float64),
                     np.ascontiguousarray(qhat, dtype=np.float64))
        lib().orc_problem_set_eig(self.h, _ptr(self._eig[0]), _ptr(self._eig[1]), _ptr(self._eig[2]))

    def compute_eig(self):
        """The host-side setup of utilities.py:160-166, with the reference's own NumPy calls."""
        n, m = self.n, self.m
        lmb = np.empty((m, n)); Q = np.empty((m, n, n)); qhat = np.empty((m, n))
        for i in range(m):
            P = sp.csr_matrix(self.forms[i + 1][0])
            Psymm = (P + P.T) / 2.
            lmb[i], Q[i] = np.linalg.eigh(np.asarray(Psymm.todense()))
            qhat[i] = Q[i].T.dot(np.asarray(self.forms[i + 1][1], dtype=np.float64).ravel())
        self.set_eig(lmb, Q, qhat)
        return lmb, Q, qhat

    def onecons(self, j, z, tol=1e-6):
        z = np.ascontiguousarray(z, dtype=np.float64)
        out = np.empty(self.n)
        it = lib().orc_onecons(self.h, j, _ptr(z), tol, _ptr(out))
        return out, it

    def zinv(self, rho):
        """inverse of 2(P0 + rho m I) -- the matrix qcqp.py:226-227 hands to SuperLU."""
        P0 = np.asarray(sp.csr_matrix(self.forms[0][0]).todense())
        return np.ascontiguousarray(np.linalg.inv(2 * (P0 + rho * self.m * np.eye(self.n))))

    def improve_admm(self, x0, rho, num_iters=1000, viol_lim=1e4, tol=1e-2, phase1=True):
        if self._eig is None:
            self.compute_eig()
        x = np.array(x0, dtype=np.float64, copy=True)
        Zinv = self.zinv(rho)
        prm = AdmmParams(num_iters, viol_lim, tol, rho, int(bool(phase1)))
        st = AdmmStats()
        lib().orc_improve_admm(self.h, C.byref(prm), _ptr(Zinv), _ptr(x), C.byref(st))
        return x, st

    def improve_admm_batch(self, X0, rhos, num_iters=1000, viol_lim=1e4, tol=1e-2, phase1=True, nthreads=0):
        if self._eig is None:
            self.compute_eig()
        X0 = np.ascontiguousarray(X0, dtype=np.float64).reshape(-1, self.n)
        R = X0.shape[0]
        rhos = np.ascontiguousarray(rhos, dtype=np.float64)
        K = len(rhos)
        Zinv = np.ascontiguousarray(np.stack([self.zinv(r) for r in rhos]))
        X = np.empty((K, R, self.n)); f0 = np.empty((K, R)); mv = np.empty((K, R))
        st = (AdmmStats * (K * R))()
        prm = AdmmParams(num_iters, viol_lim, tol, 0.0, int(bool(phase1)))
        lib().orc_improve_admm_batch(self.h, C.byref(prm), K, _ptr(rhos), _ptr(Zinv), R, _ptr(X0), _ptr(X), _ptr(f0), _ptr(mv),
                                     C.byref(st), nthreads)
        return X, f0, mv, st

    # -- SDR sampler ------------------------------------------------------------------------
    def sdr_sample_eval(self, mu, F, Z, nthreads=1):
        mu = np.ascontiguousarray(mu, dtype=np.float64); F = np.ascontiguousarray(F, dtype=np.float64)
        Z = np.ascontiguousarray(Z, dtype=np.float64).reshape(-1, self.n)
        S = Z.shape[0]
        X = np.empty((S, self.n)); f0 = np.empty(S); mv = np.empty(S)
        lib().orc_sdr_sample_eval(self.h, _ptr(mu), _ptr(F), _ptr(Z), S, _ptr(X), _ptr(f0), _ptr(mv), nthreads)
        return X, f0, mv


def sdr_factor(Xstar, eps=1e-8, corrected=False):
    """mu, Sigma of qcqp.py:394-395 and the factor F NumPy's multivariate_normal builds from Sigma
    (F = sqrt(s)[:,None] * Vt, SURVEY a-7), so that a draw is mu + standard_normal(n) @ F."""
    Xs = np.asarray(Xstar, dtype=np.float64)
    n = Xs.shape[0] - 1
    mu = Xs[:-1, -1].copy()
    if corrected:
        Sigma = Xs[:-1, :-1] - np.outer(mu, mu) + eps * np.eye(n)
    else:
        Sigma = Xs[:-1, :-1] - mu * mu.T + eps * np.eye(n)  # the reference's row-broadcast
    _u, s, vt = np.linalg.svd(Sigma)
    F = np.sqrt(s)[:, None] * vt
    return mu, Sigma, np.ascontiguousarray(F)


def sdr_sample(mu, F, rng):
    n = len(mu)
    z = np.empty(n); x = np.empty(n)
    mu = np.ascontiguousarray(mu, dtype=np.float64); F = np.ascontiguousarray(F, dtype=np.float64)
    lib().orc_sdr_sample(n, _ptr(mu), _ptr(F), C.byref(rng), _ptr(z), _ptr(x))
    return x, z


def onevar_qcqp(f0, fs, s, rng=None):
    """f0 = (p, q, r); fs = [(p, q, r, relop)]. Returns x or None (utilities.py:241-288)."""
    rng = rng if rng is not None else RngState.from_seed(0)
    f0a = np.ascontiguousarray(f0, dtype=np.float64)
    fa = np.ascontiguousarray([f[:3] for f in fs], dtype=np.float64).reshape(-1, 3)
    ra = np.ascontiguousarray([RELOP_CODE[f[3]] for f in fs], dtype=np.int32)
    out = C.c_double()
    rc = lib().orc_onevar_qcqp(_ptr(f0a), _ptr(fa), _ptr(ra), len(fs), float(s), C.byref(rng), C.byref(out))
    if rc < 0:
        raise OverflowError("Range exceeds valid bounds")
    return out.value if rc == 1 else None


def get_feasible_intervals(f, s=0.0):
    out = np.empty(8)
    c = lib().orc_get_feasible_intervals(float(f[0]), float(f[1]), float(f[2]), RELOP_CODE[f[3]], float(s), _ptr(out))
    return [(out[2 * i], out[2 * i + 1]) for i in range(c)]
