"""Shared helpers for the parity tests."""
import numpy as np

from qcqp_b200 import problems as pb

GEN = {
    "bls": pb.boolean_least_squares,
    "maxcut": pb.maxcut,
    "beam": pb.beamforming,
    "circle": pb.circle_packing,
    "random": pb.random_qcqp,
}


def checksum(forms):
    acc = 0.0
    for j, (P, qv, r, _op) in enumerate(forms):
        P = P.tocsr()
        acc += (j + 1) * (float(np.abs(P.data).sum()) + float(np.abs(np.asarray(qv)).sum()) + abs(float(r)))
    return acc


def forms_of(case):
    forms, info = GEN[case["gen"]](**case["gargs"])
    if "checksum" in case:
        assert abs(checksum(forms) - case["checksum"]) <= 1e-12 * max(1.0, abs(case["checksum"])), "generator drifted from the golden file"
    return forms, info


def rel_close(a, b, rtol=1e-6, atol=1e-9):
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    return np.all(np.abs(a - b) <= atol + rtol * np.maximum(np.abs(a), np.abs(b)))


class OraclePack:
    """Stands in for engine.Pack in the CPU suite ONLY (host-logic tests of the facade): same methods, computed by the CPU
    oracle, which is test infrastructure.  The product never sees it."""

    def __init__(self, forms):
        from oracle import oracle as orc
        self._orc, self.P = orc, orc.Problem(forms)
        self.n = self.P.n
        self._mu = self._F = None

    def eval(self, X):
        X = np.asarray(X, dtype=np.float64).reshape(-1, self.n)
        return (np.array([self.P.eval(0, x) for x in X]), np.array([self.P.max_violation(x) for x in X]))

    def _cd_one(self, x0, st, num_iters, viol_tol, tol, phase1):
        return self.P.improve_cd(x0, st, num_iters=num_iters, viol_tol=viol_tol, tol=tol, phase1=phase1, fast=True)

    def cd_improve(self, X0, rng, num_iters=1000, viol_tol=1e-2, tol=1e-4, phase1=True, strict=False):
        import ctypes as C
        X0 = np.asarray(X0, dtype=np.float64).reshape(-1, self.n)
        X, stats = np.empty_like(X0), []
        for i in range(X0.shape[0]):
            st = self._orc.RngState()
            C.memmove(C.byref(st), C.byref(rng[i]), C.sizeof(st))
            X[i], s = self._cd_one(X0[i], st, num_iters, viol_tol, tol, phase1)
            C.memmove(C.byref(rng[i]), C.byref(st), C.sizeof(st))
            stats.append(s)
        f0, mv = self.eval(X)
        return X, f0, mv, stats

    def admm_improve(self, X0, rhos, num_iters=1000, viol_lim=1e4, tol=1e-2, phase1=True):
        return self.P.improve_admm_batch(X0, np.atleast_1d(rhos), num_iters=num_iters, viol_lim=viol_lim, tol=tol, phase1=phase1,
                                         nthreads=1)

    def sdr_sample_eval(self, mu, F, Z=None, S=None, seed=0):
        X, f0, mv = self.P.sdr_sample_eval(mu, F, np.ascontiguousarray(Z, dtype=np.float64))
        return X, f0, mv

    def sdr_cd_pipeline(self, seeds, mu=None, F=None, Z=None, S=None, seed=0, num_iters=1000, viol_tol=1e-2, tol=1e-4, phase1=True,
                        strict=False):
        from qcqp_b200.dist import local_best
        if mu is not None:
            self._mu, self._F = mu, F
        X0, _f, _v = self.sdr_sample_eval(self._mu, self._F, Z)
        X, stats = np.empty_like(X0), []
        for i, sd in enumerate(seeds):
            X[i], s = self._cd_one(X0[i], self._orc.RngState.from_seed(int(sd)), num_iters, viol_tol, tol, phase1)
            stats.append(s)
        f0, mv = self.eval(X)
        return dict(X=X, f0=f0, maxviol=mv, stats=stats, best=local_best(f0, mv)[2])
