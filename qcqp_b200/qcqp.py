"""The CVXPY-facing Suggest-and-Improve facade: same class, methods and return values as the reference's QCQP
(qcqp/qcqp.py:367-432), with the hot path running on the B200 engine.

    qcqp = QCQP(prob)                 # a qcqp_b200.model Problem (the cvxpy surface of the reference's examples) / a QCQPForm /
                                      # a list of (P, q, r, relop) / (if cvxpy 0.4 is importable) a cvxpy Problem
    f, v = qcqp.suggest(SDR)          # (objective, max violation) of the suggested point
    f, v = qcqp.improve(COORD_DESCENT)

Extensions (all optional): `samples=` / `restarts=` run many draws / restarts as one batch on the GPU and keep the best
in QCQPForm.better order; `seed=` gives restart r the MT19937 stream of np.random.seed(seed + r).  With a single point
and no seed the process-global np.random stream is consumed exactly as the reference consumes it.
"""
import logging
import sys

import numpy as np
import scipy.sparse as sp

from . import settings as s
from . import engine
from . import relax
from . import model
from .forms import QCQPForm, QuadraticFunction

from . import dist as _dist

log = logging.getLogger("qcqp_b200")   # the reference opens ./qcqp.log at import time (qcqp.py:39); this package does not


def _form_from_cvxpy(prob):
    """get_qcqp_form (utilities.py:318-347) for a cvxpy 0.4 Problem.  Only reachable when cvxpy 0.4 is installed."""
    try:
        from cvxpy.utilities import QuadCoeffExtractor
    except Exception:
        raise Exception("a cvxpy problem needs cvxpy 0.4 (QuadCoeffExtractor); pass a QCQPForm or a list of (P, q, r, relop) instead")
    if not prob.objective.args[0].is_quadratic():
        raise Exception("Objective is not quadratic.")
    if not all([constr._expr.is_quadratic() for constr in prob.constraints]):
        raise Exception("Not all constraints are quadratic.")
    id_map, N = {}, 0
    for x in prob.variables():
        id_map[x.id] = N
        N += x.size[0] * x.size[1]
    extractor = QuadCoeffExtractor(id_map, N)
    P0, q0, r0 = extractor.get_coeffs(prob.objective.args[0])
    P0, q0, r0 = (P0[0] + P0[0].T) / 2., q0.T.tocsc(), r0[0]
    maximize = prob.objective.NAME == "maximize"
    if maximize:
        P0, q0, r0 = -P0, -q0, -r0
    fs = []
    for constr in prob.constraints:
        sz = constr._expr.size[0] * constr._expr.size[1]
        Pc, qc, rc = extractor.get_coeffs(constr._expr)
        for i in range(sz):
            fs.append(QuadraticFunction((Pc[i] + Pc[i].T) / 2., qc[i, :].T.tocsc(), rc[i], constr.OP_NAME))
    return QCQPForm(QuadraticFunction(P0, q0, r0), fs), maximize


def _process_group():
    """torch.distributed when a process group of more than one rank is up (one process per GPU under torchrun), else None.
    torch is only looked at if the caller imported it: the facade itself never needs it."""
    torch = sys.modules.get("torch")
    if torch is None:
        return None
    d = torch.distributed
    if d.is_available() and d.is_initialized() and d.get_world_size() > 1:
        return d
    return None


class QCQP:
    def __init__(self, prob, maximize=False):
        self.prob = prob
        self._cvx_vars = None
        if isinstance(prob, QCQPForm):
            self.qcqp_form = prob
        elif isinstance(prob, (list, tuple)):
            self.qcqp_form = QCQPForm.from_tuples(list(prob))
        elif isinstance(prob, model.Problem):
            # the cvxpy-free modelling layer: same extraction and variable order as get_qcqp_form (utilities.py:318-347)
            self.qcqp_form = model.get_qcqp_form(prob)
            maximize = prob.objective.NAME == "maximize"
            self._cvx_vars = prob.variables()
        else:
            self.qcqp_form, maximize = _form_from_cvxpy(prob)
            self._cvx_vars = prob.variables()
        self.n = self.qcqp_form.n
        self.maximize_flag = bool(maximize)
        self.spectral_sol = None
        self.spectral_bound = None
        self.sdr_sol = None
        self.sdr_bound = None
        self.mu = None
        self.Sigma = None
        self._F = None
        self.x = None          # current point (the reference keeps it in the cvxpy variables)
        self.X = None          # current batch [R][n] (under torchrun: this rank's shard of it)
        self._shard = None     # (lo, hi, total) of this rank's restarts when a batch is sharded over ranks
        self._pack = engine.Pack(self.qcqp_form.forms())

    # ---- helpers -----------------------------------------------------------------------------------------
    def _sign(self, f):
        return -f if self.maximize_flag else f

    _STATUS_ERRORS = {1: (ValueError, "max() arg is an empty sequence"),          # qcqp.py:117
                      2: (OverflowError, "Range exceeds valid bounds")}           # utilities.py:267

    def _mask_failed(self, f0, maxviol, stats):
        """A restart whose kernel status mirrors a Python exception of the reference.  A single point in a single process raises
        at once, as the reference does.  In a batch (or under torchrun) a failed restart must neither throw the other results
        away nor leave the other ranks waiting in the best-pick collective: it is marked (+inf, +inf) so that it cannot win, and
        _assign raises -- on every rank, after the collective -- only if every restart of the whole batch failed."""
        bad = [r for r in range(len(stats)) if stats[r].status != 0]
        self._fail_code = stats[bad[0]].status if bad else 0
        self._fail_count = len(bad)
        if bad and len(stats) == 1 and self._shard is None:
            exc, msg = self._STATUS_ERRORS.get(self._fail_code, (Exception, "restart failed with status %d" % self._fail_code))
            raise exc(msg)
        if bad:
            f0 = np.array(f0, dtype=np.float64, copy=True); maxviol = np.array(maxviol, dtype=np.float64, copy=True)
            f0[bad] = np.inf; maxviol[bad] = np.inf
        return f0, maxviol

    def _assign(self, X, f0, maxviol):
        """Keeps the batch, selects the best point (QCQPForm.better order) and writes it back (assign_vars,
        utilities.py:298-308, column-major per variable)."""
        self.X = np.array(X, dtype=np.float64).reshape(-1, self.n)
        self.batch_f0 = np.asarray([self._sign(v) for v in f0])
        self.batch_maxviol = np.array(maxviol, dtype=np.float64)
        fail_count, fail_code = getattr(self, "_fail_count", 0), getattr(self, "_fail_code", 0)
        self._fail_count = self._fail_code = 0
        f0a = np.asarray(f0, dtype=np.float64); mva = np.asarray(maxviol, dtype=np.float64)
        usable = np.isfinite(f0a) & ~np.isnan(mva) if len(f0a) else np.zeros(0, dtype=bool)
        if usable.any():
            fsel = np.where(usable, f0a, np.inf); vsel = np.where(usable, mva, np.inf)
            b = engine.best(fsel, vsel) if len(f0a) > 1 else 0
            fb, vb = float(f0a[b]), float(mva[b])
        else:
            b, fb, vb = -1, np.inf, np.inf
        self.best_index = b
        self.x = self.X[b].copy() if b >= 0 else np.zeros(self.n)
        total, failed = len(f0a), fail_count
        if self._shard is not None:
            # the only collective -- ONE all-gather carrying each rank's best (bucket, f0, index), its point, and its
            # counts of restarts / failed restarts; every rank then takes the same pick in the `better` order
            lo, hi, _total = self._shard
            per = len(f0a) // (hi - lo) if hi > lo else 1          # rows per restart: the number of rho values after an ADMM sweep
            bucket = int(vb / 1e-4) if (b >= 0 and np.isfinite(vb)) else np.iinfo(np.int64).max
            mine = lo * per + b if b >= 0 else -1
            d = _process_group()
            payload = np.concatenate([self.x, [fb, vb, float(d.get_rank()), float(total), float(failed), float(fail_code)]])
            _gb, _gf, gi, w = _dist.global_best(bucket, fb, mine, x=payload)
            counts = _dist.last_gather_columns(self.n + 3, 3)        # (restarts, failed, first failure code) of every rank
            total, failed = int(counts[:, 0].sum()), int(counts[:, 1].sum())
            codes = counts[:, 2][counts[:, 2] > 0]
            fail_code = int(codes[0]) if len(codes) else 0
            if w is not None:
                self.x, fb, vb = w[:self.n].copy(), float(w[self.n]), float(w[self.n + 1])
                self.best_rank = int(w[self.n + 2])
            self.best_index = gi
        if total > 0 and failed >= total:
            exc, msg = self._STATUS_ERRORS.get(fail_code, (Exception, "every restart failed (status %d)" % fail_code))
            raise exc(msg)
        if self._cvx_vars is not None:
            model.assign_vars(self._cvx_vars, self.x)
        return (self._sign(fb), vb)

    def _shard_of(self, total):
        """This rank's contiguous slice of a batch of `total` restarts / draws (all of it without a process group)."""
        d = _process_group()
        if d is None or total <= 1:
            self._shard = None
            return 0, total
        lo, hi = _dist.shard_range(total, d.get_rank(), d.get_world_size())
        self._shard = (lo, hi, total)
        return lo, hi

    def set_sdr_solution(self, X, bound=None):
        """Supplies the relaxed solution X* (n+1 x n+1) of solve_sdr (qcqp.py:72-97) computed elsewhere."""
        X = np.asarray(X, dtype=np.float64)
        if X.shape != (self.n + 1, self.n + 1):
            raise Exception("X* must be (n+1) x (n+1)")
        self.sdr_sol = X
        self.sdr_bound = None if bound is None else (-bound if self.maximize_flag else bound)
        self.mu = None
        self._factor_on_device = False

    # ---- suggest (qcqp.py:378-401) -----------------------------------------------------------------------
    def suggest(self, method=s.RANDOM, eps=1e-8, *args, **kwargs):
        if method not in s.suggest_methods:
            raise Exception("Unknown suggest method: %s\n", method)
        S = int(kwargs.pop("samples", 1))
        # Under torchrun a batch is sharded: every rank draws the same S points from the (identically seeded) global stream
        # and keeps its contiguous slice, so the result does not depend on the number of GPUs.
        if method == s.RANDOM:
            X = np.stack([np.random.randn(self.n) for _ in range(S)])
            lo, hi = self._shard_of(S)
            X = X[lo:hi]
            f0, mv = self._pack.eval(X) if hi > lo else (np.zeros(0), np.zeros(0))
            return self._assign(X, f0, mv)
        if method == s.SPECTRAL:
            if self.spectral_sol is None:
                # host SDP, as in the reference (qcqp.py:384-387); solver kwargs (iters, tol) are forwarded
                self.spectral_sol, self.spectral_bound = relax.solve_spectral(self.qcqp_form, *args, **kwargs)
                if self.maximize_flag:
                    self.spectral_bound *= -1
            X = np.asarray(self.spectral_sol, dtype=np.float64).reshape(1, self.n)
            self._shard = None
            f0, mv = self._pack.eval(X)
            return self._assign(X, f0, mv)
        # SDR
        corrected = bool(kwargs.pop("corrected", False))
        device_rng = kwargs.pop("device_rng", False)
        seed = int(kwargs.pop("seed", 0))
        if self.sdr_sol is None:
            if "sdr_solution" in kwargs:
                self.set_sdr_solution(kwargs.pop("sdr_solution"), kwargs.pop("sdr_bound", None))
            else:
                # host SDP, as in the reference (qcqp.py:390-393): qcqp_b200.relax (NumPy) since cvxpy/SCS are not required
                self.sdr_sol, self.sdr_bound = relax.solve_sdr(self.qcqp_form, *args, **kwargs)
                if self.maximize_flag:
                    self.sdr_bound *= -1
        if self.mu is None:
            self.mu, self.Sigma, self._F = engine.sdr_factor(self.sdr_sol, eps=eps, corrected=corrected)
        lo, hi = self._shard_of(S)
        if hi == lo:
            X, f0, mv = np.zeros((0, self.n)), np.zeros(0), np.zeros(0)
            if not device_rng:
                [np.random.standard_normal(self.n) for _ in range(S)]      # keep the global stream in step with the other ranks
        elif device_rng:
            # one Philox stream per rank (seed + rank): unlike the host-stream mode the draws depend on the rank count
            rank = 0 if self._shard is None else _process_group().get_rank()
            X, f0, mv = self._pack.sdr_sample_eval(self.mu, self._F, Z=None, S=hi - lo, seed=seed + rank)
        else:
            # np.random.multivariate_normal draws standard_normal(n) per sample from the global stream (SURVEY a-7)
            Z = np.stack([np.random.standard_normal(self.n) for _ in range(S)])
            X, f0, mv = self._pack.sdr_sample_eval(self.mu, self._F, Z=np.ascontiguousarray(Z[lo:hi]))
        return self._assign(X, f0, mv)

    def suggest_improve(self, samples=1, seed=0, eps=1e-8, device_rng=False, corrected=False, **kwargs):
        """Batch form of the README loop `qcqp.suggest(SDR); qcqp.improve(COORD_DESCENT)` for `samples` draws in ONE engine call
        (qcqp_sdr_cd_pipeline): the draws stay on the device, restart s consumes the stream of np.random.seed(seed + s), the
        best point in the `better` order is written back.  kwargs: improve_coord_descent's (num_iters, viol_tol, tol, phase1)."""
        S = int(samples)
        if self.sdr_sol is None:
            self.sdr_sol, self.sdr_bound = relax.solve_sdr(self.qcqp_form)
            if self.maximize_flag:
                self.sdr_bound *= -1
        fresh = self.mu is None
        if fresh:
            self.mu, self.Sigma, self._F = engine.sdr_factor(self.sdr_sol, eps=eps, corrected=corrected)
        fresh = fresh or not getattr(self, "_factor_on_device", False)
        Z = None if device_rng else np.stack([np.random.standard_normal(self.n) for _ in range(S)])
        lo, hi = self._shard_of(S)          # under torchrun: this rank's draws, each with the stream of np.random.seed(seed + s)
        seeds = [(int(seed) + r) % (2 ** 32) for r in range(lo, hi)]
        kw = dict(num_iters=kwargs.get('num_iters', 1000), viol_tol=kwargs.get('viol_tol', 1e-2), tol=kwargs.get('tol', 1e-4),
                  phase1=kwargs.get('phase1', True), strict=kwargs.get('strict', False))
        if hi == lo:
            self.cd_stats = []
            return self._assign(np.zeros((0, self.n)), np.zeros(0), np.zeros(0))
        rank = 0 if self._shard is None else _process_group().get_rank()
        res = self._pack.sdr_cd_pipeline(seeds, mu=self.mu if fresh else None, F=self._F if fresh else None,
                                         Z=None if Z is None else np.ascontiguousarray(Z[lo:hi]), S=hi - lo, seed=int(seed) + rank, **kw)
        self._factor_on_device = True
        self.cd_stats = res["stats"]
        f0m, mvm = self._mask_failed(res["f0"], res["maxviol"], res["stats"])
        return self._assign(res["X"], f0m, mvm)

    # ---- improve (qcqp.py:403-432) -----------------------------------------------------------------------
    def _improve(self, method, *args, **kwargs):
        X0 = self.X
        R = X0.shape[0]
        cd_base = None
        if method == s.COORD_DESCENT:
            # under torchrun every rank must consume the identically seeded global stream identically: the base seed is drawn
            # BEFORE the empty-shard return, and a sharded run never reads or overwrites the global state per rank
            seed = kwargs.pop("seed", None)
            single_stream = seed is None and R == 1 and self._shard is None
            if not single_stream:
                cd_base = int(seed) if seed is not None else int(np.random.randint(0, 2 ** 31 - 1))
        if R == 0 and method in (s.COORD_DESCENT, s.ADMM):     # a rank whose shard of the batch is empty only joins the best-pick
            return self._assign(X0, np.zeros(0), np.zeros(0))
        if method == s.COORD_DESCENT:
            kw = dict(num_iters=kwargs.get('num_iters', 1000), viol_tol=kwargs.get('viol_tol', 1e-2), tol=kwargs.get('tol', 1e-4),
                      phase1=kwargs.get('phase1', True), strict=kwargs.get('strict', False))
            if single_stream:
                rng = engine.rng_states(states=[np.random.get_state()])   # the reference's process-global stream
            else:
                base = cd_base
                lo = self._shard[0] if self._shard is not None else 0       # restart r of the whole batch owns stream base + r
                rng = engine.rng_states(seeds=[(base + lo + r) % (2 ** 32) for r in range(R)])
            X, f0, mv, stats = self._pack.cd_improve(X0, rng, **kw)
            f0, mv = self._mask_failed(f0, mv, stats)
            if single_stream:
                np.random.set_state(engine.rng_state_tuple(rng[0]))
            self.cd_stats = stats
            return self._assign(X, f0, mv)
        if method == s.ADMM:
            num_iters = kwargs.get('num_iters', 1000)
            viol_lim = kwargs.get('viol_lim', 1e4)
            tol = kwargs.get('tol', 1e-2)
            rho = kwargs.get('rho', None)
            phase1 = kwargs.get('phase1', True)
            P0 = np.asarray(self.qcqp_form.f0.P.todense())
            lmb_min = float(np.min(np.linalg.eigh(P0)[0]))
            m = self.qcqp_form.m
            rhos = None if rho is None else np.atleast_1d(np.asarray(rho, dtype=np.float64))
            if rhos is not None:
                for rv in rhos:
                    if lmb_min + m * rv < 0:
                        log.error("rho parameter is too small, z-update not convex.")
                        raise Exception("rho parameter is too small, need at least %.3f." % rv)   # sic, qcqp.py:268
            else:
                rv = 2. * (1. - lmb_min) / m if lmb_min < 0 else 1. / m
                rv *= 50.
                log.warning("Automatically setting rho to %.3f", rv)
                rhos = np.array([rv])
            X, f0, mv, stats = self._pack.admm_improve(X0, rhos, num_iters=num_iters, viol_lim=viol_lim, tol=tol, phase1=phase1)
            self.admm_stats = stats
            self.admm_rhos = rhos
            return self._assign(X.reshape(-1, self.n), f0.ravel(), mv.ravel())
        if method == s.DCCP:
            raise Exception("DCCP package is not installed.")       # qcqp.py:289-292; third-party wrapper, out of scope
        if method == s.IPOPT:
            raise Exception("PyIpopt package is not installed.")    # qcqp.py:326-329; third-party wrapper, out of scope

    def improve(self, method, *args, **kwargs):
        if not isinstance(method, list):
            methods = [method]
        else:
            methods = method
        if not all([mm in s.improve_methods for mm in methods]):
            raise Exception("Unknown improve method(s): ", methods)
        if self._cvx_vars is not None and self._shard is None and (self.X is None or self.X.shape[0] == 1) \
                and all(v.value is not None for v in self._cvx_vars):
            # single-point mode starts from the variables' current values, as the reference does (flatten_vars, qcqp.py:404),
            # so a point the user wrote into `x.value` is the one that gets improved
            self.X = model.flatten_vars(self._cvx_vars, self.n).reshape(1, self.n)
        if self.X is None:
            # the reference means to start from suggest() when no point exists (qcqp.py:427; its test on Variable objects
            # never fires -- SURVEY H7 -- the intent is kept here)
            self.suggest(samples=int(kwargs.pop("restarts", 1)))
        restarts = kwargs.pop("restarts", None)
        held = self._shard[2] if self._shard is not None else self.X.shape[0]
        if restarts is not None and held != int(restarts):
            raise Exception("restarts=%d but the current batch holds %d points; call suggest(samples=%d) first"
                            % (int(restarts), held, int(restarts)))
        for mm in methods:
            f, v = self._improve(mm, *args, **kwargs)
        return (f, v)
