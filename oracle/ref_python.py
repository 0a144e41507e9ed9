"""The UNMODIFIED Python reference (cvxgrp/qcqp 0.8.3) as a timed CPU baseline -- TEST / BENCH INFRASTRUCTURE ONLY.

`__graft_entry__.build()` installs the reference's own package directory, untouched, under the git-ignored `baseline/_ref/`
(it travels to the GPU box with the snapshot; `pip install --target baseline/_ref` of the reference fails in this image because
its setup.py asks for `use_2to3`, which current setuptools rejects, so the pure-Python package is copied as pip would have laid it
out).  This module imports it from there behind a stub `cvxpy` (the hot path is NumPy/SciPy only; cvxpy 0.4 is not installable
here -- see oracle/ref_harness.py) and times *windows* of its coordinate-descent loop:

    coord_descent_phase2(x0, prob, num_iters=1)          (qcqp.py:152-178, the reference's own loop, unmodified)

is entered with the reference's own QCQPForm / QuadraticFunction objects; the objective is wrapped in a proxy that only COUNTS
the calls of `get_onevar_func` (one per coordinate step, qcqp.py:163) and raises after K of them, so that a window of exactly K
coordinate steps of the reference's code is timed (a whole sweep at n = 1000 costs ~110 s per core).  One worker process per
host core, each on its own restart.  Only bench.py (`--impl reference` and the `cpu_baseline` leg) and tests/ use this module.
"""
import importlib.util
import multiprocessing as mp
import os
import sys
import tempfile
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]


def ref_root():
    for d in REF_DIRS:
        if os.path.isfile(os.path.join(d, "qcqp", "qcqp.py")):
            return d
    return None


def load_reference():
    """(utilities, qcqp) modules of the reference, imported from baseline/_ref (else /root/reference)."""
    root = ref_root()
    if root is None:
        raise RuntimeError("the reference package is neither under baseline/_ref nor at /root/reference")
    if "qcqp.qcqp" in sys.modules:
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
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="qcqp_ref_"))       # the reference opens ./qcqp.log at import time (qcqp.py:39)
    sys.path.insert(0, root)
    try:
        import qcqp.utilities as u
        import qcqp.qcqp as q
    finally:
        sys.path.remove(root)
        os.chdir(cwd)
    return u, q


def plain_module(name):
    """qcqp_b200/<name>.py loaded as a stand-alone module (they are pure NumPy), WITHOUT importing the qcqp_b200 package --
    whose __init__ dlopens the product library; the CPU arms must not map it."""
    key = "_qcqp_plain_" + name
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(ROOT, "qcqp_b200", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod


class _WindowDone(Exception):
    pass


class _CountingObjective(object):
    """Delegates everything to the reference's QuadraticFunction; counts get_onevar_func calls (= coordinate steps of phase 2)."""

    def __init__(self, f, limit):
        self.__dict__["_f"], self.__dict__["_limit"], self.__dict__["_count"] = f, limit, 0
        self.__dict__["_t0"] = self.__dict__["_t1"] = None

    def get_onevar_func(self, x, k):
        d = self.__dict__
        if d["_count"] == 0:
            d["_t0"] = time.perf_counter()
        if d["_count"] == d["_limit"]:
            d["_t1"] = time.perf_counter()
            raise _WindowDone()
        d["_count"] += 1
        return d["_f"].get_onevar_func(x, k)

    def __getattr__(self, name):
        return getattr(self.__dict__["_f"], name)

    def __setattr__(self, name, value):
        setattr(self.__dict__["_f"], name, value)


_W = {}


def _worker_init(gen, gargs):
    import scipy.sparse as sp
    u, q = load_reference()
    pb = plain_module("problems")
    forms, _info = getattr(pb, gen)(**gargs)
    fs = []
    for (P, qv, r, relop) in forms:
        n = len(qv)
        fs.append(u.QuadraticFunction(sp.csr_matrix(P), sp.csc_matrix(np.asarray(qv, dtype=float).reshape(n, 1)), float(r), relop))
    _W.update(u=u, q=q, fs=fs, n=len(forms[0][1]))


def _worker_window(args):
    """K coordinate steps of the reference's coord_descent_phase2 from a +-sqrt(1 + 5e-3) point (golden G2' construction: phase 2
    moves only from a slightly infeasible point, SURVEY a-9).  Returns (steps, seconds)."""
    seed, K = args
    u, q, fs, n = _W["u"], _W["q"], _W["fs"], _W["n"]
    rs = np.random.RandomState(seed)
    x0 = np.sign(rs.randn(n)) * np.sqrt(1 + 5e-3 * (np.arange(n) % 10 + 1) / 10.0)
    f0 = _CountingObjective(fs[0], K)
    prob = u.QCQPForm(f0, fs[1:])
    np.random.seed(seed)
    try:
        q.coord_descent_phase2(x0, prob, num_iters=1)
        d = f0.__dict__
        return d["_count"], time.perf_counter() - d["_t0"]         # the sweep ended before the window did
    except _WindowDone:
        d = f0.__dict__
        return d["_count"], d["_t1"] - d["_t0"]


class ReferencePool:
    """One worker process per core, each holding the reference's objects of one generator instance."""

    def __init__(self, gen, gargs, procs=None):
        self.procs = int(procs or os.cpu_count() or 1)
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.procs, initializer=_worker_init, initargs=(gen, gargs))

    def window(self, steps_per_proc, seed0=1000):
        """Every worker runs `steps_per_proc` coordinate steps of its own restart; returns (total steps, wall seconds of the slowest
        worker's window, per-worker seconds)."""
        res = self.pool.map(_worker_window, [(seed0 + i, steps_per_proc) for i in range(self.procs)])
        steps = sum(r[0] for r in res)
        secs = [r[1] for r in res]
        return steps, max(secs), secs

    def close(self):
        self.pool.terminate()
        self.pool.join()


if __name__ == "__main__":
    t = time.time()
    pool = ReferencePool("boolean_least_squares", dict(n=int(sys.argv[1]) if len(sys.argv) > 1 else 200, m=300, seed=1), procs=2)
    print("setup %.1f s" % (time.time() - t))
    print(pool.window(8))
    pool.close()
