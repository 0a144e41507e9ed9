"""A/B of phase 1 of the separable dense path: bisections computed ahead by lpc_p1_pre_kernel (default) vs inside the sweep kernel
(QCQP_LPC_PRE=0): bit equality and the phase-1 time."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qcqp_b200 import _lib, engine, problems as pb
L = _lib.load()
for n, R in ((1000, 1024), (333, 200), (130, 64)):
    forms, _ = pb.boolean_least_squares(n, int(1.5 * n), seed=1)
    pack = engine.Pack(forms)
    mu, _S, F = engine.sdr_factor(pb.synthetic_sdr_solution(n, rank=16, seed=5))
    Z = np.random.RandomState(2).standard_normal((R, n))
    X0, _f, _v = pack.sdr_sample_eval(mu, F, Z=Z)
    X0[:, 3] = 1.0 + 1.0101e-2                     # stuck coordinates (|x^2 - 1| in (viol_tol, viol_tol + tol]): more than one sweep
    out = {}
    for mode in ("1", "0"):
        os.environ["QCQP_LPC_PRE"] = mode
        for rep in range(3):
            rng = engine.rng_states(seeds=1000 + np.arange(R))
            X, f0, mv, st = pack.cd_improve(X0, rng, num_iters=5)
        ms = (C.c_double * 4)(); cnt = C.c_int32(0)
        _lib.check(L.qcqp_cd_get_timing(pack.handle, ms, C.byref(cnt)))
        out[mode] = (X.copy(), f0.copy(), [(s.steps_p1, s.steps_p2, s.sweeps_p1, s.steps_skipped) for s in st], [r.pos for r in rng], ms[0])
    a, b = out["1"], out["0"]
    print("n=%d R=%d: phase-1 %.3f ms (ahead) vs %.3f ms (in-sweep); identical: X %s f0 %s stats %s pos %s" %
          (n, R, a[4], b[4], np.array_equal(a[0], b[0]), np.array_equal(a[1], b[1]), a[2] == b[2], a[3] == b[3]))
    pack.close()
