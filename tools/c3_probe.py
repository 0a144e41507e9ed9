"""A/B of the sparse-objective phase-2 loop of cd_lpc_kernel (MAXCUT): certified quiet-window filter + prefetched constants (default)
vs the plain loop (QCQP_LPC_QUIET=0): bit equality and time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qcqp_b200 import engine, problems as pb
for n, p, R, iters in ((2000, 0.1, 256, 1000), (300, 0.1, 64, 200), (61, 0.3, 33, 100)):
    forms, _ = pb.maxcut(n, p, seed=1)
    pack = engine.Pack(forms)
    X0 = np.random.RandomState(3).randn(R, n)
    out = {}
    for mode in ("1", "0"):
        os.environ["QCQP_LPC_QUIET"] = mode
        best = 1e9
        for rep in range(2):
            rng = engine.rng_states(seeds=1000 + np.arange(R))
            t0 = time.perf_counter()
            X, f0, mv, st = pack.cd_improve(X0, rng, num_iters=iters)
            best = min(best, time.perf_counter() - t0)
        out[mode] = (X.copy(), f0.copy(), [(s.steps_p1, s.steps_p2, s.updates_p2, s.sweeps_p2) for s in st], [r.pos for r in rng], best)
    a, b = out["1"], out["0"]
    sw = sum(s[1] for s in a[2]) / float(n)
    print("maxcut n=%d R=%d (<=%d sweeps, %.0f restart-sweeps): %.2f ms (filter) vs %.2f ms (plain); identical: X %s f0 %s stats %s pos %s"
          % (n, R, iters, sw, a[4] * 1e3, b[4] * 1e3, np.array_equal(a[0], b[0]), np.array_equal(a[1], b[1]), a[2] == b[2], a[3] == b[3]))
    pack.close()
