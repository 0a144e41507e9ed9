"""First-contact script for a GPU box: runs the CD engine on a few instances next to the oracle and prints the
differences and timings (diagnostics; the assertions live in tests/)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import oracle as orc
from qcqp_b200 import engine, problems as pb


def compare(name, forms, X0, seeds, strict, **kw):
    P = orc.Problem(forms)
    pack = engine.Pack(forms)
    rng = engine.rng_states(seeds=seeds)
    t0 = time.time()
    Xg, fg, vg, sg = pack.cd_improve(X0, rng, strict=strict, **kw)
    tg = time.time() - t0
    worst = 0.0
    for r in range(X0.shape[0]):
        st = orc.RngState.from_seed(int(seeds[r]))
        xo, so = P.improve_cd(X0[r], st, fast=True, **kw)
        fo, vo = P.eval(0, xo), P.max_violation(xo)
        d = abs(fg[r] - fo) / max(1e-300, abs(fo))
        worst = max(worst, d)
        if r < 3 or d > 1e-6:
            print("  %s r=%d strict=%d f0 gpu=%.15g orc=%.15g rel=%.2e | viol %.6g %.6g | steps gpu=(%d,%d) orc=(%d,%d) pos %d %d status %d"
                  % (name, r, strict, fg[r], fo, d, vg[r], vo, sg[r].steps_p1, sg[r].steps_p2, so.steps_p1, so.steps_p2, rng[r].pos, st.pos, sg[r].status))
    print("%s strict=%d: R=%d worst rel diff %.3e, gpu call %.3fs" % (name, strict, X0.shape[0], worst, tg))
    pack.close()


if __name__ == "__main__":
    rs = np.random.RandomState(0)
    forms, _ = pb.boolean_least_squares(10, 15)
    compare("bls10", forms, rs.randn(4, 10), 1000 + np.arange(4), True)
    compare("bls10", forms, rs.randn(4, 10), 1000 + np.arange(4), False)
    forms, _ = pb.boolean_least_squares(100, 150)
    compare("bls100", forms, rs.randn(8, 100), 2000 + np.arange(8), True)
    compare("bls100", forms, rs.randn(8, 100), 2000 + np.arange(8), False)
    forms, _ = pb.maxcut(40, 0.15, seed=2)
    compare("maxcut40", forms, rs.randn(4, 40), 3000 + np.arange(4), True, num_iters=30)
    forms, _ = pb.circle_packing(4)
    compare("circle4", forms, np.abs(rs.randn(4, 9)) * 3 + 0.5, 4000 + np.arange(4), True, num_iters=8)
    forms, _ = pb.random_qcqp(8, 6, seed=3)
    compare("random8", forms, rs.randn(4, 8), 5000 + np.arange(4), True, num_iters=10)
    # size: C2-like
    forms, _ = pb.boolean_least_squares(1000, 1500)
    pack = engine.Pack(forms)
    print("C2 pack info: n_dense", pack.info.n_dense, "bytes/sweep p2", pack.info.bytes_per_sweep_phase2, "p1", pack.info.bytes_per_sweep_phase1)
    for R in (148, 1024):
        X0 = rs.randn(R, 1000)
        rng = engine.rng_states(seeds=np.arange(R))
        t0 = time.time()
        Xg, fg, vg, sg = pack.cd_improve(X0, rng)
        dt = time.time() - t0
        steps = sum(s.steps_p1 + s.steps_p2 for s in sg)
        print("   skipped steps", sum(s.steps_skipped for s in sg), "p1 sweeps mean", np.mean([s.sweeps_p1 for s in sg]), "stuck", sum(1 for s in sg if s.steps_skipped > 0))
        sw2 = np.mean([s.sweeps_p2 for s in sg])
        print("C2 R=%d: %.3fs wall (incl copies), total steps %d = %.1f restart-sweeps -> %.0f restart-sweeps/s; mean p2 sweeps %.1f; f0 best %.6g maxviol max %.3g"
              % (R, dt, steps, steps / 1000.0, steps / 1000.0 / dt, sw2, fg.min(), vg.max()))
