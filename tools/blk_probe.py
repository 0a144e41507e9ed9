"""Small run of the CTA-per-restart CD kernel (for compute-sanitizer and quick timings): circle packing, forced through cd_blk.cu."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qcqp_b200 import engine, problems as pb

ncirc = int(sys.argv[1]) if len(sys.argv) > 1 else 40
R = int(sys.argv[2]) if len(sys.argv) > 2 else 2
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
strict = int(sys.argv[4]) if len(sys.argv) > 4 else 5
forms, _ = pb.circle_packing(ncirc)
pack = engine.Pack(forms)
n = 2 * ncirc + 1
X0 = np.abs(np.random.RandomState(1).randn(R, n)) * 3 + 0.5
p2only = len(sys.argv) > 5 and sys.argv[5] == "grid"
if p2only:
    # strictly feasible start (circles on a jittered grid, small radius): phase 2 only
    g = int(np.ceil(np.sqrt(ncirc)))
    rs = np.random.RandomState(2)
    X0 = np.zeros((R, n))
    for r in range(R):
        idx = rs.permutation(g * g)[:ncirc]
        cx = (idx % g + 0.5) * (10.0 / g) + 0.05 * rs.randn(ncirc) / g
        cy = (idx // g + 0.5) * (10.0 / g) + 0.05 * rs.randn(ncirc) / g
        X0[r, 0] = 0.2 * (10.0 / g)
        X0[r, 1::2] = cx; X0[r, 2::2] = cy
for rep in range(2):
    rng = engine.rng_states(seeds=np.arange(R))
    t0 = time.perf_counter()
    X, f0, mv, st = pack.cd_improve(X0, rng, num_iters=iters, strict=strict, phase1=not p2only)
    dt = time.perf_counter() - t0
sw = sum(s.steps_p1 + s.steps_p2 for s in st) / float(n)
sw1 = sum(s.steps_p1 for s in st) / float(n); up2 = sum(s.updates_p2 for s in st)
print("p1 sweeps %.1f p2 sweeps %.1f p2 moves %d feasible %d" % (sw1, sw - sw1, up2, int((mv < 1e-2).sum())))
print("circle %d R=%d iters=%d strict=%d: %.4f s, %.1f restart-sweeps -> %.1f /s; maxviol %.3g; f0[0]=%.12g pos %d" % (ncirc, R, iters, strict, dt, sw, sw / dt, mv.max(), f0[0], rng[0].pos))
pack.close()
