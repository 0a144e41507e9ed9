"""C4 (beamforming N=128, 32 constraints, 16 rho values): per-run iteration counts and results of the two ADMM kernels
next to the CPU oracle's (tests/golden/c4_admm_oracle.json)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qcqp_b200 import engine, problems as pb
g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "c4_admm_oracle.json")))
forms, _ = pb.beamforming(n=64, m=24, l=8, seed=1)
pack = engine.Pack(forms)
pack.compute_eig()
rhos = np.array(g["rhos"]); X0 = np.array(g["x0"])[None, :]
for kern in ("resident", "run")[:int(os.environ.get("NKERN", 2))]:
    os.environ["QCQP_ADMM_KERNEL"] = kern
    for rep in range(2):
        t0 = time.perf_counter(); X, f0, mv, st = pack.admm_improve(X0, rhos); dt = time.perf_counter() - t0
    print(kern, "%.4f s" % dt)
    for k in range(len(rhos)):
        print("  rho %8.4f  p1 %4d/%4d  p2 %4d/%4d  f0 %.12g / %.12g  mv %.3e / %.3e" % (rhos[k], st[k].iters_p1, g["iters_p1"][k], st[k].iters_p2, g["iters_p2"][k],
              f0[k, 0], g["f0"][k], mv[k, 0], g["maxviol"][k]))
