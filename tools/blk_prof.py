"""Cycle profile of cd_blk_kernel's phase 1 on the C5 instance (restart 0), from a -DBLK_PROF build of the library
(qcqp_b200/libqcqp_b200_prof.so; build recipe in tools/README.md).  Diagnostic only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if os.environ.get("BLK_LIB"):                    # another build of the library, by file name inside qcqp_b200/ (A/B timing)
    os.environ["QCQP_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qcqp_b200", os.environ["BLK_LIB"])
elif os.environ.get("BLK_PROF", "1") != "0":     # BLK_PROF=0: the shipped library (timing / ncu)
    os.environ["QCQP_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qcqp_b200", "libqcqp_b200_prof.so")
from qcqp_b200 import engine, problems as pb
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
forms, _ = pb.circle_packing(ncirc=200)
pack = engine.Pack(forms)
n = pack.n
X0 = np.stack([np.random.RandomState(r).randn(n) for r in range(R)])
for rep in range(3):
    rng = engine.rng_states(seeds=np.arange(R))
    t0 = time.perf_counter()
    X, f0, mv, st = pack.cd_improve(X0, rng, num_iters=iters)
    print("R=%d, %d sweeps: %.1f ms" % (R, iters, (time.perf_counter() - t0) * 1e3), flush=True)
