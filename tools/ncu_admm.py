"""One C4 ADMM launch (beamforming N=128, 32 constraints, 16 rho values) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qcqp_b200 import engine, problems as pb
forms, _ = pb.beamforming(n=64, m=24, l=8, seed=1)
pack = engine.Pack(forms); pack.compute_eig()
rhos = np.sqrt(32) * 2.0 ** (np.arange(-8, 8) / 2.0)
np.random.seed(4); X0 = 2 * np.random.randn(1, 128)
X, f0, mv, st = pack.admm_improve(X0, rhos, num_iters=int(os.environ.get("ITERS", 1000)))
print(sum(s.iters_p1 + s.iters_p2 for s in st))
